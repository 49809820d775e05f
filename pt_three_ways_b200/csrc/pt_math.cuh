// Device arithmetic of the B200 path tracer.
//
// Everything here computes in IEEE-754 binary64 with a FIXED rounding sequence: the library is
// compiled with --fmad=false, so a fused multiply-add happens exactly where fma() is written
// and nowhere else; / and sqrt are correctly rounded; sin/cos/acos are our own fixed
// polynomial kernels rather than the CUDA math library's.  The sequence restates the
// reference's formulas (file:line cited per function, relative to mattgodbolt/pt-three-ways
// @ a4aeda0) — the reference itself pins no rounding order because it is built with
// -funsafe-math-optimizations (CMakeLists.txt:21).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace ptb200 {

constexpr double kEpsilon = 0.000000001; // src/math/Epsilon.h:3
constexpr double kPi = 3.14159265358979323846;

struct V3 {
  double x, y, z;
};

__device__ __forceinline__ V3 mk(double x, double y, double z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 add(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 scale(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 neg(V3 a) { return mk(-a.x, -a.y, -a.z); }
// Vec3::dot (src/math/Vec3.h:83-85).
__device__ __forceinline__ double dot(V3 a, V3 b) {
  return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x));
}
// Vec3::cross (src/math/Vec3.h:87-92).
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)),
            fma(a.x, b.y, -(a.y * b.x)));
}
// Correctly rounded sqrt and division expand to ~50 SASS instructions each (MUFU seed, Newton
// steps, exponent fix-ups).  In the one-kernel forms (pt_kernels.cu) they are called out of line
// everywhere except the innermost triangle loop: same IEEE results, but the megakernel's code stays
// inside the instruction cache.  PT_INLINE_LEVEL (set by the including .cu file) inlines 1: these,
// 2: + the sin/cos pair, 3: + Philox, 4: + cone sampling; the three-kernel pipeline uses 3.
#ifndef PT_INLINE_LEVEL
#define PT_INLINE_LEVEL 0
#endif
#if PT_INLINE_LEVEL >= 1
#define PT_IEEE_ATTR __forceinline__
#else
#define PT_IEEE_ATTR __noinline__
#endif
#if PT_INLINE_LEVEL >= 2
#define PT_SINCOS_ATTR __forceinline__
#else
#define PT_SINCOS_ATTR __noinline__
#endif
#if PT_INLINE_LEVEL >= 3
#define PT_PHILOX_ATTR __forceinline__
#else
#define PT_PHILOX_ATTR __noinline__
#endif
#if PT_INLINE_LEVEL >= 4
#define PT_CONE_ATTR __forceinline__
#else
#define PT_CONE_ATTR __noinline__
#endif
static __device__ PT_IEEE_ATTR double ieeeSqrt(double x) { return sqrt(x); }
static __device__ PT_IEEE_ATTR double ieeeDiv(double a, double b) { return a / b; }
static __device__ PT_IEEE_ATTR double ieeeRcpSqrt(double x) { return 1.0 / sqrt(x); }
// Two independent square roots in one call: the same correctly rounded results, two Newton chains
// to interleave (hemisphereSample's sqrt(v) and sqrt(1 - v), Samples.cpp:23,28).
static __device__ PT_IEEE_ATTR double2 ieeeSqrtPair(double a, double b) { return make_double2(sqrt(a), sqrt(b)); }

// Vec3::normalised (src/math/Vec3.impl.h:5-7): *this / length(), and operator/ multiplies by
// the reciprocal (src/math/Vec3.h:51-54).
__device__ __forceinline__ V3 normalised(V3 a) {
  const double reciprocal = ieeeRcpSqrt(dot(a, a));
  return scale(a, reciprocal);
}
// Ray::positionAlong (src/math/Ray.h:25-27).
__device__ __forceinline__ V3 positionAlong(V3 o, V3 d, double t) {
  return mk(fma(d.x, t, o.x), fma(d.y, t, o.y), fma(d.z, t, o.z));
}

// ---- elementary functions ---------------------------------------------------------------
// Arguments on this path are bounded (angles in [-pi, 2*pi], acos on [0, 1)), so a
// two-constant quadrant reduction and the classic double-precision polynomial kernels
// suffice (< 1 ulp typical).  ~30 FP64 instructions for a sin/cos pair.
//
// A binary64 literal costs two move instructions every time it is materialised (FP64 instructions
// take no 64-bit immediates): 30 moves per sin/cos pair, 6.7 % of the sub-path kernel's issue slots
// on constants altogether (profiles/r2t).  Read from the constant bank, sixteen of them arrive in
// eight LDCU.128 and feed DFMA from uniform registers.  The same literals, so the same bits.
#ifndef PT_CONSTANTS_IN_BANK
#define PT_CONSTANTS_IN_BANK 1
#endif
#define PT_SINCOS_CONSTANTS                                                                        \
  1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06,             \
      -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01,        \
      -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07,        \
      2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02,         \
      6.36619772367581382433e-01, 1.57079632673412561417e+00, 6.07710050650619224932e-11, 0.0
#if PT_CONSTANTS_IN_BANK
static __constant__ __align__(16) double kSinCos[16] = {PT_SINCOS_CONSTANTS};
static __constant__ double kEpsilonBank = 0.000000001; // kEpsilon for the loops that would rematerialise it per trip
#define PT_EPSILON kEpsilonBank
#else
#define PT_EPSILON kEpsilon
#endif
__device__ __forceinline__ double kernelSin(double r, const double *c) { // c: six coefficients
  const double z = r * r;
  double p = fma(c[0], z, c[1]);
  p = fma(p, z, c[2]);
  p = fma(p, z, c[3]);
  p = fma(p, z, c[4]);
  p = fma(p, z, c[5]);
  return fma(r * z, p, r);
}
__device__ __forceinline__ double kernelCos(double r, const double *c) {
  const double z = r * r;
  double p = fma(c[0], z, c[1]);
  p = fma(p, z, c[2]);
  p = fma(p, z, c[3]);
  p = fma(p, z, c[4]);
  p = fma(p, z, c[5]);
  return fma(z * z, p, fma(-0.5, z, 1.0));
}
// Out of line (one copy for every caller), results returned BY VALUE: reference parameters of a
// non-inlined function travel through local memory.
static __device__ PT_SINCOS_ATTR double2 sinCosPair(double x) { // {sin x, cos x}
#if PT_CONSTANTS_IN_BANK
  const double *c = kSinCos;
#else
  const double c[16] = {PT_SINCOS_CONSTANTS};
#endif
  const double kd = rint(x * c[12]); // 2/pi; round half to even
  const int k = static_cast<int>(kd);
  double r = fma(-kd, c[13], x);     // pi/2, high part
  r = fma(-kd, c[14], r);            // pi/2, low part
  const double sr = kernelSin(r, c);
  const double cr = kernelCos(r, c + 6);
  const double a = (k & 1) ? cr : sr;
  const double b = (k & 1) ? sr : cr;
  return make_double2((k & 2) ? -a : a, ((k + 1) & 2) ? -b : b);
}
__device__ __forceinline__ void sinCos(double x, double &s, double &c) {
  const double2 pair = sinCosPair(x);
  s = pair.x;
  c = pair.y;
}
__device__ __forceinline__ double asinCore(double z) {
  double p = fma(3.47933107596021167570e-05, z, 7.91534994289814532176e-04);
  p = fma(p, z, -4.00555345006794114027e-02);
  p = fma(p, z, 2.01212532134862925881e-01);
  p = fma(p, z, -3.25565818622400915405e-01);
  p = fma(p, z, 1.66666666666666657415e-01);
  p = p * z;
  double q = fma(7.70381505559019352791e-02, z, -6.88283971605453293030e-01);
  q = fma(q, z, 2.02094576023350569471e+00);
  q = fma(q, z, -2.40339491173441421878e+00);
  q = fma(q, z, 1.0);
  return ieeeDiv(p, q);
}
__device__ __forceinline__ double arcCos(double x) { // x in [0, 1]
  if (x < 0.5) {
    const double r = asinCore(x * x);
    return 1.57079632679489655800e+00 - (x - fma(-x, r, 6.12323399573676603587e-17));
  }
  const double z = (1.0 - x) * 0.5;
  const double s = ieeeSqrt(z);
  const double r = asinCore(z);
  return 2.0 * fma(s, r, s);
}

// ---- random numbers ---------------------------------------------------------------------
// libstdc++ generate_canonical<double,53> over a 32-bit engine: two words, low word first,
// (lo + hi*2^32) / 2^64 with the >= 1 guard (bits/random.tcc:3349-3381 of GCC 13), which is
// what uniform_real_distribution<double> draws in Scene.cpp:149-161 and Camera.h:30-33,56-58.
__device__ __forceinline__ double canonicalFromWords(uint32_t lo, uint32_t hi) {
  const double sum = fma(static_cast<double>(hi), 4294967296.0, static_cast<double>(lo));
  double ret = sum * 5.42101086242752217003726400434970855712890625e-20; // 2^-64
  if (ret >= 1.0)
    ret = 0.99999999999999988897769753748434595763683319091796875;
  return ret;
}

struct Philox4 {
  uint32_t w[4];
};
constexpr uint32_t kPhiloxKeyHigh = 0xB200D0D0u;
// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1).
static __device__ PT_PHILOX_ATTR Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                      uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{{c0, c1, c2, c3}};
}

// Two blocks that differ in their last counter word only (c3 = 0 and 1: what one bounce or one camera
// ray draws), advanced together: one key schedule, one call, two independent chains to interleave.
struct Philox8 {
  uint32_t w[8];
};
static __device__ PT_PHILOX_ATTR Philox8 philox4x32_10_pair(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t k0,
                                                           uint32_t k1) {
  uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = 0u, b0 = c0, b1 = c1, b2 = c2, b3 = 1u;
#pragma unroll
  for (int round = 0; round < 10; ++round) {
    const uint32_t ahi0 = __umulhi(0xD2511F53u, a0), alo0 = 0xD2511F53u * a0;
    const uint32_t ahi1 = __umulhi(0xCD9E8D57u, a2), alo1 = 0xCD9E8D57u * a2;
    const uint32_t bhi0 = __umulhi(0xD2511F53u, b0), blo0 = 0xD2511F53u * b0;
    const uint32_t bhi1 = __umulhi(0xCD9E8D57u, b2), blo1 = 0xCD9E8D57u * b2;
    a0 = ahi1 ^ a1 ^ k0;
    a2 = ahi0 ^ a3 ^ k1;
    a1 = alo1;
    a3 = alo0;
    b0 = bhi1 ^ b1 ^ k0;
    b2 = bhi0 ^ b3 ^ k1;
    b1 = blo1;
    b3 = blo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox8{{a0, a1, a2, a3, b0, b1, b2, b3}};
}

// ---- shading helpers --------------------------------------------------------------------
// Norm3::reflectance (src/math/Norm3.cpp:7-24).  rParallel is evaluated with the
// rPerpendicular formula there, so (rPerp^2 + rPar^2)/2 == rPerp^2 exactly.
__device__ __forceinline__ double reflectance(V3 normal, V3 incoming, double iorFrom,
                                              double iorTo, double iorRatio /* iorFrom / iorTo */) {
  const double cosThetaI = -dot(normal, incoming);
  const double sinThetaTSquared = (iorRatio * iorRatio) * fma(-cosThetaI, cosThetaI, 1.0);
  if (sinThetaTSquared > 1)
    return 1.0;
  const double cosThetaT = ieeeSqrt(1 - sinThetaTSquared);
  const double a = iorFrom * cosThetaI;
  const double b = iorTo * cosThetaT;
  const double rPerpendicular = ieeeDiv(a - b, a + b);
  return rPerpendicular * rPerpendicular;
}
// Norm3::reflect (src/math/Norm3.impl.h:41-44).
__device__ __forceinline__ V3 reflect(V3 normal, V3 incoming) {
  const double k = dot(normal, incoming);
  return mk(fma(-(normal.x * 2), k, incoming.x), fma(-(normal.y * 2), k, incoming.y),
            fma(-(normal.z * 2), k, incoming.z));
}
struct Basis {
  V3 x, y, z;
};
// OrthoNormalBasis::fromZ (src/math/OrthoNormalBasis.cpp:36-51).  The helper-axis cross
// product is written out: xAxis x z = (0, -z.z, z.y), yAxis x z = (z.z, 0, -z.x).
__device__ __forceinline__ Basis basisFromZ(V3 z) {
  const V3 c = fabs(z.x) > 0.9999 ? mk(z.z, 0.0, -z.x) : mk(0.0, -z.z, z.y);
  const V3 xx = normalised(c);
  const V3 yy = normalised(cross(z, xx));
  return Basis{xx, yy, z};
}
// OrthoNormalBasis::transform (src/math/OrthoNormalBasis.h:18-20).
__device__ __forceinline__ V3 transform(const Basis &b, V3 p) {
  return mk(fma(b.z.x, p.z, fma(b.y.x, p.y, b.x.x * p.x)),
            fma(b.z.y, p.z, fma(b.y.y, p.y, b.x.y * p.x)),
            fma(b.z.z, p.z, fma(b.y.z, p.y, b.x.z * p.x)));
}
// coneSample (src/math/Samples.cpp:6-19).  Rare (specular picks only): kept out of line so the
// hot loop's instruction footprint stays inside the instruction cache.
static __device__ PT_CONE_ATTR V3 coneSample(V3 direction, double coneTheta, double u, double v) {
  if (coneTheta < kEpsilon)
    return direction;
  coneTheta = coneTheta * (1.0 - ieeeDiv(2.0 * arcCos(u), kPi));
  double radius, zScale, sinT, cosT;
  sinCos(coneTheta, radius, zScale);
  const double randomTheta = v * 2 * kPi;
  sinCos(randomTheta, sinT, cosT);
  const Basis basis = basisFromZ(direction);
  return normalised(transform(basis, mk(cosT * radius, sinT * radius, zScale)));
}
// The specular-only part of coneSample(): everything up to the final
// normalised(basis.transform(cos(t)*r, sin(t)*r, z)), which has the same shape as
// hemisphereSample()'s and is therefore executed once, by specular and diffuse lanes together,
// in the megakernel.  Returns true when the cone is degenerate and `direction` is the answer.
static __device__ PT_CONE_ATTR bool coneSampleSetup(V3 direction, double coneTheta, double u, double v,
                                                    Basis &basis, double &randomTheta, double &radius,
                                                    double &zScale) {
  if (coneTheta < kEpsilon)
    return true;
  coneTheta = coneTheta * (1.0 - ieeeDiv(2.0 * arcCos(u), kPi));
  sinCos(coneTheta, radius, zScale);
  randomTheta = v * 2 * kPi;
  basis = basisFromZ(direction);
  return false;
}

// hemisphereSample (src/math/Samples.cpp:21-30).
__device__ __forceinline__ V3 hemisphereSample(const Basis &basis, double u, double v) {
  const double theta = (2 * kPi) * u;
  const double radius = ieeeSqrt(v);
  double sinT, cosT;
  sinCos(theta, sinT, cosT);
  return normalised(transform(basis, mk(cosT * radius, sinT * radius, ieeeSqrt(1 - v))));
}

} // namespace ptb200
