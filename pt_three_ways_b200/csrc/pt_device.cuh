// Device-side scene layout and the per-ray building blocks shared by every kernel:
// the brute-force primitive sweep (Scene.cpp:14-122 of the reference), the hit epilogues,
// camera ray generation and the bounce sampler.
#pragma once

#include "pt_math.cuh"

namespace ptb200 {

constexpr int kMaxDepth = 64;        // deepest supported RenderParams::maxDepth
constexpr uint32_t kFullMask = 0xffffffffu;

// ---- layout in HBM --------------------------------------------------------------------------
// Triangle data exists in three layouts, all holding v0, e1 = v1-v0, e2 = v2-v0 (what
// TriangleVertices::uVector()/vVector(), TriangleVertices.h:25-31, recompute on every call;
// storing them is bit-identical), padded per tile with all-zero triangles (det == 0 ->
// skipped, Scene.cpp:66-68):
//   triSweep   FP64, tile-major SoA (9 arrays of `tileTris` doubles per tile): what the FP64
//              sweep variants stage into shared memory with one TMA bulk copy per tile; every
//              lane of a warp reads the same triangle, i.e. 16-byte broadcast loads;
//   triFilter  FP32 copies + error bounds, blocked by four triangles: what the default FP32
//              stage-0 sweep stages into shared memory (same TMA path);
//   triExact   FP64 AoS records for the exact test of stage-0 survivors and for the sequential
//              kernel's lane-strided sweep (global memory / L1).
struct DeviceScene {
  const double *triSweep;     // [numTiles][9][tileTris]
  const double4 *triShade;    // [numTriangles][4] {normal xyz, material} {frontX xyz, frontY.x}
                              //   {frontY yz, backX xy} {backX z, backY xyz}: the shading normal and the
                              //   OrthoNormalBasis::fromZ of +normal / -normal, precomputed at upload
  const double4 *spheres;     // [numSpheres] {centre xyz, radius^2}   (Sphere.h:7-12)
  const uint32_t *sphereMaterial;
  const double *materials;    // [numMaterials][10] MaterialSpec order + 1/indexOfRefraction
  const float *triFilter;     // [numTiles][tileTris/4][14][4] fp32 stage-0 data (buildFilterKernel)
  const float *triMoment;     // [numTiles][tileTris/4][19][4] fp32 stage-0 data in moment (Pluecker) form
  const uint32_t *fanMask;    // bit per group of four triangles (tile-major): the group is two quads in fan
                              //   order and its triMoment lanes are stored [A0, A1, B0, B1] (stage0RejectFan4)
  const double *triExact;     // [numTiles*tileTris][10] AoS copy of the 9 sweep doubles (+pad) for the
                              //   survivors' exact test: one address, five 16-byte loads
  uint32_t numTriangles;
  uint32_t numSpheres;
  uint32_t tileTris;          // triangles per tile (multiple of 4)
  uint32_t numTiles;
  double environment[3];
};

struct DeviceCamera { // PtCamera / Camera.h:11-18
  V3 centre, axisX, axisY, axisZ;
  double aspectRatio, cameraPlaneDist, reciprocalHeight, reciprocalWidth, apertureRadius,
      focalDistance;
};

struct Nearest {
  double t;     // currentNearestDist
  double det;   // determinant of the winning triangle (for `backfacing`, Scene.cpp:108)
  int prim;     // INT_MAX-like "none" = kNoPrim; >= 0 triangle; < 0 sphere -(i+1)
};
constexpr int kNoPrim = 0x7fffffff;

struct HitInfo {
  V3 position, normal;
  uint32_t material;
  int triangle;        // >= 0: the hit triangle (its precomputed basis can be loaded), else -1
  bool inside;
};

__device__ __forceinline__ double4 ldgDouble4(const double4 *p) {
  const double2 lo = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 hi = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(lo.x, lo.y, hi.x, hi.y);
}

// ---- sphere sweep (Scene.cpp:14-37), spheres in shared memory --------------------------------
__device__ __forceinline__ void sweepSpheres(const double4 *__restrict__ spheres, int numSpheres,
                                             V3 o, V3 d, Nearest &best) {
#pragma unroll 1
  for (int i = 0; i < numSpheres; ++i) {
    const double4 s = spheres[i];
    const V3 op = sub(mk(s.x, s.y, s.z), o);
    const double b = dot(op, d);
    double determinant = fma(b, b, -dot(op, op)) + s.w;
    if (determinant < 0)
      continue;
    determinant = ieeeSqrt(determinant);
    const double minusT = b - determinant;
    const double plusT = b + determinant;
    const double epsilon = PT_EPSILON;
    if (minusT < epsilon && plusT < epsilon)
      continue;
    const double t = minusT > epsilon ? minusT : plusT;
    if (t < best.t) {
      best.t = t;
      best.prim = -(i + 1);
    }
  }
}

// ---- one ray against one triangle, the reference's Moller-Trumbore (Scene.cpp:62-98) -------
// kFpWay: fp::Triangle::intersect (src/fp/Triangle.cpp:9-41) is the same arithmetic but rejects
// `t < Epsilon` where Scene.cpp:94 accepts `t > Epsilon`, i.e. t == Epsilon is a hit there.
template <bool kFpWay = false>
__device__ __forceinline__ void testTriangle(V3 v0, V3 e1, V3 e2, V3 o, V3 d, int index,
                                             Nearest &best) {
  const V3 pVec = cross(d, e2);
  const double det = dot(e1, pVec);
  const double invDet = 1.0 / det;
  const V3 tVec = sub(o, v0);
  const double u = dot(tVec, pVec) * invDet;
  const V3 qVec = cross(tVec, e1);
  const double v = dot(d, qVec) * invDet;
  const double t = dot(e2, qVec) * invDet;
  // `continue` conditions of Scene.cpp:67,89 and the acceptance test of :94, as one predicate.
  const double epsilon = PT_EPSILON; // (from the constant bank: one load instead of two moves per trip)
  const bool reject = (fabs(det) < epsilon) | (u < 0.0) | (u > 1.0) | (v < 0.0) | (u + v > 1);
  const bool accept = !reject & (kFpWay ? t >= epsilon : t > epsilon) & (t < best.t);
  if (accept) {
    best.t = t;
    best.det = det;
    best.prim = index;
  }
}

// Sweeps triangles [0, count) of a tile that sits in shared memory; `count` is even.
// Every lane of a warp reads the same addresses (broadcast); two triangles per iteration so
// each of the 9 arrays is read with one 16-byte load.
template <bool kFpWay = false>
__device__ __forceinline__ void sweepTile(const double *__restrict__ tile, int tileTris, int count,
                                          int firstIndex, V3 o, V3 d, Nearest &best) {
#pragma unroll 1
  for (int i = 0; i < count; i += 2) {
    const double2 v0x = *reinterpret_cast<const double2 *>(tile + 0 * tileTris + i);
    const double2 v0y = *reinterpret_cast<const double2 *>(tile + 1 * tileTris + i);
    const double2 v0z = *reinterpret_cast<const double2 *>(tile + 2 * tileTris + i);
    const double2 e1x = *reinterpret_cast<const double2 *>(tile + 3 * tileTris + i);
    const double2 e1y = *reinterpret_cast<const double2 *>(tile + 4 * tileTris + i);
    const double2 e1z = *reinterpret_cast<const double2 *>(tile + 5 * tileTris + i);
    const double2 e2x = *reinterpret_cast<const double2 *>(tile + 6 * tileTris + i);
    const double2 e2y = *reinterpret_cast<const double2 *>(tile + 7 * tileTris + i);
    const double2 e2z = *reinterpret_cast<const double2 *>(tile + 8 * tileTris + i);
    testTriangle<kFpWay>(mk(v0x.x, v0y.x, v0z.x), mk(e1x.x, e1y.x, e1z.x), mk(e2x.x, e2y.x, e2z.x), o, d,
                         firstIndex + i, best);
    testTriangle<kFpWay>(mk(v0x.y, v0y.y, v0z.y), mk(e1x.y, e1y.y, e1z.y), mk(e2x.y, e2y.y, e2z.y), o, d,
                         firstIndex + i + 1, best);
  }
}

// ---- two-stage sweep: division-free prefilter, exact test for the survivors -------------------
// Stage 1 evaluates, for every triangle, the reference's own numerators X = tVec.pVec and
// Y = dir.qVec and the determinant (24 FP64 instructions), and keeps the triangle unless it is
// PROVABLY rejected by Scene.cpp:67,89.  With s = sign(det), Xs = s*X, Ys = s*Y:
//   |det| < Epsilon         is only used when it is decided by the high 32 bits of |det|;
//   u = rn(X*rn(1/det)) < 0 needs Xs < 0: rounding cannot change a sign, and the 1e-300 guard
//                           (again on the high word) covers an underflow of the product to -0;
//   v < 0                   likewise for Ys;
//   u + v > 1               needs Xs + Ys > |det|*(1 + 2^-49) at least, so testing against
//                           hi = |det|*(1 + 2^-40) can only keep MORE triangles; u > 1 is
//                           implied by it once v >= 0.
// Sign handling and the three "high word" tests are integer instructions (ALU pipe), leaving
// 27 FP64 instructions per triangle.  Survivors (a handful per ray) are re-tested with the exact
// reference arithmetic (testTriangle) in index order, so the nearest-hit decision is identical
// to the one-stage sweep bit for bit; the division and the scaled products leave the hot loop.
__device__ __forceinline__ bool prefilterTriangle(V3 v0, V3 e1, V3 e2, V3 o, V3 d) {
  const V3 pVec = cross(d, e2);
  const double det = dot(e1, pVec);
  const V3 tVec = sub(o, v0);
  const double x = dot(tVec, pVec);
  const V3 qVec = cross(tVec, e1);
  const double y = dot(d, qVec);
  const uint32_t detHi = static_cast<uint32_t>(__double2hiint(det));
  const uint32_t sign = detHi & 0x80000000u;
  const uint32_t xsHi = static_cast<uint32_t>(__double2hiint(x)) ^ sign;
  const uint32_t ysHi = static_cast<uint32_t>(__double2hiint(y)) ^ sign;
  const double xs = __hiloint2double(static_cast<int>(xsHi), __double2loint(x));
  const double ys = __hiloint2double(static_cast<int>(ysHi), __double2loint(y));
  const double adet = __hiloint2double(static_cast<int>(detHi & 0x7fffffffu), __double2loint(det));
  const double hi = adet * (1.0 + 0x1p-40);
  // high words: 1e-9 = 0x3E112E0B_E826D695, 1e-300 = 0x01A56E1F_C2F8F359
  const bool detTooSmall = (detHi & 0x7fffffffu) < 0x3E112E0Bu;  // certainly |det| < Epsilon
  const bool xNegative = xsHi > 0x81A56E1Fu;                      // certainly Xs < -1e-300
  const bool yNegative = ysHi > 0x81A56E1Fu;
  return !(detTooSmall | xNegative | yNegative) & (xs + ys <= hi);
}

template <bool kFpWay = false>
__device__ __forceinline__ void sweepTilePrefiltered(const double *__restrict__ tile, int tileTris,
                                                     int count, int firstIndex, V3 o, V3 d,
                                                     Nearest &best) {
#pragma unroll 1
  for (int chunk = 0; chunk < count; chunk += 64) {
    const int chunkEnd = min(count, chunk + 64);
    unsigned long long survivors = 0;
#pragma unroll 1
    for (int i = chunk; i < chunkEnd; i += 2) {
      const double2 v0x = *reinterpret_cast<const double2 *>(tile + 0 * tileTris + i);
      const double2 v0y = *reinterpret_cast<const double2 *>(tile + 1 * tileTris + i);
      const double2 v0z = *reinterpret_cast<const double2 *>(tile + 2 * tileTris + i);
      const double2 e1x = *reinterpret_cast<const double2 *>(tile + 3 * tileTris + i);
      const double2 e1y = *reinterpret_cast<const double2 *>(tile + 4 * tileTris + i);
      const double2 e1z = *reinterpret_cast<const double2 *>(tile + 5 * tileTris + i);
      const double2 e2x = *reinterpret_cast<const double2 *>(tile + 6 * tileTris + i);
      const double2 e2y = *reinterpret_cast<const double2 *>(tile + 7 * tileTris + i);
      const double2 e2z = *reinterpret_cast<const double2 *>(tile + 8 * tileTris + i);
      const bool a = prefilterTriangle(mk(v0x.x, v0y.x, v0z.x), mk(e1x.x, e1y.x, e1z.x),
                                       mk(e2x.x, e2y.x, e2z.x), o, d);
      const bool b = prefilterTriangle(mk(v0x.y, v0y.y, v0z.y), mk(e1x.y, e1y.y, e1z.y),
                                       mk(e2x.y, e2y.y, e2z.y), o, d);
      survivors |= (static_cast<unsigned long long>(a) | (static_cast<unsigned long long>(b) << 1))
                   << (i - chunk);
    }
    while (survivors) { // ascending index: the serial loop's tie-break order
      const int i = chunk + __ffsll(static_cast<long long>(survivors)) - 1;
      survivors &= survivors - 1;
      testTriangle<kFpWay>(mk(tile[0 * tileTris + i], tile[1 * tileTris + i], tile[2 * tileTris + i]),
                           mk(tile[3 * tileTris + i], tile[4 * tileTris + i], tile[5 * tileTris + i]),
                           mk(tile[6 * tileTris + i], tile[7 * tileTris + i], tile[8 * tileTris + i]), o, d,
                           firstIndex + i, best);
    }
  }
}

// ---- three-stage sweep: conservative FP32 stage 0, exact FP64 test for the survivors ----------
// Stage 0 evaluates Moller-Trumbore's det, X = tVec.pVec, Y = dir.qVec in FP32 from FP32 copies
// of the triangle and of the ray, and rejects a triangle only when the reference's FP64 test
// CERTAINLY rejects it, using per-triangle absolute error bounds built by buildFilterKernel:
//   Ed >= |det32 - det|,  Ex >= |X32 - X|,  Ey >= |Y32 - Y|    (2^-18 x the magnitudes involved:
//   64 FP32 roundoffs, about 4x what a forward error analysis of the 24 operations needs).
// With s = sign(det32), certain rejection needs |det32| > Ed (the sign is then the reference's)
// and one of   s*X32 < -2Ex   (=> u < 0),   s*Y32 < -2Ey   (=> v < 0),
//              s*(X32+Y32) > |det32|*(1+2^-20) + Ed + Ex + Ey   (=> u + v > 1),
//              s*T32 < -2Et   (=> t = T/det < 0, so `t > Epsilon` fails; T = e2.qVec).
// NaNs keep the triangle.  Survivors are re-tested with the exact reference arithmetic
// (testTriangle on the FP64 data in global memory/L1) in index order, so results are bit-identical
// to the one-stage sweep; the FP64 pipe only sees a handful of triangles per ray.
struct Stage0Ray {
  float ox, oy, oz, dx, dy, dz;
};
__device__ __forceinline__ bool stage0Keep(float v0x, float v0y, float v0z, float e1x, float e1y,
                                           float e1z, float e2x, float e2y, float e2z, float ed,
                                           float kx, float ky, float k3, float kt, const Stage0Ray &r) {
  const float px = fmaf(r.dy, e2z, -(r.dz * e2y));
  const float py = fmaf(r.dz, e2x, -(r.dx * e2z));
  const float pz = fmaf(r.dx, e2y, -(r.dy * e2x));
  const float det = fmaf(e1z, pz, fmaf(e1y, py, e1x * px));
  const float tx = r.ox - v0x, ty = r.oy - v0y, tz = r.oz - v0z;
  const float x = fmaf(tz, pz, fmaf(ty, py, tx * px));
  const float qx = fmaf(ty, e1z, -(tz * e1y));
  const float qy = fmaf(tz, e1x, -(tx * e1z));
  const float qz = fmaf(tx, e1y, -(ty * e1x));
  const float y = fmaf(r.dz, qz, fmaf(r.dy, qy, r.dx * qx));
  const float t = fmaf(e2z, qz, fmaf(e2y, qy, e2x * qx));
  const uint32_t sign = __float_as_uint(det) & 0x80000000u;
  const float xs = __uint_as_float(__float_as_uint(x) ^ sign);
  const float ys = __uint_as_float(__float_as_uint(y) ^ sign);
  const float ts = __uint_as_float(__float_as_uint(t) ^ sign);
  const float adet = fabsf(det);
  const bool certain = (xs < -kx) | (ys < -ky) | (ts < -kt) | (xs + ys > fmaf(adet, 1.0f + 0x1p-20f, k3));
  return (adet <= ed) | !certain;
}

// The same test for TWO triangles at once in the packed FP32x2 datapath (FFMA2/FMUL2/FADD2,
// new on sm_100): .x holds triangle i, .y triangle i+1.  Returns bit 0 / bit 1.
struct Stage0Ray2 {
  float2 ox, oy, oz, dx, dy, dz; // each component duplicated into both halves
  uint32_t one;                  // 0x3f800000 held in a register (stage0Reject2)
};
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
template <bool kRejectNegativeT>
__device__ __forceinline__ unsigned stage0Keep2(float2 v0x, float2 v0y, float2 v0z, float2 e1x,
                                                float2 e1y, float2 e1z, float2 e2x, float2 e2y,
                                                float2 e2z, float2 ed, float2 kx, float2 ky,
                                                float2 k3, float2 kt, const Stage0Ray2 &r) {
  const float2 px = __ffma2_rn(r.dy, e2z, neg2(__fmul2_rn(r.dz, e2y)));
  const float2 py = __ffma2_rn(r.dz, e2x, neg2(__fmul2_rn(r.dx, e2z)));
  const float2 pz = __ffma2_rn(r.dx, e2y, neg2(__fmul2_rn(r.dy, e2x)));
  const float2 det = __ffma2_rn(e1z, pz, __ffma2_rn(e1y, py, __fmul2_rn(e1x, px)));
  const float2 tx = __fadd2_rn(r.ox, neg2(v0x)), ty = __fadd2_rn(r.oy, neg2(v0y)),
               tz = __fadd2_rn(r.oz, neg2(v0z));
  const float2 x = __ffma2_rn(tz, pz, __ffma2_rn(ty, py, __fmul2_rn(tx, px)));
  const float2 qx = __ffma2_rn(ty, e1z, neg2(__fmul2_rn(tz, e1y)));
  const float2 qy = __ffma2_rn(tz, e1x, neg2(__fmul2_rn(tx, e1z)));
  const float2 qz = __ffma2_rn(tx, e1y, neg2(__fmul2_rn(ty, e1x)));
  const float2 y = __ffma2_rn(r.dz, qz, __ffma2_rn(r.dy, qy, __fmul2_rn(r.dx, qx)));
  const uint32_t signA = __float_as_uint(det.x) & 0x80000000u, signB = __float_as_uint(det.y) & 0x80000000u;
  const float2 xs = make_float2(__uint_as_float(__float_as_uint(x.x) ^ signA), __uint_as_float(__float_as_uint(x.y) ^ signB));
  const float2 ys = make_float2(__uint_as_float(__float_as_uint(y.x) ^ signA), __uint_as_float(__float_as_uint(y.y) ^ signB));
  const float2 adet = make_float2(fabsf(det.x), fabsf(det.y));
  const float2 sum = __fadd2_rn(xs, ys);
  const float2 bound = __ffma2_rn(adet, make_float2(1.0f + 0x1p-20f, 1.0f + 0x1p-20f), k3);
  bool certainA = (xs.x < -kx.x) | (ys.x < -ky.x) | (sum.x > bound.x);
  bool certainB = (xs.y < -kx.y) | (ys.y < -ky.y) | (sum.y > bound.y);
  if (kRejectNegativeT) { // worth its 3 packed operations only when survivors dominate (small scenes)
    const float2 t = __ffma2_rn(e2z, qz, __ffma2_rn(e2y, qy, __fmul2_rn(e2x, qx)));
    certainA |= __uint_as_float(__float_as_uint(t.x) ^ signA) < -kt.x;
    certainB |= __uint_as_float(__float_as_uint(t.y) ^ signB) < -kt.y;
  }
  const bool keepA = (adet.x <= ed.x) | !certainA;
  const bool keepB = (adet.y <= ed.y) | !certainB;
  return (keepA ? 1u : 0u) | (keepB ? 2u : 0u);
}

// `filter` is the FP32 tile in shared memory, blocked by groups of four triangles:
// [group][14][4] floats (v0 xyz, e1 xyz, e2 xyz, Ed, 2Ex, 2Ey, K3, 2Et), so one base register and
// immediate offsets address all fourteen 16-byte loads; `exact` is the tile's first record in the
// AoS FP64 array (10 doubles per triangle, global memory / L1).  count is a multiple of 4.
constexpr int kFilterFloats = 14;
template <bool kPacked, bool kRejectNegativeT, bool kFpWay = false>
__device__ __forceinline__ void sweepTileStage0(const float *__restrict__ filter,
                                                const double *__restrict__ exact, int tileTris,
                                                int count, int firstIndex, V3 o, V3 d,
                                                Nearest &best) {
  const Stage0Ray r{static_cast<float>(o.x), static_cast<float>(o.y), static_cast<float>(o.z),
                    static_cast<float>(d.x), static_cast<float>(d.y), static_cast<float>(d.z)};
  const Stage0Ray2 r2{make_float2(r.ox, r.ox), make_float2(r.oy, r.oy), make_float2(r.oz, r.oz),
                      make_float2(r.dx, r.dx), make_float2(r.dy, r.dy), make_float2(r.dz, r.dz), 0u};
#pragma unroll 1
  for (int chunk = 0; chunk < count; chunk += 64) {
    const int chunkEnd = min(count, chunk + 64);
    unsigned long long survivors = 0;
#pragma unroll 1
    for (int i = chunk; i < chunkEnd; i += 4) {
      float4 a[kFilterFloats];
      const float4 *group = reinterpret_cast<const float4 *>(filter) + (i >> 2) * kFilterFloats;
#pragma unroll
      for (int k = 0; k < kFilterFloats; ++k)
        a[k] = group[k];
      unsigned keep;
      if (kPacked) {
#define PT_LO(k) make_float2(a[k].x, a[k].y)
#define PT_HI(k) make_float2(a[k].z, a[k].w)
        keep = stage0Keep2<kRejectNegativeT>(PT_LO(0), PT_LO(1), PT_LO(2), PT_LO(3), PT_LO(4), PT_LO(5), PT_LO(6), PT_LO(7),
                           PT_LO(8), PT_LO(9), PT_LO(10), PT_LO(11), PT_LO(12), PT_LO(13), r2) |
               (stage0Keep2<kRejectNegativeT>(PT_HI(0), PT_HI(1), PT_HI(2), PT_HI(3), PT_HI(4), PT_HI(5), PT_HI(6), PT_HI(7),
                            PT_HI(8), PT_HI(9), PT_HI(10), PT_HI(11), PT_HI(12), PT_HI(13), r2) << 2);
#undef PT_LO
#undef PT_HI
      } else
      keep =
          (stage0Keep(a[0].x, a[1].x, a[2].x, a[3].x, a[4].x, a[5].x, a[6].x, a[7].x, a[8].x, a[9].x, a[10].x, a[11].x, a[12].x, a[13].x, r) ? 1u : 0u) |
          (stage0Keep(a[0].y, a[1].y, a[2].y, a[3].y, a[4].y, a[5].y, a[6].y, a[7].y, a[8].y, a[9].y, a[10].y, a[11].y, a[12].y, a[13].y, r) ? 2u : 0u) |
          (stage0Keep(a[0].z, a[1].z, a[2].z, a[3].z, a[4].z, a[5].z, a[6].z, a[7].z, a[8].z, a[9].z, a[10].z, a[11].z, a[12].z, a[13].z, r) ? 4u : 0u) |
          (stage0Keep(a[0].w, a[1].w, a[2].w, a[3].w, a[4].w, a[5].w, a[6].w, a[7].w, a[8].w, a[9].w, a[10].w, a[11].w, a[12].w, a[13].w, r) ? 8u : 0u);
      survivors |= static_cast<unsigned long long>(keep) << (i - chunk);
    }
    // Survivors in ascending index: the serial loop's tie-break order.  (Prefetching the next
    // survivor's record while testing the current one was measured: 7 % slower, more spills.)
    while (survivors) {
      const int i = chunk + __ffsll(static_cast<long long>(survivors)) - 1;
      survivors &= survivors - 1;
      const double2 *record = reinterpret_cast<const double2 *>(exact + 10 * static_cast<size_t>(i));
      const double2 a0 = __ldg(record), a1 = __ldg(record + 1), a2 = __ldg(record + 2), a3 = __ldg(record + 3),
                    a4 = __ldg(record + 4);
      testTriangle<kFpWay>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d, firstIndex + i, best);
    }
  }
}

// ---- the same stage 0 with its decision kept in SIGN BITS (sweep variants 5 and 6) -------------
// stage0Keep2() spends as many instructions on its comparisons (FSETP/SEL/LOP3, one triangle at
// a time) as on its arithmetic.  Every one of its tests is the sign of a difference, and an IEEE
// subtraction never gets a sign wrong, so the tests are evaluated two triangles at a time in the
// packed datapath and combined as integers:
//   with s = copysign(1, det32) (s*x is exact):
//     a = s*X32 + 2Ex  (< 0  <=>  s*X32 < -2Ex)        b = s*Y32 + 2Ey        c = s*T32 + 2Et
//     e = K - s*(X32+Y32),  K = |det32|(1+2^-20) + K3   (< 0  <=>  s*(X32+Y32) > K)
//     f = Ed - |det32|                                  (< 0  <=>  |det32| > Ed)
//   certainly rejected  <=>  sign bit of  (a | b | c | e) & f.
// These are the decisions of stage0Keep2<kRejectNegativeT>() bit for bit (none of the sums can be
// -0: the bounds are >= +0 or, for padding triangles, -1 with det32 == 0), except that a NaN no
// longer forces "keep": a NaN needs a non-finite ray or an FP32 overflow (excluded by the
// coordinate range check at upload), and the exact test never accepts a triangle whose
// arithmetic involves a NaN (Scene.cpp:89,94 compare false).
template <bool kRejectNegativeT>
__device__ __forceinline__ void stage0Reject2(float2 v0x, float2 v0y, float2 v0z, float2 e1x, float2 e1y,
                                              float2 e1z, float2 e2x, float2 e2y, float2 e2z, float2 ed,
                                              float2 kx, float2 ky, float2 k3, float2 kt,
                                              const Stage0Ray2 &r, uint32_t &rejectA, uint32_t &rejectB) {
  const float2 px = __ffma2_rn(r.dy, e2z, neg2(__fmul2_rn(r.dz, e2y)));
  const float2 py = __ffma2_rn(r.dz, e2x, neg2(__fmul2_rn(r.dx, e2z)));
  const float2 pz = __ffma2_rn(r.dx, e2y, neg2(__fmul2_rn(r.dy, e2x)));
  const float2 det = __ffma2_rn(e1z, pz, __ffma2_rn(e1y, py, __fmul2_rn(e1x, px)));
  const float2 tx = __fadd2_rn(r.ox, neg2(v0x)), ty = __fadd2_rn(r.oy, neg2(v0y)),
               tz = __fadd2_rn(r.oz, neg2(v0z));
  const float2 x = __ffma2_rn(tz, pz, __ffma2_rn(ty, py, __fmul2_rn(tx, px)));
  const float2 qx = __ffma2_rn(ty, e1z, neg2(__fmul2_rn(tz, e1y)));
  const float2 qy = __ffma2_rn(tz, e1x, neg2(__fmul2_rn(tx, e1z)));
  const float2 qz = __ffma2_rn(tx, e1y, neg2(__fmul2_rn(ty, e1x)));
  const float2 y = __ffma2_rn(r.dz, qz, __ffma2_rn(r.dy, qy, __fmul2_rn(r.dx, qx)));
  // copysign(1, det): (det & 0x80000000) | 0x3f800000 as ONE three-input logic instruction; `one`
  // arrives in a register (a LOP3 takes a single immediate).
  uint32_t sA, sB;
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(sA) : "r"(__float_as_uint(det.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(sB) : "r"(__float_as_uint(det.y)), "r"(r.one));
  const float2 s = make_float2(__uint_as_float(sA), __uint_as_float(sB));
  const float2 adet = make_float2(fabsf(det.x), fabsf(det.y));
  const float2 a = __ffma2_rn(x, s, kx);
  const float2 b = __ffma2_rn(y, s, ky);
  const float2 bound = __ffma2_rn(adet, make_float2(1.0f + 0x1p-20f, 1.0f + 0x1p-20f), k3);
  const float2 e = __ffma2_rn(neg2(__fadd2_rn(x, y)), s, bound);
  const float2 f = __fadd2_rn(ed, neg2(adet));
  uint32_t anyA = __float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(e.x);
  uint32_t anyB = __float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(e.y);
  if (kRejectNegativeT) {
    const float2 t = __ffma2_rn(e2z, qz, __ffma2_rn(e2y, qy, __fmul2_rn(e2x, qx)));
    const float2 c = __ffma2_rn(t, s, kt);
    anyA |= __float_as_uint(c.x);
    anyB |= __float_as_uint(c.y);
  }
  rejectA = anyA & __float_as_uint(f.x);
  rejectB = anyB & __float_as_uint(f.y);
}

// Sweeps a staged FP32 tile with stage0Reject2().  The decisions of a chunk of 64 triangles are
// shifted into one 64-bit register, four at a time (first triangle ends up in the highest bit), so
// a small scene still has ONE survivor loop per cast, which the lanes of a warp walk together.
// Survivors are taken from the top bit down: ascending index, the serial loop's tie-break order.
template <bool kRejectNegativeT, bool kFpWay = false>
__device__ __forceinline__ void sweepTileStage0Signs(const float *__restrict__ filter,
                                                     const double *__restrict__ exact, int count,
                                                     int firstIndex, V3 o, V3 d, Nearest &best) {
  const Stage0Ray r{static_cast<float>(o.x), static_cast<float>(o.y), static_cast<float>(o.z),
                    static_cast<float>(d.x), static_cast<float>(d.y), static_cast<float>(d.z)};
  uint32_t one;
  asm("mov.u32 %0, 0x3f800000;" : "=r"(one)); // opaque to constant folding: see stage0Reject2
  const Stage0Ray2 r2{make_float2(r.ox, r.ox), make_float2(r.oy, r.oy), make_float2(r.oz, r.oz),
                      make_float2(r.dx, r.dx), make_float2(r.dy, r.dy), make_float2(r.dz, r.dz), one};
#pragma unroll 1
  for (int chunk = 0; chunk < count; chunk += 64) {
    const int chunkEnd = min(count, chunk + 64);
    uint32_t rejectedHi = 0xffffffffu, rejectedLo = 0xffffffffu;
    const float4 *group = reinterpret_cast<const float4 *>(filter) + (chunk >> 2) * kFilterFloats;
    const float4 *const groupEnd = reinterpret_cast<const float4 *>(filter) + (chunkEnd >> 2) * kFilterFloats;
#pragma unroll 1
    for (; group != groupEnd; group += kFilterFloats) {
      float4 a[kFilterFloats];
#pragma unroll
      for (int k = 0; k < kFilterFloats; ++k)
        a[k] = group[k];
      uint32_t r0, r1, r2bits, r3;
#define PT_LO(k) make_float2(a[k].x, a[k].y)
#define PT_HI(k) make_float2(a[k].z, a[k].w)
      stage0Reject2<kRejectNegativeT>(PT_LO(0), PT_LO(1), PT_LO(2), PT_LO(3), PT_LO(4), PT_LO(5), PT_LO(6),
                                      PT_LO(7), PT_LO(8), PT_LO(9), PT_LO(10), PT_LO(11), PT_LO(12), PT_LO(13),
                                      r2, r0, r1);
      stage0Reject2<kRejectNegativeT>(PT_HI(0), PT_HI(1), PT_HI(2), PT_HI(3), PT_HI(4), PT_HI(5), PT_HI(6),
                                      PT_HI(7), PT_HI(8), PT_HI(9), PT_HI(10), PT_HI(11), PT_HI(12), PT_HI(13),
                                      r2, r2bits, r3);
#undef PT_LO
#undef PT_HI
      rejectedHi = __funnelshift_l(rejectedLo, rejectedHi, 4);
      rejectedLo = __funnelshift_l(r0, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r1, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r2bits, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r3, rejectedLo, 1);
    }
    // left-align: triangle chunk + k at bit 63 - k; the slots past chunkEnd read "rejected"
    unsigned long long keep = ~((static_cast<unsigned long long>(rejectedHi) << 32) | rejectedLo)
                              << (64 - (chunkEnd - chunk));
    while (keep) {
      const int k = __clzll(static_cast<long long>(keep));
      keep &= ~(0x8000000000000000ull >> k);
      const int i = chunk + k;
      const double2 *record = reinterpret_cast<const double2 *>(exact + 10 * static_cast<size_t>(i));
      const double2 a0 = __ldg(record), a1 = __ldg(record + 1), a2 = __ldg(record + 2), a3 = __ldg(record + 3),
                    a4 = __ldg(record + 4);
      testTriangle<kFpWay>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d, firstIndex + i, best);
    }
  }
}

// ---- stage 0 in moment (Pluecker) form: sweep variant 7 -----------------------------------------
// Moller-Trumbore's three numerators are scalar triple products, so with the ray's moment
// m = o x d (per ray, once, in FP64) and per-triangle constants they are plain dot products:
//     det = e1 . (d x e2)            = d . (e2 x e1)
//     X   = (o - v0) . (d x e2)      = e2 . m + d . (v0 x e2)
//     Y   = d . ((o - v0) x e1)      = (-e1) . m + d . (-(v0 x e1))
// i.e. 15 packed operations per pair of triangles instead of 24 (no per-triangle cross products),
// on five FP32 vectors per triangle (built in FP64 by buildFilterKernel, rounded once).  Forward
// error analysis with u = 2^-24, |d| = 1, inputs rounded to FP32 (1 roundoff each):
//     |det32 - det| <= (gamma_3 + 2u) |e1||e2|        ~  5.1 u |e1||e2|
//     |X32 - X|     <= (gamma_6 + 2u) |e2| (|o|+|v0|) ~  8.2 u r|e2|      (|m| <= |o|, |v0 x e2| <= |v0||e2|)
//     |Y32 - Y|     <= likewise                       ~  8.2 u r|e1|
// all well inside the bounds Ed, Ex, Ey = 64 u x the same magnitudes that buildFilterKernel stores
// for the classic form (which needs 13..18 u), so the decision logic and its soundness argument
// are stage0Reject2()'s unchanged: reject only if |det32| > Ed and one of s*X32 < -2Ex,
// s*Y32 < -2Ey, s*(X32+Y32) > |det32|(1+2^-20) + Ed + Ex + Ey, all taken as sign bits.
constexpr int kMomentFloats = 19; // nn xyz, e2 xyz, a2 xyz, -e1 xyz, -a1 xyz, Ed, 2Ex, 2Ey, K3
struct MomentRay {
  float dx, dy, dz, mx, my, mz;
  uint32_t one; // 0x3f800000 held in a register (a LOP3 takes a single immediate)
};
__device__ __forceinline__ float2 splat(float a) { return make_float2(a, a); }
__device__ __forceinline__ void stage0RejectMoment(float2 nx, float2 ny, float2 nz, float2 e2x, float2 e2y,
                                                   float2 e2z, float2 a2x, float2 a2y, float2 a2z, float2 f1x,
                                                   float2 f1y, float2 f1z, float2 b1x, float2 b1y, float2 b1z,
                                                   float2 ed, float2 kx, float2 ky, float2 k3, const MomentRay &r,
                                                   uint32_t &rejectA, uint32_t &rejectB) {
  const float2 det = __ffma2_rn(splat(r.dz), nz, __ffma2_rn(splat(r.dy), ny, __fmul2_rn(splat(r.dx), nx)));
  float2 x = __ffma2_rn(splat(r.my), e2y, __fmul2_rn(splat(r.mx), e2x));
  x = __ffma2_rn(splat(r.mz), e2z, x);
  x = __ffma2_rn(splat(r.dx), a2x, x);
  x = __ffma2_rn(splat(r.dy), a2y, x);
  x = __ffma2_rn(splat(r.dz), a2z, x);
  float2 y = __ffma2_rn(splat(r.my), f1y, __fmul2_rn(splat(r.mx), f1x));
  y = __ffma2_rn(splat(r.mz), f1z, y);
  y = __ffma2_rn(splat(r.dx), b1x, y);
  y = __ffma2_rn(splat(r.dy), b1y, y);
  y = __ffma2_rn(splat(r.dz), b1z, y);
  uint32_t sA, sB; // copysign(1, det)
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(sA) : "r"(__float_as_uint(det.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(sB) : "r"(__float_as_uint(det.y)), "r"(r.one));
  const float2 s = make_float2(__uint_as_float(sA), __uint_as_float(sB));
  const float2 adet = make_float2(fabsf(det.x), fabsf(det.y));
  const float2 a = __ffma2_rn(x, s, kx);
  const float2 b = __ffma2_rn(y, s, ky);
  const float2 bound = __ffma2_rn(adet, make_float2(1.0f + 0x1p-20f, 1.0f + 0x1p-20f), k3);
  const float2 e = __ffma2_rn(neg2(__fadd2_rn(x, y)), s, bound);
  const float2 f = __fadd2_rn(ed, neg2(adet));
  rejectA = (__float_as_uint(a.x) | __float_as_uint(b.x) | __float_as_uint(e.x)) & __float_as_uint(f.x);
  rejectB = (__float_as_uint(a.y) | __float_as_uint(b.y) | __float_as_uint(e.y)) & __float_as_uint(f.y);
}
// The same decision for ONE triangle in scalar FP32 (same operations, same rounding): the audit
// kernel's view of the filter.
__device__ __forceinline__ bool stage0KeepMoment(const float *f, int stride, const MomentRay &r) {
  const float det = fmaf(r.dz, f[2 * stride], fmaf(r.dy, f[1 * stride], r.dx * f[0]));
  float x = fmaf(r.my, f[4 * stride], r.mx * f[3 * stride]);
  x = fmaf(r.mz, f[5 * stride], x);
  x = fmaf(r.dx, f[6 * stride], x);
  x = fmaf(r.dy, f[7 * stride], x);
  x = fmaf(r.dz, f[8 * stride], x);
  float y = fmaf(r.my, f[10 * stride], r.mx * f[9 * stride]);
  y = fmaf(r.mz, f[11 * stride], y);
  y = fmaf(r.dx, f[12 * stride], y);
  y = fmaf(r.dy, f[13 * stride], y);
  y = fmaf(r.dz, f[14 * stride], y);
  const float s = copysignf(1.0f, det), adet = fabsf(det);
  const float a = fmaf(x, s, f[16 * stride]);
  const float b = fmaf(y, s, f[17 * stride]);
  const float bound = fmaf(adet, 1.0f + 0x1p-20f, f[18 * stride]);
  const float e = fmaf(-(x + y), s, bound);
  const float g = f[15 * stride] - adet;
  const uint32_t reject = (__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(e)) & __float_as_uint(g);
  return (reject & 0x80000000u) == 0;
}
__device__ __forceinline__ MomentRay makeMomentRay(V3 o, V3 d) {
  const V3 m = cross(o, d);
  uint32_t one;
  asm("mov.u32 %0, 0x3f800000;" : "=r"(one)); // opaque to constant folding
  return MomentRay{static_cast<float>(d.x), static_cast<float>(d.y), static_cast<float>(d.z),
                   static_cast<float>(m.x), static_cast<float>(m.y), static_cast<float>(m.z), one};
}

// stage0RejectMoment() for FOUR triangles with the two pairs interleaved component by component, so
// that each float4 of the table is consumed right after it arrives (one 128-bit uniform load instead
// of two 64-bit ones when the table sits in the constant bank).  Same operations per triangle.
__device__ __forceinline__ void stage0RejectMoment4(const float4 (&a)[kMomentFloats], const MomentRay &r,
                                                    uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
#define PT_LO(k) make_float2(a[k].x, a[k].y)
#define PT_HI(k) make_float2(a[k].z, a[k].w)
  float2 detL = __fmul2_rn(splat(r.dx), PT_LO(0)), detH = __fmul2_rn(splat(r.dx), PT_HI(0));
  detL = __ffma2_rn(splat(r.dy), PT_LO(1), detL), detH = __ffma2_rn(splat(r.dy), PT_HI(1), detH);
  detL = __ffma2_rn(splat(r.dz), PT_LO(2), detL), detH = __ffma2_rn(splat(r.dz), PT_HI(2), detH);
  float2 xL = __fmul2_rn(splat(r.mx), PT_LO(3)), xH = __fmul2_rn(splat(r.mx), PT_HI(3));
  xL = __ffma2_rn(splat(r.my), PT_LO(4), xL), xH = __ffma2_rn(splat(r.my), PT_HI(4), xH);
  xL = __ffma2_rn(splat(r.mz), PT_LO(5), xL), xH = __ffma2_rn(splat(r.mz), PT_HI(5), xH);
  xL = __ffma2_rn(splat(r.dx), PT_LO(6), xL), xH = __ffma2_rn(splat(r.dx), PT_HI(6), xH);
  xL = __ffma2_rn(splat(r.dy), PT_LO(7), xL), xH = __ffma2_rn(splat(r.dy), PT_HI(7), xH);
  xL = __ffma2_rn(splat(r.dz), PT_LO(8), xL), xH = __ffma2_rn(splat(r.dz), PT_HI(8), xH);
  float2 yL = __fmul2_rn(splat(r.mx), PT_LO(9)), yH = __fmul2_rn(splat(r.mx), PT_HI(9));
  yL = __ffma2_rn(splat(r.my), PT_LO(10), yL), yH = __ffma2_rn(splat(r.my), PT_HI(10), yH);
  yL = __ffma2_rn(splat(r.mz), PT_LO(11), yL), yH = __ffma2_rn(splat(r.mz), PT_HI(11), yH);
  yL = __ffma2_rn(splat(r.dx), PT_LO(12), yL), yH = __ffma2_rn(splat(r.dx), PT_HI(12), yH);
  yL = __ffma2_rn(splat(r.dy), PT_LO(13), yL), yH = __ffma2_rn(splat(r.dy), PT_HI(13), yH);
  yL = __ffma2_rn(splat(r.dz), PT_LO(14), yL), yH = __ffma2_rn(splat(r.dz), PT_HI(14), yH);
  uint32_t s0, s1, s2, s3; // copysign(1, det)
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s0) : "r"(__float_as_uint(detL.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s1) : "r"(__float_as_uint(detL.y)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s2) : "r"(__float_as_uint(detH.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s3) : "r"(__float_as_uint(detH.y)), "r"(r.one));
  const float2 sL = make_float2(__uint_as_float(s0), __uint_as_float(s1));
  const float2 sH = make_float2(__uint_as_float(s2), __uint_as_float(s3));
  const float2 adetL = make_float2(fabsf(detL.x), fabsf(detL.y)), adetH = make_float2(fabsf(detH.x), fabsf(detH.y));
  const float2 fL = __fadd2_rn(PT_LO(15), neg2(adetL)), fH = __fadd2_rn(PT_HI(15), neg2(adetH));
  const float2 aL = __ffma2_rn(xL, sL, PT_LO(16)), aH = __ffma2_rn(xH, sH, PT_HI(16));
  const float2 bL = __ffma2_rn(yL, sL, PT_LO(17)), bH = __ffma2_rn(yH, sH, PT_HI(17));
  const float2 boundL = __ffma2_rn(adetL, splat(1.0f + 0x1p-20f), PT_LO(18));
  const float2 boundH = __ffma2_rn(adetH, splat(1.0f + 0x1p-20f), PT_HI(18));
  const float2 eL = __ffma2_rn(neg2(__fadd2_rn(xL, yL)), sL, boundL), eH = __ffma2_rn(neg2(__fadd2_rn(xH, yH)), sH, boundH);
#undef PT_LO
#undef PT_HI
  r0 = (__float_as_uint(aL.x) | __float_as_uint(bL.x) | __float_as_uint(eL.x)) & __float_as_uint(fL.x);
  r1 = (__float_as_uint(aL.y) | __float_as_uint(bL.y) | __float_as_uint(eL.y)) & __float_as_uint(fL.y);
  r2 = (__float_as_uint(aH.x) | __float_as_uint(bH.x) | __float_as_uint(eH.x)) & __float_as_uint(fH.x);
  r3 = (__float_as_uint(aH.y) | __float_as_uint(bH.y) | __float_as_uint(eH.y)) & __float_as_uint(fH.y);
}

// A quad face `f a b c d` reaches the scene as the fan (a,b,c), (a,c,d) (ObjLoaderImpl.h:75-85): triangle
// B shares v0 with triangle A and its first edge IS A's second edge, bit for bit.  Then B's
// Y numerator is A's X numerator negated —  Y_B = -e1_B.m - d.(v0 x e1_B) = -(e2_A.m + d.(v0 x e2_A))
// = -X_A  — through the very same operations (negation commutes with rounding), so a group of two
// such quads, packed as [A0, A1 | B0, B1], needs Y for its first pair only: 24 instead of 30 packed
// operations and 32 instead of 38 constant loads per four triangles, the same decisions bit for bit.
// The upload marks the groups that qualify (exact comparisons of the stored doubles).
__device__ __forceinline__ void stage0RejectFan4(const float4 (&a)[kMomentFloats], const MomentRay &r,
                                                 uint32_t &rA0, uint32_t &rA1, uint32_t &rB0, uint32_t &rB1) {
#define PT_LO(k) make_float2(a[k].x, a[k].y)
#define PT_HI(k) make_float2(a[k].z, a[k].w)
  float2 detL = __fmul2_rn(splat(r.dx), PT_LO(0)), detH = __fmul2_rn(splat(r.dx), PT_HI(0));
  detL = __ffma2_rn(splat(r.dy), PT_LO(1), detL), detH = __ffma2_rn(splat(r.dy), PT_HI(1), detH);
  detL = __ffma2_rn(splat(r.dz), PT_LO(2), detL), detH = __ffma2_rn(splat(r.dz), PT_HI(2), detH);
  float2 xL = __fmul2_rn(splat(r.mx), PT_LO(3)), xH = __fmul2_rn(splat(r.mx), PT_HI(3));
  xL = __ffma2_rn(splat(r.my), PT_LO(4), xL), xH = __ffma2_rn(splat(r.my), PT_HI(4), xH);
  xL = __ffma2_rn(splat(r.mz), PT_LO(5), xL), xH = __ffma2_rn(splat(r.mz), PT_HI(5), xH);
  xL = __ffma2_rn(splat(r.dx), PT_LO(6), xL), xH = __ffma2_rn(splat(r.dx), PT_HI(6), xH);
  xL = __ffma2_rn(splat(r.dy), PT_LO(7), xL), xH = __ffma2_rn(splat(r.dy), PT_HI(7), xH);
  xL = __ffma2_rn(splat(r.dz), PT_LO(8), xL), xH = __ffma2_rn(splat(r.dz), PT_HI(8), xH);
  float2 yL = __fmul2_rn(splat(r.mx), PT_LO(9));
  yL = __ffma2_rn(splat(r.my), PT_LO(10), yL);
  yL = __ffma2_rn(splat(r.mz), PT_LO(11), yL);
  yL = __ffma2_rn(splat(r.dx), PT_LO(12), yL);
  yL = __ffma2_rn(splat(r.dy), PT_LO(13), yL);
  yL = __ffma2_rn(splat(r.dz), PT_LO(14), yL);
  const float2 yH = neg2(xL); // the fan identity
  uint32_t s0, s1, s2, s3; // copysign(1, det)
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s0) : "r"(__float_as_uint(detL.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s1) : "r"(__float_as_uint(detL.y)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s2) : "r"(__float_as_uint(detH.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s3) : "r"(__float_as_uint(detH.y)), "r"(r.one));
  const float2 sL = make_float2(__uint_as_float(s0), __uint_as_float(s1));
  const float2 sH = make_float2(__uint_as_float(s2), __uint_as_float(s3));
  const float2 adetL = make_float2(fabsf(detL.x), fabsf(detL.y)), adetH = make_float2(fabsf(detH.x), fabsf(detH.y));
  const float2 fL = __fadd2_rn(PT_LO(15), neg2(adetL)), fH = __fadd2_rn(PT_HI(15), neg2(adetH));
  const float2 aL = __ffma2_rn(xL, sL, PT_LO(16)), aH = __ffma2_rn(xH, sH, PT_HI(16));
  const float2 bL = __ffma2_rn(yL, sL, PT_LO(17)), bH = __ffma2_rn(yH, sH, PT_HI(17));
  const float2 boundL = __ffma2_rn(adetL, splat(1.0f + 0x1p-20f), PT_LO(18));
  const float2 boundH = __ffma2_rn(adetH, splat(1.0f + 0x1p-20f), PT_HI(18));
  const float2 eL = __ffma2_rn(neg2(__fadd2_rn(xL, yL)), sL, boundL), eH = __ffma2_rn(neg2(__fadd2_rn(xH, yH)), sH, boundH);
#undef PT_LO
#undef PT_HI
  rA0 = (__float_as_uint(aL.x) | __float_as_uint(bL.x) | __float_as_uint(eL.x)) & __float_as_uint(fL.x);
  rA1 = (__float_as_uint(aL.y) | __float_as_uint(bL.y) | __float_as_uint(eL.y)) & __float_as_uint(fL.y);
  rB0 = (__float_as_uint(aH.x) | __float_as_uint(bH.x) | __float_as_uint(eH.x)) & __float_as_uint(fH.x);
  rB1 = (__float_as_uint(aH.y) | __float_as_uint(bH.y) | __float_as_uint(eH.y)) & __float_as_uint(fH.y);
}

// The fan flags of the (up to) 16 groups of a 64-triangle chunk, bit 0 = its first group.
__device__ __forceinline__ uint32_t fanBitsOfChunk(const uint32_t *__restrict__ fanMask, uint32_t firstGroup) {
  const uint32_t word = firstGroup >> 5, shift = firstGroup & 31u;
  const uint32_t lo = __ldg(fanMask + word), hi = __ldg(fanMask + word + 1); // (one spare word is allocated)
  return __funnelshift_r(lo, hi, shift);
}

// Sweeps a staged tile of moment-form data; same survivor bookkeeping as sweepTileStage0Signs().
template <bool kFpWay = false>
__device__ __forceinline__ void sweepTileStage0Moment(const float *__restrict__ filter,
                                                      const double *__restrict__ exact, int count,
                                                      int firstIndex, V3 o, V3 d, Nearest &best,
                                                      const uint32_t *__restrict__ fanMask) {
  const MomentRay r = makeMomentRay(o, d);
#pragma unroll 1
  for (int chunk = 0; chunk < count; chunk += 64) {
    const int chunkEnd = min(count, chunk + 64);
    uint32_t fanBits = fanBitsOfChunk(fanMask, static_cast<uint32_t>(firstIndex + chunk) >> 2);
    uint32_t rejectedHi = 0xffffffffu, rejectedLo = 0xffffffffu;
    const float4 *group = reinterpret_cast<const float4 *>(filter) + (chunk >> 2) * kMomentFloats;
    const float4 *const groupEnd = reinterpret_cast<const float4 *>(filter) + (chunkEnd >> 2) * kMomentFloats;
#pragma unroll 1 // (unrolled by two: 180 vs 193 Msamples/s, spills at 80 and 96 registers; r2d)
    for (; group != groupEnd; group += kMomentFloats) {
      float4 a[kMomentFloats];
#pragma unroll
      for (int k = 0; k < kMomentFloats; ++k)
        a[k] = group[k];
      uint32_t r0, r1, r2bits, r3; // in triangle-index order
      if (fanBits & 1u) // warp-uniform; lanes [A0, A1, B0, B1] = triangles 4g + {0, 2, 1, 3}
        stage0RejectFan4(a, r, r0, r2bits, r1, r3);
      else
        stage0RejectMoment4(a, r, r0, r1, r2bits, r3);
      fanBits >>= 1;
      rejectedHi = __funnelshift_l(rejectedLo, rejectedHi, 4);
      rejectedLo = __funnelshift_l(r0, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r1, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r2bits, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r3, rejectedLo, 1);
    }
    // left-align: triangle chunk + k at bit 63 - k; the slots past chunkEnd read "rejected"
    unsigned long long keep = ~((static_cast<unsigned long long>(rejectedHi) << 32) | rejectedLo)
                              << (64 - (chunkEnd - chunk));
    while (keep) {
      const int k = __clzll(static_cast<long long>(keep));
      keep &= ~(0x8000000000000000ull >> k);
      const int i = chunk + k;
      const double2 *record = reinterpret_cast<const double2 *>(exact + 10 * static_cast<size_t>(i));
      const double2 a0 = __ldg(record), a1 = __ldg(record + 1), a2 = __ldg(record + 2), a3 = __ldg(record + 3),
                    a4 = __ldg(record + 4);
      testTriangle<kFpWay>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d, firstIndex + i, best);
    }
  }
}

// stage0RejectMoment4() for TWO rays, interleaved component by component: each float4 of the group is
// used for both rays right after it is loaded, so only the twelve running sums (two rays x two pairs
// x {det, X, Y}) stay live instead of the whole 19-vector group.
struct Stage0Sums {
  float2 detL, detH, xL, xH, yL, yH;
};
__device__ __forceinline__ void stage0Accumulate(Stage0Sums &s, const MomentRay &r, int k, float4 a) {
  const float2 lo = make_float2(a.x, a.y), hi = make_float2(a.z, a.w);
  // k: 0-2 det (d . nn), 3-8 X (m . e2 + d . a2), 9-14 Y (m . -e1 + d . -a1)
  const float w = (k == 0 || k == 6 || k == 12) ? r.dx : (k == 1 || k == 7 || k == 13) ? r.dy
                  : (k == 2 || k == 8 || k == 14) ? r.dz : (k == 3 || k == 9) ? r.mx
                  : (k == 4 || k == 10) ? r.my : r.mz;
  if (k == 0) {
    s.detL = __fmul2_rn(splat(w), lo), s.detH = __fmul2_rn(splat(w), hi);
  } else if (k < 3) {
    s.detL = __ffma2_rn(splat(w), lo, s.detL), s.detH = __ffma2_rn(splat(w), hi, s.detH);
  } else if (k == 3) {
    s.xL = __fmul2_rn(splat(w), lo), s.xH = __fmul2_rn(splat(w), hi);
  } else if (k < 9) {
    s.xL = __ffma2_rn(splat(w), lo, s.xL), s.xH = __ffma2_rn(splat(w), hi, s.xH);
  } else if (k == 9) {
    s.yL = __fmul2_rn(splat(w), lo), s.yH = __fmul2_rn(splat(w), hi);
  } else {
    s.yL = __ffma2_rn(splat(w), lo, s.yL), s.yH = __ffma2_rn(splat(w), hi, s.yH);
  }
}
__device__ __forceinline__ void stage0Decide(const Stage0Sums &s, const MomentRay &r, float4 ed, float4 kx, float4 ky,
                                             float4 k3, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
  uint32_t s0, s1, s2, s3; // copysign(1, det)
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s0) : "r"(__float_as_uint(s.detL.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s1) : "r"(__float_as_uint(s.detL.y)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s2) : "r"(__float_as_uint(s.detH.x)), "r"(r.one));
  asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(s3) : "r"(__float_as_uint(s.detH.y)), "r"(r.one));
  const float2 sL = make_float2(__uint_as_float(s0), __uint_as_float(s1));
  const float2 sH = make_float2(__uint_as_float(s2), __uint_as_float(s3));
  const float2 adetL = make_float2(fabsf(s.detL.x), fabsf(s.detL.y)), adetH = make_float2(fabsf(s.detH.x), fabsf(s.detH.y));
  const float2 fL = __fadd2_rn(make_float2(ed.x, ed.y), neg2(adetL)), fH = __fadd2_rn(make_float2(ed.z, ed.w), neg2(adetH));
  const float2 aL = __ffma2_rn(s.xL, sL, make_float2(kx.x, kx.y)), aH = __ffma2_rn(s.xH, sH, make_float2(kx.z, kx.w));
  const float2 bL = __ffma2_rn(s.yL, sL, make_float2(ky.x, ky.y)), bH = __ffma2_rn(s.yH, sH, make_float2(ky.z, ky.w));
  const float2 boundL = __ffma2_rn(adetL, splat(1.0f + 0x1p-20f), make_float2(k3.x, k3.y));
  const float2 boundH = __ffma2_rn(adetH, splat(1.0f + 0x1p-20f), make_float2(k3.z, k3.w));
  const float2 eL = __ffma2_rn(neg2(__fadd2_rn(s.xL, s.yL)), sL, boundL), eH = __ffma2_rn(neg2(__fadd2_rn(s.xH, s.yH)), sH, boundH);
  r0 = (__float_as_uint(aL.x) | __float_as_uint(bL.x) | __float_as_uint(eL.x)) & __float_as_uint(fL.x);
  r1 = (__float_as_uint(aL.y) | __float_as_uint(bL.y) | __float_as_uint(eL.y)) & __float_as_uint(fL.y);
  r2 = (__float_as_uint(aH.x) | __float_as_uint(bH.x) | __float_as_uint(eH.x)) & __float_as_uint(fH.x);
  r3 = (__float_as_uint(aH.y) | __float_as_uint(bH.y) | __float_as_uint(eH.y)) & __float_as_uint(fH.y);
}

// The same sweep for TWO rays of one lane: every group of triangles is loaded once and tested
// against both (the loads, not the arithmetic, bound the one-ray form: see subPathDualKernel).
__device__ __forceinline__ void survivorsOfChunk(unsigned long long keep, const double *__restrict__ exact, int chunk,
                                                 int firstIndex, V3 o, V3 d, Nearest &best) {
  while (keep) {
    const int k = __clzll(static_cast<long long>(keep));
    keep &= ~(0x8000000000000000ull >> k);
    const int i = chunk + k;
    const double2 *record = reinterpret_cast<const double2 *>(exact + 10 * static_cast<size_t>(i));
    const double2 a0 = __ldg(record), a1 = __ldg(record + 1), a2 = __ldg(record + 2), a3 = __ldg(record + 3),
                  a4 = __ldg(record + 4);
    testTriangle<false>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d, firstIndex + i, best);
  }
}
__device__ __forceinline__ void sweepTileStage0Moment2(const float *__restrict__ filter, const double *__restrict__ exact,
                                                       int count, int firstIndex, V3 o0, V3 d0, bool live0, Nearest &best0,
                                                       V3 o1, V3 d1, bool live1, Nearest &best1,
                                                       const uint32_t *__restrict__ fanMask) {
  const MomentRay ray0 = makeMomentRay(o0, d0), ray1 = makeMomentRay(o1, d1);
#pragma unroll 1
  for (int chunk = 0; chunk < count; chunk += 64) {
    const int chunkEnd = min(count, chunk + 64);
    uint32_t fanBits = fanBitsOfChunk(fanMask, static_cast<uint32_t>(firstIndex + chunk) >> 2);
    uint32_t hi0 = 0xffffffffu, lo0 = 0xffffffffu, hi1 = 0xffffffffu, lo1 = 0xffffffffu;
    const float4 *group = reinterpret_cast<const float4 *>(filter) + (chunk >> 2) * kMomentFloats;
    const float4 *const groupEnd = reinterpret_cast<const float4 *>(filter) + (chunkEnd >> 2) * kMomentFloats;
#pragma unroll 1
    for (; group != groupEnd; group += kMomentFloats) {
      const bool fan = (fanBits & 1u) != 0; // warp-uniform: lanes [A0, A1, B0, B1], Y of the B pair = -X of the A pair
      fanBits >>= 1;
      Stage0Sums sums0, sums1;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float4 a = group[k];
        stage0Accumulate(sums0, ray0, k, a);
        stage0Accumulate(sums1, ray1, k, a);
      }
      if (fan) {
#pragma unroll
        for (int k = 9; k < 15; ++k) {
          const float2 lo = *reinterpret_cast<const float2 *>(group + k);
          stage0Accumulate(sums0, ray0, k, make_float4(lo.x, lo.y, 0.f, 0.f));
          stage0Accumulate(sums1, ray1, k, make_float4(lo.x, lo.y, 0.f, 0.f));
        }
        sums0.yH = neg2(sums0.xL);
        sums1.yH = neg2(sums1.xL);
      } else {
#pragma unroll
        for (int k = 9; k < 15; ++k) {
          const float4 a = group[k];
          stage0Accumulate(sums0, ray0, k, a);
          stage0Accumulate(sums1, ray1, k, a);
        }
      }
      const float4 ed = group[15], kx = group[16], ky = group[17], k3 = group[18];
      uint32_t r0, r1, r2, r3; // lane order; a fan group's triangle-index order is r0, r2, r1, r3
      stage0Decide(sums0, ray0, ed, kx, ky, k3, r0, r1, r2, r3);
      hi0 = __funnelshift_l(lo0, hi0, 4);
      lo0 = __funnelshift_l(r0, lo0, 1);
      lo0 = __funnelshift_l(fan ? r2 : r1, lo0, 1);
      lo0 = __funnelshift_l(fan ? r1 : r2, lo0, 1);
      lo0 = __funnelshift_l(r3, lo0, 1);
      stage0Decide(sums1, ray1, ed, kx, ky, k3, r0, r1, r2, r3);
      hi1 = __funnelshift_l(lo1, hi1, 4);
      lo1 = __funnelshift_l(r0, lo1, 1);
      lo1 = __funnelshift_l(fan ? r2 : r1, lo1, 1);
      lo1 = __funnelshift_l(fan ? r1 : r2, lo1, 1);
      lo1 = __funnelshift_l(r3, lo1, 1);
    }
    // left-align: triangle chunk + k at bit 63 - k; the slots past chunkEnd read "rejected"
    const int shift = 64 - (chunkEnd - chunk);
    const unsigned long long keep0 = live0 ? ~((static_cast<unsigned long long>(hi0) << 32) | lo0) << shift : 0ull;
    const unsigned long long keep1 = live1 ? ~((static_cast<unsigned long long>(hi1) << 32) | lo1) << shift : 0ull;
    survivorsOfChunk(keep0, exact, chunk, firstIndex, o0, d0, best0);
    survivorsOfChunk(keep1, exact, chunk, firstIndex, o1, d1, best1);
  }
}

// ---- stage 0 out of the CONSTANT BANK: sweep variant 9, scenes of up to 64 triangles ----------
// ncu (profiles/r2b) shows the sub-path kernel bound by the L1/shared-memory data pipe (79 % of its
// peak): a broadcast LDS.128 costs two wavefronts whether it serves one lane or 32, and stage 0
// issues 19 of them per four triangles — 380 of the ~830 wavefronts of a warp iteration.  The
// moment-form table of a small scene (76 B/triangle, 3 KB for the Cornell box) fits into the kernel's
// parameter block, i.e. constant bank 0: with the group loop fully unrolled every table address is an
// immediate, the compiler loads each float4 with ONE uniform-datapath LDCU.128 into uniform
// registers and feeds FFMA2/FMUL2 from them directly — no shared-memory wavefronts, no vector
// registers for triangle data.  Same arithmetic and decisions as sweepTileStage0Moment(), bit for bit.
constexpr int kConstGroups = 16; // 64 triangles
struct MomentTable {
  float4 group[kConstGroups][kMomentFloats];
  uint32_t fanGroups; // bit g: group g holds two FAN PAIRS, lanes ordered [A0, A1, B0, B1] (see below)
};

// `exact`: the AoS FP64 records of the survivors' test (triExact layout), here in SHARED memory —
// the gather of 80-byte records was the other big client of the L1 data pipe.
template <bool kFpWay = false, bool kUnrolled = true>
__device__ __forceinline__ void sweepConstTable(const MomentTable &table, int groups, const double *exact, V3 o, V3 d,
                                                Nearest &best) {
  const MomentRay r = makeMomentRay(o, d);
  uint32_t rejectedHi = 0xffffffffu, rejectedLo = 0xffffffffu;
  // kUnrolled: immediate table addresses (LDCU into uniform registers, straight-line code);
  // otherwise a rolled loop with register-indexed constant loads (LDC.64 into vector registers).
#pragma unroll(kUnrolled ? kConstGroups : 1)
  for (int g = 0; g < (kUnrolled ? kConstGroups : groups); ++g) {
    if (!kUnrolled || g < groups) { // warp-uniform
      const float4(&a)[kMomentFloats] = table.group[g];
      uint32_t r0, r1, r2bits, r3; // in triangle-index order
      if (table.fanGroups & (1u << g)) // warp-uniform; lanes [A0, A1, B0, B1] = triangles 4g + {0, 2, 1, 3}
        stage0RejectFan4(a, r, r0, r2bits, r1, r3);
      else
        stage0RejectMoment4(a, r, r0, r1, r2bits, r3);
      rejectedHi = __funnelshift_l(rejectedLo, rejectedHi, 4);
      rejectedLo = __funnelshift_l(r0, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r1, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r2bits, rejectedLo, 1);
      rejectedLo = __funnelshift_l(r3, rejectedLo, 1);
    }
  }
  if (groups <= 0)
    return;
  // left-align: triangle k at bit 63 - k; ascending index is the serial loop's tie-break order
  unsigned long long keep = ~((static_cast<unsigned long long>(rejectedHi) << 32) | rejectedLo) << (64 - 4 * groups);
  while (keep) {
    const int i = __clzll(static_cast<long long>(keep));
    keep &= ~(0x8000000000000000ull >> i);
    const double2 *record = reinterpret_cast<const double2 *>(exact + 10 * i);
    const double2 a0 = record[0], a1 = record[1], a2 = record[2], a3 = record[3], a4 = record[4];
    testTriangle<kFpWay>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d, i, best);
  }
}

// ---- hit epilogues (Scene.cpp:38-48 spheres, :99-112 triangles) ------------------------------
__device__ __forceinline__ HitInfo finishHit(const DeviceScene &scene,
                                             const double4 *__restrict__ spheres, V3 o, V3 d,
                                             const Nearest &best) {
  HitInfo hit;
  hit.position = positionAlong(o, d, best.t);
  if (best.prim < 0) {
    const int i = -best.prim - 1;
    const double4 s = spheres[i];
    V3 normal = normalised(sub(hit.position, mk(s.x, s.y, s.z)));
    hit.inside = dot(normal, d) > 0;
    hit.normal = hit.inside ? neg(normal) : normal;
    hit.material = __ldg(scene.sphereMaterial + i);
    hit.triangle = -1;
  } else {
    // Three equal vertex normals make Scene.cpp:99-107 independent of u,v; the value and the
    // local bases of both orientations are precomputed at upload (ptb200_shim.cu) with the
    // same arithmetic basisFromZ() uses, so loading them is bit-identical to recomputing.
    const double4 r0 = ldgDouble4(scene.triShade + 4 * static_cast<size_t>(best.prim));
    const bool backfacing = best.det < kEpsilon;
    hit.inside = backfacing;
    hit.material = static_cast<uint32_t>(r0.w);
    hit.triangle = best.prim;
    hit.normal = backfacing ? mk(-r0.x, -r0.y, -r0.z) : mk(r0.x, r0.y, r0.z);
  }
  return hit;
}

// OrthoNormalBasis::fromZ(hit.normal): loaded for triangles (precomputed for both orientations
// at upload with basisFromZ()'s arithmetic), computed for spheres.
__device__ __forceinline__ void hitBasis(const DeviceScene &scene, const HitInfo &hit, V3 &basisX,
                                         V3 &basisY) {
  if (hit.triangle >= 0) {
    const double4 *record = scene.triShade + 4 * static_cast<size_t>(hit.triangle);
    const double4 r2 = ldgDouble4(record + 2);
    if (hit.inside) { // backfacing
      const double4 r3 = ldgDouble4(record + 3);
      basisX = mk(r2.z, r2.w, r3.x);
      basisY = mk(r3.y, r3.z, r3.w);
    } else {
      const double4 r1 = ldgDouble4(record + 1);
      basisX = mk(r1.x, r1.y, r1.z);
      basisY = mk(r1.w, r2.x, r2.y);
    }
  } else {
    const Basis basis = basisFromZ(hit.normal);
    basisX = basis.x;
    basisY = basis.y;
  }
}

struct MaterialView {
  const double *m;
  __device__ __forceinline__ V3 emission() const { return mk(__ldg(m + 0), __ldg(m + 1), __ldg(m + 2)); }
  __device__ __forceinline__ V3 diffuse() const { return mk(__ldg(m + 3), __ldg(m + 4), __ldg(m + 5)); }
  __device__ __forceinline__ double indexOfRefraction() const { return __ldg(m + 6); }
  __device__ __forceinline__ double reflectivity() const { return __ldg(m + 7); }
  __device__ __forceinline__ double coneAngle() const { return __ldg(m + 8); }
  __device__ __forceinline__ double inverseIndexOfRefraction() const { return __ldg(m + 9); }
};
__device__ __forceinline__ MaterialView materialOf(const DeviceScene &scene, uint32_t index) {
  return MaterialView{scene.materials + 10 * static_cast<size_t>(index)};
}

// Reflectivity at a hit (Scene.cpp:140-146).
__device__ __forceinline__ double hitReflectivity(const MaterialView &mat, const HitInfo &hit,
                                                  V3 incoming) {
  const double fixed = mat.reflectivity();
  if (!(fixed < 0))
    return fixed;
  // iorFrom / iorTo is ior / 1.0 == ior from inside, 1.0 / ior (precomputed, same IEEE
  // division) from outside.
  const double ior = mat.indexOfRefraction();
  return reflectance(hit.normal, incoming, hit.inside ? ior : 1.0, hit.inside ? 1.0 : ior,
                     hit.inside ? ior : mat.inverseIndexOfRefraction());
}

// Camera::randomRay + rayFromUnit (Camera.h:20-37,54-60) given its four uniform draws.
__device__ __forceinline__ void cameraRay(const DeviceCamera &cam, int pixelX, int pixelY,
                                          double ux, double uy, double uAngle, double uRadius,
                                          V3 &origin, V3 &direction) {
  const double x = (pixelX + ux) * cam.reciprocalWidth;
  const double y = (pixelY + uy) * cam.reciprocalHeight;
  const double xu = 2 * x - 1;
  const double yu = 2 * y - 1;
  const V3 xContrib = scale(scale(cam.axisX, -xu), cam.aspectRatio);
  const V3 yContrib = scale(cam.axisY, -yu);
  const V3 zContrib = scale(cam.axisZ, cam.cameraPlaneDist);
  const V3 dir = normalised(add(add(xContrib, yContrib), zContrib));
  if (cam.apertureRadius == 0) {
    origin = cam.centre;
    direction = dir;
    return;
  }
  const V3 focalPoint = positionAlong(cam.centre, dir, cam.focalDistance);
  const double angle = uAngle * (2 * kPi);
  const double radius = uRadius * cam.apertureRadius;
  double sinA, cosA;
  sinCos(angle, sinA, cosA);
  origin = add(add(cam.centre, scale(scale(cam.axisX, cosA), radius)),
               scale(scale(cam.axisY, sinA), radius));
  direction = normalised(sub(focalPoint, origin));
}

// result += emission + radiance  /  result += emission + diffuse * radiance (Scene.cpp:168,172-174)
__device__ __forceinline__ V3 shadeTerm(const MaterialView &mat, bool specular, V3 incoming) {
  const V3 e = mat.emission();
  if (specular)
    return add(e, incoming);
  const V3 k = mat.diffuse();
  return mk(fma(k.x, incoming.x, e.x), fma(k.y, incoming.y, e.y), fma(k.z, incoming.z, e.z));
}

// The `fp` way adds the emission after averaging (src/fp/Render.cpp:118), so a 1x1 level is
// emission + (0 + term) * (1/1) with term = radiance(child) or diffuse * radiance(child)
// (:66-73): the product is rounded on its own, unlike shadeTerm()'s.
__device__ __forceinline__ V3 fpSubSampleTerm(const MaterialView &mat, bool specular, V3 incoming) {
  if (specular)
    return incoming;
  const V3 k = mat.diffuse();
  return mk(k.x * incoming.x, k.y * incoming.y, k.z * incoming.z);
}
__device__ __forceinline__ V3 fpLevelRadiance(const MaterialView &mat, V3 incomingLight, double reciprocal) {
  const V3 e = mat.emission();
  return mk(fma(incomingLight.x, reciprocal, e.x), fma(incomingLight.y, reciprocal, e.y),
            fma(incomingLight.z, reciprocal, e.z));
}

// The `oo` way's level value (src/oo/Renderer.cpp:90): Material::totalEmission(result / n) is
// Vec3::operator/ (a multiply by the reciprocal, Vec3.h:51-54) in the caller and
// emission + inbound inside a virtual function (src/oo/Material.cpp:19-22): rounded separately,
// unlike fpLevelRadiance()'s fused form.
__device__ __forceinline__ V3 ooLevelRadiance(const MaterialView &mat, V3 incomingLight, double reciprocal) {
  return add(mat.emission(), scale(incomingLight, reciprocal));
}

// ---- TMA / mbarrier primitives (sm_90+ PTX; SASS: UBLKCP, SYNCS) ----------------------------
__device__ __forceinline__ uint32_t smemAddress(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbarInit(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddress(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddress(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra WAIT_DONE;\n"
               "bra WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}" ::"r"(smemAddress(bar)),
               "r"(parity)
               : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tmaLoad1D(void *dstShared, const void *srcGlobal, uint32_t bytes,
                                          uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemAddress(dstShared)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddress(bar))
               : "memory");
}
__device__ __forceinline__ void fenceBarrierInit() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

} // namespace ptb200
