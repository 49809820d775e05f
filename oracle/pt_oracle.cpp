// ============================================================================================
// TEST INFRASTRUCTURE — CPU ORACLE.  NOT PRODUCT CODE.
//
// A plain C++ restatement of the `dod` renderer hot path of mattgodbolt/pt-three-ways, used
// ONLY as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs.  Nothing under pt_three_ways_b200/ may include, link, load or call
// anything in this directory; the product fails loudly without its CUDA library instead.
//
// Parity status: PINNED.  oracle/Makefile builds oracle/_ref/ref_tool from the reference's own
// unmodified sources (src/dod, src/math, src/util); tests/test_oracle_vs_reference.py and
// tests/golden/make_golden.py compare this restatement against it (per-pass raw framebuffers,
// intersection records, the reference's own known-answer tests test/dod/*.cpp).
//
// What is restated (all file:line relative to /root/reference):
//   src/dod/Scene.cpp:14-49     intersectSpheres      -> intersectSpheres()
//   src/dod/Scene.cpp:52-113    intersectTriangles    -> intersectTriangles()
//   src/dod/Scene.cpp:115-122   intersect             -> intersect()
//   src/dod/Scene.cpp:124-179   radiance              -> radiance()
//   src/dod/Scene.cpp:181-195   addTriangle/addSphere -> Scene construction in oracle_scene_create()
//   src/dod/Scene.cpp:198-254   render (per-pass walk)-> renderPass()
//   src/math/Camera.h:20-37,54-60                     -> cameraRandomRay()
//   src/math/Norm3.cpp:7-24     reflectance           -> reflectance()
//   src/math/Norm3.impl.h:41-44 reflect               -> reflect()
//   src/math/OrthoNormalBasis.cpp:40-51 fromZ         -> basisFromZ()
//   src/math/Samples.cpp:6-30   cone/hemisphereSample -> coneSample(), hemisphereSample()
//   src/util/SampledPixel.cpp:3-17, ArrayOutput.cpp:39-56 -> accumulation in oracle_render()
//   libstdc++ 13 bits/random.tcc:3349-3381 generate_canonical, bits/random.h:1903-1909
//   uniform_real_distribution (third-party, pinned GCC 13.3.0)  -> canonicalFromWords()
//   std::mt19937 (ISO C++ [rand.predef])              -> Mt19937
//
// Arithmetic contract ("canonical rounding").  The reference is built with
// `-march=native -funsafe-math-optimizations` (CMakeLists.txt:21), i.e. its FMA contraction
// is whatever GCC chose; no rounding sequence is pinned by the reference.  This oracle fixes
// one: IEEE-754 binary64, round-to-nearest-even, FMA exactly where written (std::fma), no
// other contraction (-ffp-contract=off), correctly rounded / and sqrt, and its own
// sin/cos/acos (below) instead of glibc's.  The CUDA product implements the same sequence
// with the same constants, so product-vs-oracle comparisons are expected to be BIT-EXACT,
// while oracle-vs-reference is asserted to 1e-12 (measured on every golden image: exactly 0,
// because pixel values are sums of products of material constants selected by the discrete
// hit sequence; see tests/test_oracle_golden.py).
//
// Random-number / estimator policies:
//   RNG_MT19937_SEQUENTIAL (1): one std::mt19937(seed+s) per pass consumed pixel after pixel
//       in row-major order, exactly the reference stream (Scene.cpp:211-216).
//   RNG_KEYED_PHILOX (0): Philox4x32-10 keyed by (seed+s) with counter
//       (pixel, subPath, depth+1 | 0 for the camera, call); the same two-words-to-double rule.
//       This is the throughput policy of the CUDA product; it is NOT the reference's stream
//       (SURVEY.md section 0 item 3 explains why a per-pixel generator cannot be).
//   RNG_MT19937_PER_PIXEL (2): the reference's `fp` way, one std::mt19937 per (pass, pixel)
//       (src/fp/Render.cpp:76-135) -> radianceFp(uMajor = false).
//   RNG_MT19937_SEQUENTIAL_OO (3): the reference's `oo` way (src/oo/Renderer.cpp:60-106):
//       dod's per-pass sequential stream and u-major strata with the fp way's estimator
//       (emission added after the average, t == Epsilon accepted) -> radianceFp(uMajor = true).
// ============================================================================================
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

constexpr double Epsilon = 0.000000001; // src/math/Epsilon.h:3
constexpr double Pi = 3.14159265358979323846;

// ------------------------------------------------------------------------------------------
// Vectors (src/math/Vec3.h:8-107, Norm3.h:7-49).  Canonical rounding: see header.
// ------------------------------------------------------------------------------------------
struct V3 {
  double x, y, z;
};

inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 scale(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline double dot(V3 a, V3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline V3 cross(V3 a, V3 b) { // Vec3.h:87-92
  return {std::fma(a.y, b.z, -(a.z * b.y)), std::fma(a.z, b.x, -(a.x * b.z)),
          std::fma(a.x, b.y, -(a.y * b.x))};
}
// Vec3::normalised (Vec3.impl.h:5-7) = *this / length(); operator/ multiplies by the
// reciprocal (Vec3.h:51-54).
inline V3 normalised(V3 a) {
  const double reciprocal = 1.0 / std::sqrt(dot(a, a));
  return scale(a, reciprocal);
}
// Ray::positionAlong (Ray.h:25-27): origin + direction * t.
inline V3 positionAlong(V3 o, V3 d, double t) {
  return {std::fma(d.x, t, o.x), std::fma(d.y, t, o.y), std::fma(d.z, t, o.z)};
}

// ------------------------------------------------------------------------------------------
// Elementary functions with a fixed rounding sequence (shared, by construction, with the
// CUDA product).  Arguments on this path are bounded: sin/cos see [-pi, 2*pi], acos sees
// [0, 1).  Polynomial coefficients are the classic fdlibm kernel constants.
// ------------------------------------------------------------------------------------------
inline double kernelSin(double r) { // |r| <= pi/4 (+ slack)
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = r * r;
  double p = std::fma(S6, z, S5);
  p = std::fma(p, z, S4);
  p = std::fma(p, z, S3);
  p = std::fma(p, z, S2);
  p = std::fma(p, z, S1);
  return std::fma(r * z, p, r);
}
inline double kernelCos(double r) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = r * r;
  double p = std::fma(C6, z, C5);
  p = std::fma(p, z, C4);
  p = std::fma(p, z, C3);
  p = std::fma(p, z, C2);
  p = std::fma(p, z, C1);
  // 1 - z/2 + z*z*p
  return std::fma(z * z, p, std::fma(-0.5, z, 1.0));
}
// sin and cos of x for |x| <= ~8 (quadrant reduction with a two-constant pi/2).
inline void sinCos(double x, double &s, double &c) {
  const double TwoOverPi = 6.36619772367581382433e-01;
  const double PiO2Hi = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double PiO2Lo = 6.07710050650619224932e-11;  // pi/2 - PiO2Hi
  const double kd = std::nearbyint(x * TwoOverPi);    // round half to even
  const int k = static_cast<int>(kd);
  double r = std::fma(-kd, PiO2Hi, x);
  r = std::fma(-kd, PiO2Lo, r);
  const double sr = kernelSin(r);
  const double cr = kernelCos(r);
  switch (k & 3) {
  case 0: s = sr; c = cr; break;
  case 1: s = cr; c = -sr; break;
  case 2: s = -sr; c = -cr; break;
  default: s = -cr; c = sr; break;
  }
}
// acos(x) for x in [0, 1]: pi/2 - asin(x) below 0.5, 2*asin(sqrt((1-x)/2)) above, with the
// fdlibm rational asin core R(z) = z*P(z)/Q(z).
inline double asinCore(double z) {
  const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
               pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
               pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05;
  const double qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
               qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
  double p = std::fma(pS5, z, pS4);
  p = std::fma(p, z, pS3);
  p = std::fma(p, z, pS2);
  p = std::fma(p, z, pS1);
  p = std::fma(p, z, pS0);
  p = p * z;
  double q = std::fma(qS4, z, qS3);
  q = std::fma(q, z, qS2);
  q = std::fma(q, z, qS1);
  q = std::fma(q, z, 1.0);
  return p / q;
}
inline double arcCos(double x) {
  const double PiO2Hi = 1.57079632679489655800e+00, PiO2Lo = 6.12323399573676603587e-17;
  if (x < 0.5) {
    const double r = asinCore(x * x);
    // pi/2 - (x + x*r)
    return PiO2Hi - (x - std::fma(-x, r, PiO2Lo));
  }
  const double z = (1.0 - x) * 0.5;
  const double s = std::sqrt(z);
  const double r = asinCore(z);
  return 2.0 * std::fma(s, r, s);
}

// ------------------------------------------------------------------------------------------
// Scene data (src/dod/Scene.h:21-31).  Per-triangle derived values that the reference
// recomputes on every call are computed once here with the identical expressions.
// ------------------------------------------------------------------------------------------
struct Material { // src/util/MaterialSpec.h:7-12
  V3 emission, diffuse;
  double indexOfRefraction, reflectivity, reflectionConeAngleRadians;
};
struct Triangle {
  V3 v0, e1, e2; // vertex(0), uVector() = v1-v0, vVector() = v2-v0 (TriangleVertices.h:21-31)
  V3 shadingNormal; // what Scene.cpp:99-107 evaluates to for three equal vertex normals
  uint32_t material;
};
struct Sphere { // src/dod/Sphere.h:7-12
  V3 centre;
  double radiusSquared;
  uint32_t material;
};
struct Scene {
  std::vector<Triangle> triangles;
  std::vector<Sphere> spheres;
  std::vector<Material> materials;
  V3 environment{0, 0, 0};
};
struct Camera { // src/math/Camera.h:10-18, declaration order
  V3 centre, axisX, axisY, axisZ;
  double aspectRatio, cameraPlaneDist, reciprocalHeight, reciprocalWidth, apertureRadius,
      focalDistance;
};
struct Params { // src/util/RenderParams.h:3-13
  int width, height, preview, samplesPerPixel, maxCpus, maxDepth, firstBounceUSamples,
      firstBounceVSamples, seed;
};

struct HitRecord { // src/math/Hit.h:6-11 + the material reference of IntersectionRecord.h:8-11
  double distance;
  bool inside;
  V3 position, normal;
  uint32_t material;
  int32_t primitive; // >=0 triangle index, <0: -(sphere index)-1   (oracle-only bookkeeping)
};

// Scene.cpp:181-187: the three vertex normals are three copies of faceNormal();
// Scene.cpp:99-107 then evaluates normalised(u*(n1-n0) + v*(n2-n0) + n0).  With n1==n2==n0
// the deltas are exactly +0, u*0 and v*0 are +0 for the finite non-negative u,v that reach
// that line, and (+0 + +0) + n0 == n0 except that a -0 component becomes +0; the value is
// therefore independent of u,v: normalised(n0 + 0).
inline V3 triangleShadingNormal(V3 e1, V3 e2) {
  const V3 face = normalised(cross(e1, e2)); // TriangleVertices.h:33-35
  const V3 summed = {0.0 + face.x, 0.0 + face.y, 0.0 + face.z};
  return normalised(summed);
}

// ------------------------------------------------------------------------------------------
// Intersection (Scene.cpp:14-122).
// ------------------------------------------------------------------------------------------
struct Counters {
  uint64_t casts{0};
  uint64_t rngWords{0};
};

inline bool intersectSpheres(const Scene &scene, V3 origin, V3 direction, double nearerThan,
                             HitRecord &out) { // Scene.cpp:14-49
  double currentNearestDist = nearerThan;
  int nearestIndex = -1;
  for (size_t i = 0; i < scene.spheres.size(); ++i) {
    const Sphere &sphere = scene.spheres[i];
    const V3 op = sub(sphere.centre, origin);
    const double b = dot(op, direction);
    double determinant = std::fma(b, b, -dot(op, op)) + sphere.radiusSquared;
    if (determinant < 0)
      continue;
    determinant = std::sqrt(determinant);
    const double minusT = b - determinant;
    const double plusT = b + determinant;
    if (minusT < Epsilon && plusT < Epsilon)
      continue;
    const double t = minusT > Epsilon ? minusT : plusT;
    if (t < currentNearestDist) {
      nearestIndex = static_cast<int>(i);
      currentNearestDist = t;
    }
  }
  if (nearestIndex < 0)
    return false;
  const Sphere &sphere = scene.spheres[nearestIndex];
  const V3 hitPosition = positionAlong(origin, direction, currentNearestDist);
  V3 normal = normalised(sub(hitPosition, sphere.centre));
  const bool inside = dot(normal, direction) > 0;
  if (inside)
    normal = neg(normal);
  out = HitRecord{currentNearestDist, inside, hitPosition, normal, sphere.material,
                  -nearestIndex - 1};
  return true;
}

// fpWay: fp::Triangle::intersect (src/fp/Triangle.cpp:9-41) is the same arithmetic, but rejects
// `t < Epsilon` where Scene.cpp:93 accepts `t > Epsilon`: t == Epsilon exactly is a hit there.
inline bool intersectTriangles(const Scene &scene, V3 origin, V3 direction, double nearerThan,
                               HitRecord &out, bool fpWay = false) { // Scene.cpp:52-113
  double currentNearestDist = nearerThan;
  int nearestIndex = -1;
  double nearestDet = 0;
  for (size_t i = 0; i < scene.triangles.size(); ++i) {
    const Triangle &tri = scene.triangles[i];
    const V3 pVec = cross(direction, tri.e2);
    const double det = dot(tri.e1, pVec);
    if (std::fabs(det) < Epsilon)
      continue;
    const double invDet = 1.0 / det;
    const V3 tVec = sub(origin, tri.v0);
    const double u = dot(tVec, pVec) * invDet;
    const V3 qVec = cross(tVec, tri.e1);
    const double v = dot(direction, qVec) * invDet;
    if ((u < 0.0) | (u > 1.0) | (v < 0.0) | (u + v > 1)) // Unpredictable::any, Scene.cpp:89
      continue;
    const double t = dot(tri.e2, qVec) * invDet;
    if ((fpWay ? t >= Epsilon : t > Epsilon) && t < currentNearestDist) {
      nearestIndex = static_cast<int>(i);
      nearestDet = det;
      currentNearestDist = t;
    }
  }
  if (nearestIndex < 0)
    return false;
  const Triangle &tri = scene.triangles[nearestIndex];
  const bool backfacing = nearestDet < Epsilon; // Scene.cpp:108
  const V3 normal = backfacing ? neg(tri.shadingNormal) : tri.shadingNormal;
  out = HitRecord{currentNearestDist, backfacing,
                  positionAlong(origin, direction, currentNearestDist), normal, tri.material,
                  nearestIndex};
  return true;
}

// fpWay: fp::intersect (src/fp/Render.cpp:36-46) walks ONE primitive list in insertion order and
// keeps the strictly nearest hit.  The flat scene of this interface keeps triangles and spheres
// in separate arrays, so an exact distance tie between a sphere and a triangle resolves to the
// sphere here whatever the insertion order was (ties within a kind resolve to the lower index
// in both); the nearest distance itself is order-independent.
inline bool intersect(const Scene &scene, V3 origin, V3 direction, HitRecord &out,
                      Counters &counters, bool fpWay = false) { // Scene.cpp:115-122
  counters.casts++;
  HitRecord sphereRec, triangleRec;
  const bool hitSphere = intersectSpheres(scene, origin, direction,
                                          std::numeric_limits<double>::infinity(), sphereRec);
  const bool hitTriangle = intersectTriangles(
      scene, origin, direction,
      hitSphere ? sphereRec.distance : std::numeric_limits<double>::infinity(), triangleRec, fpWay);
  if (hitTriangle) {
    out = triangleRec;
    return true;
  }
  if (hitSphere) {
    out = sphereRec;
    return true;
  }
  return false;
}

// ------------------------------------------------------------------------------------------
// Random numbers.
// ------------------------------------------------------------------------------------------
// generate_canonical<double,53>(32-bit engine): two words, low first (random.tcc:3349-3381).
inline double canonicalFromWords(uint32_t lo, uint32_t hi) {
  const double sum = std::fma(static_cast<double>(hi), 4294967296.0, static_cast<double>(lo));
  double ret = sum * 5.42101086242752217003726400434970855712890625e-20; // exact 2^-64
  if (ret >= 1.0)
    ret = 0.99999999999999988897769753748434595763683319091796875; // nextafter(1, 0)
  return ret;
}

class Mt19937 { // ISO C++ mersenne_twister_engine<uint32_t,32,624,397,31,0x9908b0df,...>
  uint32_t state_[624];
  int index_{624};

  void refill() {
    for (int i = 0; i < 624; ++i) {
      const uint32_t y = (state_[i] & 0x80000000u) | (state_[(i + 1) % 624] & 0x7fffffffu);
      state_[i] = state_[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    index_ = 0;
  }

public:
  explicit Mt19937(uint32_t seed) {
    state_[0] = seed;
    for (int i = 1; i < 624; ++i)
      state_[i] = 1812433253u * (state_[i - 1] ^ (state_[i - 1] >> 30)) + static_cast<uint32_t>(i);
  }
  uint32_t next() {
    if (index_ >= 624)
      refill();
    uint32_t y = state_[index_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1, uint32_t out[4]) {
  for (int round = 0; round < 10; ++round) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
constexpr uint32_t PhiloxKeyHigh = 0xB200D0D0u;

// One generator object per pass; a "site" (pixel, subPath, depth) is announced before each
// group of draws.  The sequential policy ignores sites.
struct Rng {
  int mode; // 0 keyed Philox, 1 mt19937 sequential (one per pass), 2 mt19937 per sample (fp way),
            // 3 mt19937 sequential with the oo way's estimator
  Mt19937 mt;
  uint32_t key0;
  uint32_t pixel{0}, subPath{0}, level{0}, call{0};
  uint32_t buffered[4];
  int bufferedLeft{0};
  Counters *counters;

  Rng(int mode_, uint32_t passSeed, Counters *c) : mode(mode_), mt(passSeed), key0(passSeed), counters(c) {}

  void reseed(uint32_t seed) { mt = Mt19937(seed); } // mode 2: a fresh engine per (pass, pixel)
  void site(uint32_t pixel_, uint32_t subPath_, uint32_t level_) {
    pixel = pixel_; subPath = subPath_; level = level_; call = 0; bufferedLeft = 0;
  }
  uint32_t word() {
    counters->rngWords++;
    if (mode != 0)
      return mt.next();
    if (bufferedLeft == 0) {
      philox4x32_10(pixel, subPath, level, call++, key0, PhiloxKeyHigh, buffered);
      bufferedLeft = 4;
    }
    return buffered[4 - bufferedLeft--];
  }
  double canonical() {
    const uint32_t lo = word();
    const uint32_t hi = word();
    return canonicalFromWords(lo, hi);
  }
  // uniform_real_distribution<double>(a, b)(rng)  (bits/random.h:1903-1909)
  double uniform(double a, double b) { return (canonical() * (b - a)) + a; }
};

// ------------------------------------------------------------------------------------------
// Shading helpers.
// ------------------------------------------------------------------------------------------
inline double reflectance(V3 normal, V3 incoming, double iorFrom, double iorTo) { // Norm3.cpp:7-24
  const double iorRatio = iorFrom / iorTo;
  const double cosThetaI = -dot(normal, incoming);
  const double sinThetaTSquared = (iorRatio * iorRatio) * std::fma(-cosThetaI, cosThetaI, 1.0);
  if (sinThetaTSquared > 1)
    return 1.0;
  const double cosThetaT = std::sqrt(1 - sinThetaTSquared);
  const double a = iorFrom * cosThetaI;
  const double b = iorTo * cosThetaT;
  const double rPerpendicular = (a - b) / (a + b);
  // Norm3.cpp:19-23 evaluates rParallel with the same formula, so
  // (rPerp^2 + rPar^2) / 2 == rPerp^2 exactly in binary arithmetic.
  return rPerpendicular * rPerpendicular;
}
inline V3 reflect(V3 normal, V3 incoming) { // Norm3.impl.h:41-44
  const double k = dot(normal, incoming);
  return {std::fma(-(normal.x * 2), k, incoming.x), std::fma(-(normal.y * 2), k, incoming.y),
          std::fma(-(normal.z * 2), k, incoming.z)};
}
struct Basis {
  V3 x, y, z;
};
inline Basis basisFromZ(V3 z) { // OrthoNormalBasis.cpp:36-51
  // z.dot(xAxis) is z.x exactly.  The helper-axis cross product written out:
  // xAxis x z = (0, -z.z, z.y), yAxis x z = (z.z, 0, -z.x) (the reference's generic cross()
  // yields the same values up to the sign of the zero component, which never survives the
  // sums it feeds).
  const V3 c = std::fabs(z.x) > 0.9999 ? V3{z.z, 0.0, -z.x} : V3{0.0, -z.z, z.y};
  const V3 xx = normalised(c);
  const V3 yy = normalised(cross(z, xx));
  return {xx, yy, z};
}
inline V3 transform(const Basis &b, V3 p) { // OrthoNormalBasis.h:18-20
  return {std::fma(b.z.x, p.z, std::fma(b.y.x, p.y, b.x.x * p.x)),
          std::fma(b.z.y, p.z, std::fma(b.y.y, p.y, b.x.y * p.x)),
          std::fma(b.z.z, p.z, std::fma(b.y.z, p.y, b.x.z * p.x))};
}
inline V3 coneSample(V3 direction, double coneTheta, double u, double v) { // Samples.cpp:6-19
  if (coneTheta < Epsilon)
    return direction;
  coneTheta = coneTheta * (1.0 - (2.0 * arcCos(u) / Pi));
  double radius, zScale, sinT, cosT;
  sinCos(coneTheta, radius, zScale);
  const double randomTheta = v * 2 * Pi;
  sinCos(randomTheta, sinT, cosT);
  const Basis basis = basisFromZ(direction);
  return normalised(transform(basis, V3{cosT * radius, sinT * radius, zScale}));
}
inline V3 hemisphereSample(const Basis &basis, double u, double v) { // Samples.cpp:21-30
  const double theta = (2 * Pi) * u;
  const double radiusSquared = v;
  const double radius = std::sqrt(radiusSquared);
  double sinT, cosT;
  sinCos(theta, sinT, cosT);
  return normalised(transform(basis, V3{cosT * radius, sinT * radius, std::sqrt(1 - radiusSquared)}));
}

// Camera::randomRay / rayFromUnit (Camera.h:20-37,54-60).
inline void cameraRandomRay(const Camera &cam, int pixelX, int pixelY, Rng &rng, V3 &origin,
                            V3 &direction) {
  const double ux = rng.uniform(0.0, 1.0);
  const double uy = rng.uniform(0.0, 1.0);
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
  const double angle = rng.uniform(0.0, 2 * Pi);
  const double radius = rng.uniform(0.0, cam.apertureRadius);
  double sinA, cosA;
  sinCos(angle, sinA, cosA);
  origin = add(add(cam.centre, scale(scale(cam.axisX, cosA), radius)),
               scale(scale(cam.axisY, sinA), radius));
  direction = normalised(sub(focalPoint, origin)); // Ray::fromTwoPoints, Ray.h:13-16
}

// Scene::radiance (Scene.cpp:124-179).  `subPath` identifies the depth-0 stratum for the
// keyed RNG policy (ignored by the sequential policy).
V3 radiance(const Scene &scene, Rng &rng, uint32_t pixel, uint32_t subPath, V3 origin,
            V3 direction, int depth, const Params &params, Counters &counters) {
  const int numUSamples = depth == 0 ? params.firstBounceUSamples : 1;
  const int numVSamples = depth == 0 ? params.firstBounceVSamples : 1;
  if (depth >= params.maxDepth)
    return V3{0, 0, 0};
  HitRecord hit;
  if (!intersect(scene, origin, direction, hit, counters))
    return scene.environment;
  const Material &mat = scene.materials[hit.material];
  if (params.preview)
    return mat.diffuse;
  const double iorFrom = hit.inside ? mat.indexOfRefraction : 1.0;
  const double iorTo = hit.inside ? 1.0 : mat.indexOfRefraction;
  const double reflectivity =
      mat.reflectivity < 0 ? reflectance(hit.normal, direction, iorFrom, iorTo) : mat.reflectivity;
  const Basis basis = basisFromZ(hit.normal);
  V3 result{0, 0, 0};
  for (int uSample = 0; uSample < numUSamples; ++uSample) {
    for (int vSample = 0; vSample < numVSamples; ++vSample) {
      const uint32_t childSubPath =
          depth == 0 ? static_cast<uint32_t>(uSample * numVSamples + vSample) : subPath;
      rng.site(pixel, childSubPath, static_cast<uint32_t>(depth) + 1);
      const double u = (static_cast<double>(uSample) + rng.uniform(0, 1.0)) /
                       static_cast<double>(numUSamples);
      const double v = (static_cast<double>(vSample) + rng.uniform(0, 1.0)) /
                       static_cast<double>(numVSamples);
      const double p = rng.uniform(0, 1.0);
      if (p < reflectivity) {
        const V3 newDir =
            coneSample(reflect(hit.normal, direction), mat.reflectionConeAngleRadians, u, v);
        const V3 incoming = radiance(scene, rng, pixel, childSubPath, hit.position, newDir,
                                     depth + 1, params, counters);
        result = add(result, add(mat.emission, incoming));
      } else {
        const V3 newDir = hemisphereSample(basis, u, v);
        const V3 incoming = radiance(scene, rng, pixel, childSubPath, hit.position, newDir,
                                     depth + 1, params, counters);
        const V3 term = {std::fma(mat.diffuse.x, incoming.x, mat.emission.x),
                         std::fma(mat.diffuse.y, incoming.y, mat.emission.y),
                         std::fma(mat.diffuse.z, incoming.z, mat.emission.z)};
        result = add(result, term);
      }
    }
  }
  const double reciprocal = 1.0 / static_cast<double>(numUSamples * numVSamples); // Vec3.h:51-54
  return scale(result, reciprocal);
}

// fp::radiance + radianceAtIntersection (src/fp/Render.cpp:48-118): the same estimator as
// Scene::radiance with three differences that change the numbers:
//   * the strata are walked v-major (cartesian_product(ints(0,numV), ints(0,numU)), :109-110),
//     each drawing u, then v (toUVSample, :97-102), then p (:116);
//   * a sub-sample contributes  radiance(child)  or  diffuse * radiance(child)  WITHOUT the
//     emission (:66-73); the emission is added once, after the average:
//     emission + sum / (numU*numV)  (:118);
//   * hits come from fp::intersect (see intersect() above).
//
// uMajor = true is oo::Renderer::radiance (src/oo/Renderer.cpp:60-91) instead: the same
// post-average emission (Material::totalEmission(result / (numU*numV)), :90 with
// src/oo/Material.cpp:19-22; the sub-sample terms are MatteMaterial/ShinyMaterial::sample,
// Material.cpp:28-67, i.e. radiance(child) or diffuse * radiance(child)) and the same
// t == Epsilon acceptance (src/oo/Triangle.cpp:31), but with the strata walked u-major
// (:79-80) like Scene.cpp:155-156.  The virtual calls (Material::sample, totalEmission) keep
// GCC from contracting across them: diffuse * child, result += term, result * (1/n) and
// emission + inbound are each rounded on their own.
V3 radianceFp(const Scene &scene, Rng &rng, V3 origin, V3 direction, int depth, const Params &params,
              Counters &counters, bool uMajor = false) {
  const int numUSamples = depth == 0 ? params.firstBounceUSamples : 1;
  const int numVSamples = depth == 0 ? params.firstBounceVSamples : 1;
  if (depth >= params.maxDepth)
    return V3{0, 0, 0};
  HitRecord hit;
  if (!intersect(scene, origin, direction, hit, counters, true))
    return scene.environment;
  const Material &mat = scene.materials[hit.material];
  if (params.preview)
    return mat.diffuse;
  const Basis basis = basisFromZ(hit.normal);
  const double iorFrom = hit.inside ? mat.indexOfRefraction : 1.0;
  const double iorTo = hit.inside ? 1.0 : mat.indexOfRefraction;
  const double reflectivity =
      mat.reflectivity < 0 ? reflectance(hit.normal, direction, iorFrom, iorTo) : mat.reflectivity;
  V3 incomingLight{0, 0, 0}; // accumulate(..., Vec3()): init = init + element, in order
  const int numStrata = numUSamples * numVSamples;
  for (int stratum = 0; stratum < numStrata; ++stratum) {
    {
      const int uSample = uMajor ? stratum / numVSamples : stratum % numUSamples;
      const int vSample = uMajor ? stratum % numVSamples : stratum / numUSamples;
      const double u = (static_cast<double>(uSample) + rng.uniform(0, 1.0)) /
                       static_cast<double>(numUSamples);
      const double v = (static_cast<double>(vSample) + rng.uniform(0, 1.0)) /
                       static_cast<double>(numVSamples);
      const double p = rng.uniform(0, 1.0);
      if (p < reflectivity) {
        const V3 newDir =
            coneSample(reflect(hit.normal, direction), mat.reflectionConeAngleRadians, u, v);
        incomingLight = add(incomingLight, radianceFp(scene, rng, hit.position, newDir, depth + 1,
                                                      params, counters, uMajor));
      } else {
        const V3 newDir = hemisphereSample(basis, u, v);
        const V3 child = radianceFp(scene, rng, hit.position, newDir, depth + 1, params, counters, uMajor);
        incomingLight = add(incomingLight, V3{mat.diffuse.x * child.x, mat.diffuse.y * child.y,
                                              mat.diffuse.z * child.z});
      }
    }
  }
  const double reciprocal = 1.0 / static_cast<double>(numUSamples * numVSamples); // Vec3.h:51-54
  if (uMajor) { // oo: Vec3::operator/ in radiance(), operator+ inside the virtual totalEmission()
    const V3 average = scale(incomingLight, reciprocal);
    return add(mat.emission, average);
  }
  return {std::fma(incomingLight.x, reciprocal, mat.emission.x),
          std::fma(incomingLight.y, reciprocal, mat.emission.y),
          std::fma(incomingLight.z, reciprocal, mat.emission.z)};
}

// The engine seed of fp::renderWholeScreen's renderOnePixel (src/fp/Render.cpp:125-126):
// height*width*seed + x*width + y  — x*width, not y*width: the reference's own indexing —
// evaluated in size_t and reduced mod 2^32 by mersenne_twister_engine::seed.
inline uint32_t fpPixelSeed(const Params &params, int passSeed, int x, int y) {
  const uint64_t frame = static_cast<uint64_t>(static_cast<int64_t>(params.height * params.width));
  const uint64_t within = static_cast<uint64_t>(static_cast<int64_t>(x * params.width + y));
  return static_cast<uint32_t>(frame * static_cast<uint64_t>(static_cast<int64_t>(passSeed)) + within);
}

// One pass of Scene::render's lambda (Scene.cpp:208-220): per-pixel colours of pass `s`,
// rows [rowBegin, height) stepping rowStep for the keyed policy (the sequential policy
// must walk every pixel; rows outside the selection are traced but not stored).
void renderPass(const Scene &scene, const Camera &cam, const Params &params, int rngMode, int pass,
                int rowBegin, int rowStep, double *colours /* W*H*3 */, Counters &counters) {
  Rng rng(rngMode, static_cast<uint32_t>(params.seed + pass), &counters);
  for (int y = 0; y < params.height; ++y) {
    const bool selected = y >= rowBegin && (y - rowBegin) % rowStep == 0;
    if (!selected && rngMode != 1 && rngMode != 3)
      continue;
    for (int x = 0; x < params.width; ++x) {
      const uint32_t pixel = static_cast<uint32_t>(x + y * params.width);
      rng.site(pixel, 0, 0);
      if (rngMode == 2) // fp::renderWholeScreen: one engine per pixel of the pass
        rng.reseed(fpPixelSeed(params, params.seed + pass, x, y));
      V3 origin, direction;
      cameraRandomRay(cam, x, y, rng, origin, direction);
      const V3 colour = rngMode == 2 || rngMode == 3
                            ? radianceFp(scene, rng, origin, direction, 0, params, counters, rngMode == 3)
                            : radiance(scene, rng, pixel, 0, origin, direction, 0, params, counters);
      if (selected) {
        colours[3 * pixel + 0] = colour.x;
        colours[3 * pixel + 1] = colour.y;
        colours[3 * pixel + 2] = colour.z;
      }
    }
  }
}

} // namespace

// ============================================================================================
// C interface for ctypes (tests/, bench.py).  Flat arrays only.
// ============================================================================================
extern "C" {

struct OracleScene {
  Scene scene;
};

// triangleVertices: T*9 (v0,v1,v2); sphereCentreRadius: S*4; materials: M*9 doubles in
// MaterialSpec order {emission, diffuse, ior, reflectivity, cone}.
OracleScene *oracle_scene_create(uint32_t numTriangles, const double *triangleVertices,
                                 const uint32_t *triangleMaterial, uint32_t numSpheres,
                                 const double *sphereCentreRadius, const uint32_t *sphereMaterial,
                                 uint32_t numMaterials, const double *materials,
                                 const double *environment) {
  auto *handle = new OracleScene;
  Scene &scene = handle->scene;
  for (uint32_t i = 0; i < numMaterials; ++i) {
    const double *m = materials + 9 * i;
    scene.materials.push_back(Material{{m[0], m[1], m[2]}, {m[3], m[4], m[5]}, m[6], m[7], m[8]});
  }
  for (uint32_t i = 0; i < numTriangles; ++i) { // Scene::addTriangle, Scene.cpp:181-187
    const double *t = triangleVertices + 9 * i;
    const V3 v0{t[0], t[1], t[2]}, v1{t[3], t[4], t[5]}, v2{t[6], t[7], t[8]};
    Triangle tri;
    tri.v0 = v0;
    tri.e1 = sub(v1, v0);
    tri.e2 = sub(v2, v0);
    tri.shadingNormal = triangleShadingNormal(tri.e1, tri.e2);
    tri.material = triangleMaterial[i];
    scene.triangles.push_back(tri);
  }
  for (uint32_t i = 0; i < numSpheres; ++i) { // Scene::addSphere, Scene.cpp:189-193
    const double *s = sphereCentreRadius + 4 * i;
    scene.spheres.push_back(Sphere{{s[0], s[1], s[2]}, s[3] * s[3], sphereMaterial[i]});
  }
  scene.environment = V3{environment[0], environment[1], environment[2]};
  return handle;
}

void oracle_scene_destroy(OracleScene *handle) { delete handle; }

// out per ray: 12 doubles {hit(0/1), distance, inside(0/1), px,py,pz, nx,ny,nz, material,
// primitive, 0}.  which: 0 = intersect, 1 = intersectSpheres, 2 = intersectTriangles
// (nearerThan applies to 1 and 2).
void oracle_intersect(const OracleScene *handle, int which, double nearerThan, uint32_t numRays,
                      const double *rays, double *out) {
  Counters counters;
  for (uint32_t i = 0; i < numRays; ++i) {
    const V3 o{rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]};
    const V3 d{rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]};
    HitRecord hit{};
    bool found;
    if (which == 1)
      found = intersectSpheres(handle->scene, o, d, nearerThan, hit);
    else if (which == 2)
      found = intersectTriangles(handle->scene, o, d, nearerThan, hit);
    else
      found = intersect(handle->scene, o, d, hit, counters);
    double *r = out + 12 * i;
    std::memset(r, 0, 12 * sizeof(double));
    r[0] = found ? 1 : 0;
    if (found) {
      r[1] = hit.distance;
      r[2] = hit.inside ? 1 : 0;
      r[3] = hit.position.x; r[4] = hit.position.y; r[5] = hit.position.z;
      r[6] = hit.normal.x; r[7] = hit.normal.y; r[8] = hit.normal.z;
      r[9] = hit.material;
      r[10] = hit.primitive;
    }
  }
}

// Renders passes [passBegin, passBegin+numPasses) and accumulates them IN PASS ORDER
// (SampledPixel::accumulate, SampledPixel.cpp:3-6; the reference with --max-cpus 1 adds
// passes in launch order, Scene.cpp:242).  sums: W*H*3 doubles (zeroed here), counts: W*H.
// perPass (optional): numPasses*W*H*3 doubles receiving each pass's own image.
// stats (optional): [0]=casts [1]=rng words.  Threads work on whole passes.
void oracle_render(const OracleScene *handle, const double *camera18, const int32_t *params9,
                   int rngMode, int passBegin, int numPasses, int rowBegin, int rowStep,
                   int numThreads, double *sums, uint64_t *counts, double *perPass,
                   uint64_t *stats) {
  Camera cam;
  static_assert(sizeof(Camera) == 18 * sizeof(double), "camera layout");
  std::memcpy(&cam, camera18, sizeof cam);
  Params params{params9[0], params9[1], params9[2], params9[3], params9[4],
                params9[5], params9[6], params9[7], params9[8]};
  const size_t pixels = static_cast<size_t>(params.width) * params.height;
  if (rowStep < 1)
    rowStep = 1;
  std::vector<double> local;
  double *passImages = perPass;
  if (!passImages) {
    local.assign(static_cast<size_t>(numPasses) * pixels * 3, 0.0);
    passImages = local.data();
  } else {
    std::memset(passImages, 0, static_cast<size_t>(numPasses) * pixels * 3 * sizeof(double));
  }
  if (numThreads < 1)
    numThreads = 1;
  std::atomic<int> nextPass{0};
  std::vector<Counters> counters(numThreads);
  auto worker = [&](int threadIndex) {
    for (;;) {
      const int p = nextPass.fetch_add(1);
      if (p >= numPasses)
        break;
      renderPass(handle->scene, cam, params, rngMode, passBegin + p, rowBegin, rowStep,
                 passImages + static_cast<size_t>(p) * pixels * 3, counters[threadIndex]);
    }
  };
  std::vector<std::thread> threads;
  for (int t = 1; t < numThreads; ++t)
    threads.emplace_back(worker, t);
  worker(0);
  for (auto &t : threads)
    t.join();

  std::memset(sums, 0, pixels * 3 * sizeof(double));
  std::memset(counts, 0, pixels * sizeof(uint64_t));
  for (int p = 0; p < numPasses; ++p) {
    const double *image = passImages + static_cast<size_t>(p) * pixels * 3;
    for (int y = rowBegin; y < params.height; y += rowStep) {
      for (int x = 0; x < params.width; ++x) {
        const size_t pixel = static_cast<size_t>(x) + static_cast<size_t>(y) * params.width;
        sums[3 * pixel + 0] += image[3 * pixel + 0];
        sums[3 * pixel + 1] += image[3 * pixel + 1];
        sums[3 * pixel + 2] += image[3 * pixel + 2];
        counts[pixel] += 1;
      }
    }
  }
  if (stats) {
    stats[0] = stats[1] = 0;
    for (const Counters &c : counters) {
      stats[0] += c.casts;
      stats[1] += c.rngWords;
    }
  }
}

// Elementary-function probes so tests can pin them against libm.
void oracle_sincos(uint32_t n, const double *x, double *s, double *c) {
  for (uint32_t i = 0; i < n; ++i)
    sinCos(x[i], s[i], c[i]);
}
void oracle_acos(uint32_t n, const double *x, double *out) {
  for (uint32_t i = 0; i < n; ++i)
    out[i] = arcCos(x[i]);
}
void oracle_mt19937(uint32_t seed, uint32_t n, uint32_t *out) {
  Mt19937 mt(seed);
  for (uint32_t i = 0; i < n; ++i)
    out[i] = mt.next();
}
void oracle_philox(const uint32_t *counter4, const uint32_t *key2, uint32_t *out4) {
  philox4x32_10(counter4[0], counter4[1], counter4[2], counter4[3], key2[0], key2[1], out4);
}
double oracle_canonical(uint32_t lo, uint32_t hi) { return canonicalFromWords(lo, hi); }

} // extern "C"
