// The keyed policy as a three-kernel pipeline (the default since round 2; the one-kernel
// megakernel of pt_kernels.cu remains for the `fp` way, whose engine is sequential within a sample).
//
//   primaryHitsKernel     one lane per (pass, pixel) SAMPLE: Camera::randomRay, the first
//                         Scene::intersect, and everything radiance() holds at depth 0 when it
//                         enters its sampling loop (Scene.cpp:124-152) written as a 144-byte record.
//                         Samples that end at the camera ray (miss, preview, maxDepth <= 1) get
//                         their colour directly.
//   subPathKernel         persistent, one lane per SUB-PATH = (record, stratum): the loop body of
//                         Scene.cpp:155-175 for one (uSample, vSample) with the recursion below it.
//                         Every lane of every iteration does the same three things — bounce, cast,
//                         then shade the hit or unwind and take the next ticket — so there is no
//                         camera ray, no stratum bookkeeping and no sample completion inside the hot
//                         loop (the one-kernel form spent half of its instructions there at 7-28
//                         active lanes, profiles/r1t_hotspots.md).
//   resolveSamplesKernel  per pixel: the strata's terms added in stratum order and averaged
//                         (Scene.cpp:168,172-174,178), samples added to the accumulator in pass
//                         order (SampledPixel.cpp:3-6).
//
// Results are bit-identical to the one-kernel form (and to the oracle): the same arithmetic in
// the same order, only distributed differently over lanes.
// The arithmetic helpers that the one-kernel forms call out of line (correctly rounded sqrt / division,
// the sin/cos pair, Philox) are INLINED here: the sub-path kernel's hot loop is small enough for the
// instruction cache, and a call costs its argument / result moves and a scheduling barrier — 230.2 ->
// 239.1 (sqrt, division) -> 242.9 (+ sin/cos) -> 244.2 Msamples/s (+ Philox) on the Cornell box, same
// session (profiles/r2y_inline_ab.txt); the fp-way megakernel and the exact-stream kernel are faster
// with the calls (pt_kernels.cu keeps level 0), cone sampling is rare and stays out of line.
#ifndef PT_INLINE_LEVEL
#define PT_INLINE_LEVEL 3 // (before the first header that pulls in pt_math.cuh)
#endif
#include "pt_kernels.h"

#include "pt_device.cuh"
#include "pt_stage.cuh"

namespace ptb200 {

// ---- the record of a camera ray's hit: 9 x 16 bytes ------------------------------------------
//   q0 pos.x pos.y | q1 pos.z nrm.x | q2 nrm.y nrm.z | q3 inc.x inc.y | q4 inc.z bX.x
//   q5 bX.y bX.z   | q6 bY.x bY.y   | q7 bY.z reflectivity
//   q8 {material | pixel << 32} {key0 | sample << 32}
constexpr int kRecordQuads = 9;

__device__ __forceinline__ double packWords(uint32_t lo, uint32_t hi) {
  return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
}
__device__ __forceinline__ uint32_t lowWord(double v) { return static_cast<uint32_t>(__double2loint(v)); }
__device__ __forceinline__ uint32_t highWord(double v) { return static_cast<uint32_t>(__double2hiint(v)); }

__device__ __forceinline__ void storeRecord(double2 *record, const Surface &s, uint32_t pixel, uint32_t key0,
                                            uint32_t sample) {
  record[0] = make_double2(s.position.x, s.position.y);
  record[1] = make_double2(s.position.z, s.normal.x);
  record[2] = make_double2(s.normal.y, s.normal.z);
  record[3] = make_double2(s.incoming.x, s.incoming.y);
  record[4] = make_double2(s.incoming.z, s.basisX.x);
  record[5] = make_double2(s.basisX.y, s.basisX.z);
  record[6] = make_double2(s.basisY.x, s.basisY.y);
  record[7] = make_double2(s.basisY.z, s.reflectivity);
  record[8] = make_double2(packWords(s.material, pixel), packWords(key0, sample));
}

// Material of the nearest hit without the rest of the hit epilogue: all the deepest level of a
// path needs (its children return Vec3(), Scene.cpp:128-129, so it contributes its emission).
__device__ __forceinline__ uint32_t hitMaterial(const DeviceScene &scene, const Nearest &best) {
  if (best.prim < 0)
    return __ldg(scene.sphereMaterial + (-best.prim - 1));
  return static_cast<uint32_t>(__ldg(reinterpret_cast<const double *>(scene.triShade + 4 * static_cast<size_t>(best.prim)) + 3));
}

// Scene::intersect for one lane: spheres first, then every tile (Scene.cpp:115-122).  Whole-CTA
// when the scene streams (every thread takes part in the tile hand-over).
template <int kSweep>
__device__ __forceinline__ Nearest castRay(const DeviceScene &scene, TileStream &stream, bool resident, bool tracing,
                                           V3 origin, V3 direction, const MomentTable &table) {
  Nearest best{__longlong_as_double(0x7ff0000000000000ll), 0.0, kNoPrim};
  if (tracing)
    sweepSpheres(stream.spheres(), static_cast<int>(scene.numSpheres), origin, direction, best);
  if (kSweep == 9 || kSweep == 8) { // stage 0 out of the constant bank (8: rolled loop), survivors' records in shared memory
    if (tracing)
      sweepConstTable<false, kSweep == 9>(table, static_cast<int>(scene.tileTris / 4),
                                          reinterpret_cast<const double *>(stream.tile(0)), origin, direction, best);
    return best;
  }
  for (uint32_t j = 0; j < scene.numTiles; ++j) {
    const unsigned char *tile = resident ? stream.tile(0) : stream.acquire();
    if (tracing)
      sweepStagedTile<kSweep, false>(scene, tile, j, origin, direction, best);
    if (!resident)
      stream.release();
  }
  return best;
}

// The Surface of a hit a bounce will leave from (Scene.cpp:135-152).
__device__ __forceinline__ Surface surfaceOfHit(const DeviceScene &scene, const double4 *spheres, V3 origin,
                                                V3 direction, const Nearest &best) {
  const HitInfo hit = finishHit(scene, spheres, origin, direction, best);
  Surface surface;
  surface.position = hit.position;
  surface.normal = hit.normal;
  surface.incoming = direction;
  surface.material = hit.material;
  surface.reflectivity = hitReflectivity(materialOf(scene, hit.material), hit, direction);
  hitBasis(scene, hit, surface.basisX, surface.basisY);
  return surface;
}

// =============================================================================================
// 1. Camera rays and their hits.
// =============================================================================================
template <int kBlock, int kSweep>
__global__ void __launch_bounds__(kBlock) primaryHitsKernel(const __grid_constant__ SplitArgs args) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceScene &scene = args.scene;
  TileStream stream = makeTileStream(smemRaw, scene, kSweep);
  stream.start();
#pragma unroll 1
  for (uint32_t i = threadIdx.x; i < scene.numSpheres; i += kBlock)
    stream.spheres()[i] = scene.spheres[i];
  __syncthreads();
  const bool resident = scene.numTiles <= 1;
  if (resident && scene.numTiles == 1)
    stream.acquire();

  const unsigned lane = threadIdx.x & 31u;
  const uint32_t numSub = args.numSub;
  const double invNumSub = 1.0 / static_cast<double>(numSub); // Vec3::operator/ (Vec3.h:51-54)
  unsigned long long casts = 0;
  const uint32_t numChunks = (args.totalSamples + kBlock - 1) / kBlock;
  for (uint32_t chunk = blockIdx.x; chunk < numChunks; chunk += gridDim.x) {
    const uint32_t sample = chunk * kBlock + threadIdx.x;
    const bool live = sample < args.totalSamples;
    const uint32_t passInBatch = sample / args.ownPixels;
    const uint32_t own = sample % args.ownPixels;
    const int px = static_cast<int>(own % args.width);
    const int py = args.rowBegin + static_cast<int>(own / args.width) * args.rowStep;
    const uint32_t pixel = static_cast<uint32_t>(px) + static_cast<uint32_t>(py) * args.width;
    const uint32_t key0 = static_cast<uint32_t>(args.seed + args.passBegin + static_cast<int>(passInBatch));
    const bool tracing = live && args.maxDepth > 0; // radiance() returns Vec3() before intersecting (Scene.cpp:128-129)
    V3 origin = mk(0, 0, 0), direction = mk(0, 0, 1);
    if (tracing)
      keyedCameraRay(args.camera, key0, pixel, px, py, origin, direction);
    const Nearest best = castRay<kSweep>(scene, stream, resident, tracing, origin, direction, args.momentTable);
    casts += tracing ? 1u : 0u;

    bool isRecord = false;
    V3 colour = mk(0, 0, 0);
    Surface surface{};
    if (tracing) {
      if (best.prim == kNoPrim) {
        colour = mk(scene.environment[0], scene.environment[1], scene.environment[2]); // Scene.cpp:132-133
      } else if (args.preview) { // Scene.cpp:137-138
        colour = materialOf(scene, hitMaterial(scene, best)).diffuse();
      } else if (args.maxDepth <= 1) {
        // The sampling loop still runs, every child returns Vec3(): numSub emission terms, averaged.
        const V3 term = shadeTerm(materialOf(scene, hitMaterial(scene, best)), true, mk(0, 0, 0));
        V3 acc = mk(0, 0, 0);
#pragma unroll 1
        for (uint32_t k = 0; k < numSub; ++k)
          acc = add(acc, term);
        colour = scale(acc, invNumSub);
      } else {
        surface = surfaceOfHit(scene, stream.spheres(), origin, direction, best);
        isRecord = true;
      }
    }
    // Records are appended in whatever order warps arrive: results are addressed by sample.
    const unsigned recordMask = __ballot_sync(kFullMask, isRecord);
    if (recordMask) {
      unsigned long long base = 0;
      const int leader = __ffs(recordMask) - 1;
      if (static_cast<int>(lane) == leader)
        base = atomicAdd(args.counters + 2, static_cast<unsigned long long>(__popc(recordMask)));
      base = __shfl_sync(kFullMask, base, leader);
      if (isRecord) {
        const unsigned long long index = base + __popc(recordMask & ((1u << lane) - 1u));
        storeRecord(args.records + kRecordQuads * index, surface, pixel, key0, sample);
        args.sampleKind[sample] = 0;
      }
    }
    if (live && !isRecord) {
      double *slot = args.terms + 3 * static_cast<size_t>(sample) * numSub;
      slot[0] = colour.x;
      slot[1] = colour.y;
      slot[2] = colour.z;
      args.sampleKind[sample] = 1;
    }
  }
  if (!resident)
    stream.drain();
#pragma unroll 1
  for (int offset = 16; offset > 0; offset >>= 1)
    casts += __shfl_down_sync(kFullMask, casts, offset);
  if (lane == 0 && casts)
    atomicAdd(args.counters + 1, casts);
}

// =============================================================================================
// 2. Sub-paths.
// =============================================================================================
constexpr uint32_t kTicketGrab = 64; // tickets a warp takes per atomic (>= 32)

// Levels 1.. of a sub-path: material index and branch taken, for the unwind.  Paths of the
// reference's default depth keep them in one 64-bit register (16 bits per level); deeper ones in
// local memory.
template <bool kDeep>
struct LevelStack;
template <>
struct LevelStack<false> {
  static constexpr int kLevels = 4;
  unsigned long long bits;
  __device__ __forceinline__ void reset() { bits = 0; }
  __device__ __forceinline__ void set(int level, uint32_t material, bool specular) { // level >= 1, once per path
    bits |= static_cast<unsigned long long>(material | (specular ? 0x8000u : 0u)) << (16 * (level - 1));
  }
  __device__ __forceinline__ void get(int level, uint32_t &material, bool &specular) const {
    const uint32_t entry = static_cast<uint32_t>(bits >> (16 * (level - 1)));
    material = entry & 0x7fffu;
    specular = (entry & 0x8000u) != 0;
  }
};
template <>
struct LevelStack<true> {
  uint16_t material_[kMaxDepth];
  bool specular_[kMaxDepth];
  __device__ __forceinline__ void reset() {}
  __device__ __forceinline__ void set(int level, uint32_t material, bool specular) {
    material_[level] = static_cast<uint16_t>(material);
    specular_[level] = specular;
  }
  __device__ __forceinline__ void get(int level, uint32_t &material, bool &specular) const {
    material = material_[level];
    specular = specular_[level];
  }
};

// Surface word flag: the local basis of this surface has not been computed (a sphere hit inside
// the sub-path kernel).  OrthoNormalBasis::fromZ(hit.normal) is evaluated before the sampling loop
// in the reference (Scene.cpp:152) but only hemisphereSample reads it (:170), so evaluating it
// on the diffuse branch only is the same arithmetic on the same values.
constexpr uint32_t kLazyBasisFlag = 0x80000000u;

// The specular-only part of coneSample() (Samples.cpp:6-19): everything up to the final
// normalised(basis.transform(cos(t)*r, sin(t)*r, z)), which the diffuse lanes share.  Returns true
// when the cone is degenerate and the mirror direction itself is the answer.
__device__ __forceinline__ bool coneSetup(V3 mirror, double coneTheta, double u, double v, V3 &bx, V3 &by, V3 &bz,
                                          double &angle, double &radius, double &zScale) {
  if (coneTheta < kEpsilon)
    return true;
  coneTheta = coneTheta * (1.0 - ieeeDiv(2.0 * arcCos(u), kPi));
  sinCos(coneTheta, radius, zScale);
  angle = v * 2 * kPi;
  const Basis basis = basisFromZ(mirror);
  bx = basis.x;
  by = basis.y;
  bz = mirror;
  return false;
}

// One sub-path's state.  Kept small on purpose: the ray, six words of bookkeeping and the level
// stack.  The Surface a bounce leaves from never lives in registers across phases: a hit writes it
// to a shared-memory slot of the thread (same 9 x 16-byte layout as a camera-hit record), a new
// sub-path points at its record in global memory, and the bounce reads whichever through one
// generic pointer.
template <bool kDeep>
struct SubPath {
  V3 origin, direction;
  int depth;              // depth of the ray in flight (>= 1 after the first bounce)
  uint32_t pixel, key0, subPath;
  uint32_t termIndex;     // sample * numSub + subPath
  uint32_t primary;       // the camera hit's material | specular pick << 31
  LevelStack<kDeep> stack;
  const double2 *surface; // what the next bounce leaves from (generic: record or slot)
  bool needItem, finished;
  __device__ __forceinline__ void init(const double2 *slot) {
    origin = mk(0, 0, 0);
    direction = mk(0, 0, 1);
    depth = 0;
    pixel = key0 = subPath = termIndex = primary = 0;
    stack.reset();
    surface = slot;
    needItem = true;
    finished = false;
  }
};
struct TicketPool { // a warp's tickets (warp-uniform)
  uint32_t next, end;
};

// ---- 1. the next sub-path for lanes whose path has ended (whole warp) ----
// (Holding one ticket ahead per lane and asking its record into L2 when the ticket is taken was
// measured: 193.3 vs 195.0 Msamples/s, profiles/sweep_cornell_r2d.jsonl — the 16 strata of a
// record start in adjacent lanes, so one miss already serves sixteen starts.)
template <bool kDeep>
__device__ __forceinline__ void takeSubPath(const SplitArgs &args, SubPath<kDeep> &path, TicketPool &pool, unsigned lane,
                                            uint32_t totalItems, unsigned int *ticket) {
  const unsigned needMask = __ballot_sync(kFullMask, path.needItem);
  if (!needMask)
    return;
  const uint32_t numSub = args.numSub;
  const uint32_t want = static_cast<uint32_t>(__popc(needMask));
  const uint32_t rank = static_cast<uint32_t>(__popc(needMask & ((1u << lane) - 1u)));
  const uint32_t available = pool.end - pool.next;
  uint32_t item;
  if (available >= want) {
    item = pool.next + rank;
    pool.next += want;
  } else { // the rest of the old pool, then a fresh one
    uint32_t base = 0;
    if (lane == 0)
      base = atomicAdd(ticket, kTicketGrab);
    base = __shfl_sync(kFullMask, base, 0);
    item = rank < available ? pool.next + rank : base + (rank - available);
    pool.next = base + (want - available);
    pool.end = base + kTicketGrab;
  }
  if (path.needItem) {
    path.needItem = false;
    if (item >= totalItems) {
      path.finished = true;
    } else {
      uint32_t recordIndex;
      if (args.numSubShift >= 0) {
        recordIndex = item >> args.numSubShift;
        path.subPath = item & (numSub - 1u);
      } else {
        recordIndex = item / numSub;
        path.subPath = item - recordIndex * numSub;
      }
      path.surface = args.records + kRecordQuads * static_cast<size_t>(recordIndex);
      const double2 q8 = path.surface[8];
      path.pixel = highWord(q8.x);
      path.key0 = lowWord(q8.y);
      path.termIndex = highWord(q8.y) * numSub + path.subPath;
      path.depth = 0;
      path.stack.reset();
    }
  }
}

// ---- 2. bounce: one (u, v, p) triple, cone or hemisphere sample (Scene.cpp:157-175) ----
template <bool kDeep>
__device__ __forceinline__ void bounceSubPath(const SplitArgs &args, SubPath<kDeep> &path) {
  const DeviceScene &scene = args.scene;
  const double2 *surface = path.surface;
  double ru, rv, rp;
  KeyedDraws{path.key0}.bounce(path.pixel, path.subPath, static_cast<uint32_t>(path.depth), ru, rv, rp);
  double u = ru, v = rv; // (0 + r) / 1 exactly, below the first bounce
  if (path.depth == 0) {
    // u-major strata (Scene.cpp:155-156); x / n == x * (1/n) exactly when n is a power of two
    uint32_t stratumU, stratumV;
    if (args.firstBounceVShift >= 0) {
      stratumU = path.subPath >> args.firstBounceVShift;
      stratumV = path.subPath & (static_cast<uint32_t>(args.firstBounceV) - 1u);
    } else {
      stratumU = path.subPath / static_cast<uint32_t>(args.firstBounceV);
      stratumV = path.subPath - stratumU * static_cast<uint32_t>(args.firstBounceV);
    }
    const double su = static_cast<double>(stratumU) + ru;
    const double sv = static_cast<double>(stratumV) + rv;
    u = args.firstBounceUPow2 ? su * args.invFirstBounceU : ieeeDiv(su, static_cast<double>(args.firstBounceU));
    v = args.firstBounceVPow2 ? sv * args.invFirstBounceV : ieeeDiv(sv, static_cast<double>(args.firstBounceV));
  }
  const double2 q7 = surface[7], q8 = surface[8];
  const uint32_t surfaceWord = lowWord(q8.x);
  const uint32_t material = surfaceWord & ~kLazyBasisFlag;
  const bool specular = rp < q7.y; // p < reflectivity
  const double2 q1 = surface[1], q2 = surface[2];
  V3 frameZ = mk(q1.y, q2.x, q2.y); // the surface normal
  V3 frameX, frameY;
  double angle = (2 * kPi) * u, radius = 0, zScale = 0;
  bool direct = false;
  V3 newDirection = mk(0, 0, 0);
  // coneSample() / hemisphereSample() (Samples.cpp:6-30) end in the same
  // normalised(basis.transform(cos(t)*r, sin(t)*r, z)): the few specular lanes only prepare
  // its inputs, then every lane runs that tail together.
  if (specular) {
    const double2 q3 = surface[3], q4 = surface[4];
    newDirection = reflect(frameZ, mk(q3.x, q3.y, q4.x));
    direct = coneSetup(newDirection, materialOf(scene, material).coneAngle(), u, v, frameX, frameY, frameZ, angle,
                       radius, zScale);
  } else {
    if (surfaceWord & kLazyBasisFlag) {
      const Basis basis = basisFromZ(frameZ);
      frameX = basis.x;
      frameY = basis.y;
    } else {
      const double2 q4 = surface[4], q5 = surface[5], q6 = surface[6];
      frameX = mk(q4.y, q5.x, q5.y);
      frameY = mk(q6.x, q6.y, q7.x);
    }
    const double2 roots = ieeeSqrtPair(v, 1 - v);
    radius = roots.x;
    zScale = roots.y;
  }
  if (!direct) {
    double sinT, cosT;
    sinCos(angle, sinT, cosT);
    newDirection = normalised(transform(Basis{frameX, frameY, frameZ}, mk(cosT * radius, sinT * radius, zScale)));
  }
  if (path.depth == 0)
    path.primary = material | (specular ? 0x80000000u : 0u);
  else
    path.stack.set(path.depth, material, specular);
  const double2 q0 = surface[0];
  path.origin = mk(q0.x, q0.y, q1.x);
  path.direction = newDirection;
  ++path.depth;
}

// ---- 4. the hit becomes the next bounce's surface (written to `slot`), or the path ends ----
template <bool kDeep>
__device__ __forceinline__ void afterCast(const SplitArgs &args, const double4 *spheres, SubPath<kDeep> &path,
                                          double2 *slot, const Nearest &best, bool tracing) {
  const DeviceScene &scene = args.scene;
  const V3 origin = path.origin, direction = path.direction;
  bool ended = false;
  V3 incoming = mk(0, 0, 0);
  if (tracing) {
    if (best.prim == kNoPrim) {
      incoming = mk(scene.environment[0], scene.environment[1], scene.environment[2]); // Scene.cpp:132-133
      ended = true;
    } else if (path.depth + 1 >= args.maxDepth) {
      // Deepest level: its sampling loop runs, but every child returns Vec3() (Scene.cpp:128-129).
      incoming = shadeTerm(materialOf(scene, hitMaterial(scene, best)), true, mk(0, 0, 0));
      ended = true;
    } else { // Scene.cpp:135-152 into the slot
      const V3 position = positionAlong(origin, direction, best.t);
      V3 normal;
      uint32_t word;
      bool inside;
      double2 q4 = make_double2(direction.z, 0.0), q5 = make_double2(0.0, 0.0), q6 = q5;
      double basisYz = 0.0;
      if (best.prim < 0) { // sphere epilogue (Scene.cpp:38-48); its basis is left to the bounce
        const int i = -best.prim - 1;
        const double4 s = spheres[i];
        const V3 outward = normalised(sub(position, mk(s.x, s.y, s.z)));
        inside = dot(outward, direction) > 0;
        normal = inside ? neg(outward) : outward;
        word = __ldg(scene.sphereMaterial + i) | kLazyBasisFlag;
      } else { // triangle epilogue (Scene.cpp:99-112), normal and bases precomputed at upload
        const double4 *record = scene.triShade + 4 * static_cast<size_t>(best.prim);
        const double4 r0 = ldgDouble4(record);
        const double4 r2 = ldgDouble4(record + 2);
        inside = best.det < kEpsilon; // backfacing
        word = static_cast<uint32_t>(r0.w);
        if (inside) {
          const double4 r3 = ldgDouble4(record + 3);
          normal = mk(-r0.x, -r0.y, -r0.z);
          q4.y = r2.z;
          q5 = make_double2(r2.w, r3.x);
          q6 = make_double2(r3.y, r3.z);
          basisYz = r3.w;
        } else {
          const double4 r1 = ldgDouble4(record + 1);
          normal = mk(r0.x, r0.y, r0.z);
          q4.y = r1.x;
          q5 = make_double2(r1.y, r1.z);
          q6 = make_double2(r1.w, r2.x);
          basisYz = r2.y;
        }
      }
      HitInfo hit;
      hit.position = position;
      hit.normal = normal;
      hit.inside = inside;
      const double reflectivity = hitReflectivity(materialOf(scene, word & ~kLazyBasisFlag), hit, direction);
      slot[0] = make_double2(position.x, position.y);
      slot[1] = make_double2(position.z, normal.x);
      slot[2] = make_double2(normal.y, normal.z);
      slot[3] = make_double2(direction.x, direction.y);
      slot[4] = q4;
      slot[5] = q5;
      slot[6] = q6;
      slot[7] = make_double2(basisYz, reflectivity);
      slot[8] = make_double2(packWords(word, 0u), 0.0);
      path.surface = slot;
    }
  }
  __syncwarp();
  if (ended) {
    // unwind levels depth-1 .. 1 (Scene.cpp:168,172-174 with a 1x1 stratum), then the camera
    // hit's own term for this stratum
#pragma unroll 1
    for (int level = path.depth - 1; level >= 1; --level) {
      uint32_t material;
      bool specular;
      path.stack.get(level, material, specular);
      incoming = shadeTerm(materialOf(scene, material), specular, incoming);
    }
    const V3 term = shadeTerm(materialOf(scene, path.primary & 0x7fffffffu), (path.primary >> 31) != 0, incoming);
    double *out = args.terms + 3 * static_cast<size_t>(path.termIndex);
    out[0] = term.x;
    out[1] = term.y;
    out[2] = term.z;
    path.needItem = true;
  }
  __syncwarp();
}

template <int kBlock, int kMinBlocks, int kSweep, bool kDeep>
__global__ void __launch_bounds__(kBlock, kMinBlocks) subPathKernel(const __grid_constant__ SplitArgs args) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceScene &scene = args.scene;
  TileStream stream = makeTileStream(smemRaw, scene, kSweep);
  stream.start();
#pragma unroll 1
  for (uint32_t i = threadIdx.x; i < scene.numSpheres; i += kBlock)
    stream.spheres()[i] = scene.spheres[i];
  __syncthreads();
  const bool resident = scene.numTiles <= 1;
  if (resident && scene.numTiles == 1)
    stream.acquire(); // the one tile stays in buffer 0 for the whole launch
  double2 *const slot = reinterpret_cast<double2 *>(smemRaw + smemAfterTiles(scene.numSpheres, scene.tileTris,
                                                                             scene.numTiles, kSweep)) +
                        kRecordQuads * threadIdx.x;

  const unsigned lane = threadIdx.x & 31u;
  const uint32_t totalItems = static_cast<uint32_t>(args.counters[2]) * args.numSub; // records x strata, < 2^31
  unsigned int *const ticket = reinterpret_cast<unsigned int *>(args.counters);
  SubPath<kDeep> path;
  path.init(slot);
  TicketPool pool{0, 0};
  uint32_t casts = 0; // per lane and launch: far below 2^32

  for (;;) {
    takeSubPath(args, path, pool, lane, totalItems, ticket);
    __syncwarp();
    if (resident) {
      if (__all_sync(kFullMask, path.finished))
        break;
    } else {
      if (__syncthreads_and(path.finished))
        break;
    }
    const bool tracing = !path.finished;
    if (tracing) {
      bounceSubPath(args, path);
      ++casts;
    }
    __syncwarp();
    const Nearest best = castRay<kSweep>(scene, stream, resident, tracing, path.origin, path.direction, args.momentTable);
    afterCast(args, stream.spheres(), path, slot, best, tracing);
  }

  if (!resident)
    stream.drain();
  unsigned long long warpCasts = casts;
#pragma unroll 1
  for (int offset = 16; offset > 0; offset >>= 1)
    warpCasts += __shfl_down_sync(kFullMask, warpCasts, offset);
  if (lane == 0 && warpCasts)
    atomicAdd(args.counters + 1, warpCasts);
}

// ---- two sub-paths per lane -------------------------------------------------------------------
// Scenes whose stage-0 tile sits in shared memory are bound by the shared-memory data pipe, not by
// issue slots: a broadcast LDS.128 costs two wavefronts, stage 0 needs 19 of them per four
// triangles, i.e. 38 pipe cycles against ~20 issue cycles per SM (suzanne: 9 200 wavefronts per warp
// iteration).  Here every lane carries TWO sub-paths and sweeps both rays against each group of
// triangles it loads, which halves the wavefronts per ray; one CTA of kBlock threads per SM so that
// the tile is staged once.
template <int kBlock, bool kDeep>
__global__ void __launch_bounds__(kBlock, 1) subPathDualKernel(const __grid_constant__ SplitArgs args) {
  constexpr int kSweep = 7;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceScene &scene = args.scene;
  TileStream stream = makeTileStream(smemRaw, scene, kSweep);
  stream.start();
#pragma unroll 1
  for (uint32_t i = threadIdx.x; i < scene.numSpheres; i += kBlock)
    stream.spheres()[i] = scene.spheres[i];
  __syncthreads();
  const bool resident = scene.numTiles <= 1;
  if (resident && scene.numTiles == 1)
    stream.acquire();
  double2 *const slot0 = reinterpret_cast<double2 *>(smemRaw + smemAfterTiles(scene.numSpheres, scene.tileTris,
                                                                              scene.numTiles, kSweep)) +
                         2 * kRecordQuads * threadIdx.x;
  double2 *const slot1 = slot0 + kRecordQuads;

  const unsigned lane = threadIdx.x & 31u;
  const uint32_t totalItems = static_cast<uint32_t>(args.counters[2]) * args.numSub;
  unsigned int *const ticket = reinterpret_cast<unsigned int *>(args.counters);
  SubPath<kDeep> path0, path1;
  path0.init(slot0);
  path1.init(slot1);
  TicketPool pool{0, 0};
  uint32_t casts = 0;

  for (;;) {
    takeSubPath(args, path0, pool, lane, totalItems, ticket);
    takeSubPath(args, path1, pool, lane, totalItems, ticket);
    __syncwarp();
    const bool done = path0.finished && path1.finished;
    if (resident) {
      if (__all_sync(kFullMask, done))
        break;
    } else {
      if (__syncthreads_and(done))
        break;
    }
    const bool tracing0 = !path0.finished, tracing1 = !path1.finished;
    if (tracing0) {
      bounceSubPath(args, path0);
      ++casts;
    }
    __syncwarp();
    if (tracing1) {
      bounceSubPath(args, path1);
      ++casts;
    }
    __syncwarp();

    // Scene::intersect for both rays: spheres first, then every tile (Scene.cpp:115-122)
    Nearest best0{__longlong_as_double(0x7ff0000000000000ll), 0.0, kNoPrim}, best1 = best0;
    if (tracing0)
      sweepSpheres(stream.spheres(), static_cast<int>(scene.numSpheres), path0.origin, path0.direction, best0);
    if (tracing1)
      sweepSpheres(stream.spheres(), static_cast<int>(scene.numSpheres), path1.origin, path1.direction, best1);
    for (uint32_t j = 0; j < scene.numTiles; ++j) {
      const unsigned char *tile = resident ? stream.tile(0) : stream.acquire();
      const int first = static_cast<int>(j * scene.tileTris);
      sweepTileStage0Moment2(reinterpret_cast<const float *>(tile), scene.triExact + static_cast<size_t>(first) * 10,
                             static_cast<int>(scene.tileTris), first, path0.origin, path0.direction, tracing0, best0,
                             path1.origin, path1.direction, tracing1, best1, scene.fanMask);
      if (!resident)
        stream.release();
    }
    afterCast(args, stream.spheres(), path0, slot0, best0, tracing0);
    afterCast(args, stream.spheres(), path1, slot1, best1, tracing1);
  }

  if (!resident)
    stream.drain();
  unsigned long long warpCasts = casts;
#pragma unroll 1
  for (int offset = 16; offset > 0; offset >>= 1)
    warpCasts += __shfl_down_sync(kFullMask, warpCasts, offset);
  if (lane == 0 && warpCasts)
    atomicAdd(args.counters + 1, warpCasts);
}

// =============================================================================================
// 3. Strata -> samples -> pixels, both sums in the reference's order.
// =============================================================================================
// kStrata > 0: the strata count is a compile-time even number (the reference's default 4x4 = 16):
// the 24-byte terms of a sample are read as 16-byte pairs, the loop is unrolled.
template <int kStrata>
__global__ void resolveSamplesKernel(const __grid_constant__ SplitArgs args) {
  const uint32_t own = blockIdx.x * blockDim.x + threadIdx.x;
  if (own >= args.ownPixels)
    return;
  const uint32_t px = own % args.width;
  const uint32_t py = static_cast<uint32_t>(args.rowBegin) + (own / args.width) * static_cast<uint32_t>(args.rowStep);
  PtPixelDevice *dst = args.accumulator + (px + static_cast<size_t>(py) * args.width);
  const uint32_t numSub = kStrata > 0 ? static_cast<uint32_t>(kStrata) : args.numSub;
  const double invNumSub = 1.0 / static_cast<double>(numSub);
  double r = dst->sum[0], g = dst->sum[1], b = dst->sum[2];
#pragma unroll 1
  for (uint32_t p = 0; p < args.numPasses; ++p) {
    const size_t sample = static_cast<size_t>(p) * args.ownPixels + own;
    const double *t = args.terms + 3 * sample * numSub;
    V3 colour;
    if (args.sampleKind[sample]) {
      colour = mk(t[0], t[1], t[2]);
    } else {
      V3 acc = mk(0, 0, 0);
      if (kStrata > 0) {
        const double2 *pairs = reinterpret_cast<const double2 *>(t); // 3 * kStrata doubles, 16-byte aligned
#pragma unroll
        for (int k = 0; k < kStrata; k += 2) { // two terms = three pairs: (x0 y0) (z0 x1) (y1 z1)
          const double2 a = pairs[3 * (k / 2)], c = pairs[3 * (k / 2) + 1], e = pairs[3 * (k / 2) + 2];
          acc = add(acc, mk(a.x, a.y, c.x));
          acc = add(acc, mk(c.y, e.x, e.y));
        }
      } else {
        for (uint32_t k = 0; k < numSub; ++k)
          acc = add(acc, mk(t[3 * k], t[3 * k + 1], t[3 * k + 2]));
      }
      colour = scale(acc, invNumSub); // Scene.cpp:178
    }
    r += colour.x;
    g += colour.y;
    b += colour.z;
  }
  dst->sum[0] = r;
  dst->sum[1] = g;
  dst->sum[2] = b;
  dst->numSamples += args.numPasses;
}

// =============================================================================================
// Host-side launchers.
// =============================================================================================
bool constTableFits(uint32_t numTriangles, uint32_t numTiles) {
  return numTiles == 1 && numTriangles <= 4u * kConstGroups;
}

size_t splitBytesPerSample(uint32_t numSub) {
  return kRecordQuads * sizeof(double2) + 3 * sizeof(double) * static_cast<size_t>(numSub) + 1;
}

template <int kBlock, int kSweep>
static cudaError_t launchPrimary(const SplitArgs &args, int numSms, cudaStream_t stream) {
  auto kernel = primaryHitsKernel<kBlock, kSweep>;
  const size_t smemBytes = smemAfterTiles(args.scene.numSpheres, args.scene.tileTris, args.scene.numTiles, kSweep);
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  int perSm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBlock, smemBytes);
  if (err != cudaSuccess)
    return err;
  if (perSm < 1)
    return cudaErrorInvalidConfiguration;
  const unsigned long long chunks = (static_cast<unsigned long long>(args.totalSamples) + kBlock - 1) / kBlock;
  unsigned long long grid = static_cast<unsigned long long>(numSms) * perSm;
  if (chunks < grid)
    grid = chunks ? chunks : 1;
  kernel<<<static_cast<unsigned>(grid), kBlock, smemBytes, stream>>>(args);
  return cudaGetLastError();
}

template <int kBlock, int kMinBlocks, int kSweep, bool kDeep>
static cudaError_t launchSubPaths(const SplitArgs &args, int numSms, cudaStream_t stream) {
  auto kernel = subPathKernel<kBlock, kMinBlocks, kSweep, kDeep>;
  const size_t smemBytes = smemAfterTiles(args.scene.numSpheres, args.scene.tileTris, args.scene.numTiles, kSweep) +
                           static_cast<size_t>(kBlock) * kRecordQuads * sizeof(double2); // one Surface slot per thread
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  int perSm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBlock, smemBytes);
  if (err != cudaSuccess)
    return err;
  if (perSm < 1)
    return cudaErrorInvalidConfiguration;
  // Persistent grid: every SM holds `perSm` CTAs for the whole launch (148 x perSm on B200); the
  // number of sub-paths is only known on the device (records x strata).
  const unsigned long long wanted = (static_cast<unsigned long long>(args.totalSamples) * args.numSub + kBlock - 1) / kBlock;
  unsigned long long grid = static_cast<unsigned long long>(numSms) * perSm;
  if (wanted < grid)
    grid = wanted ? wanted : 1;
  kernel<<<static_cast<unsigned>(grid), kBlock, smemBytes, stream>>>(args);
  return cudaGetLastError();
}

template <int kBlock, bool kDeep>
static cudaError_t launchSubPathsDual(const SplitArgs &args, int numSms, cudaStream_t stream) {
  auto kernel = subPathDualKernel<kBlock, kDeep>;
  const size_t smemBytes = smemAfterTiles(args.scene.numSpheres, args.scene.tileTris, args.scene.numTiles, 7) +
                           static_cast<size_t>(kBlock) * 2 * kRecordQuads * sizeof(double2); // two Surface slots per thread
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  int perSm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBlock, smemBytes);
  if (err != cudaSuccess)
    return err;
  if (perSm < 1)
    return cudaErrorInvalidConfiguration;
  const unsigned long long wanted = (static_cast<unsigned long long>(args.totalSamples) * args.numSub + 2 * kBlock - 1) / (2 * kBlock);
  unsigned long long grid = static_cast<unsigned long long>(numSms) * perSm;
  if (wanted < grid)
    grid = wanted ? wanted : 1;
  kernel<<<static_cast<unsigned>(grid), kBlock, smemBytes, stream>>>(args);
  return cudaGetLastError();
}

// The pipeline with two sub-paths per lane in the middle kernel (configurations 200 + 10 * shape + 7).
template <int kBlock>
static cudaError_t launchSplitDual(const SplitArgs &args, int numSms, cudaStream_t stream) {
  cudaError_t err = launchPrimary<256, 7>(args, numSms, stream);
  if (err != cudaSuccess)
    return err;
  const bool deep = args.maxDepth - 2 > LevelStack<false>::kLevels || args.numMaterials > 0x8000u;
  return deep ? launchSubPathsDual<kBlock, true>(args, numSms, stream) : launchSubPathsDual<kBlock, false>(args, numSms, stream);
}

template <int kBlock, int kMinBlocks, int kSweep>
static cudaError_t launchSplitShape(const SplitArgs &args, int numSms, cudaStream_t stream) {
  cudaError_t err = launchPrimary<256, kSweep >= 8 ? 7 : kSweep>(args, numSms, stream);
  if (err != cudaSuccess)
    return err;
  const bool deep = args.maxDepth - 2 > LevelStack<false>::kLevels || args.numMaterials > 0x8000u;
  return deep ? launchSubPaths<kBlock, kMinBlocks, kSweep, true>(args, numSms, stream)
              : launchSubPaths<kBlock, kMinBlocks, kSweep, false>(args, numSms, stream);
}

// The third kernel, on its own: blocks of 64 threads (2 304 registers) fit next to the three resident
// CTAs of the NEXT batch's sub-path kernel, so that resolving one batch overlaps tracing the next.
cudaError_t launchSplitResolve(const SplitArgs &args, cudaStream_t stream) {
  const int block = 64;
  const unsigned grid = (args.ownPixels + block - 1) / block;
  if (args.numSub == 16)
    resolveSamplesKernel<16><<<grid, block, 0, stream>>>(args);
  else
    resolveSamplesKernel<0><<<grid, block, 0, stream>>>(args);
  return cudaGetLastError();
}

// A pipeline configuration is 100 + 10 * launchShape + sweepVariant (the megakernel's numbering
// plus 100; 200 + ... = two sub-paths per lane): sweep variants 1 (two-stage FP64), 6 (sign-bit FP32
// stage 0 + exact), 7 (the same in moment form), 8 (7 out of the constant bank, scenes of up to 64
// triangles) and 9 (8 with the group loop unrolled); launch
// shapes of the sub-path kernel 0 = 256 threads x 2 CTAs/SM, 2 = 256 x 3, 3 = 192 x 4, 4 = 128 x 5,
// 6 = 256 x 4.
// Camera hits + sub-paths of one batch (two launches); launchSplitResolve() finishes it.
cudaError_t launchSplitTrace(const SplitArgs &args, int numSms, int config, cudaStream_t stream) {
  switch (config) {
  case 207: return launchSplitDual<384>(args, numSms, stream);
  case 217: return launchSplitDual<512>(args, numSms, stream);
  case 227: return launchSplitDual<256>(args, numSms, stream);
  case 101: return launchSplitShape<256, 2, 1>(args, numSms, stream);
  case 106: return launchSplitShape<256, 2, 6>(args, numSms, stream);
  case 121: return launchSplitShape<256, 3, 1>(args, numSms, stream);
  case 126: return launchSplitShape<256, 3, 6>(args, numSms, stream);
  case 128: return launchSplitShape<256, 3, 8>(args, numSms, stream);
  case 168: return launchSplitShape<256, 4, 8>(args, numSms, stream);
  case 148: return launchSplitShape<128, 5, 8>(args, numSms, stream);
  case 129: return launchSplitShape<256, 3, 9>(args, numSms, stream); // kept as the measured counter-example
  case 107: return launchSplitShape<256, 2, 7>(args, numSms, stream);
  case 127: return launchSplitShape<256, 3, 7>(args, numSms, stream);
  case 137: return launchSplitShape<192, 4, 7>(args, numSms, stream);
  case 147: return launchSplitShape<128, 5, 7>(args, numSms, stream);
  default: return cudaErrorInvalidValue;
  }
}

} // namespace ptb200
