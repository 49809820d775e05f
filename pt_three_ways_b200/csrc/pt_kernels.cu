// sm_100a kernels of the B200 path tracer.
//
//   renderKeyedKernel       persistent megakernel, one path per LANE, keyed Philox RNG
//                           (throughput mode; replaces the pass/pixel loops of
//                           dod::Scene::render + radiance + intersect, Scene.cpp:115-254);
//                           its kWay = 1 instantiation renders the reference's `fp` way instead:
//                           one mt19937 per (pass, pixel), src/fp/Render.cpp:76-135
//   renderSequentialKernel  one pass per WARP or per group of 16 / 8 / 4 lanes, the reference's own
//                           mt19937 stream walked in row-major order (Scene.cpp:208-220), primitives
//                           spread over the group's lanes
//   reducePassesKernel      per-pixel accumulation of per-pass samples IN PASS ORDER
//                           (SampledPixel::accumulate, SampledPixel.cpp:3-6)
//   buildFilterKernel       FP32 copies + per-triangle error bounds for the stage-0 sweep
//   intersectKernel         Scene::intersect/intersectSpheres/intersectTriangles for tests,
//                           through every sweep variant
//   auditStage0Kernel       test hook: stage 0 vs the exact test on every (ray, triangle) pair
//   fp64PeakKernel          DFMA throughput probe (roofline denominator)
#include "pt_kernels.h"

#include "pt_device.cuh"
#include "pt_stage.cuh"
#include "pt_mt19937.cuh"

#include <cstdlib>

namespace ptb200 {


// The megakernel additionally parks each thread's primary-hit Surface (16 doubles) in shared
// memory between the strata of a sample, [component][thread] so consecutive threads hit
// consecutive banks; that keeps ~32 registers per thread free for a third resident CTA.
constexpr uint32_t kPrimaryDoubles = 16;
constexpr uint32_t kPendingDoubles = 8;   // the prefetched next sample, see the megakernel
constexpr uint32_t kPendingDoublesFp = 9; //   ... plus its engine's two running seed words
constexpr int kPrefetchBatch = 12;        // refill the prefetch slots when this many lanes' are empty
__host__ size_t keyedSmemBytes(uint32_t numSpheres, uint32_t tileTris, uint32_t numTiles, int sweep,
                               uint32_t threadsForPrimarySlots, int way) {
  return smemAfterTiles(numSpheres, tileTris, numTiles, sweep) +
         static_cast<size_t>(threadsForPrimarySlots) *
             (kPrimaryDoubles + (way == 1 ? kPendingDoublesFp : kPendingDoubles)) * sizeof(double);
}


// =============================================================================================
// Keyed (Philox) megakernel: one path per lane, persistent CTAs, work pulled from a ticket.
// =============================================================================================

enum LaneMode : int { kNeedWork = 0, kTracing = 1, kFinished = 2 };


// ---- the `fp` way: one std::mt19937 per (pass, pixel) ---------------------------------------
// Engine seed of renderOnePixel (src/fp/Render.cpp:125-126): height*width*seed + x*width + y
// (x*width, the reference's own indexing), evaluated in size_t and reduced mod 2^32 by
// mersenne_twister_engine::seed — i.e. plain uint32 arithmetic.
__device__ __forceinline__ uint32_t fpEngineSeed(uint32_t width, uint32_t height, int passSeed, int px, int py) {
  return height * width * static_cast<uint32_t>(passSeed) + static_cast<uint32_t>(px) * width +
         static_cast<uint32_t>(py);
}
// Seeds the engine of a prefetched sample and draws its camera ray (Camera::randomRay draws two
// doubles, rayFromUnit two more only when the aperture is open, Camera.h:30-33,56-58).  Always
// executed by a batch of lanes together (see the megakernel), so out of line.
static __device__ __noinline__ void fpCameraRay(const DeviceCamera &camera, uint32_t engineSeed, int px, int py,
                                         uint32_t *history, uint32_t &seedWordA, uint32_t &seedWordB,
                                         V3 &origin, V3 &direction) {
  LaneMt19937 rng;
  rng.seed(engineSeed);
  double u[4] = {0, 0, 0, 0};
  const int draws = camera.apertureRadius == 0 ? 2 : 4;
#pragma unroll 1
  for (int i = 0; i < draws; ++i) {
    const uint32_t lo = rng.word<true>(history, 0u);
    const uint32_t hi = rng.word<true>(history, 0u);
    u[i] = canonicalFromWords(lo, hi);
  }
  cameraRay(camera, px, py, u[0], u[1], u[2], u[3], origin, direction);
  seedWordA = rng.a;
  seedWordB = rng.b;
}


template <int kBlock, int kMinBlocks, int kSweep, int kWay>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
    renderKeyedKernel(const __grid_constant__ KeyedArgs args) {
  constexpr bool kFp = kWay == 1;
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

  const unsigned lane = threadIdx.x & 31u;
  const int numSub = args.firstBounceU * args.firstBounceV;
  const double invNumSub = 1.0 / static_cast<double>(numSub); // Vec3::operator/ (Vec3.h:51-54)
  const V3 environment = mk(scene.environment[0], scene.environment[1], scene.environment[2]);

  // ---- per-lane path state ----
  int mode = kNeedWork;
  uint32_t casts = 0;            // per lane and launch: far below 2^32
  uint64_t sampleSlot = 0;       // where this sample's colour goes
  uint32_t pixel = 0;            // x + y*width (RNG key and framebuffer index)
  uint32_t key0 = 0;
  V3 origin = mk(0, 0, 0), direction = mk(0, 0, 1);
  int depth = 0;                 // depth of the ray in flight
  // The camera ray's hit stays alive across its numSub sub-paths: in shared memory.
  double *const primarySlot = reinterpret_cast<double *>(smemRaw + smemAfterTiles(scene.numSpheres, scene.tileTris,
                                                                                  scene.numTiles, kSweep)) + threadIdx.x;
  // ... and so does this lane's prefetched next sample (6 doubles of ray + slot + key/pixel).
  double *const pendingSlot = primarySlot + kPrimaryDoubles * kBlock;
  bool pendingValid = false, exhausted = false;
  uint32_t primaryMaterial = 0;
  bool primarySpecular = false;
  int subPath = 0;
  V3 acc = mk(0, 0, 0);
  // levels 1.. of the current sub-path: material index and branch taken
  uint16_t stackMaterial[kMaxDepth];
  bool stackSpecular[kMaxDepth];
  // fp way: this lane's engine, and the words it has generated (its slice of the scratch buffer).
  LaneMt19937 rng{0u, 0u, 0u};
  // (32-bit offset: the compiler re-derives this address at every use rather than hold it)
  uint32_t *const history = kFp ? args.mtHistory + (blockIdx.x * kBlock + threadIdx.x) * kMtHistoryStride : nullptr;
  const uint32_t cameraWords = args.camera.apertureRadius == 0 ? 4u : 8u;
  const uint32_t storeLimit = args.mtStoreLimit;

  for (;;) {
    // ---- 1. tickets and camera rays, prefetched in batches ----
    // A camera ray costs ~400 instructions whether 1 or 32 lanes need one.  So every lane keeps
    // ONE prefetched sample (ticket + camera ray) parked in shared memory; slots are refilled
    // together once kPrefetchBatch of them are empty (or a lane is out of work right now).
    const bool slotEmpty = !pendingValid && !exhausted;
    const unsigned emptyMask = __ballot_sync(kFullMask, slotEmpty);
    const unsigned starvingMask = __ballot_sync(kFullMask, slotEmpty && mode == kNeedWork);
    if (emptyMask && (starvingMask || __popc(emptyMask) >= kPrefetchBatch)) {
      unsigned long long base = 0;
      const int leader = __ffs(emptyMask) - 1;
      if (static_cast<int>(lane) == leader)
        base = atomicAdd(args.ticket, static_cast<unsigned long long>(__popc(emptyMask)));
      base = __shfl_sync(kFullMask, base, leader);
      if (slotEmpty) {
        const unsigned long long item = base + __popc(emptyMask & ((1u << lane) - 1u));
        if (item < args.totalItems) {
          const uint32_t passInBatch = static_cast<uint32_t>(item / args.ownPixels);
          const uint32_t own = static_cast<uint32_t>(item % args.ownPixels);
          const int px = static_cast<int>(own % args.width);
          const int py = args.rowBegin + static_cast<int>(own / args.width) * args.rowStep;
          const uint32_t nextPixel = static_cast<uint32_t>(px) + static_cast<uint32_t>(py) * args.width;
          const uint32_t nextKey = static_cast<uint32_t>(args.seed + args.passBegin + static_cast<int>(passInBatch));
          if (args.maxDepth <= 0) { // radiance() returns Vec3() before intersecting (Scene.cpp:128-129)
            args.samples[3 * item + 0] = 0.0;
            args.samples[3 * item + 1] = 0.0;
            args.samples[3 * item + 2] = 0.0;
          } else {
            V3 o, d;
            if (kFp) {
              uint32_t seedWordA, seedWordB;
              fpCameraRay(args.camera, fpEngineSeed(args.width, args.height, static_cast<int>(nextKey), px, py),
                          px, py, history, seedWordA, seedWordB, o, d);
              pendingSlot[8 * kBlock] = __hiloint2double(static_cast<int>(seedWordA), static_cast<int>(seedWordB));
            } else {
              keyedCameraRay(args.camera, nextKey, nextPixel, px, py, o, d);
            }
            pendingSlot[0 * kBlock] = o.x;
            pendingSlot[1 * kBlock] = o.y;
            pendingSlot[2 * kBlock] = o.z;
            pendingSlot[3 * kBlock] = d.x;
            pendingSlot[4 * kBlock] = d.y;
            pendingSlot[5 * kBlock] = d.z;
            pendingSlot[6 * kBlock] = __longlong_as_double(static_cast<long long>(item));
            pendingSlot[7 * kBlock] = __hiloint2double(static_cast<int>(nextKey), static_cast<int>(nextPixel));
            pendingValid = true;
          }
        } else {
          exhausted = true;
        }
      }
    }
    __syncwarp();
    if (mode == kNeedWork) {
      if (pendingValid) { // start the prefetched sample
        origin = mk(pendingSlot[0 * kBlock], pendingSlot[1 * kBlock], pendingSlot[2 * kBlock]);
        direction = mk(pendingSlot[3 * kBlock], pendingSlot[4 * kBlock], pendingSlot[5 * kBlock]);
        sampleSlot = static_cast<uint64_t>(__double_as_longlong(pendingSlot[6 * kBlock]));
        const double packed = pendingSlot[7 * kBlock];
        key0 = static_cast<uint32_t>(__double2hiint(packed));
        pixel = static_cast<uint32_t>(__double2loint(packed));
        if (kFp) { // take over the prefetched engine: its camera words move to the front
          const double words = pendingSlot[8 * kBlock];
          rng.a = static_cast<uint32_t>(__double2hiint(words));
          rng.b = static_cast<uint32_t>(__double2loint(words));
          rng.k = cameraWords;
#pragma unroll
          for (uint32_t i = 0; i < kMtPrefetchWords; ++i)
            history[i] = history[kMtWords + i];
        }
        pendingValid = false;
        depth = 0;
        mode = kTracing;
      } else if (exhausted) {
        mode = kFinished;
      }
    }
    const bool tracing = mode == kTracing;
    if (resident) {
      if (__all_sync(kFullMask, mode == kFinished))
        break;
    } else {
      if (__syncthreads_and(mode == kFinished))
        break;
    }

    // ---- 2. cast: spheres first, then every triangle (Scene.cpp:115-122); ONE sweep site ----
    Nearest best{__longlong_as_double(0x7ff0000000000000ll), 0.0, kNoPrim};
    if (tracing) {
      ++casts;
      sweepSpheres(stream.spheres(), static_cast<int>(scene.numSpheres), origin, direction, best);
    }
    for (uint32_t j = 0; j < scene.numTiles; ++j) {
      const unsigned char *tile = resident ? stream.tile(0) : stream.acquire();
      if (tracing)
        sweepStagedTile<kSweep, kFp>(scene, tile, j, origin, direction, best);
      if (!resident)
        stream.release();
    }

    // ---- 3. what the cast means for this lane's path ----
    // The phases below are separated by __syncwarp() so that lanes arriving from different
    // branches execute each (expensive) phase together instead of one convergence group at a
    // time: profiling showed the bounce code running at ~8 of 32 lanes without them.
    bool ended = false;          // the ray in flight has its radiance (`incoming`)
    bool bounce = false;         // launch a bounce from `surface` (or from `primary` at depth 0)
    bool terminalPrimary = false;
    bool needSurface = false;
    bool skippedTriple = false;
    V3 incoming = mk(0, 0, 0);
    HitInfo hit{};
    if (tracing) {
      if (best.prim == kNoPrim) {
        incoming = environment; // Scene.cpp:132-133
        ended = true;
      } else {
        hit = finishHit(scene, stream.spheres(), origin, direction, best);
        if (depth == 0 && args.preview) { // Scene.cpp:137-138
          incoming = materialOf(scene, hit.material).diffuse();
          ended = true;
        } else if (depth + 1 >= args.maxDepth) {
          // Deepest level: its bounce loop still runs, but every child returns Vec3()
          // (Scene.cpp:128-129), so the level contributes its emission only.
          incoming = shadeTerm(materialOf(scene, hit.material), true, mk(0, 0, 0));
          ended = true;
          terminalPrimary = depth == 0;
          // fp way: ... and its (u, v, p) draws still advance the sample's engine.  They are
          // taken at the bounce site below, right before the next stratum's own draws, where the
          // lanes of the warp are together (this branch runs at ~5 of 32 lanes).
          skippedTriple = kFp && depth > 0;
        } else {
          needSurface = true;
        }
      }
    }
    __syncwarp();

    Surface surface{};
    if (needSurface) { // Scene.cpp:135-152: reflectivity and the local basis at the hit
      surface.position = hit.position;
      surface.normal = hit.normal;
      surface.incoming = direction;
      surface.material = hit.material;
      surface.reflectivity = hitReflectivity(materialOf(scene, hit.material), hit, direction);
      hitBasis(scene, hit, surface.basisX, surface.basisY);
      if (depth == 0) {
        const double values[kPrimaryDoubles] = {
            surface.position.x, surface.position.y, surface.position.z, surface.normal.x, surface.normal.y,
            surface.normal.z, surface.incoming.x, surface.incoming.y, surface.incoming.z, surface.basisX.x,
            surface.basisX.y, surface.basisX.z, surface.basisY.x, surface.basisY.y, surface.basisY.z,
            surface.reflectivity};
#pragma unroll
        for (uint32_t c = 0; c < kPrimaryDoubles; ++c)
          primarySlot[c * kBlock] = values[c];
        primaryMaterial = surface.material;
        acc = mk(0, 0, 0);
        subPath = 0;
      }
      bounce = true;
    }
    __syncwarp();

    if (ended) {
      bool sampleDone = false;
      V3 colour = incoming;
      if (depth == 0) {
        if (terminalPrimary && !kFp) { // maxDepth == 1: numSub children, each Vec3()
          acc = mk(0, 0, 0);
#pragma unroll 1
          for (int k = 0; k < numSub; ++k)
            acc = add(acc, incoming);
          colour = scale(acc, invNumSub);
        } // fp way: emission + (sum of zero terms) / numSub, which `incoming` already is
        sampleDone = true; // camera ray missed / preview / maxDepth == 1
      } else {
        // unwind levels depth-1 .. 1 (Scene.cpp:168,172-174 with a 1x1 stratum), then the
        // primary hit's own term, in the reference's summation order
#pragma unroll 1
        for (int level = depth - 1; level >= 1; --level) {
          const MaterialView mat = materialOf(scene, stackMaterial[level]);
          incoming = kFp ? fpLevelRadiance(mat, fpSubSampleTerm(mat, stackSpecular[level], incoming), 1.0)
                         : shadeTerm(mat, stackSpecular[level], incoming);
        }
        const MaterialView primaryMat = materialOf(scene, primaryMaterial);
        acc = add(acc, kFp ? fpSubSampleTerm(primaryMat, primarySpecular, incoming)
                           : shadeTerm(primaryMat, primarySpecular, incoming));
        ++subPath;
        if (subPath >= numSub) {
          colour = kFp ? fpLevelRadiance(primaryMat, acc, invNumSub) // src/fp/Render.cpp:118
                       : scale(acc, invNumSub);                      // Scene.cpp:178
          sampleDone = true;
        } else {
          depth = 0;
          bounce = true; // next stratum of the camera hit
        }
      }
      if (sampleDone) {
        args.samples[3 * sampleSlot + 0] = colour.x;
        args.samples[3 * sampleSlot + 1] = colour.y;
        args.samples[3 * sampleSlot + 2] = colour.z;
        mode = kNeedWork;
      }
    }
    __syncwarp();

    // ---- 4. ONE bounce site (Scene.cpp:155-175) ----
    if (bounce) {
      const bool fromPrimary = depth == 0;
      if (fromPrimary && !needSurface) { // a later stratum: reload what the camera hit stored
        surface.position = mk(primarySlot[0 * kBlock], primarySlot[1 * kBlock], primarySlot[2 * kBlock]);
        surface.normal = mk(primarySlot[3 * kBlock], primarySlot[4 * kBlock], primarySlot[5 * kBlock]);
        surface.incoming = mk(primarySlot[6 * kBlock], primarySlot[7 * kBlock], primarySlot[8 * kBlock]);
        surface.basisX = mk(primarySlot[9 * kBlock], primarySlot[10 * kBlock], primarySlot[11 * kBlock]);
        surface.basisY = mk(primarySlot[12 * kBlock], primarySlot[13 * kBlock], primarySlot[14 * kBlock]);
        surface.reflectivity = primarySlot[15 * kBlock];
        surface.material = primaryMaterial;
      }
      double ru, rv, rp;
      if (kFp) { // toUVSample draws u then v, then p (src/fp/Render.cpp:97-102,116)
        uint32_t words[6] = {0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll 1
        for (int triple = skippedTriple ? 0 : 1; triple < 2; ++triple) // one copy of the generator code
          rng.six(history, storeLimit, words);
        ru = canonicalFromWords(words[0], words[1]);
        rv = canonicalFromWords(words[2], words[3]);
        rp = canonicalFromWords(words[4], words[5]);
      } else {
        KeyedDraws{key0}.bounce(pixel, static_cast<uint32_t>(subPath), static_cast<uint32_t>(depth), ru, rv, rp);
      }
      double u = ru, v = rv; // (0 + r) / 1 exactly, below the first bounce
      if (fromPrimary) {
        // Strata: u-major in Scene.cpp:155-156, v-major in the fp way (src/fp/Render.cpp:109-110).
        // x / n == x * (1/n) exactly when n is a power of two (the default 4x4 strata)
        const double su = static_cast<double>(kFp ? subPath % args.firstBounceU : subPath / args.firstBounceV) + ru;
        const double sv = static_cast<double>(kFp ? subPath / args.firstBounceU : subPath % args.firstBounceV) + rv;
        u = args.firstBounceUPow2 ? su * args.invFirstBounceU : ieeeDiv(su, static_cast<double>(args.firstBounceU));
        v = args.firstBounceVPow2 ? sv * args.invFirstBounceV : ieeeDiv(sv, static_cast<double>(args.firstBounceV));
      }
      // coneSample() / hemisphereSample() (Samples.cpp:6-30) end in the same
      // normalised(basis.transform(cos(t)*r, sin(t)*r, z)): the few specular lanes only prepare
      // its inputs, then every bouncing lane runs that tail together.
      const bool specular = rp < surface.reflectivity;
      Basis frame{surface.basisX, surface.basisY, surface.normal};
      double angle = (2 * kPi) * u, radius = 0, zScale = 0;
      bool direct = false;
      V3 newDirection = mk(0, 0, 0);
      if (specular) {
        newDirection = reflect(surface.normal, surface.incoming);
        direct = coneSampleSetup(newDirection, materialOf(scene, surface.material).coneAngle(), u, v, frame,
                                 angle, radius, zScale);
      } else {
        radius = ieeeSqrt(v);
        zScale = ieeeSqrt(1 - v);
      }
      if (!direct) {
        double sinT, cosT;
        sinCos(angle, sinT, cosT);
        newDirection = normalised(transform(frame, mk(cosT * radius, sinT * radius, zScale)));
      }
      if (fromPrimary) {
        primarySpecular = specular;
      } else {
        stackMaterial[depth] = static_cast<uint16_t>(surface.material);
        stackSpecular[depth] = specular;
      }
      origin = surface.position;
      direction = newDirection;
      ++depth;
    }
    __syncwarp();
  }

  if (!resident)
    stream.drain();
  // one atomic per warp for the cast counter
  unsigned long long warpCasts = casts;
#pragma unroll 1
  for (int offset = 16; offset > 0; offset >>= 1)
    warpCasts += __shfl_down_sync(kFullMask, warpCasts, offset);
  if (lane == 0 && warpCasts)
    atomicAdd(args.castCounter, warpCasts);
}

// =============================================================================================
// Sequential (mt19937) kernel: the reference's exact stream, one pass per GROUP of kGroup lanes
// (a whole warp, or 16 / 8 / 4 lanes so that a warp walks 2 / 4 / 8 passes side by side).
// =============================================================================================
template <int kGroup>
__device__ __forceinline__ unsigned groupMaskOf(unsigned lane) {
  return kGroup == 32 ? kFullMask : ((1u << (kGroup & 31)) - 1u) << (lane & ~static_cast<unsigned>(kGroup - 1));
}

// uniform_real_distribution<double> draws from a pass's engine (pt_mt19937.cuh: GroupMt19937).  Two /
// three canonical doubles at once: when they do not straddle the end of a generation the state words
// are read independently instead of through dependent next() calls.
template <int kGroup>
__device__ __forceinline__ double mtCanonical(GroupMt19937<kGroup> &rng) {
  const uint32_t lo = rng.next();
  const uint32_t hi = rng.next();
  return canonicalFromWords(lo, hi);
}
template <int kGroup>
__device__ __forceinline__ void mtCanonical2(GroupMt19937<kGroup> &rng, double &a, double &b) {
  if (rng.index + 4 <= 624) {
    const uint32_t *w = rng.state + rng.index;
    const uint32_t w0 = rng.temper(w[0]), w1 = rng.temper(w[1]), w2 = rng.temper(w[2]), w3 = rng.temper(w[3]);
    rng.index += 4;
    a = canonicalFromWords(w0, w1);
    b = canonicalFromWords(w2, w3);
  } else {
    a = mtCanonical(rng);
    b = mtCanonical(rng);
  }
}
template <int kGroup>
__device__ __forceinline__ void mtCanonical3(GroupMt19937<kGroup> &rng, double &a, double &b, double &c) {
  if (rng.index + 6 <= 624) {
    const uint32_t *w = rng.state + rng.index;
    const uint32_t w0 = rng.temper(w[0]), w1 = rng.temper(w[1]), w2 = rng.temper(w[2]), w3 = rng.temper(w[3]),
                   w4 = rng.temper(w[4]), w5 = rng.temper(w[5]);
    rng.index += 6;
    a = canonicalFromWords(w0, w1);
    b = canonicalFromWords(w2, w3);
    c = canonicalFromWords(w4, w5);
  } else {
    a = mtCanonical(rng);
    b = mtCanonical(rng);
    c = mtCanonical(rng);
  }
}

// Scene::intersect by a group of kGroup lanes: the lanes stride the primitive lists (groupSweep), then
// an argmin over the group (groupArgmin).  kAcceptEpsilon: oo::Triangle::intersect (src/oo/Triangle.cpp:31) rejects
// `t < Epsilon` where Scene.cpp:94 accepts `t > Epsilon`.
template <int kGroup, bool kAcceptEpsilon = false>
__device__ __forceinline__ Nearest groupSweep(const DeviceScene &scene, V3 o, V3 d, unsigned glane, bool useSpheres,
                                              bool useTriangles, double nearerThanLimit) {
  Nearest best{nearerThanLimit, 0.0, kNoPrim};
  if (useSpheres) {
    for (uint32_t i = glane; i < scene.numSpheres; i += kGroup) {
      const double4 s = ldgDouble4(scene.spheres + i); // loop body of Scene.cpp:17-36
      const V3 op = sub(mk(s.x, s.y, s.z), o);
      const double b = dot(op, d);
      double determinant = fma(b, b, -dot(op, op)) + s.w;
      if (determinant < 0)
        continue;
      determinant = sqrt(determinant);
      const double minusT = b - determinant;
      const double plusT = b + determinant;
      if (minusT < kEpsilon && plusT < kEpsilon)
        continue;
      const double t = minusT > kEpsilon ? minusT : plusT;
      if (t < best.t) {
        best.t = t;
        best.prim = -(static_cast<int>(i) + 1);
      }
    }
  }
  if (useTriangles) {
    // A triangle only replaces a sphere hit when strictly nearer, which the argmin's ordering
    // below encodes; per lane the strict `<` keeps the lowest index among equals.
    // One 80-byte AoS record per triangle (triExact), tiles back to back: only the last one is padded,
    // so the real triangles are the first numTriangles slots (a padding record would reject itself,
    // det == 0, after a trip through the division's slow path).
    // (Testing two or four records per lane and trip side by side was measured: no gain, r2q — the
    // kernel waits on the shading chain, not on the trips.)
#pragma unroll 1
    for (uint32_t index = glane; index < scene.numTriangles; index += kGroup) {
      const double2 *record = reinterpret_cast<const double2 *>(scene.triExact + 10 * static_cast<size_t>(index));
      const double2 a0 = __ldg(record), a1 = __ldg(record + 1), a2 = __ldg(record + 2), a3 = __ldg(record + 3),
                    a4 = __ldg(record + 4);
      testTriangle<kAcceptEpsilon>(mk(a0.x, a0.y, a1.x), mk(a1.y, a2.x, a2.y), mk(a3.x, a3.y, a4.x), o, d,
                                   static_cast<int>(index), best);
    }
  }
  return best;
}

// The group's nearest hit in the serial scans' order (t, sphere-before-triangle, index); hit distances
// are positive, so their bit patterns order like unsigned integers.  Every lane of the WARP must call
// it (a lane without a ray passes Nearest{inf, 0, kNoPrim}).
template <int kGroup>
__device__ __forceinline__ Nearest groupArgmin(Nearest best, unsigned mask) {
  const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(best.t));
  const uint32_t order = best.prim == kNoPrim ? 0xffffffffu
                         : best.prim < 0      ? static_cast<uint32_t>(-best.prim - 1)
                                              : 0x40000000u + static_cast<uint32_t>(best.prim);
  int winner;
  if (kGroup == 32) { // three integer min-reductions (REDUX)
    const uint32_t hi = static_cast<uint32_t>(bits >> 32), lo = static_cast<uint32_t>(bits);
    const uint32_t minHi = __reduce_min_sync(kFullMask, hi);
    const uint32_t minLo = __reduce_min_sync(kFullMask, hi == minHi ? lo : 0xffffffffu);
    const bool nearest = hi == minHi && lo == minLo;
    const uint32_t minOrder = __reduce_min_sync(kFullMask, nearest ? order : 0xffffffffu);
    winner = __ffs(__ballot_sync(kFullMask, nearest && order == minOrder)) - 1;
  } else { // a butterfly inside each group: REDUX on a partial mask is emulated, shuffles are not
    unsigned long long minBits = bits;
    uint32_t minOrder = order;
#pragma unroll
    for (int offset = 1; offset < kGroup; offset <<= 1) {
      const unsigned long long otherBits = __shfl_xor_sync(kFullMask, minBits, offset);
      const uint32_t otherOrder = __shfl_xor_sync(kFullMask, minOrder, offset);
      const bool less = otherBits < minBits || (otherBits == minBits && otherOrder < minOrder);
      minBits = less ? otherBits : minBits;
      minOrder = less ? otherOrder : minOrder;
    }
    winner = __ffs(__ballot_sync(kFullMask, bits == minBits && order == minOrder) & mask) - 1; // a lane of the warp
  }
  best.t = __shfl_sync(kFullMask, best.t, winner);
  best.det = __shfl_sync(kFullMask, best.det, winner);
  best.prim = __shfl_sync(kFullMask, best.prim, winner);
  return best;
}

template <int kGroup, bool kAcceptEpsilon = false>
__device__ __forceinline__ Nearest groupIntersect(const DeviceScene &scene, V3 o, V3 d, unsigned glane, unsigned mask,
                                                  bool useSpheres, bool useTriangles, double nearerThanLimit) {
  return groupArgmin<kGroup>(groupSweep<kGroup, kAcceptEpsilon>(scene, o, d, glane, useSpheres, useTriangles, nearerThanLimit),
                             mask);
}

template <bool kAcceptEpsilon = false>
__device__ __forceinline__ Nearest warpIntersect(const DeviceScene &scene, V3 o, V3 d, unsigned lane,
                                                 bool useSpheres, bool useTriangles,
                                                 double nearerThanLimit) {
  return groupIntersect<32, kAcceptEpsilon>(scene, o, d, lane, kFullMask, useSpheres, useTriangles, nearerThanLimit);
}

// One pass = one std::mt19937(seed + s) walked over the frame in row-major order, camera ray then
// radiance() per pixel (Scene.cpp:208-216): the work within a pass is one dependent chain, so the only
// parallelism is passes x primitives.  A group's lanes share the pass: they split the primitive
// list in groupIntersect() and compute everything else redundantly (group-uniform); the groups of a
// warp run the same state machine on different passes and diverge like the keyed kernel's lanes do.
// Smaller groups mean less redundant shading per pass and more sweep trips per cast: the launcher
// picks the group from the number of passes (sequentialLanesPerPass()).
//
// kOo: the reference's `oo` way (src/oo/Renderer.cpp:60-107) walks the same stream in the same
// order; its estimator differs from Scene::radiance in where the emission is added
// (Material::totalEmission after the average, :90) and in accepting t == Epsilon.
template <int kWarps, int kGroup, bool kOo>
__global__ void __launch_bounds__(kWarps * 32)
    renderSequentialKernel(const __grid_constant__ SequentialArgs args) {
  constexpr int kPassesPerWarp = 32 / kGroup;
  __shared__ uint32_t mtState[kWarps * kPassesPerWarp][624];
  __shared__ Surface primarySlots[kWarps * kPassesPerWarp];
  const DeviceScene &scene = args.scene;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;
  const unsigned glane = lane & static_cast<unsigned>(kGroup - 1);
  const unsigned groupInWarp = lane / kGroup;
  const unsigned mask = groupMaskOf<kGroup>(lane);
  const int passInBatch = (static_cast<int>(blockIdx.x) * kWarps + static_cast<int>(warp)) * kPassesPerWarp +
                          static_cast<int>(groupInWarp);
  // A group without a pass (the tail of the last CTA) stays with its warp: the loop below is uniform
  // over the WARP, so that its groups meet again every iteration and sweep together.
  const bool hasPass = passInBatch < args.numPasses;
  GroupMt19937<kGroup> rng{mtState[warp * kPassesPerWarp + groupInWarp], 0, 0, mask, glane};
  if (hasPass)
    rng.seed(static_cast<uint32_t>(args.seed + args.passBegin + passInBatch));

  const int numSub = args.firstBounceU * args.firstBounceV;
  const double invNumSub = 1.0 / static_cast<double>(numSub);
  const bool uPow2 = (args.firstBounceU & (args.firstBounceU - 1)) == 0;
  const bool vPow2 = (args.firstBounceV & (args.firstBounceV - 1)) == 0;
  const double invFirstBounceU = 1.0 / static_cast<double>(args.firstBounceU);
  const double invFirstBounceV = 1.0 / static_cast<double>(args.firstBounceV);
  const V3 environment = mk(scene.environment[0], scene.environment[1], scene.environment[2]);
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  double *passSamples = args.samples + static_cast<size_t>(passInBatch) * args.width * args.height * 3;
  uint64_t casts = 0;
  uint16_t stackMaterial[kMaxDepth];
  bool stackSpecular[kMaxDepth];

  // Everything below is uniform over the group: its lanes only differ inside groupIntersect().
  const int numPixels = args.width * args.height;
  int pixelIndex = -1;
  int px = -1, py = 0;
  V3 origin = mk(0, 0, 0), direction = mk(0, 0, 1);
  V3 colour = mk(0, 0, 0);
  int depth = 0;
  int subPath = 0;
  // The camera ray's surface waits in shared memory while the sample's sub-paths are traced: written by
  // the group's first lane once per sample, read by all of them at every sub-path's first bounce.
  Surface &primary = primarySlots[warp * kPassesPerWarp + groupInWarp];
  bool primarySpecular = false;
  V3 acc = mk(0, 0, 0);
  bool sampleDone = true;
  bool finished = !hasPass;

  for (;;) { // the keyed megakernel's state machine, one path at a time per group
    if (sampleDone && !finished) { // store the finished pixel, start the next one: row-major (Scene.cpp:212-213)
      if (pixelIndex >= 0 && glane == 0) {
        double *dst = passSamples + 3 * static_cast<size_t>(pixelIndex);
        dst[0] = colour.x;
        dst[1] = colour.y;
        dst[2] = colour.z;
      }
      if (++pixelIndex >= numPixels) {
        finished = true;
      } else {
        if (++px == args.width) {
          px = 0;
          ++py;
        }
        // Camera::randomRay draws before radiance() looks at the depth (Scene.cpp:214-215).
        double ux, uy, ua = 0, ur = 0;
        mtCanonical2(rng, ux, uy);
        if (args.camera.apertureRadius != 0)
          mtCanonical2(rng, ua, ur);
        cameraRay(args.camera, px, py, ux, uy, ua, ur, origin, direction);
        colour = mk(0, 0, 0);
        depth = 0;
        subPath = 0;
        sampleDone = args.maxDepth <= 0; // radiance() returns Vec3() at once (Scene.cpp:128-129)
      }
    }
    if (__all_sync(kFullMask, finished))
      break;
    // The warp's groups are together here: they twist their consumed words, sweep and take the group
    // minimum side by side; a group between two pixels (maxDepth <= 0) or past its last one idles.
    const bool casting = !(finished || sampleDone);
    __syncwarp(); // last iteration's reads of the shared slots precede this one's writes
    rng.advance();
    Nearest best{inf, 0.0, kNoPrim};
    if (casting)
      best = groupSweep<kGroup, kOo>(scene, origin, direction, glane, true, true, inf);
    best = groupArgmin<kGroup>(best, mask);
    if (!casting)
      continue;
    ++casts;
    bool ended = false, bounce = false, terminalPrimary = false;
    V3 incoming = mk(0, 0, 0);
    Surface surface{};
    if (best.prim == kNoPrim) {
      incoming = environment;
      ended = true;
    } else {
      const HitInfo hit = finishHit(scene, scene.spheres, origin, direction, best);
      const MaterialView mat = materialOf(scene, hit.material);
      if (depth == 0 && args.preview) {
        incoming = mat.diffuse();
        ended = true;
      } else if (depth + 1 >= args.maxDepth) {
        // The deepest level draws its (u, v, p) triples and builds rays whose radiance is
        // Vec3() (Scene.cpp:128-129,157-175): consume the stream, contribute the emission.
        rng.skip(6u * static_cast<uint32_t>(depth == 0 ? numSub : 1));
        incoming = shadeTerm(mat, true, mk(0, 0, 0)); // oo: emission + (0 + .. + 0) * (1/n), the same value
        ended = true;
        terminalPrimary = depth == 0;
      } else {
        surface.position = hit.position;
        surface.normal = hit.normal;
        surface.incoming = direction;
        surface.material = hit.material;
        surface.reflectivity = hitReflectivity(mat, hit, direction);
        hitBasis(scene, hit, surface.basisX, surface.basisY);
        if (depth == 0) {
          if (glane == 0)
            primary = surface;
          __syncwarp(mask);
          acc = mk(0, 0, 0);
          subPath = 0;
        }
        bounce = true;
      }
    }
    if (ended) {
      colour = incoming;
      if (depth == 0) {
        if (terminalPrimary && !kOo) {
          acc = mk(0, 0, 0);
          for (int k = 0; k < numSub; ++k)
            acc = add(acc, incoming);
          colour = scale(acc, invNumSub);
        }
        sampleDone = true;
      } else {
        for (int level = depth - 1; level >= 1; --level) {
          const MaterialView mat = materialOf(scene, stackMaterial[level]);
          incoming = kOo ? ooLevelRadiance(mat, fpSubSampleTerm(mat, stackSpecular[level], incoming), 1.0)
                         : shadeTerm(mat, stackSpecular[level], incoming);
        }
        const MaterialView primaryMat = materialOf(scene, primary.material);
        acc = add(acc, kOo ? fpSubSampleTerm(primaryMat, primarySpecular, incoming)
                           : shadeTerm(primaryMat, primarySpecular, incoming));
        ++subPath;
        if (subPath >= numSub) {
          colour = kOo ? ooLevelRadiance(primaryMat, acc, invNumSub) // src/oo/Renderer.cpp:90
                       : scale(acc, invNumSub);                      // Scene.cpp:178
          sampleDone = true;
        } else {
          depth = 0;
          bounce = true;
        }
      }
    }
    if (bounce) {
      const bool fromPrimary = depth == 0;
      if (fromPrimary)
        surface = primary;
      double ru, rv, rp; // u, v, p in this order (Scene.cpp:157-161)
      mtCanonical3(rng, ru, rv, rp);
      double u = ru, v = rv;
      if (fromPrimary) {
        // a division by a power of two is the multiplication by its (exact) reciprocal, bit for bit
        const double su = static_cast<double>(subPath / args.firstBounceV) + ru;
        const double sv = static_cast<double>(subPath % args.firstBounceV) + rv;
        u = uPow2 ? su * invFirstBounceU : ieeeDiv(su, static_cast<double>(args.firstBounceU));
        v = vPow2 ? sv * invFirstBounceV : ieeeDiv(sv, static_cast<double>(args.firstBounceV));
      }
      const MaterialView mat = materialOf(scene, surface.material);
      bool specular;
      V3 newDirection;
      if (rp < surface.reflectivity) {
        newDirection = coneSample(reflect(surface.normal, surface.incoming), mat.coneAngle(), u, v);
        specular = true;
      } else {
        newDirection = hemisphereSample(Basis{surface.basisX, surface.basisY, surface.normal}, u, v);
        specular = false;
      }
      if (fromPrimary) {
        primarySpecular = specular;
      } else {
        stackMaterial[depth] = static_cast<uint16_t>(surface.material);
        stackSpecular[depth] = specular;
      }
      origin = surface.position;
      direction = newDirection;
      ++depth;
    }
  }
  if (glane == 0)
    atomicAdd(args.castCounter, static_cast<unsigned long long>(casts));
}

// =============================================================================================
// Pass-ordered accumulation.
// =============================================================================================
__global__ void reducePassesKernel(const __grid_constant__ ReduceArgs args) {
  const uint32_t own = blockIdx.x * blockDim.x + threadIdx.x;
  if (own >= args.ownPixels)
    return;
  const uint32_t row = own / args.width;
  const uint32_t px = own % args.width;
  const uint32_t py = args.rowBegin + row * args.rowStep;
  const size_t samplePixel = args.samplesAreFullFrame ? (px + static_cast<size_t>(py) * args.width) : own;
  PtPixelDevice *dst = args.accumulator + (px + static_cast<size_t>(py) * args.width);
  double r = dst->sum[0], g = dst->sum[1], b = dst->sum[2];
  for (uint32_t p = 0; p < args.numPasses; ++p) {
    const double *s = args.samples + 3 * (static_cast<size_t>(p) * args.samplePassStride + samplePixel);
    r += s[0];
    g += s[1];
    b += s[2];
  }
  dst->sum[0] = r;
  dst->sum[1] = g;
  dst->sum[2] = b;
  dst->numSamples += args.numPasses;
}

// =============================================================================================
// Scene::intersect for tests: one ray per lane (block-wide sweep from shared memory, the
// megakernel's path) or one ray per warp (the sequential kernel's path).
// =============================================================================================
__global__ void intersectKernel(const __grid_constant__ IntersectArgs args) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  const DeviceScene &scene = args.scene;
  TileStream stream = makeTileStream(smemRaw, scene, args.sweep);
  stream.start();
  for (uint32_t i = threadIdx.x; i < scene.numSpheres; i += blockDim.x)
    stream.spheres()[i] = scene.spheres[i];
  __syncthreads();
  const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = ray < args.numRays;
  V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
  if (live) {
    o = mk(args.rays[6 * ray + 0], args.rays[6 * ray + 1], args.rays[6 * ray + 2]);
    d = mk(args.rays[6 * ray + 3], args.rays[6 * ray + 4], args.rays[6 * ray + 5]);
  }
  const double inf = __longlong_as_double(0x7ff0000000000000ll);
  Nearest best{args.which == 0 ? inf : args.nearerThan, 0.0, kNoPrim};
  if (args.warpCooperative) {
    // every lane of the warp walks the warp's 32 rays one at a time
    const unsigned lane = threadIdx.x & 31u;
    for (int r = 0; r < 32; ++r) {
      const V3 ro = mk(__shfl_sync(kFullMask, o.x, r), __shfl_sync(kFullMask, o.y, r), __shfl_sync(kFullMask, o.z, r));
      const V3 rd = mk(__shfl_sync(kFullMask, d.x, r), __shfl_sync(kFullMask, d.y, r), __shfl_sync(kFullMask, d.z, r));
      const Nearest n = warpIntersect(scene, ro, rd, lane, args.which != 2, args.which != 1,
                                      args.which == 0 ? inf : args.nearerThan);
      if (static_cast<int>(lane) == r)
        best = n;
    }
    if (scene.numTiles > 1)
      stream.drain();
    else if (scene.numTiles == 1)
      stream.acquire();
  } else {
    if (args.which != 2)
      sweepSpheres(stream.spheres(), static_cast<int>(scene.numSpheres), o, d, best);
    if (args.which != 1) {
      for (uint32_t j = 0; j < scene.numTiles; ++j) {
        const unsigned char *tile = stream.acquire();
        if (args.sweep == 7)
          sweepStagedTile<7>(scene, tile, j, o, d, best);
        else if (args.sweep == 6)
          sweepStagedTile<6>(scene, tile, j, o, d, best);
        else if (args.sweep == 5)
          sweepStagedTile<5>(scene, tile, j, o, d, best);
        else if (args.sweep == 4)
          sweepStagedTile<4>(scene, tile, j, o, d, best);
        else if (args.sweep == 3)
          sweepStagedTile<3>(scene, tile, j, o, d, best);
        else if (args.sweep == 2)
          sweepStagedTile<2>(scene, tile, j, o, d, best);
        else if (args.sweep == 1)
          sweepStagedTile<1>(scene, tile, j, o, d, best);
        else
          sweepStagedTile<0>(scene, tile, j, o, d, best);
        if (scene.numTiles > 1)
          stream.release();
      }
      if (scene.numTiles > 1)
        stream.drain();
    } else if (scene.numTiles > 1) {
      stream.drain();
    } else if (scene.numTiles == 1) {
      stream.acquire();
    }
  }
  if (!live)
    return;
  PtHitDevice out{};
  if (best.prim != kNoPrim) {
    const HitInfo hit = finishHit(scene, stream.spheres(), o, d, best);
    out.hit = 1;
    out.inside = hit.inside ? 1 : 0;
    out.material = static_cast<int32_t>(hit.material);
    out.primitive = best.prim;
    out.distance = best.t;
    out.position[0] = hit.position.x; out.position[1] = hit.position.y; out.position[2] = hit.position.z;
    out.normal[0] = hit.normal.x; out.normal[1] = hit.normal.y; out.normal[2] = hit.normal.z;
  }
  args.out[ray] = out;
}

// =============================================================================================
// FP32 stage-0 data: float copies of {v0, e1, e2} and the per-triangle error bounds.
// =============================================================================================
__global__ void buildFilterKernel(const __grid_constant__ BuildFilterArgs args) {
  const DeviceScene &scene = args.scene;
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= scene.numTiles * scene.tileTris)
    return;
  const uint32_t tile = slot / scene.tileTris, within = slot % scene.tileTris;
  const double *src = scene.triSweep + static_cast<size_t>(tile) * 9 * scene.tileTris + within;
  // blocked layout: [tile][group of 4][14][4]
  float *dst = args.out + (static_cast<size_t>(tile) * (scene.tileTris / 4) + within / 4) * (kFilterFloats * 4) +
               within % 4;
  double v[9];
  for (int k = 0; k < 9; ++k) {
    v[k] = src[static_cast<size_t>(k) * scene.tileTris];
    dst[k * 4] = __double2float_rn(v[k]);
  }
  const bool padding = slot >= scene.numTriangles;
  const double lenV0 = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  const double lenE1 = sqrt(v[3] * v[3] + v[4] * v[4] + v[5] * v[5]);
  const double lenE2 = sqrt(v[6] * v[6] + v[7] * v[7] + v[8] * v[8]);
  // 2^-18 = 64 FP32 unit roundoffs (2^-24).  Forward error analysis of the 24 (+3) operations
  // on inputs rounded to FP32 gives |det32 - det| <= ~13u |e1||e2|, |X32 - X| <= ~16u r|e2|,
  // |Y32 - Y| <= ~18u r|e1|, |T32 - T| <= ~18u r|e1||e2| with r = |o| + |v0|; the rest is safety
  // margin, which also covers the rounding of the comparisons themselves.
  const double c = 0x1p-18;
  const double reach = args.originBound + lenV0;
  const double ed = c * lenE1 * lenE2;
  const double ex = c * lenE2 * reach;
  const double ey = c * lenE1 * reach;
  const double et = c * lenE1 * lenE2 * reach;
  const double slack = 1.0 + 0x1p-10;
  float fed = __double2float_ru(ed * slack);
  float fkx = __double2float_ru(2.0 * ex * slack);
  float fky = __double2float_ru(2.0 * ey * slack);
  float fk3 = __double2float_ru((ed * (1.0 + 0x1p-20) + ex + ey) * slack);
  float fkt = __double2float_ru(2.0 * et * slack);
  if (padding) { // all-zero padding triangles: make stage 0 reject them outright
    fed = -1.0f;
    fkx = -1.0f;
    fky = -1.0f;
    fk3 = 0.0f;
    fkt = 0.0f;
  }
  dst[9 * 4] = fed;
  dst[10 * 4] = fkx;
  dst[11 * 4] = fky;
  dst[12 * 4] = fk3;
  dst[13 * 4] = fkt;
  // Moment (Pluecker) form for sweep variant 7, blocked [tile][group of 4][19][4]: the per-triangle
  // constants of stage0RejectMoment(), evaluated in FP64 and rounded once; the same bounds.
  const uint32_t groupIndex = tile * (scene.tileTris / 4) + within / 4;
  const bool fan = (scene.fanMask[groupIndex >> 5] >> (groupIndex & 31u)) & 1u;
  const uint32_t lanePosition = fan ? ((within & 1u) << 1 | ((within >> 1) & 1u)) : within % 4; // [A0, A1, B0, B1]
  float *mom = args.outMoment + static_cast<size_t>(groupIndex) * (kMomentFloats * 4) + lanePosition;
  const V3 v0 = mk(v[0], v[1], v[2]), e1 = mk(v[3], v[4], v[5]), e2 = mk(v[6], v[7], v[8]);
  const V3 nn = cross(e2, e1), a2 = cross(v0, e2), a1 = cross(v0, e1);
  const double moment[15] = {nn.x, nn.y, nn.z, e2.x, e2.y, e2.z, a2.x, a2.y, a2.z,
                             -e1.x, -e1.y, -e1.z, -a1.x, -a1.y, -a1.z};
  for (int k = 0; k < 15; ++k)
    mom[k * 4] = __double2float_rn(moment[k]);
  mom[15 * 4] = fed;
  mom[16 * 4] = fkx;
  mom[17 * 4] = fky;
  mom[18 * 4] = fk3;
}

// For every (ray, triangle): does stage 0 keep it, does the exact test accept it (with no
// nearer-than limit), and is there any triangle the exact test accepts but stage 0 rejects?
__global__ void auditStage0Kernel(const __grid_constant__ AuditArgs args) {
  const DeviceScene &scene = args.scene;
  const uint32_t ray = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long pairs = 0, kept = 0, accepted = 0, violations = 0;
  if (ray < args.numRays) {
    const V3 o = mk(args.rays[6 * ray + 0], args.rays[6 * ray + 1], args.rays[6 * ray + 2]);
    const V3 d = mk(args.rays[6 * ray + 3], args.rays[6 * ray + 4], args.rays[6 * ray + 5]);
    const Stage0Ray r{static_cast<float>(o.x), static_cast<float>(o.y), static_cast<float>(o.z),
                      static_cast<float>(d.x), static_cast<float>(d.y), static_cast<float>(d.z)};
    const MomentRay momentRay = makeMomentRay(o, d);
    for (uint32_t index = 0; index < scene.numTriangles; ++index) {
      const uint32_t tile = index / scene.tileTris, i = index % scene.tileTris;
      const float *f = scene.triFilter +
                       (static_cast<size_t>(tile) * (scene.tileTris / 4) + i / 4) * (kFilterFloats * 4) + i % 4;
      const double *e = scene.triSweep + static_cast<size_t>(tile) * 9 * scene.tileTris + i;
      const uint32_t n = scene.tileTris;
      const uint32_t groupIndex = tile * (scene.tileTris / 4) + i / 4;
      const bool fan = (scene.fanMask[groupIndex >> 5] >> (groupIndex & 31u)) & 1u;
      // (a fan group's second triangles keep their own Y rows: the audit checks the plain decision,
      // which the fan identity reproduces bit for bit)
      const float *g = scene.triMoment + static_cast<size_t>(groupIndex) * (kMomentFloats * 4) +
                       (fan ? ((i & 1u) << 1 | ((i >> 1) & 1u)) : i % 4);
      const bool keep = args.momentForm
                            ? stage0KeepMoment(g, 4, momentRay)
                            : stage0Keep(f[0], f[4], f[8], f[12], f[16], f[20], f[24], f[28], f[32], f[36], f[40], f[44],
                                         f[48], f[52], r);
      Nearest best{__longlong_as_double(0x7ff0000000000000ll), 0.0, kNoPrim};
      testTriangle(mk(e[0], e[n], e[2 * n]), mk(e[3 * n], e[4 * n], e[5 * n]), mk(e[6 * n], e[7 * n], e[8 * n]),
                   o, d, static_cast<int>(index), best);
      const bool accept = best.prim != kNoPrim;
      ++pairs;
      kept += keep;
      accepted += accept;
      violations += accept && !keep;
    }
  }
  atomicAdd(args.counters + 0, pairs);
  atomicAdd(args.counters + 1, kept);
  atomicAdd(args.counters + 2, accepted);
  atomicAdd(args.counters + 3, violations);
}

// =============================================================================================
// DFMA throughput probe: 8 independent accumulator chains per thread.
// =============================================================================================
__global__ void fp64PeakKernel(double *sink, int iterations) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iterations; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double total = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (total == 12345.678)
    sink[0] = total;
}

// Packed-FP32 throughput probe (FFMA2, the instruction of the stage-0 sweep): 8 independent chains.
__global__ void fp32PeakKernel(float *sink, int iterations) {
  float2 a0 = make_float2(threadIdx.x * 1e-6f, 1.f), a1 = a0, a2 = a0, a3 = a0, a4 = a0, a5 = a0, a6 = a0, a7 = a0;
  const float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
  for (int i = 0; i < iterations; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      a0 = __ffma2_rn(a0, m, c); a1 = __ffma2_rn(a1, m, c); a2 = __ffma2_rn(a2, m, c); a3 = __ffma2_rn(a3, m, c);
      a4 = __ffma2_rn(a4, m, c); a5 = __ffma2_rn(a5, m, c); a6 = __ffma2_rn(a6, m, c); a7 = __ffma2_rn(a7, m, c);
    }
  }
  const float total = (a0.x + a1.y) + (a2.x + a3.y) + (a4.x + a5.y) + (a6.x + a7.y);
  if (total == 12345.678f)
    sink[0] = total;
}

// =============================================================================================
// Host-side launchers (called from ptb200_shim.cu).
// =============================================================================================
constexpr int kSequentialWarps = 2;

template <int kBlock, int kMinBlocks, int kSweep, int kWay = 0>
cudaError_t launchKeyedConfig(const KeyedArgs &args, int numSms, cudaStream_t stream) {
  auto kernel = renderKeyedKernel<kBlock, kMinBlocks, kSweep, kWay>;
  const size_t smemBytes = keyedSmemBytes(args.scene.numSpheres, args.scene.tileTris, args.scene.numTiles, kSweep, kBlock, kWay);
  cudaError_t err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  int perSm = 0;
  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, kBlock, smemBytes);
  if (err != cudaSuccess)
    return err;
  if (perSm < 1)
    return cudaErrorInvalidConfiguration;
  // Persistent grid: every SM holds `perSm` CTAs for the whole launch (148 x perSm on B200).
  unsigned long long wanted = (args.totalItems + kBlock - 1) / kBlock;
  unsigned long long grid = static_cast<unsigned long long>(numSms) * perSm;
  if (wanted < grid)
    grid = wanted ? wanted : 1;
  if (kWay == 1 && (args.mtHistory == nullptr || grid * kBlock > args.mtHistoryThreads ||
                    args.mtHistoryThreads * kMtHistoryStride > 0xffffffffull))
    return cudaErrorInvalidValue; // scratch for the per-lane engines: see mtHistoryThreadsFor()
  kernel<<<static_cast<unsigned>(grid), kBlock, smemBytes, stream>>>(args);
  return cudaGetLastError();
}

// A configuration is 10 * launchShape + sweepVariant, + 100 for the three-kernel pipeline.
//   sweep variants: 0 one-stage FP64; 1 two-stage FP64 (prefilter + exact); 2 FP32 stage 0 + exact;
//                   3 the same with the packed FP32x2 datapath (FFMA2); 4 = 3 + stage 0 also rejects
//                   triangles certainly behind the ray; 5 = 4 and 6 = 3 with the stage-0 decisions kept
//                   in sign bits; 7 = 6 in moment (Pluecker) form.  All of them stay reachable through
//                   ptb200_intersect (tests); the render kernels are instantiated for 1, 6 and 7 only —
//                   the others lost every measurement (profiles/README.md).
//   launch shapes:  0 = 256 threads x 2 CTAs/SM; 2 = 256 x 3; 3 = 192 x 4; 4 = 128 x 5.
// Default (measured on B200, profiles/r2b..r2d): the pipeline with the moment-form stage 0 wherever the
// FP32 filter is usable, three CTAs per SM for small scenes (shading latency rather than the sweep
// limits them), two for scenes whose tile needs the shared memory.  PTB200_KEYED_CONFIG overrides it
// (tools/sweep_configs.py).
int chooseKeyedConfig(uint32_t numTriangles, bool filterUsable, int way) {
  static const int forced = [] {
    const char *env = getenv("PTB200_KEYED_CONFIG");
    return env ? atoi(env) : -1;
  }();
  if (forced >= 0)
    return forced;
  const bool small = numTriangles <= 512;
  if (way == 1)
    return filterUsable ? 6 : 1; // the fp way: the megakernel, two CTAs per SM
  const int sweep = filterUsable ? 7 : 1;
  // the dod estimator with keyed draws: the three-kernel pipeline of pt_split.cu (100 + ...); scenes of
  // up to 64 triangles keep their stage-0 table in the constant bank (variant 8: 220.6 vs 197.5
  // Msamples/s on the Cornell box, profiles/README.md r2g)
  if (filterUsable && numTriangles > 0 && constTableFits(numTriangles, 1))
    return 128;
  // scenes whose tile fills the shared memory are bound by its data pipe: two sub-paths per lane share
  // every group of triangles they load (suzanne 40.0 -> 46.4, ce 3.65 -> 4.08 Msamples/s; one CTA of 512
  // threads per SM at 128 registers)
  if (filterUsable && !small)
    return 217;
  return 100 + 10 * (small ? 2 : 0) + sweep;
}

// Threads a persistent fp-way grid can have: its instantiations are 256 threads x <= 3 CTAs/SM.
size_t mtHistoryThreadsFor(int numSms) { return static_cast<size_t>(numSms) * 768; }

cudaError_t launchRenderKeyed(const KeyedArgs &args, int numSms, int config, cudaStream_t stream) {
  if (args.way == 1) { // the fp way: FP64 fallback, sign-bit stage 0 at two / three CTAs per SM
    switch (config) {
    case 1: return launchKeyedConfig<256, 2, 1, 1>(args, numSms, stream);
    case 6: return launchKeyedConfig<256, 2, 6, 1>(args, numSms, stream);
    case 26: return launchKeyedConfig<256, 3, 6, 1>(args, numSms, stream);
    default: return cudaErrorInvalidValue;
    }
  }
  // The dod estimator with keyed draws in ONE kernel: round 1's form, kept as the A/B reference
  // of the three-kernel pipeline (pt_split.cu, configurations 100 + ...) that replaced it.
  switch (config) {
  case 1: return launchKeyedConfig<256, 2, 1>(args, numSms, stream);
  case 6: return launchKeyedConfig<256, 2, 6>(args, numSms, stream);
  case 26: return launchKeyedConfig<256, 3, 6>(args, numSms, stream);
  default: return cudaErrorInvalidValue;
  }
}

template <int kGroup>
static int sequentialOccupancy() {
  int ctas = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, renderSequentialKernel<kSequentialWarps, kGroup, false>,
                                                    kSequentialWarps * 32, 0) != cudaSuccess || ctas < 1)
    ctas = 5;
  return ctas * kSequentialWarps;
}
static int sequentialWarpsPerSm(int group) {
  static const int warps[4] = {sequentialOccupancy<4>(), sequentialOccupancy<8>(), sequentialOccupancy<16>(),
                               sequentialOccupancy<32>()};
  return warps[group == 4 ? 0 : group == 8 ? 1 : group == 16 ? 2 : 3];
}

// Lanes per pass of the sequential kernel.  A pass is one dependent chain, so the machine is filled by
// passes alone and a launch is latency-bound until every SM holds its `residentWarps`.  With few passes
// every pass gets a whole warp: the sweep is split 32 ways and the chain is as short as it gets (2.2 us
// per cast on the Cornell box; 3.0 / 3.9 / 5.0 us with 16 / 8 / 4 lanes, whose 2 / 4 / 8 passes per warp
// diverge in the shading).  More passes than resident warps would run in waves, so the group is the
// largest one that keeps the launch in ONE wave: 4096 passes on 8 lanes are 1024 warps side by side
// (21.5 Msamples/s) where a warp per pass takes 2.3 waves (12.3).  Measured: profiles/README.md r2r.
// `requested` (PtRenderOptions.lanesPerPass or the environment variable PTB200_SEQUENTIAL_LANES)
// overrides the choice.
int sequentialLanesPerPass(int numPasses, int numSms, int requested) {
  if (requested <= 0) {
    const char *env = getenv("PTB200_SEQUENTIAL_LANES");
    requested = env ? atoi(env) : 0;
  }
  if (requested == 4 || requested == 8 || requested == 16 || requested == 32)
    return requested;
  // the largest group whose launch fits the warps its own kernel keeps resident
  int group = 32;
  while (group > 4 && static_cast<long long>(numPasses) * group / 32 >
                          static_cast<long long>(numSms) * sequentialWarpsPerSm(group))
    group /= 2;
  return group;
}

template <int kGroup>
static cudaError_t launchSequentialGroup(const SequentialArgs &args, cudaStream_t stream) {
  constexpr int kPassesPerCta = kSequentialWarps * (32 / kGroup);
  const int blocks = (args.numPasses + kPassesPerCta - 1) / kPassesPerCta;
  if (args.way == 2)
    renderSequentialKernel<kSequentialWarps, kGroup, true><<<blocks, kSequentialWarps * 32, 0, stream>>>(args);
  else if (args.way == 0)
    renderSequentialKernel<kSequentialWarps, kGroup, false><<<blocks, kSequentialWarps * 32, 0, stream>>>(args);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launchRenderSequential(const SequentialArgs &args, int lanesPerPass, cudaStream_t stream) {
  switch (lanesPerPass) {
  case 32: return launchSequentialGroup<32>(args, stream);
  case 16: return launchSequentialGroup<16>(args, stream);
  case 8: return launchSequentialGroup<8>(args, stream);
  case 4: return launchSequentialGroup<4>(args, stream);
  default: return cudaErrorInvalidValue;
  }
}

cudaError_t launchReducePasses(const ReduceArgs &args, cudaStream_t stream) {
  const int block = 256;
  const int grid = static_cast<int>((args.ownPixels + block - 1) / block);
  reducePassesKernel<<<grid, block, 0, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launchBuildFilter(const BuildFilterArgs &args, cudaStream_t stream) {
  const uint32_t slots = args.scene.numTiles * args.scene.tileTris;
  if (slots == 0)
    return cudaSuccess;
  buildFilterKernel<<<(slots + 127) / 128, 128, 0, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launchAuditStage0(const AuditArgs &args, cudaStream_t stream) {
  auditStage0Kernel<<<(args.numRays + 127) / 128, 128, 0, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launchIntersect(const IntersectArgs &args, cudaStream_t stream) {
  const size_t smemBytes = keyedSmemBytes(args.scene.numSpheres, args.scene.tileTris, args.scene.numTiles, args.sweep, 0, 0);
  cudaError_t err = cudaFuncSetAttribute(intersectKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smemBytes));
  if (err != cudaSuccess)
    return err;
  const int block = 128;
  const int grid = static_cast<int>((args.numRays + block - 1) / block);
  intersectKernel<<<grid, block, smemBytes, stream>>>(args);
  return cudaGetLastError();
}

cudaError_t launchFp32Peak(float *sink, int iterations, int blocks, int threads, cudaStream_t stream) {
  fp32PeakKernel<<<blocks, threads, 0, stream>>>(sink, iterations);
  return cudaGetLastError();
}

cudaError_t launchFp64Peak(double *sink, int iterations, int blocks, int threads, cudaStream_t stream) {
  fp64PeakKernel<<<blocks, threads, 0, stream>>>(sink, iterations);
  return cudaGetLastError();
}

} // namespace ptb200
