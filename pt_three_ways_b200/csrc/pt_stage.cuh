// Pieces shared by the path-tracing kernels (pt_kernels.cu: the one-kernel megakernel, the
// sequential kernel; pt_split.cu: the primary / sub-path / resolve pipeline): shared-memory
// staging of the primitive list by TMA, the per-tile sweep dispatch, the keyed Philox draws.
#pragma once

#include "pt_device.cuh"

namespace ptb200 {

// =============================================================================================
// Shared-memory staging of the scene.
// =============================================================================================
// Dynamic shared memory of a sweeping CTA (offsets from the 128-byte aligned base):
//   [0, 16)            two mbarriers
//   [32, 32 + 32 S)    spheres {centre, r^2}
//   [tileOffset, ...)  tile buffer 0, then tile buffer 1 when the scene streams (numTiles > 1)
// Pointers are always formed as  smemBase + offset  so the compiler keeps them in the shared
// address space (LDS, not generic LD).
__host__ __device__ inline uint32_t smemTileOffset(uint32_t numSpheres) {
  return 128u + ((numSpheres * 32u + 127u) & ~127u);
}

constexpr uint32_t kExactBytesPerTriangle = 72;  // 9 doubles
constexpr uint32_t kFilterBytesPerTriangle = 56; // 14 floats
constexpr uint32_t kMomentBytesPerTriangle = 76; // 19 floats (sweep variant 7)
__host__ __device__ inline uint32_t sweepBytesPerTriangle(int sweep) {
  return sweep >= 8   ? 80u // variants 8/9 stage the survivors' AoS records (triExact), not a stage-0 tile
         : sweep == 7 ? kMomentBytesPerTriangle
         : sweep >= 2 ? kFilterBytesPerTriangle
                      : kExactBytesPerTriangle;
}

__host__ __device__ inline uint32_t smemAfterTiles(uint32_t numSpheres, uint32_t tileTris, uint32_t numTiles,
                                                   int sweep) {
  return smemTileOffset(numSpheres) + tileTris * sweepBytesPerTriangle(sweep) * (numTiles > 1 ? 2u : 1u);
}
// Streams tiles cyclically (0,1,..,n-1,0,1,..) through two buffers with TMA bulk copies.
// `consumed` counts tiles this CTA has swept so far; buffer = consumed & 1,
// mbarrier parity = (consumed >> 1) & 1.
struct TileStream {
  unsigned char *smemBase;
  const unsigned char *source; // tile-major array in HBM: triSweep (FP64) or triFilter (FP32)
  uint32_t tileBytes, numTiles, tileOffset;
  uint32_t consumed;

  __device__ __forceinline__ uint64_t *bar(uint32_t buffer) const {
    return reinterpret_cast<uint64_t *>(smemBase) + buffer;
  }
  __device__ __forceinline__ double4 *spheres() const {
    return reinterpret_cast<double4 *>(smemBase + 32);
  }
  __device__ __forceinline__ const unsigned char *tile(uint32_t buffer) const {
    return smemBase + tileOffset + buffer * tileBytes;
  }
  __device__ __forceinline__ void issue(uint32_t sequence) const { // one thread
    const uint32_t buffer = sequence & 1u;
    const uint32_t tileIndex = sequence % numTiles;
    mbarExpectTx(bar(buffer), tileBytes);
    tmaLoad1D(const_cast<unsigned char *>(tile(buffer)), source + static_cast<size_t>(tileIndex) * tileBytes,
              tileBytes, bar(buffer));
  }
  __device__ __forceinline__ void start() { // whole CTA, once
    consumed = 0;
    if (threadIdx.x == 0) {
      mbarInit(bar(0), 1);
      mbarInit(bar(1), 1);
      fenceBarrierInit();
    }
    __syncthreads();
    if (threadIdx.x == 0 && numTiles > 0) {
      issue(0);
      if (numTiles > 1)
        issue(1);
    }
  }
  __device__ __forceinline__ const unsigned char *acquire() const {
    mbarWait(bar(consumed & 1u), (consumed >> 1) & 1u);
    return tile(consumed & 1u);
  }
  __device__ __forceinline__ void release() { // whole CTA; only for numTiles > 1
    __syncthreads();
    if (threadIdx.x == 0)
      issue(consumed + 2);
    ++consumed;
  }
  __device__ __forceinline__ void drain() const { // numTiles > 1: two copies are still in flight
    mbarWait(bar(consumed & 1u), (consumed >> 1) & 1u);
    mbarWait(bar((consumed + 1) & 1u), ((consumed + 1) >> 1) & 1u);
  }
};

__device__ __forceinline__ TileStream makeTileStream(unsigned char *smemBase, const DeviceScene &scene,
                                                     int sweep) {
  return TileStream{smemBase,
                    sweep >= 8   ? reinterpret_cast<const unsigned char *>(scene.triExact)
                    : sweep == 7 ? reinterpret_cast<const unsigned char *>(scene.triMoment)
                    : sweep >= 2 ? reinterpret_cast<const unsigned char *>(scene.triFilter)
                               : reinterpret_cast<const unsigned char *>(scene.triSweep),
                    scene.tileTris * sweepBytesPerTriangle(sweep), scene.numTiles,
                    smemTileOffset(scene.numSpheres), 0};
}

// One tile of the sweep, whichever variant: `tile` is what the TileStream staged.
template <int kSweep, bool kFpWay = false>
__device__ __forceinline__ void sweepStagedTile(const DeviceScene &scene, const unsigned char *tile,
                                                uint32_t tileIndex, V3 o, V3 d, Nearest &best) {
  const int tileTris = static_cast<int>(scene.tileTris);
  const int first = static_cast<int>(tileIndex * scene.tileTris);
  if (kSweep == 7)
    sweepTileStage0Moment<kFpWay>(reinterpret_cast<const float *>(tile), scene.triExact + static_cast<size_t>(first) * 10,
                                  tileTris, first, o, d, best, scene.fanMask);
  else if (kSweep >= 5)
    sweepTileStage0Signs<kSweep == 5, kFpWay>(reinterpret_cast<const float *>(tile),
                    scene.triExact + static_cast<size_t>(first) * 10, tileTris, first, o, d, best);
  else if (kSweep >= 2)
    sweepTileStage0<kSweep >= 3, kSweep == 4, kFpWay>(reinterpret_cast<const float *>(tile),
                    scene.triExact + static_cast<size_t>(first) * 10, tileTris, tileTris, first, o, d, best);
  else if (kSweep == 1)
    sweepTilePrefiltered<kFpWay>(reinterpret_cast<const double *>(tile), tileTris, tileTris, first, o, d, best);
  else
    sweepTile<kFpWay>(reinterpret_cast<const double *>(tile), tileTris, tileTris, first, o, d, best);
}

// =============================================================================================
// Keyed (Philox) random numbers: counter = (pixel, sub-path, depth + 1 | 0 for the camera, call).
// =============================================================================================
struct KeyedDraws {
  uint32_t key0;
  __device__ __forceinline__ void camera(uint32_t pixel, double &a, double &b, double &c,
                                         double &d) const {
    const Philox8 w = philox4x32_10_pair(pixel, 0u, 0u, key0, kPhiloxKeyHigh);
    a = canonicalFromWords(w.w[0], w.w[1]);
    b = canonicalFromWords(w.w[2], w.w[3]);
    c = canonicalFromWords(w.w[4], w.w[5]);
    d = canonicalFromWords(w.w[6], w.w[7]);
  }
  // The (u, v, p) triple of the radiance() call at `depth` in sub-path `subPath`.
  __device__ __forceinline__ void bounce(uint32_t pixel, uint32_t subPath, uint32_t depth,
                                         double &u, double &v, double &p) const {
    const Philox8 w = philox4x32_10_pair(pixel, subPath, depth + 1u, key0, kPhiloxKeyHigh);
    u = canonicalFromWords(w.w[0], w.w[1]);
    v = canonicalFromWords(w.w[2], w.w[3]);
    p = canonicalFromWords(w.w[4], w.w[5]);
  }
};

// Camera::randomRay for the keyed policy; once per ~45 casts, so out of line.
static __device__ __noinline__ void keyedCameraRay(const DeviceCamera &camera, uint32_t key0, uint32_t pixel,
                                            int px, int py, V3 &origin, V3 &direction) {
  double ux, uy, ua, ur;
  KeyedDraws{key0}.camera(pixel, ux, uy, ua, ur);
  cameraRay(camera, px, py, ux, uy, ua, ur, origin, direction);
}

// A surface a bounce leaves from: what radiance() holds between its intersect() and its
// sampling loop (Scene.cpp:135-152).
struct Surface {
  V3 position, normal, incoming, basisX, basisY;
  double reflectivity;
  uint32_t material;
};

} // namespace ptb200
