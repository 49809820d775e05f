// extern "C" shim: the C ABI of include/ptb200.h on top of the sm_100a kernels.
//
// Host-side duties only: validate arguments, build the device scene layout (tile-major SoA,
// per-triangle shading normal), upload it once, launch kernels on one stream, time them with
// CUDA events on that stream, and copy the framebuffer back.  No rendering arithmetic runs on
// the host and there is no CPU fallback: every compute entry point needs a CUDA device.
#include "ptb200.h"

#include "pt_kernels.h"
#include "pt_mt19937.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

using namespace ptb200;

static_assert(sizeof(PtPixel) == sizeof(PtPixelDevice), "PtPixel layout");
static_assert(sizeof(PtHit) == sizeof(PtHitDevice), "PtHit layout");
static_assert(sizeof(PtCamera) == 18 * sizeof(double), "PtCamera layout");
static_assert(sizeof(DeviceCamera) == sizeof(PtCamera), "DeviceCamera layout");
static_assert(sizeof(PtMaterial) == 9 * sizeof(double), "PtMaterial layout");

namespace {

thread_local char gLastError[512] = "";

int fail(int code, const char *format, ...) {
  va_list args;
  va_start(args, format);
  vsnprintf(gLastError, sizeof gLastError, format, args);
  va_end(args);
  return code;
}

#define PT_CUDA(call)                                                                            \
  do {                                                                                           \
    const cudaError_t ptErr_ = (call);                                                           \
    if (ptErr_ != cudaSuccess)                                                                   \
      return fail(PTB200_ECUDA, "%s failed: %s", #call, cudaGetErrorString(ptErr_));             \
  } while (0)

// Tile plan.  Budget: two CTAs per SM must fit in 227 KB, each with its tile buffer(s) — up to
// 76 B/triangle, the moment-form stage-0 data of sweep variant 7 — plus its per-thread slots
// (36 KB of Surface slots in the sub-path kernel, 50 KB in the one-kernel form whose tiles are
// 56 B/triangle).
constexpr uint32_t kResidentTileTriangles = 976; // one resident tile per CTA up to here
constexpr uint32_t kStreamTileTriangles = 480;   // larger scenes: two buffers of <= this
// Per-pass sample buffer of the one-kernel forms (24 B per sample).  The sequential stream modes
// are parallel over passes only, so they want every pass of a call in one launch (up to 4 GiB:
// 1080p x 86 passes); the fp way is parallel over pixels and batches at 1 GiB.
constexpr size_t kSequentialSampleBufferBytes = size_t(4) << 30;
constexpr size_t kSampleBufferBytes = size_t(1) << 30;
// Keyed pipeline: records + strata terms of one batch of passes (pt_split.cu), ~530 B per sample
// at 4x4 strata.  Batches of ~2 M samples keep the persistent sub-path kernel's tail below 1 %.
constexpr size_t kSplitBufferBytes = size_t(1) << 30; // per buffer set; there are two

// ---- host restatement of the per-triangle values the reference derives in addTriangle ----
struct H3 {
  double x, y, z;
};
H3 hsub(H3 a, H3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
double hdot(H3 a, H3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
H3 hcross(H3 a, H3 b) {
  return {std::fma(a.y, b.z, -(a.z * b.y)), std::fma(a.z, b.x, -(a.x * b.z)),
          std::fma(a.x, b.y, -(a.y * b.x))};
}
H3 hnormalised(H3 a) {
  const double reciprocal = 1.0 / std::sqrt(hdot(a, a));
  return {a.x * reciprocal, a.y * reciprocal, a.z * reciprocal};
}
// Scene::addTriangle stores faceNormal() three times (Scene.cpp:181-187,
// TriangleVertices.h:33-35); intersectTriangles then evaluates
// normalised(u*(n1-n0) + v*(n2-n0) + n0) (Scene.cpp:99-107), which for n0==n1==n2 is
// normalised(0 + n0) whatever u and v are.
H3 shadingNormal(H3 e1, H3 e2) {
  const H3 face = hnormalised(hcross(e1, e2));
  return hnormalised(H3{0.0 + face.x, 0.0 + face.y, 0.0 + face.z});
}

// OrthoNormalBasis::fromZ (OrthoNormalBasis.cpp:36-51) with the arithmetic of basisFromZ() in
// pt_math.cuh, for the per-triangle bases stored next to the shading normal.
void basisFromZ(H3 z, H3 &xx, H3 &yy) {
  const H3 c = std::fabs(z.x) > 0.9999 ? H3{z.z, 0.0, -z.x} : H3{0.0, -z.z, z.y};
  xx = hnormalised(c);
  yy = hnormalised(hcross(z, xx));
}

// Owns the CUDA events of one call so that early error returns do not leak them.
struct EventList {
  std::vector<cudaEvent_t> events;
  EventList() = default;
  EventList(const EventList &) = delete;
  EventList &operator=(const EventList &) = delete;
  ~EventList() {
    for (cudaEvent_t e : events)
      cudaEventDestroy(e);
  }
  cudaError_t make(cudaEvent_t *out) {
    const cudaError_t err = cudaEventCreate(out);
    if (err == cudaSuccess)
      events.push_back(*out);
    return err;
  }
};

template <typename T>
struct DeviceBuffer {
  T *ptr{nullptr};
  size_t count{0};
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr)
      cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  cudaError_t ensure(size_t n) {
    if (n <= count && ptr)
      return cudaSuccess;
    release();
    const cudaError_t err = cudaMalloc(reinterpret_cast<void **>(&ptr), std::max<size_t>(n, 1) * sizeof(T));
    if (err == cudaSuccess)
      count = std::max<size_t>(n, 1);
    return err;
  }
};

} // namespace

struct PtContext {
  int device{0};
  int numSms{0};
  cudaStream_t stream{nullptr};
  bool haveScene{false};
  DeviceScene scene{};
  DeviceBuffer<double> triSweep;
  DeviceBuffer<double4> triShade;
  DeviceBuffer<double4> spheres;
  DeviceBuffer<uint32_t> sphereMaterial;
  DeviceBuffer<double> materials;
  DeviceBuffer<float> triFilter;
  DeviceBuffer<float> triMoment;
  DeviceBuffer<uint32_t> fanMask;
  MomentTable momentHost{};    // host copy of triMoment for scenes whose table rides in the kernel parameters
  bool momentHostValid{false};
  uint32_t fanGroups{0};       // groups of four triangles that are two quads in fan order (pt_device.cuh)
  DeviceBuffer<double> triExact;
  double sceneRadius{0};       // >= |p| for every vertex / sphere surface point
  double filterOriginBound{-1}; // origin bound the current triFilter contents were built for
  bool filterUsable{false};    // FP32 stage 0 allowed (coordinates comfortably inside FP32 range)
  DeviceBuffer<PtPixelDevice> accumulator;
  DeviceBuffer<double> samples;
  DeviceBuffer<uint32_t> mtHistory;          // fp way: scratch of the per-lane engines (pt_mt19937.cuh)
  // Keyed pipeline (pt_split.cu): batches alternate between two buffer sets on two streams, so that
  // the next batch's tracing kernels overlap this batch's tail and its resolve kernel.
  struct SplitSet {
    DeviceBuffer<double2> records;             // camera-ray hits of one batch
    DeviceBuffer<double> terms;                // one term per (sample, stratum)
    DeviceBuffer<uint8_t> sampleKind;          // strata terms / colour
    DeviceBuffer<unsigned long long> counters; // [0] ticket, [1] casts, [2] records of the batch
    cudaStream_t stream{nullptr};
    cudaEvent_t traced{nullptr}, resolved{nullptr};
  } split[2];
  cudaEvent_t splitReady{nullptr};
  uint32_t numMaterials{0};
  DeviceBuffer<unsigned long long> counters; // [0] ticket, [1] casts (one-kernel forms)
  int accWidth{0}, accHeight{0};
};

namespace {

int validateScene(const PtScene *scene) {
  if (!scene)
    return fail(PTB200_EINVAL, "scene is null");
  if (scene->numTriangles && (!scene->triangleVertices || !scene->triangleMaterial))
    return fail(PTB200_EINVAL, "triangle arrays are null");
  if (scene->numSpheres && (!scene->sphereCentreRadius || !scene->sphereMaterial))
    return fail(PTB200_EINVAL, "sphere arrays are null");
  if (!scene->materials || scene->numMaterials == 0)
    return fail(PTB200_EINVAL, "material palette is empty");
  if (scene->numMaterials > 65535)
    return fail(PTB200_EINVAL, "more than 65535 materials");
  for (uint32_t i = 0; i < scene->numTriangles; ++i)
    if (scene->triangleMaterial[i] >= scene->numMaterials)
      return fail(PTB200_EINVAL, "triangle %u: material index out of range", i);
  for (uint32_t i = 0; i < scene->numSpheres; ++i)
    if (scene->sphereMaterial[i] >= scene->numMaterials)
      return fail(PTB200_EINVAL, "sphere %u: material index out of range", i);
  if (static_cast<size_t>(scene->numSpheres) * 32 > 64 * 1024)
    return fail(PTB200_EINVAL, "more than 2048 spheres");
  return PTB200_OK;
}

int validateParams(const PtRenderParams *p, const PtRenderOptions *o) {
  if (!p)
    return fail(PTB200_EINVAL, "params is null");
  if (p->width <= 0 || p->height <= 0)
    return fail(PTB200_EINVAL, "image size %dx%d", p->width, p->height);
  if (p->samplesPerPixel < 0)
    return fail(PTB200_EINVAL, "negative samplesPerPixel");
  if (p->maxDepth > kMaxDepth)
    return fail(PTB200_EINVAL, "maxDepth %d exceeds the supported %d", p->maxDepth, kMaxDepth);
  if (p->firstBounceUSamples <= 0 || p->firstBounceVSamples <= 0)
    return fail(PTB200_EINVAL, "first-bounce sample counts must be positive");
  if (o) {
    if (o->rngMode != PTB200_RNG_KEYED_PHILOX && o->rngMode != PTB200_RNG_MT19937_SEQUENTIAL &&
        o->rngMode != PTB200_RNG_MT19937_PER_PIXEL && o->rngMode != PTB200_RNG_MT19937_SEQUENTIAL_OO)
      return fail(PTB200_EINVAL, "unknown rngMode %d", o->rngMode);
    if (o->rowStep < 0 || o->rowBegin < 0)
      return fail(PTB200_EINVAL, "negative row partition");
    if ((o->rngMode == PTB200_RNG_MT19937_SEQUENTIAL || o->rngMode == PTB200_RNG_MT19937_SEQUENTIAL_OO) &&
        (o->rowBegin != 0 || o->rowStep > 1))
      return fail(PTB200_EINVAL,
                  "the sequential mt19937 stream cannot be partitioned by rows; partition passes");
    if (o->lanesPerPass != 0 && o->lanesPerPass != 4 && o->lanesPerPass != 8 && o->lanesPerPass != 16 &&
        o->lanesPerPass != 32)
      return fail(PTB200_EINVAL, "lanesPerPass must be 0 (automatic), 4, 8, 16 or 32, not %d", o->lanesPerPass);
  }
  return PTB200_OK;
}

void planTiles(uint32_t numTriangles, uint32_t &tileTris, uint32_t &numTiles) {
  if (numTriangles == 0) {
    tileTris = 0;
    numTiles = 0;
    return;
  }
  numTiles = numTriangles <= kResidentTileTriangles
                 ? 1u
                 : (numTriangles + kStreamTileTriangles - 1) / kStreamTileTriangles;
  tileTris = (numTriangles + numTiles - 1) / numTiles;
  tileTris = (tileTris + 3u) & ~3u; // 16-byte loads of 2 doubles / 4 floats
  numTiles = (numTriangles + tileTris - 1) / tileTris;
}

DeviceCamera toDeviceCamera(const PtCamera &c) {
  DeviceCamera d;
  std::memcpy(&d, &c, sizeof d);
  return d;
}



// One idle context per device is kept between one-shot calls so that repeated ptb200_render()
// calls reuse the stream and the device allocations (the scene is still uploaded every call).
std::mutex gPoolMutex;
std::vector<PtContext *> gPool;

PtContext *poolTake(int device) {
  std::lock_guard<std::mutex> lock(gPoolMutex);
  for (size_t i = 0; i < gPool.size(); ++i) {
    if (gPool[i]->device == device) {
      PtContext *ctx = gPool[i];
      gPool.erase(gPool.begin() + static_cast<long>(i));
      return ctx;
    }
  }
  return nullptr;
}

void poolGive(PtContext *ctx) {
  std::lock_guard<std::mutex> lock(gPoolMutex);
  gPool.push_back(ctx);
}

} // namespace

extern "C" {

const char *ptb200_last_error(void) { return gLastError; }

int ptb200_device_count(int32_t *count) {
  if (!count)
    return fail(PTB200_EINVAL, "count is null");
  int n = 0;
  const cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  *count = n;
  return PTB200_OK;
}

int ptb200_context_create(int32_t device, PtContext **out) {
  if (!out)
    return fail(PTB200_EINVAL, "out is null");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PTB200_ECUDA, "no CUDA device available (this backend has no CPU fallback)");
  }
  if (device < 0 || device >= n)
    return fail(PTB200_EINVAL, "device %d out of range (have %d)", device, n);
  PT_CUDA(cudaSetDevice(device));
  auto *ctx = new (std::nothrow) PtContext;
  if (!ctx)
    return fail(PTB200_ENOMEM, "out of host memory");
  ctx->device = device;
  cudaDeviceProp prop{};
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return fail(PTB200_ECUDA, "cudaGetDeviceProperties failed");
  }
  ctx->numSms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    return fail(PTB200_ECUDA, "cudaStreamCreate failed");
  }
  bool ok = ctx->counters.ensure(3) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->splitReady, cudaEventDisableTiming) == cudaSuccess;
  for (auto &set : ctx->split)
    ok = ok && set.counters.ensure(3) == cudaSuccess &&
         cudaStreamCreateWithFlags(&set.stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&set.traced, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&set.resolved, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    ptb200_context_destroy(ctx);
    return fail(PTB200_ENOMEM, "device allocation failed");
  }
  *out = ctx;
  return PTB200_OK;
}

void ptb200_context_destroy(PtContext *ctx) {
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  for (auto &set : ctx->split) {
    if (set.stream) {
      cudaStreamSynchronize(set.stream);
      cudaStreamDestroy(set.stream);
    }
    if (set.traced)
      cudaEventDestroy(set.traced);
    if (set.resolved)
      cudaEventDestroy(set.resolved);
  }
  if (ctx->splitReady)
    cudaEventDestroy(ctx->splitReady);
  if (ctx->stream) {
    cudaStreamSynchronize(ctx->stream);
    cudaStreamDestroy(ctx->stream);
  }
  delete ctx;
}

int ptb200_context_upload_scene(PtContext *ctx, const PtScene *scene) {
  if (!ctx)
    return fail(PTB200_EINVAL, "context is null");
  if (const int rc = validateScene(scene))
    return rc;
  PT_CUDA(cudaSetDevice(ctx->device));
  DeviceScene d{};
  d.numTriangles = scene->numTriangles;
  d.numSpheres = scene->numSpheres;
  planTiles(scene->numTriangles, d.tileTris, d.numTiles);

  const size_t sweepDoubles = static_cast<size_t>(d.numTiles) * 9 * d.tileTris;
  std::vector<double> sweep(sweepDoubles, 0.0);
  std::vector<double> exact(static_cast<size_t>(d.numTiles) * d.tileTris * 10, 0.0);
  std::vector<double4> shade(static_cast<size_t>(scene->numTriangles) * 4);
  for (uint32_t i = 0; i < scene->numTriangles; ++i) {
    const double *t = scene->triangleVertices + 9 * static_cast<size_t>(i);
    const H3 v0{t[0], t[1], t[2]}, v1{t[3], t[4], t[5]}, v2{t[6], t[7], t[8]};
    const H3 e1 = hsub(v1, v0), e2 = hsub(v2, v0);
    const uint32_t tile = i / d.tileTris, within = i % d.tileTris;
    double *base = sweep.data() + static_cast<size_t>(tile) * 9 * d.tileTris + within;
    const double values[9] = {v0.x, v0.y, v0.z, e1.x, e1.y, e1.z, e2.x, e2.y, e2.z};
    for (int a = 0; a < 9; ++a) {
      base[static_cast<size_t>(a) * d.tileTris] = values[a];
      exact[10 * static_cast<size_t>(i) + a] = values[a];
    }
    const H3 n = shadingNormal(e1, e2);
    H3 fx, fy, bx, by;
    basisFromZ(n, fx, fy);
    basisFromZ(H3{-n.x, -n.y, -n.z}, bx, by);
    double4 *record = shade.data() + 4 * static_cast<size_t>(i);
    record[0] = make_double4(n.x, n.y, n.z, static_cast<double>(scene->triangleMaterial[i]));
    record[1] = make_double4(fx.x, fx.y, fx.z, fy.x);
    record[2] = make_double4(fy.y, fy.z, bx.x, bx.y);
    record[3] = make_double4(bx.z, by.x, by.y, by.z);
  }
  std::vector<double4> spheres(scene->numSpheres);
  double radius = 0, longestEdge = 0;
  for (uint32_t i = 0; i < scene->numSpheres; ++i) {
    const double *s = scene->sphereCentreRadius + 4 * static_cast<size_t>(i);
    spheres[i] = make_double4(s[0], s[1], s[2], s[3] * s[3]); // Sphere.h:11
    radius = std::max(radius, std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) + std::fabs(s[3]));
  }
  for (uint32_t i = 0; i < scene->numTriangles; ++i) {
    const double *t = scene->triangleVertices + 9 * static_cast<size_t>(i);
    for (int k = 0; k < 3; ++k)
      radius = std::max(radius, std::sqrt(t[3 * k] * t[3 * k] + t[3 * k + 1] * t[3 * k + 1] + t[3 * k + 2] * t[3 * k + 2]));
    for (int k = 1; k < 3; ++k) {
      const double ex = t[3 * k] - t[0], ey = t[3 * k + 1] - t[1], ez = t[3 * k + 2] - t[2];
      longestEdge = std::max(longestEdge, std::sqrt(ex * ex + ey * ey + ez * ez));
    }
  }
  // Groups of four consecutive triangles (of a tile) that are two quads in fan order: (v0, e1, e2),
  // (v0, e2, e3) twice, compared as stored doubles (all-zero padding qualifies too).
  const uint32_t numGroups = d.numTiles * (d.tileTris / 4);
  std::vector<uint32_t> fanMask(numGroups / 32 + 2, 0u);
  if (!std::getenv("PTB200_NO_FAN_GROUPS")) {
    auto value = [&](uint32_t slot, int k) { return exact[10 * static_cast<size_t>(slot) + k]; };
    auto fanPair = [&](uint32_t a, uint32_t b) {
      for (int k = 0; k < 3; ++k)
        if (value(a, k) != value(b, k) || value(a, 6 + k) != value(b, 3 + k)) // same v0; e2 of A == e1 of B
          return false;
      return true;
    };
    for (uint32_t g = 0; g < numGroups; ++g)
      if (fanPair(4 * g, 4 * g + 1) && fanPair(4 * g + 2, 4 * g + 3))
        fanMask[g >> 5] |= 1u << (g & 31u);
  }
  ctx->fanGroups = fanMask[0] & 0xffffu;
  ctx->sceneRadius = radius;
  ctx->filterUsable = std::isfinite(radius) && radius < 1e6 && longestEdge < 1e6;
  ctx->filterOriginBound = -1;
  ctx->momentHostValid = false;

  PT_CUDA(ctx->triSweep.ensure(sweepDoubles));
  PT_CUDA(ctx->triShade.ensure(shade.size()));
  PT_CUDA(ctx->triFilter.ensure(static_cast<size_t>(d.numTiles) * 14 * d.tileTris));
  PT_CUDA(ctx->triMoment.ensure(static_cast<size_t>(d.numTiles) * 19 * d.tileTris));
  PT_CUDA(ctx->fanMask.ensure(fanMask.size()));
  PT_CUDA(ctx->triExact.ensure(exact.size()));
  PT_CUDA(ctx->spheres.ensure(spheres.size()));
  PT_CUDA(ctx->sphereMaterial.ensure(scene->numSpheres));
  PT_CUDA(ctx->materials.ensure(static_cast<size_t>(scene->numMaterials) * 10));
  std::vector<double> materials(static_cast<size_t>(scene->numMaterials) * 10);
  for (uint32_t i = 0; i < scene->numMaterials; ++i) {
    std::memcpy(&materials[10 * static_cast<size_t>(i)], &scene->materials[i], 72);
    materials[10 * static_cast<size_t>(i) + 9] = 1.0 / scene->materials[i].indexOfRefraction;
  }
  if (sweepDoubles) {
    PT_CUDA(cudaMemcpyAsync(ctx->triSweep.ptr, sweep.data(), sweepDoubles * 8, cudaMemcpyHostToDevice, ctx->stream));
    PT_CUDA(cudaMemcpyAsync(ctx->triExact.ptr, exact.data(), exact.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  }
  PT_CUDA(cudaMemcpyAsync(ctx->fanMask.ptr, fanMask.data(), fanMask.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!shade.empty())
    PT_CUDA(cudaMemcpyAsync(ctx->triShade.ptr, shade.data(), shade.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (!spheres.empty()) {
    PT_CUDA(cudaMemcpyAsync(ctx->spheres.ptr, spheres.data(), spheres.size() * 32, cudaMemcpyHostToDevice, ctx->stream));
    PT_CUDA(cudaMemcpyAsync(ctx->sphereMaterial.ptr, scene->sphereMaterial, scene->numSpheres * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  PT_CUDA(cudaMemcpyAsync(ctx->materials.ptr, materials.data(), materials.size() * 8,
                          cudaMemcpyHostToDevice, ctx->stream));
  PT_CUDA(cudaStreamSynchronize(ctx->stream)); // staging vectors go out of scope
  d.triSweep = ctx->triSweep.ptr;
  d.triShade = ctx->triShade.ptr;
  d.spheres = ctx->spheres.ptr;
  d.sphereMaterial = ctx->sphereMaterial.ptr;
  d.materials = ctx->materials.ptr;
  d.triFilter = ctx->triFilter.ptr;
  d.triMoment = ctx->triMoment.ptr;
  d.fanMask = ctx->fanMask.ptr;
  d.triExact = ctx->triExact.ptr;
  d.environment[0] = scene->environment[0];
  d.environment[1] = scene->environment[1];
  d.environment[2] = scene->environment[2];
  ctx->scene = d;
  ctx->numMaterials = scene->numMaterials;
  ctx->haveScene = true;
  return PTB200_OK;
}

// (Re)builds the FP32 stage-0 arrays for ray origins within `originBound` of the world origin.
static int ensureFilter(PtContext *ctx, double originBound, uint64_t *launches) {
  if (ctx->filterOriginBound >= originBound)
    return PTB200_OK;
  BuildFilterArgs b{};
  b.scene = ctx->scene;
  b.out = ctx->triFilter.ptr;
  b.outMoment = ctx->triMoment.ptr;
  b.originBound = originBound;
  PT_CUDA(launchBuildFilter(b, ctx->stream));
  ctx->filterOriginBound = originBound;
  ctx->momentHostValid = false;
  if (constTableFits(ctx->scene.numTriangles, ctx->scene.numTiles)) { // sweep variant 9: table into constant bank 0
    std::memset(&ctx->momentHost, 0, sizeof ctx->momentHost);
    PT_CUDA(cudaMemcpyAsync(&ctx->momentHost, ctx->triMoment.ptr, static_cast<size_t>(ctx->scene.tileTris) * 19 * sizeof(float),
                            cudaMemcpyDeviceToHost, ctx->stream));
    PT_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->momentHost.fanGroups = ctx->fanGroups; // (buildFilterKernel already stored those groups as [A0, A1, B0, B1])
    ctx->momentHostValid = true;
  }
  if (launches && ctx->scene.numTiles)
    *launches += 1;
  return PTB200_OK;
}

// Launches the batches of one render call on ctx->stream; does not synchronise.
static int enqueueRender(PtContext *ctx, const PtCamera *camera, const PtRenderParams *params,
                         const PtRenderOptions *options, int passBegin, int numPasses,
                         EventList *owner, std::vector<cudaEvent_t> *events, uint64_t *launches) {
  const PtRenderOptions defaults{};
  const PtRenderOptions &opt = options ? *options : defaults;
  const int rowStep = opt.rowStep > 0 ? opt.rowStep : 1;
  const int rowBegin = opt.rowBegin;
  const uint32_t ownRows = rowBegin >= params->height
                               ? 0u
                               : static_cast<uint32_t>((params->height - rowBegin + rowStep - 1) / rowStep);
  const uint32_t ownPixels = ownRows * static_cast<uint32_t>(params->width);
  if (ownPixels == 0 || numPasses == 0)
    return PTB200_OK;
  const bool ooWay = opt.rngMode == PTB200_RNG_MT19937_SEQUENTIAL_OO;
  const bool sequential = opt.rngMode == PTB200_RNG_MT19937_SEQUENTIAL || ooWay;
  const bool fpWay = opt.rngMode == PTB200_RNG_MT19937_PER_PIXEL;
  int keyedConfig = chooseKeyedConfig(ctx->scene.numTriangles, ctx->filterUsable, fpWay ? 1 : 0);
  if (fpWay) { // the megakernel's instantiations for the per-lane engines
    const int sweep = keyedConfig % 10, shape = (keyedConfig % 100) / 10;
    keyedConfig = sweep == 6 ? (shape == 2 ? 26 : 6) : 1;
  }
  const bool split = !sequential && !fpWay && keyedConfig >= 100;
  const uint64_t numSub = static_cast<uint64_t>(params->firstBounceUSamples) * static_cast<uint64_t>(params->firstBounceVSamples);
  const size_t pixelsPerPass = sequential ? static_cast<size_t>(params->width) * params->height : ownPixels;
  size_t passesPerBatch;
  if (split) {
    if (numSub * ownPixels >= (uint64_t(1) << 31))
      return fail(PTB200_EINVAL, "%llu strata x %u pixels exceed the sub-path index range",
                  static_cast<unsigned long long>(numSub), ownPixels);
    const size_t perSample = splitBytesPerSample(static_cast<uint32_t>(numSub));
    passesPerBatch = std::max<size_t>(1, kSplitBufferBytes / (pixelsPerPass * perSample));
    passesPerBatch = std::min<size_t>(passesPerBatch, ((uint64_t(1) << 31) - 1) / (numSub * ownPixels));
  } else {
    passesPerBatch = std::max<size_t>(1, (sequential ? kSequentialSampleBufferBytes : kSampleBufferBytes) / (pixelsPerPass * 24));
  }
  if (opt.passesPerBatch > 0)
    passesPerBatch = std::min<size_t>(passesPerBatch, static_cast<size_t>(opt.passesPerBatch));
  passesPerBatch = std::min<size_t>(passesPerBatch, static_cast<size_t>(numPasses));
  if (split) {
    const size_t samples = passesPerBatch * pixelsPerPass;
    const int sets = static_cast<size_t>(numPasses) > passesPerBatch ? 2 : 1;
    for (int k = 0; k < sets; ++k) {
      PT_CUDA(ctx->split[k].records.ensure(samples * 9));
      PT_CUDA(ctx->split[k].terms.ensure(samples * numSub * 3));
      PT_CUDA(ctx->split[k].sampleKind.ensure(samples));
    }
  } else {
    PT_CUDA(ctx->samples.ensure(passesPerBatch * pixelsPerPass * 3));
  }
  size_t mtThreads = 0;
  uint32_t mtLimit = 0;
  if (fpWay) {
    mtThreads = mtHistoryThreadsFor(ctx->numSms);
    PT_CUDA(ctx->mtHistory.ensure(mtThreads * kMtHistoryStride));
    // An engine draws the camera's words, then three doubles per stratum at depth 0 and per
    // sub-path and level below it.
    const uint64_t strata = static_cast<uint64_t>(params->firstBounceUSamples) * static_cast<uint64_t>(params->firstBounceVSamples);
    mtLimit = mtStoreLimit(8u + 6u * strata * static_cast<uint64_t>(std::max(params->maxDepth, 0)));
  }
  if (!sequential && keyedConfig % 10 >= 2) {
    if (!ctx->filterUsable)
      return fail(PTB200_EINVAL, "scene coordinates exceed the range the FP32 stage-0 sweep supports; "
                                 "select a FP64 sweep (PTB200_KEYED_CONFIG=1)");
    // Ray origins are the camera (centre + aperture disc) or points on primitives.
    const double cameraReach = std::sqrt(camera->centre[0] * camera->centre[0] + camera->centre[1] * camera->centre[1] +
                                         camera->centre[2] * camera->centre[2]) + std::fabs(camera->apertureRadius);
    const double bound = std::max(ctx->sceneRadius, cameraReach) * (1.0 + 1e-6);
    if (const int rc = ensureFilter(ctx, bound, launches))
      return rc;
  }

  int splitBatch = 0;
  for (int done = 0; done < numPasses;) {
    const int batch = static_cast<int>(std::min<size_t>(passesPerBatch, static_cast<size_t>(numPasses - done)));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (events) {
      PT_CUDA(owner->make(&e0));
      PT_CUDA(owner->make(&e1));
      PT_CUDA(cudaEventRecord(e0, ctx->stream));
    }
    if (sequential) {
      SequentialArgs a{};
      a.scene = ctx->scene;
      a.camera = toDeviceCamera(*camera);
      a.width = params->width;
      a.height = params->height;
      a.seed = params->seed;
      a.passBegin = passBegin + done;
      a.numPasses = batch;
      a.maxDepth = params->maxDepth;
      a.firstBounceU = params->firstBounceUSamples;
      a.firstBounceV = params->firstBounceVSamples;
      a.preview = params->preview;
      a.way = ooWay ? 2 : 0;
      a.samples = ctx->samples.ptr;
      a.castCounter = ctx->counters.ptr + 1;
      PT_CUDA(launchRenderSequential(a, sequentialLanesPerPass(batch, ctx->numSms, opt.lanesPerPass), ctx->stream));
    } else if (split) {
      // Batch b traces on the stream of buffer set b % 2 — after the resolve that last read the set
      // — and resolves on ctx->stream, in batch order: passes are added in pass order.
      PtContext::SplitSet &set = ctx->split[splitBatch & 1];
      if (splitBatch == 0)
        PT_CUDA(cudaEventRecord(ctx->splitReady, ctx->stream)); // scene, filter, accumulator are ready
      PT_CUDA(cudaStreamWaitEvent(set.stream, splitBatch < 2 ? ctx->splitReady : set.resolved, 0));
      PT_CUDA(cudaMemsetAsync(set.counters.ptr, 0, sizeof(unsigned long long), set.stream));     // ticket
      PT_CUDA(cudaMemsetAsync(set.counters.ptr + 2, 0, sizeof(unsigned long long), set.stream)); // records
      SplitArgs a{};
      a.scene = ctx->scene;
      a.camera = toDeviceCamera(*camera);
      a.width = static_cast<uint32_t>(params->width);
      a.height = static_cast<uint32_t>(params->height);
      a.rowBegin = rowBegin;
      a.rowStep = rowStep;
      a.ownPixels = ownPixels;
      a.totalSamples = ownPixels * static_cast<uint32_t>(batch);
      a.numPasses = static_cast<uint32_t>(batch);
      a.numSub = static_cast<uint32_t>(numSub);
      a.numSubShift = (numSub & (numSub - 1)) == 0 ? __builtin_ctzll(numSub) : -1;
      a.firstBounceVShift = (params->firstBounceVSamples & (params->firstBounceVSamples - 1)) == 0
                                ? __builtin_ctz(static_cast<unsigned>(params->firstBounceVSamples))
                                : -1;
      a.numMaterials = ctx->numMaterials;
      a.seed = params->seed;
      a.passBegin = passBegin + done;
      a.maxDepth = params->maxDepth;
      a.firstBounceU = params->firstBounceUSamples;
      a.firstBounceV = params->firstBounceVSamples;
      a.preview = params->preview;
      a.firstBounceUPow2 = (a.firstBounceU & (a.firstBounceU - 1)) == 0;
      a.firstBounceVPow2 = (a.firstBounceV & (a.firstBounceV - 1)) == 0;
      a.invFirstBounceU = 1.0 / static_cast<double>(a.firstBounceU);
      a.invFirstBounceV = 1.0 / static_cast<double>(a.firstBounceV);
      a.records = set.records.ptr;
      a.terms = set.terms.ptr;
      a.sampleKind = set.sampleKind.ptr;
      a.counters = set.counters.ptr;
      a.accumulator = ctx->accumulator.ptr;
      if (keyedConfig % 10 >= 8) {
        if (!ctx->momentHostValid)
          return fail(PTB200_EINVAL, "sweep variants 8/9 need a scene of at most 64 triangles");
        a.momentTable = ctx->momentHost;
      }
      PT_CUDA(launchSplitTrace(a, ctx->numSms, keyedConfig, set.stream));
      PT_CUDA(cudaEventRecord(set.traced, set.stream));
      PT_CUDA(cudaStreamWaitEvent(ctx->stream, set.traced, 0));
      PT_CUDA(launchSplitResolve(a, ctx->stream));
      PT_CUDA(cudaEventRecord(set.resolved, ctx->stream));
      if (events) { // the three kernels of the pipeline are the path-tracing time
        PT_CUDA(cudaEventRecord(e1, ctx->stream));
        events->push_back(e0);
        events->push_back(e1);
      }
      if (launches)
        *launches += 3;
      ++splitBatch;
      done += batch;
      continue;
    } else {
      PT_CUDA(cudaMemsetAsync(ctx->counters.ptr, 0, sizeof(unsigned long long), ctx->stream));
      KeyedArgs a{};
      a.scene = ctx->scene;
      a.camera = toDeviceCamera(*camera);
      a.width = static_cast<uint32_t>(params->width);
      a.height = static_cast<uint32_t>(params->height);
      a.way = fpWay ? 1 : 0;
      a.rowBegin = rowBegin;
      a.rowStep = rowStep;
      a.ownPixels = ownPixels;
      a.totalItems = static_cast<unsigned long long>(ownPixels) * static_cast<unsigned long long>(batch);
      a.seed = params->seed;
      a.passBegin = passBegin + done;
      a.maxDepth = params->maxDepth;
      a.firstBounceU = params->firstBounceUSamples;
      a.firstBounceV = params->firstBounceVSamples;
      a.preview = params->preview;
      a.firstBounceUPow2 = (a.firstBounceU & (a.firstBounceU - 1)) == 0;
      a.firstBounceVPow2 = (a.firstBounceV & (a.firstBounceV - 1)) == 0;
      a.invFirstBounceU = 1.0 / static_cast<double>(a.firstBounceU);
      a.invFirstBounceV = 1.0 / static_cast<double>(a.firstBounceV);
      a.samples = ctx->samples.ptr;
      a.mtHistory = ctx->mtHistory.ptr;
      a.mtHistoryThreads = mtThreads;
      a.mtStoreLimit = mtLimit;
      a.ticket = ctx->counters.ptr;
      a.castCounter = ctx->counters.ptr + 1;
      PT_CUDA(launchRenderKeyed(a, ctx->numSms, keyedConfig, ctx->stream));
    }
    if (events) {
      PT_CUDA(cudaEventRecord(e1, ctx->stream));
      events->push_back(e0);
      events->push_back(e1);
    }
    ReduceArgs r{};
    r.samples = ctx->samples.ptr;
    r.accumulator = ctx->accumulator.ptr;
    r.width = static_cast<uint32_t>(params->width);
    r.rowBegin = static_cast<uint32_t>(rowBegin);
    r.rowStep = static_cast<uint32_t>(rowStep);
    r.ownPixels = ownPixels;
    r.numPasses = static_cast<uint32_t>(batch);
    r.samplePassStride = pixelsPerPass;
    r.samplesAreFullFrame = sequential ? 1 : 0;
    PT_CUDA(launchReducePasses(r, ctx->stream));
    if (launches)
      *launches += 2;
    done += batch;
  }
  return PTB200_OK;
}

int ptb200_context_render(PtContext *ctx, const PtCamera *camera, const PtRenderParams *params,
                          const PtRenderOptions *options, int32_t accumulate, PtStats *stats) {
  if (!ctx || !camera)
    return fail(PTB200_EINVAL, "null argument");
  if (const int rc = validateParams(params, options))
    return rc;
  if (!ctx->haveScene)
    return fail(PTB200_ESTATE, "no scene uploaded");
  PT_CUDA(cudaSetDevice(ctx->device));
  const size_t pixels = static_cast<size_t>(params->width) * params->height;
  const bool sameSize = ctx->accWidth == params->width && ctx->accHeight == params->height;
  PT_CUDA(ctx->accumulator.ensure(pixels));
  if (!accumulate || !sameSize)
    PT_CUDA(cudaMemsetAsync(ctx->accumulator.ptr, 0, pixels * sizeof(PtPixelDevice), ctx->stream));
  ctx->accWidth = params->width;
  ctx->accHeight = params->height;
  PT_CUDA(cudaMemsetAsync(ctx->counters.ptr + 1, 0, sizeof(unsigned long long), ctx->stream));
  for (auto &set : ctx->split) // the cast counters of this call (the sets' streams start after ctx->stream's event)
    PT_CUDA(cudaMemsetAsync(set.counters.ptr + 1, 0, sizeof(unsigned long long), ctx->stream));

  EventList owned;
  cudaEvent_t begin = nullptr, end = nullptr;
  PT_CUDA(owned.make(&begin));
  PT_CUDA(owned.make(&end));
  std::vector<cudaEvent_t> events; // (start, stop) of every path-tracing kernel, owned by `owned`
  uint64_t launches = 0;
  PT_CUDA(cudaEventRecord(begin, ctx->stream));
  const int rc = enqueueRender(ctx, camera, params, options, options ? options->passBegin : 0,
                               params->samplesPerPixel, &owned, &events, &launches);
  if (rc == PTB200_OK) {
    PT_CUDA(cudaEventRecord(end, ctx->stream));
    PT_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (rc == PTB200_OK && stats) {
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, begin, end));
    double sweepMs = 0;
    for (size_t i = 0; i + 1 < events.size(); i += 2) {
      float part = 0;
      PT_CUDA(cudaEventElapsedTime(&part, events[i], events[i + 1]));
      sweepMs += part;
    }
    unsigned long long casts = 0;
    for (const unsigned long long *counter : {ctx->counters.ptr + 1, ctx->split[0].counters.ptr + 1, ctx->split[1].counters.ptr + 1}) {
      unsigned long long part = 0;
      PT_CUDA(cudaMemcpy(&part, counter, sizeof part, cudaMemcpyDeviceToHost));
      casts += part;
    }
    const PtRenderOptions defaults{};
    const PtRenderOptions &opt = options ? *options : defaults;
    const int rowStep = opt.rowStep > 0 ? opt.rowStep : 1;
    const uint64_t ownRows = opt.rowBegin >= params->height
                                 ? 0
                                 : static_cast<uint64_t>((params->height - opt.rowBegin + rowStep - 1) / rowStep);
    stats->samples = ownRows * static_cast<uint64_t>(params->width) * static_cast<uint64_t>(params->samplesPerPixel);
    stats->casts = casts;
    stats->kernelLaunches = launches;
    stats->kernelMs = ms;
    stats->sweepKernelMs = sweepMs;
  }
  return rc;
}

int ptb200_context_download(PtContext *ctx, PtPixel *out) {
  if (!ctx || !out)
    return fail(PTB200_EINVAL, "null argument");
  if (!ctx->accumulator.ptr || ctx->accWidth == 0)
    return fail(PTB200_ESTATE, "nothing rendered yet");
  PT_CUDA(cudaSetDevice(ctx->device));
  const size_t pixels = static_cast<size_t>(ctx->accWidth) * ctx->accHeight;
  PT_CUDA(cudaMemcpyAsync(out, ctx->accumulator.ptr, pixels * sizeof(PtPixel), cudaMemcpyDeviceToHost, ctx->stream));
  PT_CUDA(cudaStreamSynchronize(ctx->stream));
  return PTB200_OK;
}

// D2H of the rows a row-partitioned call rendered (every row when the call was not partitioned):
// one strided copy, rows y = rowBegin + k * rowStep; the other rows of `out` are not written.
static int downloadRows(PtContext *ctx, PtPixel *out, const PtRenderOptions *options) {
  if (!ctx->accumulator.ptr || ctx->accWidth == 0)
    return fail(PTB200_ESTATE, "nothing rendered yet");
  const int rowStep = options && options->rowStep > 0 ? options->rowStep : 1;
  const int rowBegin = options ? options->rowBegin : 0;
  if (rowBegin >= ctx->accHeight)
    return PTB200_OK;
  const size_t rowBytes = static_cast<size_t>(ctx->accWidth) * sizeof(PtPixel);
  const size_t ownRows = static_cast<size_t>((ctx->accHeight - rowBegin + rowStep - 1) / rowStep);
  const size_t first = static_cast<size_t>(rowBegin) * ctx->accWidth;
  PT_CUDA(cudaMemcpy2DAsync(out + first, rowBytes * rowStep, ctx->accumulator.ptr + first, rowBytes * rowStep,
                            rowBytes, ownRows, cudaMemcpyDeviceToHost, ctx->stream));
  PT_CUDA(cudaStreamSynchronize(ctx->stream));
  return PTB200_OK;
}

int ptb200_render(const PtScene *scene, const PtCamera *camera, const PtRenderParams *params,
                  const PtRenderOptions *options, PtPixel *out, PtProgressFn progress, void *user,
                  PtStats *stats) {
  if (!camera || !out)
    return fail(PTB200_EINVAL, "null argument");
  if (const int rc = validateParams(params, options))
    return rc;
  const int device = options ? options->device : 0;
  PtContext *ctx = poolTake(device);
  int rc = PTB200_OK;
  if (!ctx) {
    rc = ptb200_context_create(device, &ctx);
    if (rc)
      return rc;
  }
  rc = ptb200_context_upload_scene(ctx, scene);
  PtStats total{};
  if (rc == PTB200_OK) {
    if (!progress || params->samplesPerPixel == 0) {
      // (spp == 0 still sizes and zeroes the accumulator: the download below must not see the
      // previous call's frame of a pooled context)
      rc = ptb200_context_render(ctx, camera, params, options, 0, &total);
    } else {
      // Progressive: render in slices of passes, handing the partial framebuffer to the
      // caller on this thread between slices (Scene.cpp:242-245 does so per collected pass).
      PtRenderOptions slice = options ? *options : PtRenderOptions{};
      const int spp = params->samplesPerPixel;
      const bool passParallelOnly = slice.rngMode == PTB200_RNG_MT19937_SEQUENTIAL ||
                                    slice.rngMode == PTB200_RNG_MT19937_SEQUENTIAL_OO;
      // The sequential stream modes are parallel over passes only (one warp per pass): slicing
      // them would leave the device idle, so they render in one slice unless the caller asks.
      const int step = slice.passesPerBatch > 0 ? slice.passesPerBatch
                       : passParallelOnly       ? spp
                                                : std::max(1, (spp + 19) / 20);
      const int firstPass = slice.passBegin;
      for (int done = 0; done < spp && rc == PTB200_OK;) {
        PtRenderParams part = *params;
        part.samplesPerPixel = std::min(step, spp - done);
        slice.passBegin = firstPass + done;
        PtStats one{};
        rc = ptb200_context_render(ctx, camera, &part, &slice, done > 0, &one);
        if (rc)
          break;
        total.samples += one.samples;
        total.casts += one.casts;
        total.kernelLaunches += one.kernelLaunches;
        total.kernelMs += one.kernelMs;
        total.sweepKernelMs += one.sweepKernelMs;
        done += part.samplesPerPixel;
        rc = downloadRows(ctx, out, options);
        if (rc == PTB200_OK && progress(user, out, done, spp) != 0)
          break;
      }
    }
  }
  if (rc == PTB200_OK && (ctx->accWidth != params->width || ctx->accHeight != params->height))
    rc = fail(PTB200_ESTATE, "accumulator is %dx%d, expected %dx%d", ctx->accWidth, ctx->accHeight,
              params->width, params->height);
  if (rc == PTB200_OK)
    rc = downloadRows(ctx, out, options);
  if (rc == PTB200_OK)
    poolGive(ctx);
  else
    ptb200_context_destroy(ctx);
  if (rc == PTB200_OK && stats)
    *stats = total;
  return rc;
}

int ptb200_render_multi(const PtScene *scene, const PtCamera *camera, const PtRenderParams *params,
                        const PtRenderOptions *options, const int32_t *devices, int32_t numDevices,
                        PtPixel *out, PtStats *stats) {
  return ptb200_render_multi_progress(scene, camera, params, options, devices, numDevices, out, nullptr, nullptr, stats);
}

int ptb200_render_multi_progress(const PtScene *scene, const PtCamera *camera, const PtRenderParams *params,
                                 const PtRenderOptions *options, const int32_t *devices, int32_t numDevices,
                                 PtPixel *out, PtProgressFn progress, void *user, PtStats *stats) {
  if (!camera || !out)
    return fail(PTB200_EINVAL, "null argument");
  if (const int rc = validateParams(params, options))
    return rc;
  std::vector<int32_t> list;
  if (devices && numDevices > 0) {
    list.assign(devices, devices + numDevices);
  } else {
    int32_t n = 0;
    ptb200_device_count(&n);
    for (int32_t i = 0; i < n; ++i)
      list.push_back(i);
  }
  if (list.empty())
    return fail(PTB200_ECUDA, "no CUDA device available (this backend has no CPU fallback)");
  const PtRenderOptions base = options ? *options : PtRenderOptions{};
  if (base.rowStep > 1 || base.rowBegin != 0)
    return fail(PTB200_EINVAL, "render_multi partitions rows itself; leave rowBegin/rowStep zero");
  const int n = static_cast<int>(list.size());
  const size_t pixels = static_cast<size_t>(params->width) * params->height;
  const int spp = params->samplesPerPixel;
  const bool sequential = base.rngMode == PTB200_RNG_MT19937_SEQUENTIAL ||
                          base.rngMode == PTB200_RNG_MT19937_SEQUENTIAL_OO;
  PtStats total{};

  if (sequential) {
    // One engine per pass walks the whole frame: only passes can be shared out (contiguous blocks,
    // one per device), each device renders a full frame and the frames are added in device order —
    // the operator+= the reference applies to per-pass outputs (ArrayOutput.cpp:48-56).
    struct Part {
      std::vector<PtPixel> pixels;
      PtStats stats{};
      int rc{PTB200_OK};
      char error[512]{};
    };
    std::vector<Part> parts(n);
    std::vector<std::thread> threads;
    for (int g = 0; g < n; ++g) {
      threads.emplace_back([&, g] {
        Part &part = parts[g];
        part.pixels.assign(pixels, PtPixel{});
        PtRenderOptions opt = base;
        PtRenderParams prm = *params;
        opt.device = list[g];
        const int lo = static_cast<int>(static_cast<long long>(spp) * g / n);
        const int hi = static_cast<int>(static_cast<long long>(spp) * (g + 1) / n);
        opt.passBegin = base.passBegin + lo;
        prm.samplesPerPixel = hi - lo;
        part.rc = ptb200_render(scene, camera, &prm, &opt, part.pixels.data(), nullptr, nullptr, &part.stats);
        if (part.rc)
          snprintf(part.error, sizeof part.error, "%s", ptb200_last_error());
      });
    }
    for (auto &t : threads)
      t.join();
    for (int g = 0; g < n; ++g) {
      if (parts[g].rc)
        return fail(parts[g].rc, "device %d: %s", list[g], parts[g].error);
      total.samples += parts[g].stats.samples;
      total.casts += parts[g].stats.casts;
      total.kernelLaunches += parts[g].stats.kernelLaunches;
      total.kernelMs = std::max(total.kernelMs, parts[g].stats.kernelMs);
      total.sweepKernelMs = std::max(total.sweepKernelMs, parts[g].stats.sweepKernelMs);
    }
    std::memset(out, 0, pixels * sizeof(PtPixel));
    for (int g = 0; g < n; ++g) {
      for (size_t i = 0; i < pixels; ++i) {
        out[i].sum[0] += parts[g].pixels[i].sum[0];
        out[i].sum[1] += parts[g].pixels[i].sum[1];
        out[i].sum[2] += parts[g].pixels[i].sum[2];
        out[i].numSamples += parts[g].pixels[i].numSamples;
      }
    }
    if (progress)
      progress(user, out, spp, spp);
    if (stats)
      *stats = total;
    return PTB200_OK;
  }

  // Pixel-parallel policies: device g owns the rows y % n == g for every pass and copies exactly
  // those rows into the caller's frame — the final gather needs no arithmetic.  With a progress
  // callback the passes are rendered in slices: all devices render a slice concurrently, then the
  // callback sees the whole partial frame on the CALLING thread (Scene.cpp:242-245).
  struct Device {
    PtContext *ctx{nullptr};
    PtStats slice{};
    int rc{PTB200_OK};
    char error[512]{};
  };
  std::vector<Device> devs(n);
  auto onEveryDevice = [&](auto &&body) {
    std::vector<std::thread> threads;
    for (int g = 0; g < n; ++g)
      threads.emplace_back([&, g] {
        devs[g].rc = body(g);
        if (devs[g].rc)
          snprintf(devs[g].error, sizeof devs[g].error, "%s", ptb200_last_error());
      });
    for (auto &t : threads)
      t.join();
    for (int g = 0; g < n; ++g)
      if (devs[g].rc)
        return g;
    return -1;
  };
  auto release = [&](bool ok) {
    for (Device &d : devs) {
      if (!d.ctx)
        continue;
      if (ok)
        poolGive(d.ctx);
      else
        ptb200_context_destroy(d.ctx);
    }
  };
  int failed = onEveryDevice([&](int g) {
    devs[g].ctx = poolTake(list[g]);
    if (!devs[g].ctx) {
      if (const int rc = ptb200_context_create(list[g], &devs[g].ctx))
        return rc;
    }
    return ptb200_context_upload_scene(devs[g].ctx, scene);
  });
  const int step = progress && spp > 0 ? (base.passesPerBatch > 0 ? base.passesPerBatch : std::max(1, (spp + 19) / 20))
                                       : std::max(spp, 1);
  for (int done = 0; failed < 0 && done < std::max(spp, 1); done += step) {
    const int passes = std::min(step, spp - done);
    failed = onEveryDevice([&](int g) {
      PtRenderOptions opt = base;
      PtRenderParams prm = *params;
      opt.device = list[g];
      opt.rowBegin = g;
      opt.rowStep = n;
      opt.passBegin = base.passBegin + done;
      prm.samplesPerPixel = passes;
      devs[g].slice = PtStats{};
      if (const int rc = ptb200_context_render(devs[g].ctx, camera, &prm, &opt, done > 0, &devs[g].slice))
        return rc;
      return downloadRows(devs[g].ctx, out, &opt);
    });
    if (failed >= 0)
      break;
    double kernelMs = 0, sweepMs = 0;
    for (const Device &d : devs) {
      total.samples += d.slice.samples;
      total.casts += d.slice.casts;
      total.kernelLaunches += d.slice.kernelLaunches;
      kernelMs = std::max(kernelMs, d.slice.kernelMs);
      sweepMs = std::max(sweepMs, d.slice.sweepKernelMs);
    }
    total.kernelMs += kernelMs;
    total.sweepKernelMs += sweepMs;
    if (progress && progress(user, out, done + passes, spp) != 0)
      break;
  }
  if (failed >= 0) {
    const int rc = devs[failed].rc;
    char message[512];
    snprintf(message, sizeof message, "%s", devs[failed].error);
    release(false);
    return fail(rc, "device %d: %s", list[failed], message);
  }
  release(true);
  if (stats)
    *stats = total;
  return PTB200_OK;
}

int ptb200_intersect(const PtScene *scene, int32_t device, int32_t which, double nearerThan,
                     uint32_t numRays, const double *rays, PtHit *out) {
  if (numRays == 0)
    return PTB200_OK;
  if (!rays || !out)
    return fail(PTB200_EINVAL, "null argument");
  // Test hooks in `which`: bit 8 = the warp-cooperative sweep of the sequential kernel;
  // bits 9-11 = per-lane sweep variant + 1 (0 -> default two-stage FP64); bit 12 = stage-0
  // audit: `out` then receives four uint64 counters (pairs, stage-0 survivors, exact accepts,
  // VIOLATIONS) instead of hits; bit 13 = audit the moment-form filter (variant 7); bit 14 = bit 3
  // of the sweep-variant field.
  const int32_t mode = which & 0xff;
  if (mode < 0 || mode > 2)
    return fail(PTB200_EINVAL, "which must be 0, 1 or 2");
  PtContext *ctx = nullptr;
  int rc = ptb200_context_create(device, &ctx);
  if (rc)
    return rc;
  rc = ptb200_context_upload_scene(ctx, scene);
  if (rc) {
    ptb200_context_destroy(ctx);
    return rc;
  }
  DeviceBuffer<double> dRays;
  DeviceBuffer<PtHitDevice> dOut;
  auto body = [&]() -> int {
    PT_CUDA(dRays.ensure(static_cast<size_t>(numRays) * 6));
    PT_CUDA(dOut.ensure(numRays));
    PT_CUDA(cudaMemcpyAsync(dRays.ptr, rays, static_cast<size_t>(numRays) * 48, cudaMemcpyHostToDevice, ctx->stream));
    IntersectArgs a{};
    a.scene = ctx->scene;
    a.rays = dRays.ptr;
    a.out = dOut.ptr;
    a.numRays = numRays;
    a.which = mode;
    a.nearerThan = nearerThan;
    a.warpCooperative = (which & 0x100) ? 1 : 0;
    const int variant = ((which >> 9) & 7) | ((which >> 11) & 8); // bit 14 extends the field
    a.sweep = variant ? variant - 1 : 1;
    const bool audit = (which & 0x1000) != 0;
    if (a.sweep >= 2 || audit) {
      if (!ctx->filterUsable)
        return fail(PTB200_EINVAL, "scene outside the FP32 stage-0 range");
      double bound = ctx->sceneRadius;
      for (uint32_t i = 0; i < numRays; ++i)
        bound = std::max(bound, std::sqrt(rays[6 * i] * rays[6 * i] + rays[6 * i + 1] * rays[6 * i + 1] +
                                          rays[6 * i + 2] * rays[6 * i + 2]));
      if (const int rc2 = ensureFilter(ctx, bound * (1.0 + 1e-6), nullptr))
        return rc2;
    }
    if (audit) {
      DeviceBuffer<unsigned long long> counters;
      PT_CUDA(counters.ensure(4));
      PT_CUDA(cudaMemsetAsync(counters.ptr, 0, 32, ctx->stream));
      AuditArgs au{};
      au.scene = ctx->scene;
      au.rays = dRays.ptr;
      au.numRays = numRays;
      au.momentForm = (which & 0x2000) ? 1 : 0;
      au.counters = counters.ptr;
      PT_CUDA(launchAuditStage0(au, ctx->stream));
      PT_CUDA(cudaMemcpyAsync(out, counters.ptr, 32, cudaMemcpyDeviceToHost, ctx->stream));
      PT_CUDA(cudaStreamSynchronize(ctx->stream));
      return PTB200_OK;
    }
    PT_CUDA(launchIntersect(a, ctx->stream));
    PT_CUDA(cudaMemcpyAsync(out, dOut.ptr, static_cast<size_t>(numRays) * sizeof(PtHit), cudaMemcpyDeviceToHost, ctx->stream));
    PT_CUDA(cudaStreamSynchronize(ctx->stream));
    return PTB200_OK;
  };
  rc = body();
  dRays.release();
  dOut.release();
  ptb200_context_destroy(ctx);
  return rc;
}

int ptb200_measure_fp32_peak(int32_t device, double *tflops, double *milliseconds) {
  if (!tflops)
    return fail(PTB200_EINVAL, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PTB200_ECUDA, "no CUDA device available");
  }
  PT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop{};
  PT_CUDA(cudaGetDeviceProperties(&prop, device));
  DeviceBuffer<float> sink;
  PT_CUDA(sink.ensure(1));
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iterations = 4096;
  EventList owned;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PT_CUDA(owned.make(&e0));
  PT_CUDA(owned.make(&e1));
  double best = 0, bestMs = 0;
  for (int rep = 0; rep < 5; ++rep) { // first repetitions warm up
    PT_CUDA(cudaEventRecord(e0, nullptr));
    PT_CUDA(launchFp32Peak(sink.ptr, iterations, blocks, threads, nullptr));
    PT_CUDA(cudaEventRecord(e1, nullptr));
    PT_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // one FFMA2 = two lanes x (multiply + add) = 4 flop; 8 chains x 16 per iteration
    const double flops = 4.0 * 8 * 16 * static_cast<double>(iterations) * threads * blocks;
    const double rate = flops / (ms * 1e-3) / 1e12;
    if (rate > best) {
      best = rate;
      bestMs = ms;
    }
  }
  *tflops = best;
  if (milliseconds)
    *milliseconds = bestMs;
  return PTB200_OK;
}

int ptb200_measure_fp64_peak(int32_t device, double *tflops, double *milliseconds) {
  if (!tflops)
    return fail(PTB200_EINVAL, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PTB200_ECUDA, "no CUDA device available");
  }
  PT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop{};
  PT_CUDA(cudaGetDeviceProperties(&prop, device));
  DeviceBuffer<double> sink;
  PT_CUDA(sink.ensure(1));
  const int threads = 512, blocks = prop.multiProcessorCount * 4, iterations = 4096;
  EventList owned;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  PT_CUDA(owned.make(&e0));
  PT_CUDA(owned.make(&e1));
  double best = 0, bestMs = 0;
  for (int rep = 0; rep < 5; ++rep) { // first repetitions warm up
    PT_CUDA(cudaEventRecord(e0, nullptr));
    PT_CUDA(launchFp64Peak(sink.ptr, iterations, blocks, threads, nullptr));
    PT_CUDA(cudaEventRecord(e1, nullptr));
    PT_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8 * 16 * static_cast<double>(iterations) * threads * blocks;
    const double rate = flops / (ms * 1e-3) / 1e12;
    if (rate > best) {
      best = rate;
      bestMs = ms;
    }
  }
  *tflops = best;
  if (milliseconds)
    *milliseconds = bestMs;
  return PTB200_OK;
}

} // extern "C"
