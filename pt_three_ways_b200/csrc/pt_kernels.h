// Argument blocks and host-callable launchers of the sm_100a kernels (pt_kernels.cu).
#pragma once

#include "pt_device.cuh"

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace ptb200 {

struct PtPixelDevice { // == PtPixel (include/ptb200.h)
  double sum[3];
  unsigned long long numSamples;
};

struct PtHitDevice { // == PtHit (include/ptb200.h)
  int32_t hit, inside, material, primitive;
  double distance;
  double position[3];
  double normal[3];
};

struct KeyedArgs {
  DeviceScene scene;
  DeviceCamera camera;
  uint32_t width, height;
  int32_t way;                     // 0: dod estimator + keyed Philox; 1: the `fp` way, mt19937 per (pass, pixel)
  int32_t rowBegin, rowStep;
  uint32_t ownPixels;              // pixels of the selected rows
  unsigned long long totalItems;   // ownPixels * passes of this batch
  int32_t seed, passBegin;         // pass s of the batch uses key seed + passBegin + s
  int32_t maxDepth, firstBounceU, firstBounceV, preview;
  int32_t firstBounceUPow2, firstBounceVPow2; // strata counts are powers of two:
  double invFirstBounceU, invFirstBounceV;    //   divide by multiplying with the exact reciprocal
  double *samples;                 // [passInBatch][ownPixel][3]
  uint32_t *mtHistory;             // fp way: kMtHistoryStride words per thread of the grid (pt_mt19937.cuh)
  size_t mtHistoryThreads;         //   ... how many threads it was sized for
  uint32_t mtStoreLimit;           //   ... generated words >= this are never read back: not stored
  unsigned long long *ticket;      // work counter, zeroed before launch
  unsigned long long *castCounter;
};

// The keyed policy as a pipeline (pt_split.cu): camera rays + first hits, sub-paths, resolve.
struct SplitArgs {
  DeviceScene scene;
  DeviceCamera camera;
  uint32_t width, height;
  int32_t rowBegin, rowStep;
  uint32_t ownPixels;              // pixels of the selected rows
  uint32_t totalSamples;           // ownPixels * passes of this batch; totalSamples * numSub < 2^31
  uint32_t numPasses;              // passes of this batch
  uint32_t numSub;                 // firstBounceU * firstBounceV
  int32_t numSubShift;             // log2(numSub) when it is a power of two, else -1
  int32_t firstBounceVShift;       // likewise for firstBounceV
  uint32_t numMaterials;
  int32_t seed, passBegin;         // pass s of the batch uses key seed + passBegin + s
  int32_t maxDepth, firstBounceU, firstBounceV, preview;
  int32_t firstBounceUPow2, firstBounceVPow2; // strata counts are powers of two:
  double invFirstBounceU, invFirstBounceV;    //   divide by multiplying with the exact reciprocal
  double2 *records;                // [<= totalSamples][9]: what radiance() holds at a camera ray's hit
  double *terms;                   // [totalSamples][numSub][3]: one term per stratum (or, for a sample
                                   //   that ended at the camera ray, its colour in the first slot)
  uint8_t *sampleKind;             // [totalSamples] 0: strata terms, 1: colour
  unsigned long long *counters;    // [0] sub-path ticket, [1] casts, [2] records appended
  PtPixelDevice *accumulator;      // full frame
  MomentTable momentTable;         // sweep variant 9: the stage-0 table of a scene of <= 64 triangles, so that
                                   //   it lives in constant bank 0 (pt_device.cuh: sweepConstTable)
};

struct SequentialArgs {
  DeviceScene scene;
  DeviceCamera camera;
  int32_t width, height;
  int32_t seed, passBegin, numPasses;
  int32_t maxDepth, firstBounceU, firstBounceV, preview;
  int32_t way;                     // 0: the dod estimator (Scene.cpp:124-179); 2: the `oo` way's
                                   //    (src/oo/Renderer.cpp:60-91) on the same stream
  double *samples;                 // [passInBatch][height*width][3]
  unsigned long long *castCounter;
};

struct ReduceArgs {
  const double *samples;
  PtPixelDevice *accumulator;      // full frame
  uint32_t width;
  uint32_t rowBegin, rowStep;
  uint32_t ownPixels;
  uint32_t numPasses;
  size_t samplePassStride;         // pixels per pass in `samples`
  int32_t samplesAreFullFrame;     // sequential kernel writes whole frames
};

struct BuildFilterArgs {
  DeviceScene scene;
  float *out;                      // [numTiles][tileTris/4][14][4]
  float *outMoment;                // [numTiles][tileTris/4][19][4] (sweep variant 7)
  double originBound;              // >= |o| for every ray origin of the coming launch
};

struct AuditArgs {                 // test hook: stage 0 must never reject what the exact test accepts
  DeviceScene scene;
  const double *rays;
  uint32_t numRays;
  int32_t momentForm;              // audit the moment-form filter of sweep variant 7 instead
  unsigned long long *counters;    // [0] (ray,triangle) pairs, [1] stage-0 survivors,
                                   // [2] exact accepts, [3] VIOLATIONS: exact accept but stage-0 reject
};

struct IntersectArgs {
  DeviceScene scene;
  const double *rays;
  PtHitDevice *out;
  uint32_t numRays;
  int32_t which;                   // 0 intersect, 1 spheres only, 2 triangles only
  double nearerThan;
  int32_t warpCooperative;
  int32_t sweep;                   // 0 one-stage, 1 two-stage FP64, 2 FP32 stage 0 + exact
};

size_t keyedSmemBytes(uint32_t numSpheres, uint32_t tileTris, uint32_t numTiles, int sweep,
                      uint32_t threadsForPrimarySlots, int way);
size_t mtHistoryThreadsFor(int numSms);
int chooseKeyedConfig(uint32_t numTriangles, bool filterUsable, int way); // 10 * launchShape + sweepVariant
cudaError_t launchBuildFilter(const BuildFilterArgs &args, cudaStream_t stream);
cudaError_t launchAuditStage0(const AuditArgs &args, cudaStream_t stream);
cudaError_t launchRenderKeyed(const KeyedArgs &args, int numSms, int config, cudaStream_t stream);
size_t splitBytesPerSample(uint32_t numSub);
bool constTableFits(uint32_t numTriangles, uint32_t numTiles); // sweep variant 9 applies
cudaError_t launchSplitTrace(const SplitArgs &args, int numSms, int config, cudaStream_t stream);   // 2 launches
cudaError_t launchSplitResolve(const SplitArgs &args, cudaStream_t stream);                         // 1 launch
int sequentialLanesPerPass(int numPasses, int numSms, int requested);
cudaError_t launchRenderSequential(const SequentialArgs &args, int lanesPerPass, cudaStream_t stream);
cudaError_t launchReducePasses(const ReduceArgs &args, cudaStream_t stream);
cudaError_t launchIntersect(const IntersectArgs &args, cudaStream_t stream);
cudaError_t launchFp64Peak(double *sink, int iterations, int blocks, int threads,
                           cudaStream_t stream);
cudaError_t launchFp32Peak(float *sink, int iterations, int blocks, int threads, cudaStream_t stream);

} // namespace ptb200
