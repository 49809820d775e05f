/*
 * ptb200.h — C ABI of the B200-native `dod` path-tracing backend.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  Everything
 * the reference's data-oriented renderer does between "scene arrays are built" and "an
 * ArrayOutput comes back" happens behind these entry points, on the GPU.  All file:line
 * citations are relative to the reference tree (mattgodbolt/pt-three-ways @ a4aeda0).
 *
 *   reference interface                                     replaced by
 *   ------------------------------------------------------  ---------------------------------
 *   dod::Scene::addTriangle/addSphere/setEnvironmentColour  PtScene (flat arrays the host
 *     (src/dod/Scene.h:37-42, Scene.cpp:181-195)              adaptor records) + upload
 *   dod::Scene::render(camera, params, updateFunc)           ptb200_render()
 *     (src/dod/Scene.h:44-46, Scene.cpp:198-254)
 *   dod::Scene::intersect / intersectSpheres /               ptb200_intersect()
 *     intersectTriangles (Scene.h:48-56, "visible for tests")
 *   ArrayOutput / SampledPixel accumulation                  PtPixel[] written by the device
 *     (src/util/ArrayOutput.h:9-54, SampledPixel.h:5-17)
 *
 * Ownership: the caller owns every host buffer for the duration of a call; the library owns
 * all device memory; no pointer is retained after a call returns (a PtContext keeps its own
 * device copies).  Errors: every function returns 0 on success or a PTB200_E* code, and
 * ptb200_last_error() returns a thread-local description; no exception crosses this line.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * PTB200_ECUDA.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PTB200_OK 0
#define PTB200_EINVAL 1   /* bad argument (null pointer, zero size, unsupported depth, ...) */
#define PTB200_ECUDA 2    /* CUDA runtime/driver error, or no device */
#define PTB200_ENOMEM 3   /* host or device allocation failed */
#define PTB200_ESTATE 4   /* call order (e.g. render before a scene upload) */

/* MaterialSpec (src/util/MaterialSpec.h:7-12), same field order: 9 doubles. */
typedef struct PtMaterial {
  double emission[3];
  double diffuse[3];
  double indexOfRefraction;
  double reflectivity;                 /* < 0: use Fresnel reflectance (Scene.cpp:142-145) */
  double reflectionConeAngleRadians;
} PtMaterial;

/* The SoA scene of dod::Scene (src/dod/Scene.h:24-31).  The reference stores one
 * MaterialSpec per primitive by value; here primitives index a palette. */
typedef struct PtScene {
  uint32_t numTriangles;
  uint32_t numSpheres;
  uint32_t numMaterials;
  uint32_t reserved;
  const double *triangleVertices;     /* numTriangles x 9: v0 v1 v2 (TriangleVertices.h:14) */
  const uint32_t *triangleMaterial;   /* numTriangles palette indices */
  const double *sphereCentreRadius;   /* numSpheres x 4: centre, radius (Sphere.h:7-12) */
  const uint32_t *sphereMaterial;     /* numSpheres palette indices */
  const PtMaterial *materials;        /* numMaterials */
  double environment[3];              /* Scene::environment_ (Scene.h:31) */
} PtScene;

/* Camera's private state in declaration order (src/math/Camera.h:11-18): 18 doubles. */
typedef struct PtCamera {
  double centre[3];
  double axisX[3];
  double axisY[3];
  double axisZ[3];
  double aspectRatio;
  double cameraPlaneDist;
  double reciprocalHeight;
  double reciprocalWidth;
  double apertureRadius;
  double focalDistance;
} PtCamera;

/* RenderParams (src/util/RenderParams.h:3-13), same order; bool widened to int32. */
typedef struct PtRenderParams {
  int32_t width;
  int32_t height;
  int32_t preview;
  int32_t samplesPerPixel;
  int32_t maxCpus;                    /* accepted and ignored: the GPUs are the "cpus" */
  int32_t maxDepth;
  int32_t firstBounceUSamples;
  int32_t firstBounceVSamples;
  int32_t seed;
} PtRenderParams;

/* Random-number policies (DESIGN.md "RNG"). */
#define PTB200_RNG_KEYED_PHILOX 0        /* counter-based, parallel over pixels: throughput mode */
#define PTB200_RNG_MT19937_SEQUENTIAL 1  /* the reference's own stream (Scene.cpp:211-216): one
                                            mt19937(seed+s) per pass walked row-major */
#define PTB200_RNG_MT19937_PER_PIXEL 2   /* the reference's `fp` way (`--way fp`, src/fp/Render.cpp:
                                            76-135): pass s seeds one mt19937(H*W*(seed+s) + x*W + y)
                                            per pixel, strata are walked v-major and the emission is
                                            added after the average.  Exact AND parallel over pixels;
                                            the image is fp::render's with --max-cpus 1 */
#define PTB200_RNG_MT19937_SEQUENTIAL_OO 3 /* the reference's `oo` way (`--way oo`, src/oo/Renderer.cpp:
                                            60-107): the sequential stream and u-major strata of mode 1
                                            with the estimator of mode 2 (sub-samples contribute
                                            radiance(child) or diffuse*radiance(child); the emission is
                                            added after the average; t == Epsilon is a hit,
                                            src/oo/Triangle.cpp:31).  Parallel over passes only, like
                                            mode 1; the image is oo::Renderer::render's with --max-cpus 1 */

/* What a backend-specific caller may set beyond RenderParams.  Zero-initialise for defaults. */
typedef struct PtRenderOptions {
  int32_t rngMode;        /* PTB200_RNG_* */
  int32_t device;         /* CUDA ordinal for single-device calls */
  int32_t passBegin;      /* first pass index s (seed+s); passes [passBegin, passBegin+spp) */
  int32_t rowBegin;       /* framebuffer partition: this call renders rows y with          */
  int32_t rowStep;        /*   y >= rowBegin and (y-rowBegin) % rowStep == 0; 0 means 1     */
  int32_t passesPerBatch; /* 0 = library default; progress callback runs between batches   */
  int32_t lanesPerPass;   /* sequential policies: lanes that share one pass (4, 8, 16 or 32);    */
                          /*   0 = chosen from the number of passes                           */
  int32_t reserved[1];
} PtRenderOptions;

/* SampledPixel (src/util/SampledPixel.h:5-7): sum of colours and the sample count. */
typedef struct PtPixel {
  double sum[3];
  uint64_t numSamples;
} PtPixel;

/* Hit + material of an IntersectionRecord (src/math/Hit.h:6-11, dod/IntersectionRecord.h). */
typedef struct PtHit {
  int32_t hit;            /* 0 = no intersection (std::nullopt) */
  int32_t inside;
  int32_t material;       /* palette index */
  int32_t primitive;      /* >= 0 triangle index; < 0: -(sphere index) - 1 */
  double distance;
  double position[3];
  double normal[3];
} PtHit;

typedef struct PtStats {
  uint64_t samples;       /* camera rays = sum of numSamples written                     */
  uint64_t casts;         /* Scene::intersect calls, counted on the device               */
  uint64_t kernelLaunches;
  double kernelMs;        /* device time of all kernels, CUDA events on the launch stream */
  double sweepKernelMs;   /* device time of the path-tracing kernel(s) only              */
} PtStats;

/* Called on the caller's thread after each collected batch of passes (the reference calls
 * updateFunc(output) after each collected pass, Scene.cpp:242-245).  `pixels` is the whole
 * framebuffer of this call so far, valid only during the callback; return non-zero to stop. */
typedef int (*PtProgressFn)(void *user, const PtPixel *pixels, int32_t passesDone,
                            int32_t passesTotal);

/* Thread-local message for the last failing call on this thread. */
const char *ptb200_last_error(void);

/* Number of CUDA devices visible; 0 with *no* error means "none" (still no CPU fallback). */
int ptb200_device_count(int32_t *count);

/* ---- one-shot, host buffers in, host buffers out (the reference-facing call) ------------ */

/* Replaces dod::Scene::render.  `out` holds width*height PtPixel, row-major, index
 * x + y*width (ArrayOutput.h:14-17); rows outside the rowBegin/rowStep selection are not
 * written (zero-initialise the buffer to read them as "no samples").  Uses options->device
 * only.  Returns exactly samplesPerPixel samples for every selected pixel (the reference drops
 * its last in-flight passes, Scene.cpp:251; we do not). */
int ptb200_render(const PtScene *scene, const PtCamera *camera, const PtRenderParams *params,
                  const PtRenderOptions *options, PtPixel *out, PtProgressFn progress,
                  void *user, PtStats *stats);

/* Same call spread over several devices of this process: keyed mode partitions the
 * framebuffer rows round-robin, sequential mode partitions the passes; the merge is a
 * host-side gather (no collective).  devices == NULL means all visible devices. */
int ptb200_render_multi(const PtScene *scene, const PtCamera *camera,
                        const PtRenderParams *params, const PtRenderOptions *options,
                        const int32_t *devices, int32_t numDevices, PtPixel *out,
                        PtStats *stats);

/* ptb200_render_multi with the progress callback of ptb200_render: in the pixel-parallel policies
 * the passes are rendered in slices, every device renders a slice concurrently and `progress`
 * sees the whole partial frame on the calling thread between slices (Scene.cpp:242-245); in the
 * sequential policies (passes shared out between the devices) it is called once, at the end. */
int ptb200_render_multi_progress(const PtScene *scene, const PtCamera *camera,
                                 const PtRenderParams *params, const PtRenderOptions *options,
                                 const int32_t *devices, int32_t numDevices, PtPixel *out,
                                 PtProgressFn progress, void *user, PtStats *stats);

/* Replaces dod::Scene::intersect (which=0), intersectSpheres (1), intersectTriangles (2);
 * nearerThan applies to 1 and 2.  rays: numRays x 6 doubles (origin, unit direction). */
int ptb200_intersect(const PtScene *scene, int32_t device, int32_t which, double nearerThan,
                     uint32_t numRays, const double *rays, PtHit *out);

/* ---- resident API: scene uploaded once, accumulator stays in HBM ------------------------ */

typedef struct PtContext PtContext;

int ptb200_context_create(int32_t device, PtContext **out);
void ptb200_context_destroy(PtContext *ctx);
/* H2D of the SoA arrays, once (copies; the caller's buffers are free after return). */
int ptb200_context_upload_scene(PtContext *ctx, const PtScene *scene);
/* Kernels only: renders into the context's device accumulator (zeroed first unless
 * `accumulate` is non-zero).  Synchronises before returning. */
int ptb200_context_render(PtContext *ctx, const PtCamera *camera, const PtRenderParams *params,
                          const PtRenderOptions *options, int32_t accumulate, PtStats *stats);
/* D2H of the accumulator: width*height PtPixel. */
int ptb200_context_download(PtContext *ctx, PtPixel *out);

/* ---- self-measurement helpers used by bench.py for the roofline denominators ----------- */

/* Runs a dependent-free DFMA loop on every SM and returns the measured fp64 FMA rate in
 * TFLOP/s (2 flops per FMA) and the kernel time. */
int ptb200_measure_fp64_peak(int32_t device, double *tflops, double *milliseconds);

/* The same for the packed FP32 FMA (FFMA2, 4 flops each): the roof of the FP32 stage-0 sweep that
 * binds the triangle-heavy scenes. */
int ptb200_measure_fp32_peak(int32_t device, double *tflops, double *milliseconds);

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
