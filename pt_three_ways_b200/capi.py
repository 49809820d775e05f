"""ctypes binding of libptb200.so — the C ABI declared in include/ptb200.h.

This is plumbing for tests and bench.py: every call goes through the same extern "C" entry
points a C++ or cgo/JNI caller would bind.  There is no Python or CPU implementation behind
it: if the shared library is missing the import-time error says how to build it, and if no
CUDA device is present the entry points return PTB200_ECUDA.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PACKAGE_DIR, "libptb200.so")

RNG_KEYED_PHILOX = 0
RNG_MT19937_SEQUENTIAL = 1
RNG_MT19937_PER_PIXEL = 2  # the reference's `fp` way
RNG_MT19937_SEQUENTIAL_OO = 3  # the reference's `oo` way

EXPORTS = [
    "ptb200_last_error", "ptb200_device_count", "ptb200_render", "ptb200_render_multi",
    "ptb200_render_multi_progress",
    "ptb200_intersect", "ptb200_context_create", "ptb200_context_destroy",
    "ptb200_context_upload_scene", "ptb200_context_render", "ptb200_context_download",
    "ptb200_measure_fp64_peak", "ptb200_measure_fp32_peak",
]


class PtMaterial(C.Structure):
    _fields_ = [("emission", C.c_double * 3), ("diffuse", C.c_double * 3),
                ("indexOfRefraction", C.c_double), ("reflectivity", C.c_double),
                ("reflectionConeAngleRadians", C.c_double)]


class PtScene(C.Structure):
    _fields_ = [("numTriangles", C.c_uint32), ("numSpheres", C.c_uint32),
                ("numMaterials", C.c_uint32), ("reserved", C.c_uint32),
                ("triangleVertices", C.c_void_p), ("triangleMaterial", C.c_void_p),
                ("sphereCentreRadius", C.c_void_p), ("sphereMaterial", C.c_void_p),
                ("materials", C.c_void_p), ("environment", C.c_double * 3)]


class PtCamera(C.Structure):
    _fields_ = [("v", C.c_double * 18)]


class PtRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("preview", C.c_int32),
                ("samplesPerPixel", C.c_int32), ("maxCpus", C.c_int32), ("maxDepth", C.c_int32),
                ("firstBounceUSamples", C.c_int32), ("firstBounceVSamples", C.c_int32),
                ("seed", C.c_int32)]


class PtRenderOptions(C.Structure):
    _fields_ = [("rngMode", C.c_int32), ("device", C.c_int32), ("passBegin", C.c_int32),
                ("rowBegin", C.c_int32), ("rowStep", C.c_int32), ("passesPerBatch", C.c_int32),
                ("lanesPerPass", C.c_int32), ("reserved", C.c_int32 * 1)]


class PtStats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("casts", C.c_uint64), ("kernelLaunches", C.c_uint64),
                ("kernelMs", C.c_double), ("sweepKernelMs", C.c_double)]


PIXEL_DTYPE = np.dtype([("sum", "<f8", 3), ("n", "<u8")])
HIT_DTYPE = np.dtype([("hit", "<i4"), ("inside", "<i4"), ("material", "<i4"),
                      ("primitive", "<i4"), ("distance", "<f8"), ("position", "<f8", 3),
                      ("normal", "<f8", 3)])
PROGRESS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32)

_lib = None


class Ptb200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"ptb200 error {code}: {message}")
        self.code = code


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # Build in-tree if a CUDA toolchain is at hand (a fresh checkout); otherwise fail
            # loudly: there is no CPU or Python implementation to fall back to.
            import shutil
            import subprocess
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                env = dict(os.environ, PATH=os.environ.get("PATH", "") + ":/usr/local/cuda/bin")
                subprocess.run(["make", "-C", os.path.join(PACKAGE_DIR, "csrc")], check=False,
                               capture_output=True, env=env)
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C pt_three_ways_b200/csrc` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.ptb200_last_error.restype = C.c_char_p
        _lib.ptb200_context_destroy.restype = None
        _lib.ptb200_context_destroy.argtypes = [C.c_void_p]
        _lib.ptb200_context_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        _lib.ptb200_context_upload_scene.argtypes = [C.c_void_p, C.POINTER(PtScene)]
        _lib.ptb200_context_render.argtypes = [C.c_void_p, C.POINTER(PtCamera),
                                               C.POINTER(PtRenderParams),
                                               C.POINTER(PtRenderOptions), C.c_int32,
                                               C.POINTER(PtStats)]
        _lib.ptb200_context_download.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ptb200_render.argtypes = [C.POINTER(PtScene), C.POINTER(PtCamera),
                                       C.POINTER(PtRenderParams), C.POINTER(PtRenderOptions),
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]
        _lib.ptb200_render_multi.argtypes = [C.POINTER(PtScene), C.POINTER(PtCamera),
                                             C.POINTER(PtRenderParams),
                                             C.POINTER(PtRenderOptions), C.c_void_p, C.c_int32,
                                             C.c_void_p, C.POINTER(PtStats)]
        _lib.ptb200_render_multi_progress.argtypes = [C.POINTER(PtScene), C.POINTER(PtCamera),
                                                      C.POINTER(PtRenderParams),
                                                      C.POINTER(PtRenderOptions), C.c_void_p, C.c_int32,
                                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(PtStats)]
        _lib.ptb200_intersect.argtypes = [C.POINTER(PtScene), C.c_int32, C.c_int32, C.c_double,
                                          C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.ptb200_measure_fp64_peak.argtypes = [C.c_int32, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double)]
        _lib.ptb200_measure_fp32_peak.argtypes = [C.c_int32, C.POINTER(C.c_double),
                                                  C.POINTER(C.c_double)]
        _lib.ptb200_device_count.argtypes = [C.POINTER(C.c_int32)]
    return _lib


def _check(code: int) -> None:
    if code != 0:
        raise Ptb200Error(code, lib().ptb200_last_error().decode())


def device_count() -> int:
    n = C.c_int32(0)
    _check(lib().ptb200_device_count(C.byref(n)))
    return n.value


class MarshalledScene:
    """Keeps the numpy buffers alive that a PtScene points into."""

    def __init__(self, scene):
        self.tv = np.ascontiguousarray(scene.triangle_vertices, dtype=np.float64)
        self.tm = np.ascontiguousarray(scene.triangle_material, dtype=np.uint32)
        self.sc = np.ascontiguousarray(scene.sphere_centre_radius, dtype=np.float64)
        self.sm = np.ascontiguousarray(scene.sphere_material, dtype=np.uint32)
        self.mats = np.ascontiguousarray(scene.materials, dtype=np.float64)
        env = np.asarray(scene.environment, dtype=np.float64)
        self.abi = PtScene(self.tm.shape[0], self.sm.shape[0], self.mats.shape[0], 0,
                           self.tv.ctypes.data, self.tm.ctypes.data, self.sc.ctypes.data,
                           self.sm.ctypes.data, self.mats.ctypes.data,
                           (C.c_double * 3)(*env))


def make_camera(camera18) -> PtCamera:
    cam = PtCamera()
    for i, v in enumerate(np.asarray(camera18, dtype=np.float64)):
        cam.v[i] = float(v)
    return cam


def make_params(width, height, spp=1, seed=1, max_depth=5, first_u=4, first_v=4, preview=0,
                max_cpus=1) -> PtRenderParams:
    return PtRenderParams(width, height, preview, spp, max_cpus, max_depth, first_u, first_v, seed)


def make_options(rng_mode=RNG_KEYED_PHILOX, device=0, pass_begin=0, row_begin=0, row_step=0,
                 passes_per_batch=0, lanes_per_pass=0) -> PtRenderOptions:
    return PtRenderOptions(rng_mode, device, pass_begin, row_begin, row_step, passes_per_batch, lanes_per_pass)


def _stats_dict(st: PtStats) -> dict:
    return dict(samples=st.samples, casts=st.casts, kernel_launches=st.kernelLaunches,
                kernel_ms=st.kernelMs, sweep_kernel_ms=st.sweepKernelMs)


def render(scene, camera18, params: PtRenderParams, options: PtRenderOptions | None = None,
           progress=None, devices=None, out=None):
    """One-shot host-buffer render through ptb200_render (or ptb200_render_multi when
    `devices` is a list / "all").  Returns (pixels structured array (H,W), stats dict).
    `out`: an existing C-contiguous (H,W) PIXEL_DTYPE array to render into (e.g. a frame shared
    between ranks: a row-partitioned call writes only its own rows)."""
    m = scene if isinstance(scene, MarshalledScene) else MarshalledScene(scene)
    cam = make_camera(camera18)
    opts = options or make_options()
    if out is None:
        out = np.zeros((params.height, params.width), dtype=PIXEL_DTYPE)
    elif out.shape != (params.height, params.width) or out.dtype != PIXEL_DTYPE or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous (height, width) PIXEL_DTYPE array")
    st = PtStats()
    if devices is not None:
        cb = PROGRESS_FN(progress) if progress else None
        arr = None if devices == "all" else np.asarray(devices, dtype=np.int32)
        _check(lib().ptb200_render_multi_progress(C.byref(m.abi), C.byref(cam), C.byref(params), C.byref(opts),
                                                  None if arr is None else arr.ctypes.data,
                                                  0 if arr is None else arr.shape[0], out.ctypes.data, cb, None,
                                                  C.byref(st)))
    else:
        cb = PROGRESS_FN(progress) if progress else None
        _check(lib().ptb200_render(C.byref(m.abi), C.byref(cam), C.byref(params), C.byref(opts),
                                   out.ctypes.data, cb, None, C.byref(st)))
    return out, _stats_dict(st)


SWEEP_ONE_STAGE, SWEEP_TWO_STAGE_FP64, SWEEP_FP32_STAGE0, SWEEP_FP32X2_STAGE0, SWEEP_FP32X2_STAGE0_T = 0, 1, 2, 3, 4
SWEEP_FP32X2_SIGNS_T, SWEEP_FP32X2_SIGNS = 5, 6  # 4 and 3 with the stage-0 decisions kept in sign bits
SWEEP_FP32X2_MOMENT = 7  # 6 in moment (Pluecker) form: dot products against per-triangle constants


def intersect(scene, rays, which=0, nearer_than=float("inf"), device=0, warp_cooperative=False,
              sweep=None):
    m = scene if isinstance(scene, MarshalledScene) else MarshalledScene(scene)
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    flags = 0x100 if warp_cooperative else 0
    if sweep is not None:
        flags |= (((sweep + 1) & 7) << 9) | (((sweep + 1) & 8) << 11)
    _check(lib().ptb200_intersect(C.byref(m.abi), device, which | flags,
                                  nearer_than, rays.shape[0], rays.ctypes.data, out.ctypes.data))
    return out


def audit_stage0(scene, rays, device=0, moment_form=False) -> dict:
    """Test hook: runs the FP32 stage-0 filter and the exact test on every (ray, triangle) pair
    and counts pairs / stage-0 survivors / exact accepts / violations (accepted but filtered)."""
    m = scene if isinstance(scene, MarshalledScene) else MarshalledScene(scene)
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    out = np.zeros(max(1, rays.shape[0]), dtype=HIT_DTYPE)
    _check(lib().ptb200_intersect(C.byref(m.abi), device, 0x1000 | (0x2000 if moment_form else 0), float("inf"), rays.shape[0],
                                  rays.ctypes.data, out.ctypes.data))
    counters = out.view(np.uint64)[:4]
    return dict(pairs=int(counters[0]), survivors=int(counters[1]), accepts=int(counters[2]),
                violations=int(counters[3]))


class Context:
    """Resident API: scene uploaded once, accumulator stays in HBM."""

    def __init__(self, device=0):
        self.handle = C.c_void_p()
        _check(lib().ptb200_context_create(device, C.byref(self.handle)))
        self._scene = None
        self._shape = None

    def close(self):
        if self.handle:
            lib().ptb200_context_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_scene(self, scene):
        self._scene = scene if isinstance(scene, MarshalledScene) else MarshalledScene(scene)
        _check(lib().ptb200_context_upload_scene(self.handle, C.byref(self._scene.abi)))

    def render(self, camera18, params, options=None, accumulate=False) -> dict:
        cam = make_camera(camera18)
        opts = options or make_options()
        st = PtStats()
        _check(lib().ptb200_context_render(self.handle, C.byref(cam), C.byref(params),
                                           C.byref(opts), 1 if accumulate else 0, C.byref(st)))
        self._shape = (params.height, params.width)
        return _stats_dict(st)

    def download(self, out=None):
        if out is None:
            out = np.zeros(self._shape, dtype=PIXEL_DTYPE)
        _check(lib().ptb200_context_download(self.handle, out.ctypes.data))
        return out


def measure_fp64_peak(device=0):
    t, ms = C.c_double(0), C.c_double(0)
    _check(lib().ptb200_measure_fp64_peak(device, C.byref(t), C.byref(ms)))
    return t.value, ms.value


def measure_fp32_peak(device=0):
    t, ms = C.c_double(0), C.c_double(0)
    _check(lib().ptb200_measure_fp32_peak(device, C.byref(t), C.byref(ms)))
    return t.value, ms.value
