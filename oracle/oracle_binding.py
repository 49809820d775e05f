"""TEST INFRASTRUCTURE — ctypes binding of oracle/liboracle.so (the CPU restatement) and a thin
runner for oracle/_ref/ref_tool (the reference's own sources).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the product
package never imports this module."""
from __future__ import annotations

import ctypes
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_TOOL = os.path.join(HERE, "_ref", "ref_tool")

RNG_KEYED_PHILOX = 0
RNG_MT19937_SEQUENTIAL = 1
RNG_FP_PER_PIXEL = 2  # the reference's `fp` way: mt19937 per (pass, pixel), src/fp/Render.cpp
RNG_OO_SEQUENTIAL = 3  # the reference's `oo` way: dod's stream, fp's estimator, src/oo/Renderer.cpp

_lib = None


def build(force: bool = False) -> None:
    """Compiles liboracle.so (and, where the reference tree is mounted, _ref/ref_tool)."""
    if force or not os.path.exists(LIB_PATH) or (
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "pt_oracle.cpp"))):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        try:
            _lib = ctypes.CDLL(LIB_PATH)
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(LIB_PATH)
        c = ctypes
        _lib.oracle_scene_create.restype = c.c_void_p
        _lib.oracle_scene_create.argtypes = [c.c_uint32, c.c_void_p, c.c_void_p, c.c_uint32,
                                             c.c_void_p, c.c_void_p, c.c_uint32, c.c_void_p,
                                             c.c_void_p]
        _lib.oracle_scene_destroy.argtypes = [c.c_void_p]
        _lib.oracle_intersect.argtypes = [c.c_void_p, c.c_int, c.c_double, c.c_uint32,
                                          c.c_void_p, c.c_void_p]
        _lib.oracle_render.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_int,
                                       c.c_int, c.c_int, c.c_int, c.c_int, c.c_void_p,
                                       c.c_void_p, c.c_void_p, c.c_void_p]
        _lib.oracle_sincos.argtypes = [c.c_uint32, c.c_void_p, c.c_void_p, c.c_void_p]
        _lib.oracle_acos.argtypes = [c.c_uint32, c.c_void_p, c.c_void_p]
        _lib.oracle_mt19937.argtypes = [c.c_uint32, c.c_uint32, c.c_void_p]
        _lib.oracle_philox.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p]
        _lib.oracle_canonical.restype = c.c_double
        _lib.oracle_canonical.argtypes = [c.c_uint32, c.c_uint32]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def params_array(width, height, spp=1, seed=1, max_depth=5, first_u=4, first_v=4, preview=0,
                 max_cpus=1) -> np.ndarray:
    """RenderParams in declaration order (src/util/RenderParams.h:3-13)."""
    return np.array([width, height, preview, spp, max_cpus, max_depth, first_u, first_v, seed],
                    dtype=np.int32)


class OracleScene:
    def __init__(self, scene):
        self._keep = [np.ascontiguousarray(scene.triangle_vertices, dtype=np.float64),
                      np.ascontiguousarray(scene.triangle_material, dtype=np.uint32),
                      np.ascontiguousarray(scene.sphere_centre_radius, dtype=np.float64),
                      np.ascontiguousarray(scene.sphere_material, dtype=np.uint32),
                      np.ascontiguousarray(scene.materials, dtype=np.float64),
                      np.ascontiguousarray(scene.environment, dtype=np.float64)]
        tv, tm, sc, sm, mats, env = self._keep
        self.handle = lib().oracle_scene_create(tm.shape[0], _ptr(tv), _ptr(tm), sm.shape[0],
                                                _ptr(sc), _ptr(sm), mats.shape[0], _ptr(mats),
                                                _ptr(env))

    def __del__(self):
        if getattr(self, "handle", None):
            lib().oracle_scene_destroy(self.handle)
            self.handle = None

    def intersect(self, rays: np.ndarray, which: int = 0, nearer_than: float = float("inf")):
        """rays (N,6) -> (N,12): hit, distance, inside, pos3, normal3, material, primitive, 0."""
        rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
        out = np.zeros((rays.shape[0], 12), dtype=np.float64)
        lib().oracle_intersect(self.handle, which, nearer_than, rays.shape[0], _ptr(rays), _ptr(out))
        return out

    def render(self, camera18, params9, rng_mode, pass_begin=0, num_passes=None, row_begin=0,
               row_step=1, threads=1, per_pass=False):
        """Returns dict(sums (H,W,3), counts (H,W), casts, rng_words[, per_pass (P,H,W,3)])."""
        params9 = np.ascontiguousarray(params9, dtype=np.int32)
        camera18 = np.ascontiguousarray(camera18, dtype=np.float64)
        w, h = int(params9[0]), int(params9[1])
        if num_passes is None:
            num_passes = int(params9[3])
        sums = np.zeros((h, w, 3), dtype=np.float64)
        counts = np.zeros((h, w), dtype=np.uint64)
        stats = np.zeros(2, dtype=np.uint64)
        pp = np.zeros((num_passes, h, w, 3), dtype=np.float64) if per_pass else None
        lib().oracle_render(self.handle, _ptr(camera18), _ptr(params9), rng_mode, pass_begin,
                            num_passes, row_begin, row_step, threads, _ptr(sums), _ptr(counts),
                            _ptr(pp) if per_pass else None, _ptr(stats))
        out = dict(sums=sums, counts=counts, casts=int(stats[0]), rng_words=int(stats[1]))
        if per_pass:
            out["per_pass"] = pp
        return out


# ---- the reference itself ---------------------------------------------------------------------

def have_ref_tool() -> bool:
    return os.access(REF_TOOL, os.X_OK)


def have_reference_tree() -> bool:
    return os.path.isdir(os.environ.get("PT_REFERENCE_ROOT", "/root/reference") + "/scenes")


def ref_tool(*args, cwd=None) -> str:
    res = subprocess.run([REF_TOOL, *map(str, args)], check=True, capture_output=True, cwd=cwd)
    return res.stdout.decode()


def ref_pass(scene, width, height, seed, pass_index, tmpdir, first_u=4, first_v=4, max_depth=5,
             preview=0) -> np.ndarray:
    """One pass image (H,W,3) from the reference's own radiance()/randomRay()."""
    out = os.path.join(tmpdir, f"ref_pass_{width}x{height}_{seed}_{pass_index}.f64")
    ref_tool("pass", scene, width, height, seed, pass_index, first_u, first_v, max_depth, preview, out)
    return np.fromfile(out, dtype=np.float64).reshape(height, width, 3)


def ref_fp_pass(scene, width, height, seed, tmpdir, first_u=4, first_v=4, max_depth=5, preview=0) -> np.ndarray:
    """One pass image (H,W,3) from the reference's `fp` way (fp::render with spp=1, maxCpus=1)."""
    out = os.path.join(tmpdir, f"ref_fp_pass_{width}x{height}_{seed}.f64")
    ref_tool("fp-pass", scene, width, height, seed, first_u, first_v, max_depth, preview, out)
    return np.fromfile(out, dtype=np.float64).reshape(height, width, 3)


def ref_oo_pass(scene, width, height, seed, pass_index, tmpdir, first_u=4, first_v=4, max_depth=5,
                preview=0) -> np.ndarray:
    """One pass image (H,W,3) from the reference's own oo::Renderer::radiance()/randomRay()."""
    out = os.path.join(tmpdir, f"ref_oo_pass_{width}x{height}_{seed}_{pass_index}.f64")
    ref_tool("oo-pass", scene, width, height, seed, pass_index, first_u, first_v, max_depth, preview, out)
    return np.fromfile(out, dtype=np.float64).reshape(height, width, 3)


def ref_oo_render(scene, width, height, spp, max_cpus, seed, out="-", first_u=4, first_v=4, max_depth=5) -> dict:
    """Runs the unmodified oo::Renderer::render; returns its JSON line (seconds, total_samples)."""
    return json.loads(ref_tool("oo-render", scene, width, height, spp, max_cpus, seed, first_u,
                               first_v, max_depth, out))


def ref_fp_render(scene, width, height, spp, max_cpus, seed, out="-", first_u=4, first_v=4, max_depth=5) -> dict:
    """Runs the unmodified fp::render; returns its JSON line (seconds, total_samples)."""
    return json.loads(ref_tool("fp-render", scene, width, height, spp, max_cpus, seed, first_u,
                               first_v, max_depth, out))


def ref_render(scene, width, height, spp, max_cpus, seed, out="-", first_u=4, first_v=4,
               max_depth=5) -> dict:
    """Runs the unmodified dod::Scene::render; returns its JSON line (seconds, total_samples)."""
    return json.loads(ref_tool("render", scene, width, height, spp, max_cpus, seed, first_u,
                               first_v, max_depth, out))


def ref_passes(scene, width, height, spp, threads, seed, out="-", first_u=4, first_v=4,
               max_depth=5) -> dict:
    """The reference's own radiance()/randomRay() for every pass, passes spread fairly over
    `threads` workers and ALL of them kept (ref_tool `passes`); returns its JSON line."""
    return json.loads(ref_tool("passes", scene, width, height, spp, threads, seed, first_u,
                               first_v, max_depth, out))


def ref_intersect(scene, rays, tmpdir, which=0, nearer_than="inf") -> np.ndarray:
    """(N,18): hit, distance, inside, pos3, normal3, material 9 doubles."""
    rays = np.ascontiguousarray(rays, dtype=np.float64).reshape(-1, 6)
    rin = os.path.join(tmpdir, "rays.f64")
    rout = os.path.join(tmpdir, "hits.f64")
    rays.tofile(rin)
    ref_tool("intersect", scene, which, nearer_than, rin, rout)
    return np.fromfile(rout, dtype=np.float64).reshape(-1, 18)


def ref_camera(scene, width, height) -> np.ndarray:
    return np.array([float.fromhex(tok) for tok in ref_tool("camera", scene, width, height).split()])


def read_raw(path):
    """ArrayOutput raw file (src/util/ArrayOutput.cpp:65-81) -> (sums (H,W,3), counts (H,W))."""
    data = open(path, "rb").read()
    sig, ver, h, w = np.frombuffer(data, dtype="<u4", count=4)
    assert sig == 1 and ver == 1
    rec = np.frombuffer(data, dtype=np.dtype([("c", "<f8", 3), ("n", "<u4")]), offset=16)
    return rec["c"].reshape(h, w, 3).copy(), rec["n"].reshape(h, w).copy()
