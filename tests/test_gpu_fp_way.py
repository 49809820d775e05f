"""GPU parity for PTB200_RNG_MT19937_PER_PIXEL — the reference's `fp` way (src/fp/Render.cpp:76-135):
one std::mt19937 per (pass, pixel), v-major strata, emission added after the average.  Exact AND
parallel over pixels, so the megakernel renders it; checked bit for bit against the oracle's
restatement and directly against images of fp::render built from the reference's own sources
(tests/golden/fp_pass_*.npy, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import FP_PASS_CASES

pytestmark = pytest.mark.gpu

RENDER_CASES = [
    # scene, width, height, spp, seed, kwargs
    ("cornell", 40, 30, 3, 1, {}),
    ("cornell", 33, 17, 2, 5, dict(first_u=2, first_v=3, max_depth=3)),
    ("cornell", 24, 18, 2, 9, dict(max_depth=1)),
    ("cornell", 24, 18, 2, 9, dict(max_depth=2)),
    ("cornell", 24, 18, 1, 9, dict(preview=1)),
    ("cornell", 12, 9, 2, 5, dict(first_u=8, first_v=8, max_depth=6)),  # > 624 words per engine
    ("cornell", 10, 8, 1, 3, dict(first_u=1, first_v=1, max_depth=40)),
    ("suzanne", 32, 24, 2, 2, {}),
    ("single-sphere", 32, 24, 2, 3, {}),
    ("multi-sphere", 32, 24, 2, 4, {}),
    ("example1", 32, 24, 2, 5, {}),
    ("bbc-owl", 32, 24, 2, 6, {}),
    ("ce", 16, 9, 1, 7, {}),
]


@pytest.mark.parametrize("case", RENDER_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-{c[3]}spp-{len(c[5])}")
def test_fp_way_render_matches_oracle(case, scenes, oracle, capi):
    name, w, h, spp, seed, kw = case
    scene = scenes[name]
    camera = scene.camera(w, h)
    pixels, stats = capi.render(scene, camera, capi.make_params(w, h, spp=spp, seed=seed, **kw),
                                capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    want = oracle.OracleScene(scene).render(camera, oracle.params_array(w, h, spp=spp, seed=seed, **kw),
                                            oracle.RNG_FP_PER_PIXEL, threads=4)
    assert np.array_equal(pixels["n"], want["counts"])
    assert stats["casts"] == want["casts"]
    assert stats["samples"] == w * h * spp
    assert np.array_equal(pixels["sum"], want["sums"])  # bit-exact


@pytest.mark.parametrize("case", FP_PASS_CASES, ids=lambda c: c[0])
def test_fp_way_pass_equals_the_reference_image(case, scenes, capi, golden_dir):
    """The CUDA path against fp::render of the reference itself, no oracle in between.  Tolerance:
    the reference build contracts FMAs as GCC pleases and calls glibc's sin/cos; equal paths give
    equal sums of products of material constants up to rounding: 1e-12 absolute on values <= ~20."""
    name, scene_name, w, h, seed, fu, fv, depth, preview = case
    want = np.load(os.path.join(golden_dir, f"fp_pass_{name}.npy"))
    scene = scenes[scene_name]
    pixels, _ = capi.render(scene, scene.camera(w, h),
                            capi.make_params(w, h, spp=1, seed=seed, first_u=fu, first_v=fv, max_depth=depth,
                                             preview=preview),
                            capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    assert (pixels["n"] == 1).all()
    assert np.abs(pixels["sum"] - want).max() <= 1e-12


def test_fp_way_partitions_and_batches(scenes, capi):
    """Per-pixel engines: rows, passes and batches can be split any way without changing a bit."""
    scene = scenes["cornell"]
    w, h, spp, seed = 96, 72, 12, 3
    cam = scene.camera(w, h)
    fp = capi.RNG_MT19937_PER_PIXEL
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(rng_mode=fp))
    whole = ctx.download().copy()
    assert (whole["n"] == spp).all() and st["samples"] == w * h * spp
    ctx.render(cam, capi.make_params(w, h, spp=5, seed=seed), capi.make_options(rng_mode=fp))
    ctx.render(cam, capi.make_params(w, h, spp=7, seed=seed), capi.make_options(rng_mode=fp, pass_begin=5),
               accumulate=True)
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    for r in range(3):
        ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed),
                   capi.make_options(rng_mode=fp, row_begin=r, row_step=3))
        part = ctx.download()
        assert np.array_equal(part["sum"][r::3], whole["sum"][r::3])
    ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(rng_mode=fp, passes_per_batch=5))
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    # not the keyed image, but the same estimator: close in the mean
    keyed, _ = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=seed))
    assert not np.array_equal(keyed["sum"], whole["sum"])
    assert abs(keyed["sum"].mean() / whole["sum"].mean() - 1) < 0.05
    ctx.close()


@pytest.mark.parametrize("seed,kw", [(4, {}), (5, dict(first_u=3, first_v=2, max_depth=7)), (6, dict(max_depth=2)),
                                     (8, dict(first_u=1, first_v=5, max_depth=9)),
                                     (9, dict(first_u=8, first_v=8, max_depth=6))])
def test_fp_way_random_scenes_match_oracle(seed, kw, oracle, capi):
    from tests import random_scenes
    scene = random_scenes.random_scene(seed, num_triangles=25, num_spheres=3)
    w, h, spp = 28, 21, 3
    cam = scene.camera(w, h)
    got, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=13, **kw),
                          capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=spp, seed=13, **kw),
                                            oracle.RNG_FP_PER_PIXEL, threads=4)
    assert st["casts"] == want["casts"]
    assert np.array_equal(got["sum"], want["sums"])


def test_fp_way_camera_without_aperture(oracle, capi):
    """Two camera draws instead of four per engine (Camera.h:26-27)."""
    from tests import random_scenes
    scene = random_scenes.random_scene(9, num_triangles=10, num_spheres=2)
    scene.camera64x48 = random_scenes.camera18((0, 0, -5), (0, 0, 0), (0, 1, 0), 64, 48, 40.0)
    w, h = 20, 15
    cam = scene.camera(w, h)
    assert cam[16] == 0.0
    got, _ = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=5),
                         capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=2, seed=5), oracle.RNG_FP_PER_PIXEL)
    assert np.array_equal(got["sum"], want["sums"])


def test_fp_way_every_instantiation_renders_identically(tmp_path):
    """The fp way is instantiated for nine megakernel configurations (ptb200_shim.cu maps every
    PTB200_KEYED_CONFIG onto one of them)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from pt_three_ways_b200 import capi, scenefile\n"
            "s = scenefile.load(%r)\n"
            "px, st = capi.render(s, s.camera(64, 48), capi.make_params(64, 48, spp=3, seed=11),\n"
            "                     capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))\n"
            "np.save(sys.argv[1], px['sum']); print(st['casts'])\n") % (
                root, os.path.join(root, "tests/golden/scenes/suzanne.ptscene"))
    outs = []
    for config in ("1", "3", "4", "24", "5", "25", "45", "6", "26"):
        out = str(tmp_path / f"c{config}.npy")
        res = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True,
                             env=dict(os.environ, PTB200_KEYED_CONFIG=config), timeout=300)
        assert res.returncode == 0, res.stderr[-1500:]
        outs.append((np.load(out), res.stdout.strip()))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and outs[0][1] == other[1]


def test_fp_way_full_size_properties(scenes, capi):
    """CornellBox 640x480 (the BASELINE frame) at 16 passes: exactly spp samples everywhere, the
    counted casts per sample of SURVEY.md 8d, and the same image in the mean as the keyed policy."""
    scene = scenes["cornell"]
    w, h, spp = 640, 480, 16
    cam = scene.camera(w, h)
    fp, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=1),
                         capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    assert (fp["n"] == spp).all() and st["samples"] == w * h * spp
    assert abs(st["casts"] / st["samples"] - 44.6) < 0.5
    assert np.isfinite(fp["sum"]).all() and (fp["sum"] >= 0).all()
    keyed, _ = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=1))
    assert abs(fp["sum"].mean() / keyed["sum"].mean() - 1) < 0.01


def test_fp_way_on_every_device_of_the_box(scenes, capi):
    if capi.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    scene = scenes["cornell"]
    w, h, spp = 64, 50, 4
    cam = scene.camera(w, h)
    opts = capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL)
    single, st1 = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4), opts)
    multi, stn = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4), opts, devices="all")
    assert np.array_equal(multi["sum"], single["sum"]) and stn["casts"] == st1["casts"]
