"""GPU parity for PTB200_RNG_MT19937_SEQUENTIAL_OO — the reference's `oo` way
(src/oo/Renderer.cpp:60-107): dod's per-pass mt19937 stream and u-major strata with the emission
added after the average (Material::totalEmission, src/oo/Material.cpp:19-22) and t == Epsilon
accepted (src/oo/Triangle.cpp:31).  Rendered by the sequential kernel's kOo instantiation; checked
bit for bit against the oracle's restatement and directly against images of the reference's own
oo::Renderer::radiance (tests/golden/oo_pass_*.npy, tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from tests.golden.make_golden import OO_PASS_CASES

pytestmark = pytest.mark.gpu

RENDER_CASES = [
    # scene, width, height, spp, seed, kwargs
    ("cornell", 40, 30, 3, 1, {}),
    ("cornell", 33, 17, 2, 5, dict(first_u=2, first_v=3, max_depth=3)),
    ("cornell", 24, 18, 2, 9, dict(max_depth=1)),
    ("cornell", 24, 18, 2, 9, dict(max_depth=2)),
    ("cornell", 24, 18, 1, 9, dict(preview=1)),
    ("cornell", 10, 8, 1, 3, dict(first_u=1, first_v=1, max_depth=40)),
    ("suzanne", 16, 12, 2, 2, {}),
    ("single-sphere", 32, 24, 2, 3, {}),
    ("multi-sphere", 32, 24, 2, 4, {}),
    ("example1", 32, 24, 2, 5, {}),
    ("bbc-owl", 32, 24, 2, 6, {}),
    ("ce", 8, 6, 1, 7, {}),
]


@pytest.mark.parametrize("case", RENDER_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-{c[3]}spp-{len(c[5])}")
def test_oo_way_render_matches_oracle(case, scenes, oracle, capi):
    name, w, h, spp, seed, kw = case
    scene = scenes[name]
    camera = scene.camera(w, h)
    pixels, stats = capi.render(scene, camera, capi.make_params(w, h, spp=spp, seed=seed, **kw),
                                capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO))
    want = oracle.OracleScene(scene).render(camera, oracle.params_array(w, h, spp=spp, seed=seed, **kw),
                                            oracle.RNG_OO_SEQUENTIAL, threads=4)
    assert np.array_equal(pixels["n"], want["counts"])
    assert stats["casts"] == want["casts"]
    assert stats["samples"] == w * h * spp
    assert np.array_equal(pixels["sum"], want["sums"])  # bit-exact


@pytest.mark.parametrize("case", OO_PASS_CASES, ids=lambda c: c[0])
def test_oo_way_pass_equals_the_reference_image(case, scenes, capi, golden_dir):
    """The CUDA path against the reference's own oo::Renderer::radiance, no oracle in between.
    Tolerance: the reference build contracts FMAs as GCC pleases and calls glibc's sin/cos; equal
    paths give equal sums of products of material constants up to rounding: 1e-12 absolute on
    values <= ~20."""
    name, scene_name, w, h, seed, p, fu, fv, depth, preview = case
    want = np.load(os.path.join(golden_dir, f"oo_pass_{name}.npy"))
    scene = scenes[scene_name]
    pixels, _ = capi.render(scene, scene.camera(w, h),
                            capi.make_params(w, h, spp=1, seed=seed, first_u=fu, first_v=fv, max_depth=depth,
                                             preview=preview),
                            capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO, pass_begin=p))
    assert (pixels["n"] == 1).all()
    assert np.abs(pixels["sum"] - want).max() <= 1e-12


def test_oo_way_partitions_by_passes_only(scenes, capi):
    """One engine per pass: pass blocks add up bit for bit; a row partition is refused."""
    scene = scenes["cornell"]
    w, h, spp, seed = 24, 18, 5, 3
    cam = scene.camera(w, h)
    oo = capi.RNG_MT19937_SEQUENTIAL_OO
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(rng_mode=oo))
    whole = ctx.download().copy()
    assert (whole["n"] == spp).all()
    ctx.render(cam, capi.make_params(w, h, spp=2, seed=seed), capi.make_options(rng_mode=oo))
    ctx.render(cam, capi.make_params(w, h, spp=3, seed=seed), capi.make_options(rng_mode=oo, pass_begin=2),
               accumulate=True)
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    with pytest.raises(capi.Ptb200Error):
        ctx.render(cam, capi.make_params(w, h, spp=1, seed=seed),
                   capi.make_options(rng_mode=oo, row_begin=1, row_step=2))
    ctx.close()


def test_oo_way_differs_from_the_dod_stream_mode_only_in_rounding(scenes, capi):
    scene = scenes["cornell"]
    w, h = 32, 24
    cam = scene.camera(w, h)
    oo, st_oo = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=1),
                            capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL_OO))
    dod, st_dod = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=1),
                              capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL))
    assert st_oo["casts"] == st_dod["casts"]
    diff = np.abs(oo["sum"] - dod["sum"]).max()
    assert 0 < diff <= 1e-12
