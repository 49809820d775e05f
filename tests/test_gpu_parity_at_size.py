"""GPU parity AT THE SIZES BASELINE.json names.

The oracle is too slow for whole frames of configs[1..3], but in the keyed policy every pixel has
its own random stream, so a ROW SUBSET rendered at the full sample count is an exact check of
those pixels of the full render (`rowBegin/rowStep` exist on both sides).  configs[0] is small
enough to compare whole, in both the keyed and the reference's own sequential policy.  The
sequential policy is also compared directly with images of the reference's own
radiance()/Camera::randomRay() (tests/golden/pass_*.npy), no oracle in between."""
import os
import subprocess

import numpy as np
import pytest

from tests.golden.make_golden import PASS_CASES

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 4
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode_name", ["keyed", "sequential"])
def test_config0_cornell_256x256_8spp_matches_oracle(mode_name, scenes, oracle, capi):
    """BASELINE configs[0] (the reference's own smoke size, scripts/bench-st-cornell.sh:9-13 at 8 spp),
    whole frame, bit for bit."""
    mode = capi.RNG_KEYED_PHILOX if mode_name == "keyed" else capi.RNG_MT19937_SEQUENTIAL
    scene = scenes["cornell"]
    w, h, spp, seed = 256, 256, 8, 1
    cam = scene.camera(w, h)
    pixels, stats = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(rng_mode=mode))
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=spp, seed=seed), mode, threads=THREADS)
    assert stats["casts"] == want["casts"]
    assert abs(stats["casts"] / stats["samples"] - 55.5) < 0.3  # 55.50 at 256x256
    assert np.array_equal(pixels["n"], want["counts"])
    assert np.array_equal(pixels["sum"], want["sums"])  # bit-exact


@pytest.mark.parametrize("scene_name,w,h,spp,row_step", [
    ("cornell", 640, 480, 256, 48),   # BASELINE configs[1]: 10 rows x 640 px x 256 spp
    ("suzanne", 640, 480, 256, 160),  # BASELINE configs[2]: 3 rows
])
def test_full_spp_row_subset_matches_oracle_and_the_whole_frame(scene_name, w, h, spp, row_step, scenes, oracle, capi):
    scene = scenes[scene_name]
    cam = scene.camera(w, h)
    seed = 1
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(row_begin=0, row_step=row_step))
    subset = ctx.download().copy()
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=spp, seed=seed), oracle.RNG_KEYED_PHILOX,
                                            row_begin=0, row_step=row_step, threads=THREADS)
    assert st["casts"] == want["casts"]
    assert np.array_equal(subset["n"], want["counts"]) and (subset["n"][::row_step] == spp).all()
    assert np.array_equal(subset["sum"], want["sums"])  # bit-exact at the full sample count
    # ... and those rows of the WHOLE frame at full size are the same numbers
    ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed))
    whole = ctx.download()
    assert (whole["n"] == spp).all()
    assert np.array_equal(whole["sum"][::row_step], subset["sum"][::row_step])
    ctx.close()


def test_config3_ce_1280x720_rows(scenes, oracle, capi):
    """BASELINE configs[3]: ce is a closed world whose camera sits inside a zero-albedo light, so
    every sample is exactly 65 casts and every pixel the light's emission.  One row
    at the full 1024 spp through the properties, the same row at 32 spp against the oracle."""
    scene = scenes["ce"]
    w, h = 1280, 720
    cam = scene.camera(w, h)
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    st = ctx.render(cam, capi.make_params(w, h, spp=1024, seed=1), capi.make_options(row_begin=360, row_step=720))
    row = ctx.download()[360]
    assert st["samples"] == w * 1024 and st["casts"] == 65 * w * 1024
    assert (row["n"] == 1024).all()
    mean = row["sum"] / 1024.0
    np.testing.assert_allclose(mean, np.broadcast_to(np.array([2.27, 3, 2.97]) * 0.25, mean.shape), rtol=1e-13)
    st = ctx.render(cam, capi.make_params(w, h, spp=32, seed=1), capi.make_options(row_begin=360, row_step=720))
    got = ctx.download()
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=32, seed=1), oracle.RNG_KEYED_PHILOX,
                                            row_begin=360, row_step=720, threads=THREADS)
    assert st["casts"] == want["casts"]
    assert np.array_equal(got["sum"], want["sums"]) and np.array_equal(got["n"], want["counts"])
    ctx.close()


@pytest.mark.parametrize("case", PASS_CASES, ids=lambda c: c[0])
def test_dod_stream_pass_equals_the_reference_image(case, scenes, capi, golden_dir):
    """PTB200_RNG_MT19937_SEQUENTIAL — the path north_star names, at the reference's own random
    numbers — against the reference's own dod::Scene::radiance()/Camera::randomRay() per-pass
    images (ref_tool `pass`), no oracle in between.  Tolerance: the reference build contracts FMAs
    as GCC pleases and calls glibc's sin/cos/acos; equal paths give equal sums of products of
    material constants up to rounding: 1e-12 absolute on values <= ~20 (north_star asks 1e-4)."""
    name, scene_name, w, h, seed, p, fu, fv, depth, preview = case
    want = np.load(os.path.join(golden_dir, f"pass_{name}.npy"))
    scene = scenes[scene_name]
    pixels, _ = capi.render(scene, scene.camera(w, h),
                            capi.make_params(w, h, spp=1, seed=seed, first_u=fu, first_v=fv, max_depth=depth,
                                             preview=preview),
                            capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, pass_begin=p))
    assert (pixels["n"] == 1).all()
    assert np.abs(pixels["sum"] - want).max() <= 1e-12


def test_the_backend_drops_into_the_reference_itself(scenes, oracle, capi, tmp_path):
    """oracle/_ref/b200_dropin = include/ptb200_scene.hpp (b200::Scene) compiled
    against the REFERENCE's own Camera / ArrayOutput / MaterialSpec / loadObjFile and driven by its
    own createScene<SB> recipes (src/main/main.cpp:69-309) as doRender drives dod::Scene (:360-363),
    linked against libptb200.so.  Its raw file — written by the reference's ArrayOutput::save —
    must equal the ctypes render of the fixture scene bit for bit.  Built in the CPU container
    only (needs the reference tree); it travels to the GPU box with the snapshot."""
    exe = os.path.join(ROOT, "oracle", "_ref", "b200_dropin")
    if not os.access(exe, os.X_OK):
        pytest.skip("oracle/_ref/b200_dropin not built (needs /root/reference)")
    for name, w, h, spp, seed, mode in (("cornell", 48, 36, 3, 7, capi.RNG_KEYED_PHILOX),
                                        ("cornell", 24, 18, 2, 5, capi.RNG_MT19937_SEQUENTIAL),
                                        ("suzanne", 32, 24, 2, 3, capi.RNG_KEYED_PHILOX),
                                        ("bbc-owl", 32, 24, 2, 4, capi.RNG_KEYED_PHILOX)):
        out = str(tmp_path / f"{name}_{mode}.raw")
        res = subprocess.run([exe, "render", name, str(w), str(h), str(spp), str(seed), out, str(mode)],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        sums, counts = oracle.read_raw(out)
        scene = scenes[name]
        want, st = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=seed),
                               capi.make_options(rng_mode=mode))
        assert (counts == spp).all()
        assert np.array_equal(sums, want["sum"])
        assert f'"casts": {st["casts"]}' in res.stdout and '"updates": ' in res.stdout
