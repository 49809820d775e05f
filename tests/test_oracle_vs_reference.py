"""Live comparison of the oracle with the reference's own binary (oracle/_ref/ref_tool), on
scenes the committed goldens do not cover.  Runs only where ref_tool was built (the build
container); skipped on the GPU box and wherever the reference tree is absent."""
import os

import numpy as np
import pytest

from tests import random_scenes
from pt_three_ways_b200 import scenefile


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref_tool():
        pytest.skip("oracle/_ref/ref_tool not built (needs /root/reference)")
    return oracle


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scene_intersections_match_the_reference_binary(seed, ref, tmp_path):
    scene = random_scenes.random_scene(seed)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    rays = random_scenes.rays_for(scene, 3000, seed + 10)
    want = ref.ref_intersect(path, rays, str(tmp_path))
    got = ref.OracleScene(scene).intersect(rays)
    assert np.array_equal(got[:, 0], want[:, 0])
    hit = want[:, 0] != 0
    assert hit.sum() > 1000
    assert np.array_equal(got[hit, 2], want[hit, 2])
    np.testing.assert_allclose(got[hit, 1], want[hit, 1], rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(got[hit, 3:9], want[hit, 3:9], rtol=0, atol=1e-12)
    assert np.array_equal(scene.materials[got[hit, 9].astype(int)], want[hit, 9:18])


@pytest.mark.parametrize("seed,kw", [(4, {}), (5, dict(first_u=3, first_v=2, max_depth=7)), (6, dict(max_depth=2))])
def test_random_scene_pass_images_match_the_reference_binary(seed, kw, ref, tmp_path):
    """Per-pass images of the reference's radiance()/randomRay() on random scenes with glossy,
    mirror, Fresnel and emissive materials and rays that start inside spheres."""
    scene = random_scenes.random_scene(seed, num_triangles=20, num_spheres=3)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    w, h = 24, 18
    for p in (0, 1):
        want = ref.ref_pass(path, w, h, 11, p, str(tmp_path), kw.get("first_u", 4), kw.get("first_v", 4),
                            kw.get("max_depth", 5))
        got = ref.OracleScene(scene).render(scene.camera(w, h), ref.params_array(w, h, spp=1, seed=11, **kw),
                                            ref.RNG_MT19937_SEQUENTIAL, pass_begin=p, num_passes=1,
                                            per_pass=True)["per_pass"][0]
        # Same paths => same sums of products of material constants; the reference build's FMA
        # contraction and libm differ from the oracle's fixed sequence by rounding only.
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, float(np.abs(want).max()))


def test_unmodified_render_entry_point_sums_passes(ref, tmp_path):
    """dod::Scene::render itself (seed_tests.sh configuration on a random scene): the raw file
    equals the sum of the passes it kept."""
    scene = random_scenes.random_scene(7, num_triangles=12, num_spheres=2)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    out = str(tmp_path / "asis.raw")
    ref.ref_render(path, 16, 16, 6, 1, 3, out)
    sums, counts = ref.read_raw(out)
    kept = int(counts[0, 0])
    assert (counts == kept).all() and 1 <= kept <= 6
    got = ref.OracleScene(scene).render(scene.camera(16, 16), ref.params_array(16, 16, spp=kept, seed=3),
                                        ref.RNG_MT19937_SEQUENTIAL)
    assert np.abs(got["sums"] - sums).max() <= 1e-10 * max(1.0, float(np.abs(sums).max()))


def test_every_pass_kept_driver_equals_the_oracle(ref, tmp_path):
    """`ref_tool passes` (bench.py's CPU baseline: the reference's radiance()/randomRay() for every
    pass, fairly scheduled over threads, nothing dropped) renders the oracle's image: spp samples
    everywhere, sums equal up to the thread-dependent order of `+=` over passes."""
    scene = random_scenes.random_scene(7, num_triangles=12, num_spheres=2)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    out = str(tmp_path / "passes.raw")
    info = ref.ref_passes(path, 16, 16, 6, 3, 3, out)
    assert info["total_samples"] == 16 * 16 * 6
    sums, counts = ref.read_raw(out)
    assert (counts == 6).all()
    got = ref.OracleScene(scene).render(scene.camera(16, 16), ref.params_array(16, 16, spp=6, seed=3),
                                        ref.RNG_MT19937_SEQUENTIAL)
    assert np.abs(got["sums"] - sums).max() <= 1e-10 * max(1.0, float(np.abs(sums).max()))


@pytest.mark.parametrize("seed,kw", [(4, {}), (5, dict(first_u=3, first_v=2, max_depth=7)), (6, dict(max_depth=2)),
                                     (8, dict(first_u=1, first_v=5, max_depth=9)),
                                     (9, dict(first_u=8, first_v=8, max_depth=6))])
def test_fp_way_pass_images_match_the_reference_binary(seed, kw, ref, tmp_path):
    """The reference's `fp` way (fp::render built from src/fp/*.cpp with the range-v3 stand-in of
    oracle/shims) against the oracle's RNG_FP_PER_PIXEL policy, on random scenes."""
    scene = random_scenes.random_scene(seed, num_triangles=20, num_spheres=3)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    w, h = 24, 18
    for pass_seed in (11, 12):
        want = ref.ref_fp_pass(path, w, h, pass_seed, str(tmp_path), kw.get("first_u", 4), kw.get("first_v", 4),
                               kw.get("max_depth", 5))
        got = ref.OracleScene(scene).render(scene.camera(w, h), ref.params_array(w, h, spp=1, seed=pass_seed, **kw),
                                            ref.RNG_FP_PER_PIXEL, num_passes=1, per_pass=True)["per_pass"][0]
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, float(np.abs(want).max()))


def test_fp_way_unmodified_render_entry_point(ref, tmp_path):
    scene = random_scenes.random_scene(7, num_triangles=12, num_spheres=2)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    out = str(tmp_path / "fp.raw")
    ref.ref_fp_render(path, 16, 16, 6, 1, 3, out)
    sums, counts = ref.read_raw(out)
    assert (counts == 6).all()
    got = ref.OracleScene(scene).render(scene.camera(16, 16), ref.params_array(16, 16, spp=6, seed=3),
                                        ref.RNG_FP_PER_PIXEL)
    assert np.abs(got["sums"] - sums).max() <= 1e-10 * max(1.0, float(np.abs(sums).max()))


@pytest.mark.parametrize("seed,kw", [(4, {}), (5, dict(first_u=3, first_v=2, max_depth=7)), (6, dict(max_depth=2)),
                                     (8, dict(first_u=1, first_v=5, max_depth=9))])
def test_oo_way_pass_images_match_the_reference_binary(seed, kw, ref, tmp_path):
    """The reference's `oo` way (oo::Renderer::radiance built from src/oo/*.cpp) against the
    oracle's RNG_OO_SEQUENTIAL policy, on random scenes."""
    scene = random_scenes.random_scene(seed, num_triangles=20, num_spheres=3)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    w, h = 24, 18
    for p in (0, 3):
        want = ref.ref_oo_pass(path, w, h, 11, p, str(tmp_path), kw.get("first_u", 4), kw.get("first_v", 4),
                               kw.get("max_depth", 5))
        got = ref.OracleScene(scene).render(scene.camera(w, h), ref.params_array(w, h, spp=1, seed=11, **kw),
                                            ref.RNG_OO_SEQUENTIAL, pass_begin=p, num_passes=1,
                                            per_pass=True)["per_pass"][0]
        assert np.abs(got - want).max() <= 1e-11 * max(1.0, float(np.abs(want).max()))


def test_oo_way_unmodified_render_entry_point(ref, tmp_path):
    scene = random_scenes.random_scene(7, num_triangles=12, num_spheres=2)
    path = str(tmp_path / "scene.ptscene")
    scenefile.save(scene, path)
    out = str(tmp_path / "oo.raw")
    ref.ref_oo_render(path, 16, 16, 6, 1, 3, out)
    sums, counts = ref.read_raw(out)
    kept = int(counts[0, 0])
    assert (counts == kept).all() and 1 <= kept <= 6
    got = ref.OracleScene(scene).render(scene.camera(16, 16), ref.params_array(16, 16, spp=kept, seed=3),
                                        ref.RNG_OO_SEQUENTIAL)
    assert np.abs(got["sums"] - sums).max() <= 1e-10 * max(1.0, float(np.abs(sums).max()))


@pytest.mark.parametrize("name", ["cornell", "suzanne", "ce", "single-sphere", "multi-sphere", "example1", "bbc-owl"])
def test_dropin_adaptor_marshals_what_the_reference_recipes_build(name, tmp_path, scenes):
    """CPU half of the drop-in proof (the GPU half is tests/test_gpu_parity_at_size.py):
    include/ptb200_scene.hpp compiled against the reference's own types, fed by the reference's own
    createScene<SB> + loadObjFile, hands the C ABI exactly the arrays of the committed fixture —
    triangles, spheres, per-primitive materials, environment and the 18 camera doubles."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "b200_dropin")
    if not os.access(exe, os.X_OK):
        pytest.skip("oracle/_ref/b200_dropin not built (needs /root/reference)")
    out = str(tmp_path / "arrays.bin")
    res = subprocess.run([exe, "arrays", name, "640", "480", out], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr
    data = open(out, "rb").read()
    t, s, m, _ = np.frombuffer(data, dtype="<u4", count=4)
    off = 16
    def take(dtype, count):
        nonlocal off
        arr = np.frombuffer(data, dtype=dtype, count=count, offset=off)
        off += arr.nbytes
        return arr
    env = take("<f8", 3)
    tri = take("<f8", int(t) * 9).reshape(-1, 9)
    tri_mat = take("<u4", int(t))
    sph = take("<f8", int(s) * 4).reshape(-1, 4)
    sph_mat = take("<u4", int(s))
    mats = take("<f8", int(m) * 9).reshape(-1, 9)
    cam = take("<f8", 18)
    want = scenes[name]
    assert np.array_equal(tri, want.triangle_vertices) and np.array_equal(sph, want.sphere_centre_radius)
    assert np.array_equal(env, want.environment)
    assert np.array_equal(mats[tri_mat], want.materials[want.triangle_material])
    assert np.array_equal(mats[sph_mat], want.materials[want.sphere_material])
    assert np.array_equal(cam, want.camera(640, 480))
