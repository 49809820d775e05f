"""The oracle (oracle/pt_oracle.cpp) pinned against the REFERENCE: golden vectors produced by
tests/golden/make_golden.py from the reference's own unmodified sources, and the reference's
own known-answer tests (test/dod/*.cpp) restated without Catch2.  CPU only."""
import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import FP_PASS_CASES, OO_PASS_CASES, PASS_CASES, SCENES


def load_npy(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---- rendered pixels: per-pass images of the reference's radiance()/randomRay() ---------------
@pytest.mark.parametrize("case", PASS_CASES, ids=lambda c: c[0])
def test_pass_image_matches_reference(case, scenes, oracle, golden_dir):
    name, scene_name, w, h, seed, p, fu, fv, depth, preview = case
    want = load_npy(golden_dir, f"pass_{name}.npy")
    scene = scenes[scene_name]
    osc = oracle.OracleScene(scene)
    params = oracle.params_array(w, h, spp=1, seed=seed, max_depth=depth, first_u=fu, first_v=fv,
                                 preview=preview)
    got = osc.render(scene.camera(w, h), params, oracle.RNG_MT19937_SEQUENTIAL, pass_begin=p,
                     num_passes=1, per_pass=True)["per_pass"][0]
    # Tolerance: the reference build contracts FMAs as GCC pleases (-funsafe-math-optimizations)
    # and uses glibc sin/cos; the oracle fixes one rounding sequence.  Pixel values are sums of
    # products of material constants, so equal paths give equal pixels up to summation
    # rounding: 1e-12 absolute on values <= ~20.  (Measured: exactly 0.)
    assert np.abs(got - want).max() <= 1e-12


def test_unmodified_render_is_sum_of_first_passes(scenes, oracle, golden_dir):
    """dod::Scene::render as driven by test/seed_tests.sh (16x16, 16 spp, --max-cpus 1, seed 1)
    returns the sum of the passes it collected; with one worker those are passes 0..n-1
    (Scene.cpp:234-251 drops the in-flight tail)."""
    sums, counts = oracle.read_raw(os.path.join(golden_dir, "render_asis_cornell.raw"))
    kept = int(counts[0, 0])
    assert (counts == kept).all() and 1 <= kept <= 16
    scene = scenes["cornell"]
    got = oracle.OracleScene(scene).render(scene.camera(16, 16), oracle.params_array(16, 16, spp=kept, seed=1),
                                           oracle.RNG_MT19937_SEQUENTIAL)
    assert np.array_equal(got["counts"], counts.astype(np.uint64))
    assert np.abs(got["sums"] - sums).max() <= 1e-11  # 15 passes of values <= ~20


# ---- the reference's `fp` way (src/fp/Render.cpp): mt19937 per (pass, pixel) -------------------
@pytest.mark.parametrize("case", FP_PASS_CASES, ids=lambda c: c[0])
def test_fp_way_pass_image_matches_reference(case, scenes, oracle, golden_dir):
    """Golden: fp::render(spp=1, maxCpus=1) of the reference built from its own sources."""
    name, scene_name, w, h, seed, fu, fv, depth, preview = case
    want = load_npy(golden_dir, f"fp_pass_{name}.npy")
    scene = scenes[scene_name]
    params = oracle.params_array(w, h, spp=1, seed=seed, max_depth=depth, first_u=fu, first_v=fv, preview=preview)
    got = oracle.OracleScene(scene).render(scene.camera(w, h), params, oracle.RNG_FP_PER_PIXEL, num_passes=1,
                                           per_pass=True)["per_pass"][0]
    assert np.abs(got - want).max() <= 1e-12  # measured: exactly 0


def test_fp_way_unmodified_render_is_the_sum_of_its_passes(scenes, oracle, golden_dir):
    """fp::render keeps every pass (no dropped tail, src/fp/Render.cpp:146-164); with --max-cpus 1
    pass s uses seed + s and the passes are added in order."""
    sums, counts = oracle.read_raw(os.path.join(golden_dir, "fp_render_cornell.raw"))
    assert (counts == 5).all()
    scene = scenes["cornell"]
    got = oracle.OracleScene(scene).render(scene.camera(16, 16), oracle.params_array(16, 16, spp=5, seed=3),
                                           oracle.RNG_FP_PER_PIXEL)
    assert np.array_equal(got["counts"], counts.astype(np.uint64))
    assert np.abs(got["sums"] - sums).max() <= 1e-11


def test_fp_way_engine_seed_uses_the_reference_indexing(scenes, oracle):
    """x*width + y (not y*width + x): pixels (x, y) and (y', x') with x*W + y == x'*W + y' share an
    engine within a pass, e.g. (0, 1) and ... none inside a W x H frame with H <= W; but pass s+1
    starts H*W further on, so (x, y, s) and (x - H, y, s + 1) collide when W > H: same engine, but
    a different pixel position, hence a different camera ray.  Here: the colours differ, and the
    whole image differs from the dod way's."""
    scene = scenes["cornell"]
    w, h = 16, 8
    cam = scene.camera(w, h)
    osc = oracle.OracleScene(scene)
    fp = osc.render(cam, oracle.params_array(w, h, spp=2, seed=1), oracle.RNG_FP_PER_PIXEL, per_pass=True)
    dod = osc.render(cam, oracle.params_array(w, h, spp=2, seed=1), oracle.RNG_MT19937_SEQUENTIAL, per_pass=True)
    assert not np.array_equal(fp["per_pass"][0], dod["per_pass"][0])
    assert fp["casts"] > 0 and abs(fp["casts"] / dod["casts"] - 1) < 0.2


# ---- the reference's `oo` way (src/oo/Renderer.cpp): dod's stream, post-average emission -------
@pytest.mark.parametrize("case", OO_PASS_CASES, ids=lambda c: c[0])
def test_oo_way_pass_image_matches_reference(case, scenes, oracle, golden_dir):
    """Golden: the per-pass lambda of oo::Renderer::render (Renderer.cpp:97-107) driving the
    reference's own oo::Renderer::radiance(), built from src/oo/*.cpp."""
    name, scene_name, w, h, seed, p, fu, fv, depth, preview = case
    want = load_npy(golden_dir, f"oo_pass_{name}.npy")
    scene = scenes[scene_name]
    params = oracle.params_array(w, h, spp=1, seed=seed, max_depth=depth, first_u=fu, first_v=fv, preview=preview)
    got = oracle.OracleScene(scene).render(scene.camera(w, h), params, oracle.RNG_OO_SEQUENTIAL, pass_begin=p,
                                           num_passes=1, per_pass=True)["per_pass"][0]
    assert np.abs(got - want).max() <= 1e-12  # measured: exactly 0


def test_oo_way_unmodified_render_is_the_sum_of_the_passes_it_kept(scenes, oracle, golden_dir):
    """oo::Renderer::render has dod's scheduler (Renderer.cpp:109-141): it leaves its loop once the
    last pass has been LAUNCHED, so with --max-cpus 1 it returns passes 0..k-1 for some k <= spp."""
    sums, counts = oracle.read_raw(os.path.join(golden_dir, "oo_render_asis_cornell.raw"))
    kept = int(counts[0, 0])
    assert (counts == kept).all() and 1 <= kept <= 6
    scene = scenes["cornell"]
    got = oracle.OracleScene(scene).render(scene.camera(16, 16), oracle.params_array(16, 16, spp=kept, seed=2),
                                           oracle.RNG_OO_SEQUENTIAL)
    assert np.array_equal(got["counts"], counts.astype(np.uint64))
    assert np.abs(got["sums"] - sums).max() <= 1e-11


def test_oo_way_differs_from_dod_only_in_rounding(scenes, oracle):
    """Same stream, same strata order, same hits: the two estimators are algebraically equal
    (emission + mean(terms) vs mean(emission + terms)), so the images agree to rounding but not
    bit for bit — which is why `oo` is its own policy and not an alias of the sequential one."""
    scene = scenes["cornell"]
    w, h = 32, 24
    cam = scene.camera(w, h)
    osc = oracle.OracleScene(scene)
    oo = osc.render(cam, oracle.params_array(w, h, spp=2, seed=1), oracle.RNG_OO_SEQUENTIAL)
    dod = osc.render(cam, oracle.params_array(w, h, spp=2, seed=1), oracle.RNG_MT19937_SEQUENTIAL)
    assert oo["casts"] == dod["casts"] and oo["rng_words"] == dod["rng_words"]
    diff = np.abs(oo["sums"] - dod["sums"]).max()
    assert 0 < diff <= 1e-12


# ---- intersection records --------------------------------------------------------------------
@pytest.mark.parametrize("name", SCENES)
def test_intersections_match_reference(name, scenes, oracle, golden_dir):
    data = np.load(os.path.join(golden_dir, f"hits_{name}.npz"))
    rays, want = data["rays"], data["hits"]
    scene = scenes[name]
    got = oracle.OracleScene(scene).intersect(rays)
    assert np.array_equal(got[:, 0], want[:, 0])  # same hit / miss decisions
    hit = want[:, 0] != 0
    assert np.array_equal(got[hit, 2], want[hit, 2])  # inside
    np.testing.assert_allclose(got[hit, 1], want[hit, 1], rtol=1e-13, atol=1e-13)  # distance
    np.testing.assert_allclose(got[hit, 3:9], want[hit, 3:9], rtol=0, atol=1e-12)  # position, normal
    mats = scene.materials[got[hit, 9].astype(int)]
    assert np.array_equal(mats, want[hit, 9:18])  # the material the reference returned


# ---- the reference's own known-answer tests (test/dod/SphereTests.cpp, SceneTests.cpp,
# ---- TriangleTests.cpp), tolerances as Catch's Approx / ApproxVec3 (1e-4) --------------------
def ray_to(origin, target):
    o = np.array(origin, dtype=float)
    d = np.array(target, dtype=float) - o
    return np.concatenate([o, d / np.linalg.norm(d)])


def test_reference_sphere_kats(oracle):
    from pt_three_ways_b200 import scenefile
    inf = float("inf")
    s = oracle.OracleScene(scenefile.make(spheres=[((10, 20, 30), 15, 0)]))
    # test/dod/SphereTests.cpp:15-33
    assert s.intersect(ray_to((0, 0, 0), (0, 1, 0)), which=1)[0, 0] == 0
    assert s.intersect(ray_to((0, 0, 0), (-10, -20, -30)), which=1)[0, 0] == 0
    r = s.intersect(ray_to((0, 0, 0), (10, 20, 30)), which=1)[0]
    assert r[0] == 1 and r[2] == 0
    assert r[1] == pytest.approx(22.416738, rel=1e-5)
    np.testing.assert_allclose(r[3:6], (5.99108, 11.9822, 17.9732), atol=1e-4)
    np.testing.assert_allclose(r[6:9], (-0.267261, -0.534522, -0.801784), atol=1e-4)
    assert s.intersect(ray_to((0, 0, 0), (10, 20, 30)), which=1, nearer_than=22.0)[0, 0] == 0
    # SphereTests.cpp:35-52
    s = oracle.OracleScene(scenefile.make(spheres=[((0, 0, 30), 10, 0)]))
    r = s.intersect(ray_to((0, 0, 0), (0, 0, 2)), which=1, nearer_than=inf)[0]
    assert r[1] == 20 and r[2] == 0
    np.testing.assert_allclose(r[3:6], (0, 0, 20), atol=1e-4)
    np.testing.assert_allclose(r[6:9], (0, 0, -1), atol=1e-4)
    r = s.intersect(ray_to((0, 0, 30), (0, 0, 2)), which=1)[0]
    assert r[1] == 10 and r[2] == 1
    np.testing.assert_allclose(r[3:6], (0, 0, 20), atol=1e-4)
    np.testing.assert_allclose(r[6:9], (0, 0, 1), atol=1e-4)


def test_reference_scene_kats_nearer_of_two_spheres(oracle):
    from pt_three_ways_b200 import scenefile
    mats = [[0, 0, 0, 1, 1, 1, 1.0, -1.0, 0.0], [0, 0, 0, 1, 0, 0, 1.0, -1.0, 0.0]]
    # test/dod/SceneTests.cpp:53-79: either insertion order, the nearer sphere's material
    for first, second, expect in (((0, 0, 30), (0, 0, 90), 0), ((0, 0, 90), (0, 0, 30), 1)):
        s = oracle.OracleScene(scenefile.make(spheres=[(first, 10, 0), (second, 10, 1)], materials=mats))
        r = s.intersect(ray_to((0, 0, 0), (0, 0, 2)))[0]
        assert r[0] == 1 and r[1] == 20 and int(r[9]) == expect


def test_reference_triangle_kats(oracle):
    from pt_three_ways_b200 import scenefile
    # test/dod/TriangleTests.cpp:15-42: both windings
    for tri in (((0, 0, 3), (0, 1, 3), (1, 1, 3)), ((0, 0, 3), (1, 1, 3), (0, 1, 3))):
        s = oracle.OracleScene(scenefile.make(triangles=[(*tri, 0)]))
        assert s.intersect(ray_to((0, 0, 0), (0, 1, 0)), which=2)[0, 0] == 0
        assert s.intersect(ray_to((0, 0, 0), (0, 0, -1)), which=2)[0, 0] == 0
        r = s.intersect(ray_to((0, 0, 0), (0, 0, 1)), which=2)[0]
        assert r[0] == 1 and r[1] == pytest.approx(3.0)
        np.testing.assert_allclose(r[3:6], (0, 0, 3), atol=1e-4)
        np.testing.assert_allclose(r[6:9], (0, 0, -1), atol=1e-4)
    s = oracle.OracleScene(scenefile.make(triangles=[((0, 0, 3), (0, 1, 3), (1, 1, 3), 0)]))
    assert s.intersect(ray_to((0, 0, 0), (0, 0, 1)), which=2, nearer_than=2.999)[0, 0] == 0


# ---- third-party arithmetic the path relies on ------------------------------------------------
def test_mt19937_known_answer(oracle):
    out = np.zeros(10000, dtype=np.uint32)
    oracle.lib().oracle_mt19937(5489, 10000, out.ctypes.data)
    assert int(out[-1]) == 4123659995  # ISO C++ [rand.predef]: 10000th value of default mt19937


def test_canonical_double_formula(oracle):
    # libstdc++ generate_canonical<double,53>: (lo + hi*2^32) / 2^64, rounded once, < 1
    rng = np.random.default_rng(0)
    words = rng.integers(0, 2**32, size=(2000, 2), dtype=np.uint64)
    words[0] = (0xFFFFFFFF, 0xFFFFFFFF)
    words[1] = (0, 0)
    for lo, hi in words[:200]:
        want = float((int(hi) << 32) | int(lo)) / 2.0**64
        if want >= 1.0:
            want = np.nextafter(1.0, 0.0)
        assert oracle.lib().oracle_canonical(int(lo), int(hi)) == want


def test_philox_known_answers(oracle):
    # Random123 known-answer vectors for philox4x32-10
    cases = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
             ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
             ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
              (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in cases:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        out = np.zeros(4, dtype=np.uint32)
        oracle.lib().oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(v) for v in out) == want


def test_elementary_functions_close_to_libm(oracle):
    x = np.concatenate([np.linspace(-np.pi, 2 * np.pi, 20001), np.random.default_rng(1).uniform(0, 2 * np.pi, 20000)])
    s, c = np.zeros_like(x), np.zeros_like(x)
    oracle.lib().oracle_sincos(x.shape[0], x.ctypes.data, s.ctypes.data, c.ctypes.data)
    xl = x.astype(np.longdouble)
    # absolute error in units of 2^-53 (results are <= 1): below 1.5 "ulp of one"
    assert np.abs(s - np.sin(xl)).max() < 1.5 * 2.0**-53
    assert np.abs(c - np.cos(xl)).max() < 1.5 * 2.0**-53
    u = np.concatenate([np.linspace(0, 1, 20001)[:-1], np.random.default_rng(2).uniform(0, 1, 20000)])
    a = np.zeros_like(u)
    oracle.lib().oracle_acos(u.shape[0], u.ctypes.data, a.ctypes.data)
    assert np.abs(a - np.arccos(u.astype(np.longdouble))).max() < 4 * 2.0**-52


# ---- cameras: the PTSCENE2 camera block resized == the reference constructor ------------------
def test_camera_resize_matches_reference_constructor(scenes, golden_dir):
    cameras = json.load(open(os.path.join(golden_dir, "cameras.json")))
    for key, tokens in cameras.items():
        name, size = key.split(":")
        w, h = map(int, size.split("x"))
        want = np.array([float.fromhex(t) for t in tokens])
        assert np.array_equal(scenes[name].camera(w, h), want), key


def test_scene_counts(scenes):
    # 38/1, 970/2, 3442/3 through the reference loader
    assert (scenes["cornell"].num_triangles, scenes["cornell"].num_spheres) == (38, 1)
    assert (scenes["suzanne"].num_triangles, scenes["suzanne"].num_spheres) == (970, 2)
    assert (scenes["ce"].num_triangles, scenes["ce"].num_spheres) == (3442, 3)
    assert scenes["cornell"].sweep_bytes() == 2768 and scenes["cornell"].sweep_flops() == 1764


# ---- keyed policy sanity: an unbiased estimate of the same image ------------------------------
def test_keyed_policy_agrees_statistically_with_sequential(scenes, oracle):
    scene = scenes["cornell"]
    w, h, spp = 24, 18, 24
    osc = oracle.OracleScene(scene)
    cam = scene.camera(w, h)
    a = osc.render(cam, oracle.params_array(w, h, spp=spp, seed=1), oracle.RNG_MT19937_SEQUENTIAL,
                   threads=4, per_pass=True)
    b = osc.render(cam, oracle.params_array(w, h, spp=spp, seed=1), oracle.RNG_KEYED_PHILOX,
                   threads=4, per_pass=True)
    # same estimator, different random numbers: the per-pass image means are two samples of
    # one distribution -> two-sample z-test on their averages at 4.5 sigma
    pa = a["per_pass"].reshape(spp, -1).mean(axis=1)
    pb = b["per_pass"].reshape(spp, -1).mean(axis=1)
    sigma = np.sqrt(pa.var(ddof=1) / spp + pb.var(ddof=1) / spp)
    assert abs(pa.mean() - pb.mean()) < 4.5 * sigma
    assert abs(a["casts"] - b["casts"]) < 0.03 * a["casts"]
    # and the keyed policy is deterministic and independent of thread count / row partition
    c = osc.render(cam, oracle.params_array(w, h, spp=spp, seed=1), oracle.RNG_KEYED_PHILOX, threads=1)
    assert np.array_equal(b["sums"], c["sums"])
    even = osc.render(cam, oracle.params_array(w, h, spp=spp, seed=1), oracle.RNG_KEYED_PHILOX, row_begin=0, row_step=2)
    odd = osc.render(cam, oracle.params_array(w, h, spp=spp, seed=1), oracle.RNG_KEYED_PHILOX, row_begin=1, row_step=2)
    assert np.array_equal(even["sums"][0::2], b["sums"][0::2]) and np.array_equal(odd["sums"][1::2], b["sums"][1::2])
    assert (even["sums"][1::2] == 0).all()
