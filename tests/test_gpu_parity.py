"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Integer/index results must be identical; floating-point results are compared
bit-for-bit where the oracle and the kernels share one rounding sequence (everything here),
with the tolerance written next to each assertion."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SCENE_NAMES = ["cornell", "suzanne", "ce", "single-sphere", "multi-sphere", "example1", "bbc-owl"]


def random_rays(scene, n, seed):
    """Rays that start inside the scene's bounding region and point anywhere, plus rays that
    start ON primitives (the self-intersection case the 1e-9 epsilon exists for)."""
    rng = np.random.default_rng(seed)
    pts = []
    if scene.num_triangles:
        pts.append(scene.triangle_vertices.reshape(-1, 3))
    if scene.num_spheres:
        c = scene.sphere_centre_radius
        small = c[c[:, 3] < 50]
        pts.append(small[:, :3] + small[:, 3:4])
        pts.append(small[:, :3] - small[:, 3:4])
    pts = np.concatenate(pts)
    lo, hi = pts.min(0) - 0.5, pts.max(0) + 0.5
    origins = rng.uniform(lo, hi, size=(n, 3))
    if scene.num_triangles:
        # a third of the rays start on a random point of a random triangle
        k = n // 3
        tri = scene.triangle_vertices[rng.integers(0, scene.num_triangles, k)].reshape(k, 3, 3)
        a, b = rng.uniform(size=(2, k))
        flip = a + b > 1
        a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
        origins[:k] = tri[:, 0] + a[:, None] * (tri[:, 1] - tri[:, 0]) + b[:, None] * (tri[:, 2] - tri[:, 0])
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([origins, d], axis=1)


def assert_hits_equal(got, want):
    hit = want[:, 0] != 0
    assert np.array_equal(got["hit"] != 0, hit)
    g = got[hit]
    w = want[hit]
    assert np.array_equal(g["primitive"], w[:, 10].astype(np.int32))
    assert np.array_equal(g["material"], w[:, 9].astype(np.int32))
    assert np.array_equal(g["inside"] != 0, w[:, 2] != 0)
    # bit-exact: same IEEE operation sequence on both sides
    assert np.array_equal(g["distance"], w[:, 1])
    assert np.array_equal(g["position"], w[:, 3:6])
    assert np.array_equal(g["normal"], w[:, 6:9])


@pytest.mark.parametrize("name", SCENE_NAMES)
@pytest.mark.parametrize("sweep", ["two-stage", "one-stage", "fp32-stage0", "fp32x2-stage0", "fp32x2-stage0-t",
                                   "fp32x2-signs-t", "fp32x2-signs", "fp32x2-moment", "warp-cooperative"])
def test_intersect_matches_oracle(name, sweep, scenes, oracle, capi):
    """Every sweep implementation (two-stage FP64 prefilter + exact, plain one-stage, FP32 stage 0 +
    exact, the sequential kernel's lane-strided sweep) returns the oracle's nearest hit bit for bit."""
    scene = scenes[name]
    rays = random_rays(scene, 3000 if name != "ce" else 1500, seed=sum(map(ord, name)))
    want = oracle.OracleScene(scene).intersect(rays)
    variant = {"two-stage": capi.SWEEP_TWO_STAGE_FP64, "one-stage": capi.SWEEP_ONE_STAGE,
               "fp32-stage0": capi.SWEEP_FP32_STAGE0, "fp32x2-stage0": capi.SWEEP_FP32X2_STAGE0,
               "fp32x2-stage0-t": capi.SWEEP_FP32X2_STAGE0_T, "fp32x2-signs-t": capi.SWEEP_FP32X2_SIGNS_T,
               "fp32x2-signs": capi.SWEEP_FP32X2_SIGNS, "fp32x2-moment": capi.SWEEP_FP32X2_MOMENT}.get(sweep)
    got = capi.intersect(scene, rays, warp_cooperative=sweep == "warp-cooperative", sweep=variant)
    assert (want[:, 0] != 0).sum() > 100
    assert_hits_equal(got, want)


@pytest.mark.parametrize("name", [n for n in SCENE_NAMES if n not in ("single-sphere", "multi-sphere")])
def test_fp32_stage0_never_rejects_an_exact_hit(name, scenes, capi):
    """The conservative FP32 filter audited on every (ray, triangle) pair: zero pairs may be
    accepted by the exact FP64 test yet rejected by stage 0, and the filter must be selective."""
    scene = scenes[name]
    n = 20000 if scene.num_triangles < 100 else 1500
    rays = random_rays(scene, n, seed=1234)
    # grazing rays: directions almost inside triangle planes stress the |det| ~ 0 branch
    tri = scene.triangle_vertices[np.random.default_rng(5).integers(0, scene.num_triangles, n // 4)].reshape(-1, 3, 3)
    inplane = tri[:, 1] - tri[:, 0] + 1e-7 * (tri[:, 2] - tri[:, 0])
    rays[: n // 4, 3:] = inplane / np.linalg.norm(inplane, axis=1, keepdims=True)
    for moment_form in (False, True):  # the classic FP32 form (variants 2-6) and the moment form (7)
        res = capi.audit_stage0(scene, rays, moment_form=moment_form)
        assert res["pairs"] == n * scene.num_triangles
        assert res["violations"] == 0
        assert res["accepts"] > 0 and res["survivors"] >= res["accepts"]
        assert res["survivors"] < 0.5 * res["pairs"]


@pytest.mark.parametrize("which", [1, 2])
def test_intersect_primitive_kinds_and_nearer_than(which, scenes, oracle, capi):
    scene = scenes["example1"]
    rays = random_rays(scene, 2000, seed=3)
    for limit in (float("inf"), 2.5):
        want = oracle.OracleScene(scene).intersect(rays, which=which, nearer_than=limit)
        for cooperative in (False, True):
            got = capi.intersect(scene, rays, which=which, nearer_than=limit,
                                 warp_cooperative=cooperative)
            assert_hits_equal(got, want)


RENDER_CASES = [
    # scene, width, height, spp, seed, kwargs
    ("cornell", 40, 30, 3, 1, {}),
    ("cornell", 33, 17, 2, 5, dict(first_u=2, first_v=3, max_depth=3)),
    ("cornell", 24, 18, 2, 9, dict(max_depth=1)),
    ("cornell", 24, 18, 1, 9, dict(preview=1)),
    ("cornell", 20, 15, 2, 4, dict(max_depth=6)),   # deeper than the reference default
    ("cornell", 20, 15, 2, 4, dict(max_depth=7)),   
    ("suzanne", 32, 24, 2, 2, {}),
    ("single-sphere", 32, 24, 2, 3, {}),
    ("multi-sphere", 32, 24, 2, 4, {}),
    ("example1", 32, 24, 2, 5, {}),
    ("bbc-owl", 32, 24, 2, 6, {}),
    ("ce", 16, 9, 1, 7, {}),
]


@pytest.mark.parametrize("mode_name", ["keyed", "sequential"])
@pytest.mark.parametrize("case", RENDER_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-{c[3]}spp-{len(c[5])}{c[5].get('max_depth', '')}")
def test_render_matches_oracle(case, mode_name, scenes, oracle, capi):
    name, w, h, spp, seed, kw = case
    mode = capi.RNG_KEYED_PHILOX if mode_name == "keyed" else capi.RNG_MT19937_SEQUENTIAL
    scene = scenes[name]
    camera = scene.camera(w, h)
    pixels, stats = capi.render(scene, camera, capi.make_params(w, h, spp=spp, seed=seed, **kw),
                                capi.make_options(rng_mode=mode))
    want = oracle.OracleScene(scene).render(camera, oracle.params_array(w, h, spp=spp, seed=seed, **kw),
                                            mode, threads=4)
    assert np.array_equal(pixels["n"], want["counts"])
    assert stats["casts"] == want["casts"]
    assert stats["samples"] == w * h * spp
    # bit-exact sums: identical paths, identical rounding sequence, passes added in order
    assert np.array_equal(pixels["sum"], want["sums"])


# The exact-stream policies with a pass shared by 4 / 8 / 16 / 32 lanes (PtRenderOptions.lanesPerPass):
# pass counts that leave the last warp and the last CTA partly filled, every estimator branch.
LANE_GROUP_CASES = [
    ("cornell", 24, 18, 9, 1, {}),
    ("cornell", 20, 15, 17, 5, dict(first_u=2, first_v=3, max_depth=3)),
    ("cornell", 16, 12, 5, 9, dict(max_depth=1)),
    ("cornell", 16, 12, 3, 9, dict(preview=1)),
    ("cornell", 12, 9, 6, 4, dict(max_depth=7)),
    ("multi-sphere", 24, 18, 7, 4, {}),
    ("suzanne", 16, 12, 3, 2, {}),
    ("ce", 8, 6, 2, 7, {}),
    # radiance() at the deepest level discards 6 x 121 words per hit: more than one mt19937 generation
    ("cornell", 8, 6, 5, 3, dict(first_u=11, first_v=11, max_depth=1)),
    ("cornell", 8, 6, 3, 3, dict(first_u=11, first_v=7, max_depth=2)),
    ("cornell", 8, 6, 3, 3, dict(max_depth=0)),  # camera draws only
]


@pytest.mark.parametrize("lanes", [4, 8, 16, 32])
@pytest.mark.parametrize("mode_name", ["sequential", "oo"])
@pytest.mark.parametrize("case", LANE_GROUP_CASES, ids=lambda c: f"{c[0]}-{c[3]}passes-{len(c[5])}{c[5].get('max_depth', '')}{c[5].get('first_u', '')}")
def test_sequential_lane_groups_match_oracle(case, mode_name, lanes, scenes, oracle, capi):
    name, w, h, spp, seed, kw = case
    mode, oracle_mode = ((capi.RNG_MT19937_SEQUENTIAL, oracle.RNG_MT19937_SEQUENTIAL) if mode_name == "sequential"
                         else (capi.RNG_MT19937_SEQUENTIAL_OO, oracle.RNG_OO_SEQUENTIAL))
    scene = scenes[name]
    camera = scene.camera(w, h)
    pixels, stats = capi.render(scene, camera, capi.make_params(w, h, spp=spp, seed=seed, **kw),
                                capi.make_options(rng_mode=mode, lanes_per_pass=lanes))
    want = oracle.OracleScene(scene).render(camera, oracle.params_array(w, h, spp=spp, seed=seed, **kw),
                                            oracle_mode, threads=4)
    assert np.array_equal(pixels["n"], want["counts"])
    assert stats["casts"] == want["casts"]
    assert np.array_equal(pixels["sum"], want["sums"])  # bit-exact


def test_sequential_lane_group_is_chosen_from_the_pass_count(scenes, capi):
    """Many passes on a small frame: the library picks a sub-warp group by itself, and the image is
    the one a warp per pass renders."""
    scene = scenes["cornell"]
    w, h, spp = 6, 4, 2500
    cam = scene.camera(w, h)
    params = capi.make_params(w, h, spp=spp, seed=11)
    auto, sa = capi.render(scene, cam, params, capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL))
    full, sf = capi.render(scene, cam, params, capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, lanes_per_pass=32))
    assert sa["casts"] == sf["casts"]
    assert np.array_equal(auto["sum"], full["sum"]) and (auto["n"] == spp).all()
    with pytest.raises(capi.Ptb200Error):
        capi.render(scene, cam, params, capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, lanes_per_pass=5))


# ---- size-independent properties at BASELINE.json's full sizes ---------------------------------
def test_full_size_config1_properties(scenes, capi):
    """CornellBox 640x480 @ 256 spp (BASELINE configs[1]) is too big for the CPU oracle, so it is
    checked through properties the domain offers."""
    scene = scenes["cornell"]
    w, h, spp, seed = 640, 480, 256, 1
    cam = scene.camera(w, h)
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed))
    whole = ctx.download().copy()
    assert (whole["n"] == spp).all()                       # exactly spp samples everywhere
    assert st["samples"] == w * h * spp
    per_sample = st["casts"] / st["samples"]
    assert 1.0 <= per_sample <= 65.0 and abs(per_sample - 44.6) < 0.3  # 44.60
    assert np.isfinite(whole["sum"]).all() and (whole["sum"] >= 0).all()
    # pass additivity: [0,100) then [100,256) accumulated == [0,256) in one go, bit for bit
    ctx.render(cam, capi.make_params(w, h, spp=100, seed=seed))
    ctx.render(cam, capi.make_params(w, h, spp=156, seed=seed), capi.make_options(pass_begin=100),
               accumulate=True)
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    # framebuffer partition: rows y % 4 == r rendered separately are the same pixels
    for r in range(4):
        ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(row_begin=r, row_step=4))
        part = ctx.download()
        assert np.array_equal(part["sum"][r::4], whole["sum"][r::4])
        others = np.ones(h, dtype=bool)
        others[r::4] = False
        assert (part["n"][others] == 0).all()
    # pass batching does not change results (ordered per-pixel accumulation)
    ctx.render(cam, capi.make_params(w, h, spp=spp, seed=seed), capi.make_options(passes_per_batch=37))
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    # seed semantics of test/seed_tests.sh: same seed -> identical, other seed -> different
    ctx.render(cam, capi.make_params(w, h, spp=8, seed=1))
    a = ctx.download().copy()
    ctx.render(cam, capi.make_params(w, h, spp=8, seed=1))
    assert np.array_equal(ctx.download()["sum"], a["sum"])
    ctx.render(cam, capi.make_params(w, h, spp=8, seed=2))
    assert not np.array_equal(ctx.download()["sum"], a["sum"])
    ctx.close()


def test_ce_is_the_constant_image(scenes, capi):
    """Config 3's scene: the camera sits inside a zero-albedo light, so every pixel is exactly
    that light's emission and every sample is exactly 65 casts."""
    scene = scenes["ce"]
    w, h, spp = 64, 36, 2
    for mode in (capi.RNG_KEYED_PHILOX, capi.RNG_MT19937_SEQUENTIAL):
        px, st = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=5),
                             capi.make_options(rng_mode=mode))
        assert st["casts"] == 65 * w * h * spp
        mean = px["sum"] / spp
        # one value everywhere (16 equal terms summed and scaled per sample, then spp samples)
        assert (mean == mean[0, 0]).all()
        np.testing.assert_allclose(mean[0, 0], np.array([2.27, 3, 2.97]) * 0.25, rtol=1e-15)


def test_suzanne_640x480_statistics(scenes, capi):
    scene = scenes["suzanne"]
    px, st = capi.render(scene, scene.camera(640, 480), capi.make_params(640, 480, spp=4, seed=1))
    assert abs(st["casts"] / st["samples"] - 21.7) < 0.3  # 21.70
    assert (px["n"] == 4).all()


def test_sequential_mode_pass_partition_matches_single_call(scenes, capi):
    scene = scenes["cornell"]
    w, h, spp = 32, 24, 6
    cam = scene.camera(w, h)
    opts = capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL)
    whole, _ = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=3), opts)
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    ctx.render(cam, capi.make_params(w, h, spp=2, seed=3), opts)
    ctx.render(cam, capi.make_params(w, h, spp=4, seed=3),
               capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL, pass_begin=2), accumulate=True)
    assert np.array_equal(ctx.download()["sum"], whole["sum"])
    ctx.close()


def test_progress_callback_runs_on_calling_thread_with_partial_frames(scenes, capi):
    import threading
    scene = scenes["cornell"]
    w, h, spp = 32, 24, 10
    seen = []

    def progress(user, pixels, done, total):
        arr = np.ctypeslib.as_array((capi.C.c_uint8 * (w * h * 32)).from_address(pixels)).view(capi.PIXEL_DTYPE)
        seen.append((threading.get_ident(), done, total, int(arr["n"].min()), int(arr["n"].max())))
        return 0

    final, _ = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=1),
                           capi.make_options(passes_per_batch=3), progress=progress)
    assert [s[1] for s in seen] == [3, 6, 9, 10] and all(s[2] == spp for s in seen)
    assert all(s[0] == threading.get_ident() for s in seen)
    assert all(s[3] == s[4] == s[1] for s in seen)
    ref, _ = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=1))
    assert np.array_equal(final["sum"], ref["sum"])


def test_zero_samples_with_a_progress_callback_returns_an_empty_frame(scenes, capi):
    """spp == 0 through the progressive path of a pooled context: the frame of THIS call,
    all zero, not the previous call's accumulator."""
    scene = scenes["cornell"]
    big, _ = capi.render(scene, scene.camera(40, 30), capi.make_params(40, 30, spp=2, seed=1))
    assert (big["n"] == 2).all()
    calls = []
    px, st = capi.render(scene, scene.camera(16, 12), capi.make_params(16, 12, spp=0, seed=1),
                         progress=lambda user, pixels, done, total: calls.append(done) or 0)
    assert st["samples"] == 0 and (px["n"] == 0).all() and (px["sum"] == 0).all()


@pytest.mark.parametrize("scene_name,w,h", [("suzanne", 96, 72), ("cornell", 128, 96), ("ce", 16, 9)])
def test_every_megakernel_configuration_renders_identically(scene_name, w, h, scenes, tmp_path):
    """Every render-kernel instantiation — the one-kernel form (1, 6, 26) and the three-kernel
    pipeline (100 + launch shape x sweep variant), PTB200_KEYED_CONFIG — must give the same
    framebuffer and cast count bit for bit."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "from pt_three_ways_b200 import capi, scenefile\n"
            "s = scenefile.load(%r)\n"
            "px, st = capi.render(s, s.camera(%d, %d), capi.make_params(%d, %d, spp=4, seed=11))\n"
            "np.save(sys.argv[1], px['sum']); print(st['casts'])\n") % (
                root, os.path.join(root, "tests/golden/scenes/%s.ptscene" % scene_name), w, h, w, h)
    outs = []
    configs = ["1", "6", "26", "101", "106", "121", "126", "107", "127", "137", "147",
               "207", "217", "227"]  # 2xx: two sub-paths per lane
    if scene_name == "cornell":  # stage 0 out of the constant bank: scenes of up to 64 triangles
        configs += ["128", "148", "168", "129"]
    for config in configs:
        out = str(tmp_path / f"c{config}.npy")
        res = subprocess.run([sys.executable, "-c", code, out], capture_output=True, text=True,
                             env=dict(os.environ, PTB200_KEYED_CONFIG=config), timeout=300)
        assert res.returncode == 0, res.stderr[-1500:]
        outs.append((np.load(out), res.stdout.strip()))
    for other in outs[1:]:
        assert np.array_equal(outs[0][0], other[0]) and outs[0][1] == other[1]


def test_cpp_host_adaptor_and_cli_render_identically(scenes, capi, tmp_path):
    """The C++ host side (ptb200::Scene adaptor + CLI driver with the reference's flags) drives
    the same C ABI: its raw output equals the ctypes render of the same scene, bit for bit."""
    import os
    import subprocess
    from oracle import oracle_binding as ob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "pt_three_ways_b200", "host")
    exe = os.path.join(host, "pt_b200")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", host], check=True, capture_output=True)
    fixture = os.path.join(root, "tests", "golden", "scenes", "cornell.ptscene")
    for flags, mode in ((("--rng", "keyed"), capi.RNG_KEYED_PHILOX), (("--rng", "exact"), capi.RNG_MT19937_SEQUENTIAL),
                        (("--way", "fp"), capi.RNG_MT19937_PER_PIXEL), (("--way", "oo"), capi.RNG_MT19937_SEQUENTIAL_OO)):
        out = str(tmp_path / f"cli_{flags[1]}.raw")
        res = subprocess.run([exe, "--ptscene", fixture, "-w", "48", "-h", "36", "--spp", "3", "--seed", "7",
                              "--save-every", "0", *flags, "--raw", out],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stderr
        assert "Scene contains 38 triangles and 1 spheres." in res.stdout
        assert "Total samples: %d" % (48 * 36 * 3) in res.stdout and "Samples/ms:" in res.stdout
        sums, counts = ob.read_raw(out)
        scene = scenes["cornell"]
        want, _ = capi.render(scene, scene.camera(48, 36), capi.make_params(48, 36, spp=3, seed=7),
                              capi.make_options(rng_mode=mode))
        assert np.array_equal(sums, want["sum"]) and (counts == 3).all()


def test_reference_dod_tests_through_the_cpp_adaptor():
    """test/dod/{Sphere,Scene,Triangle}Tests.cpp restated against ptb200::Scene (the drop-in for
    dod::Scene), executed on the GPU through the C ABI, plus render() semantics."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "pt_three_ways_b200", "host")
    subprocess.run(["make", "-C", host], check=True, capture_output=True)
    res = subprocess.run([os.path.join(host, "host_tests"), "--fixtures", os.path.join(root, "tests/golden/scenes")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " 0 failures" in res.stdout and "GPU KATs ran" in res.stdout


# ---- random scenes: every shading branch (mirror, glossy, Fresnel, emitters, inside spheres) ---
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scene_intersections_match_oracle(seed, oracle, capi):
    from tests import random_scenes
    scene = random_scenes.random_scene(seed, num_triangles=70)
    rays = random_scenes.rays_for(scene, 4000, seed + 20)
    want = oracle.OracleScene(scene).intersect(rays)
    for sweep in (capi.SWEEP_ONE_STAGE, capi.SWEEP_TWO_STAGE_FP64, capi.SWEEP_FP32_STAGE0,
                  capi.SWEEP_FP32X2_STAGE0, capi.SWEEP_FP32X2_STAGE0_T, capi.SWEEP_FP32X2_SIGNS_T,
                  capi.SWEEP_FP32X2_SIGNS, capi.SWEEP_FP32X2_MOMENT):
        assert_hits_equal(capi.intersect(scene, rays, sweep=sweep), want)
    assert_hits_equal(capi.intersect(scene, rays, warp_cooperative=True), want)
    for moment_form in (False, True):
        audit = capi.audit_stage0(scene, rays, moment_form=moment_form)
        assert audit["violations"] == 0 and audit["accepts"] > 0


@pytest.mark.parametrize("mode_name", ["keyed", "sequential"])
@pytest.mark.parametrize("seed,kw", [(4, {}), (5, dict(first_u=3, first_v=2, max_depth=7)), (6, dict(max_depth=2)),
                                     (8, dict(first_u=1, first_v=1, max_depth=12))])
def test_random_scene_renders_match_oracle(seed, kw, mode_name, oracle, capi):
    from tests import random_scenes
    mode = capi.RNG_KEYED_PHILOX if mode_name == "keyed" else capi.RNG_MT19937_SEQUENTIAL
    scene = random_scenes.random_scene(seed, num_triangles=25, num_spheres=3)
    w, h, spp = 28, 21, 3
    cam = scene.camera(w, h)
    got, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=13, **kw), capi.make_options(rng_mode=mode))
    want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=spp, seed=13, **kw), mode, threads=4)
    assert st["casts"] == want["casts"]
    assert np.array_equal(got["sum"], want["sums"])  # bit-exact
    assert np.array_equal(got["n"], want["counts"])


def test_camera_without_aperture_draws_two_numbers(oracle, capi):
    """Camera::rayFromUnit returns before drawing the lens sample when apertureRadius == 0
    (Camera.h:26-27): the sequential stream must then be two draws shorter per pixel."""
    from tests import random_scenes
    scene = random_scenes.random_scene(9, num_triangles=10, num_spheres=2)
    scene.camera64x48 = random_scenes.camera18((0, 0, -5), (0, 0, 0), (0, 1, 0), 64, 48, 40.0)
    w, h = 20, 15
    cam = scene.camera(w, h)
    assert cam[16] == 0.0
    for mode in (capi.RNG_KEYED_PHILOX, capi.RNG_MT19937_SEQUENTIAL):
        got, _ = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=5), capi.make_options(rng_mode=mode))
        want = oracle.OracleScene(scene).render(cam, oracle.params_array(w, h, spp=2, seed=5), mode)
        assert np.array_equal(got["sum"], want["sums"])
