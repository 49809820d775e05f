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
@pytest.mark.parametrize("cooperative", [False, True])
def test_intersect_matches_oracle(name, cooperative, scenes, oracle, capi):
    scene = scenes[name]
    rays = random_rays(scene, 3000 if name != "ce" else 1500, seed=hash(name) % 1000)
    want = oracle.OracleScene(scene).intersect(rays)
    got = capi.intersect(scene, rays, warp_cooperative=cooperative)
    assert (want[:, 0] != 0).sum() > 100
    assert_hits_equal(got, want)


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
    ("suzanne", 32, 24, 2, 2, {}),
    ("single-sphere", 32, 24, 2, 3, {}),
    ("multi-sphere", 32, 24, 2, 4, {}),
    ("example1", 32, 24, 2, 5, {}),
    ("bbc-owl", 32, 24, 2, 6, {}),
    ("ce", 16, 9, 1, 7, {}),
]


@pytest.mark.parametrize("mode_name", ["keyed", "sequential"])
@pytest.mark.parametrize("case", RENDER_CASES, ids=lambda c: f"{c[0]}-{c[1]}x{c[2]}-{c[3]}spp")
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
