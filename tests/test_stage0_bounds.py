"""CPU check of the FP32 stage-0 filter's ARITHMETIC CONTRACT (csrc/pt_device.cuh: stage0Keep2 /
stage0Reject2, bounds built by buildFilterKernel in csrc/pt_kernels.cu), restated in numpy:

* soundness: a triangle the reference's FP64 Moller-Trumbore accepts (Scene.cpp:62-98, with no
  nearer-than limit) is never rejected by the FP32 filter with its per-triangle error bounds;
* the sign-bit formulation (variants 5/6: sign of (a|b|c|e) & f) takes exactly the decisions of
  the comparison formulation (variants 3/4);
* the moment (Pluecker) form of variant 7 (stage0RejectMoment) is sound under the same bounds.

float32 FMAs are emulated as float64 products/sums rounded once to float32 (products of two
float32 values are exact in float64; the sum adds one double rounding the GPU does not have, far
inside the 4x margin the bounds carry).  The GPU-side counterpart, on the device's own arithmetic,
is auditStage0Kernel (tests/test_gpu_parity.py)."""
import numpy as np
import pytest

EPS = 1e-9
F = np.float32


def fma32(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def mul32(a, b):
    return (a.astype(np.float64) * b.astype(np.float64)).astype(F)


def bounds(v0, e1, e2, origin_bound):
    """buildFilterKernel: Ed, 2Ex, 2Ey, K3, 2Et, rounded up to float32 with its slack."""
    c = 2.0 ** -18
    reach = origin_bound + np.linalg.norm(v0, axis=1)
    l1, l2 = np.linalg.norm(e1, axis=1), np.linalg.norm(e2, axis=1)
    ed, ex, ey, et = c * l1 * l2, c * l2 * reach, c * l1 * reach, c * l1 * l2 * reach
    slack = 1.0 + 2.0 ** -10
    up = lambda x: np.nextafter((x * slack).astype(F), F(np.inf))  # >= __double2float_ru
    return up(ed), up(2 * ex), up(2 * ey), up(ed * (1 + 2.0 ** -20) + ex + ey), up(2 * et)


def stage0(v0, e1, e2, o, d, origin_bound):
    """Returns (keep by comparisons, keep by sign bits) for variant 4/5 (with the negative-t test)
    and (…, …) for variant 3/6 (without), each an (N,) bool array; one ray per triangle row."""
    ed, kx, ky, k3, kt = bounds(v0, e1, e2, origin_bound)
    v0, e1, e2, o, d = (a.astype(F) for a in (v0, e1, e2, o, d))
    X, Y, Z = 0, 1, 2
    px = fma32(d[:, Y], e2[:, Z], -mul32(d[:, Z], e2[:, Y]))
    py = fma32(d[:, Z], e2[:, X], -mul32(d[:, X], e2[:, Z]))
    pz = fma32(d[:, X], e2[:, Y], -mul32(d[:, Y], e2[:, X]))
    det = fma32(e1[:, Z], pz, fma32(e1[:, Y], py, mul32(e1[:, X], px)))
    t = (o - v0).astype(F)
    x = fma32(t[:, Z], pz, fma32(t[:, Y], py, mul32(t[:, X], px)))
    qx = fma32(t[:, Y], e1[:, Z], -mul32(t[:, Z], e1[:, Y]))
    qy = fma32(t[:, Z], e1[:, X], -mul32(t[:, X], e1[:, Z]))
    qz = fma32(t[:, X], e1[:, Y], -mul32(t[:, Y], e1[:, X]))
    y = fma32(d[:, Z], qz, fma32(d[:, Y], qy, mul32(d[:, X], qx)))
    tt = fma32(e2[:, Z], qz, fma32(e2[:, Y], qy, mul32(e2[:, X], qx)))
    s = np.where(np.signbit(det), F(-1), F(1))
    adet = np.abs(det)
    xs, ys, ts = x * s, y * s, tt * s
    bound = fma32(adet, np.full_like(adet, F(1) + F(2.0 ** -20)), k3)
    # variants 3/4: comparisons
    certain = (xs < -kx) | (ys < -ky) | ((xs + ys).astype(F) > bound)
    keep_cmp = (adet <= ed) | ~certain
    keep_cmp_t = (adet <= ed) | ~(certain | (ts < -kt))
    # variants 5/6: sign bits of packed differences
    a, b, c = fma32(x, s, kx), fma32(y, s, ky), fma32(tt, s, kt)
    e = fma32(-(x + y).astype(F), s, bound)
    f = (ed - adet).astype(F)
    sign = lambda v: np.signbit(v)
    keep_sign = ~((sign(a) | sign(b) | sign(e)) & sign(f))
    keep_sign_t = ~((sign(a) | sign(b) | sign(c) | sign(e)) & sign(f))
    return keep_cmp, keep_sign, keep_cmp_t, keep_sign_t


def stage0_moment(v0, e1, e2, o, d, origin_bound):
    """Sweep variant 7 (stage0RejectMoment): det, X, Y as dot products of the ray's (d, m = o x d)
    with per-triangle constants evaluated in float64 and rounded once; the bounds and the sign-bit
    decision are the classic form's."""
    ed, kx, ky, k3, _ = bounds(v0, e1, e2, origin_bound)
    nn, a2, a1 = np.cross(e2, e1), np.cross(v0, e2), np.cross(v0, e1)
    m = np.cross(o, d)
    nn, e2f, a2, f1, b1, m, d = (a.astype(F) for a in (nn, e2, a2, -e1, -a1, m, d))
    X, Y, Z = 0, 1, 2
    det = fma32(d[:, Z], nn[:, Z], fma32(d[:, Y], nn[:, Y], mul32(d[:, X], nn[:, X])))

    def six(p, q):  # p . m + q . d in the kernel's operation order
        acc = fma32(m[:, Y], p[:, Y], mul32(m[:, X], p[:, X]))
        acc = fma32(m[:, Z], p[:, Z], acc)
        acc = fma32(d[:, X], q[:, X], acc)
        acc = fma32(d[:, Y], q[:, Y], acc)
        return fma32(d[:, Z], q[:, Z], acc)

    x, y = six(e2f, a2), six(f1, b1)
    s = np.where(np.signbit(det), F(-1), F(1))
    adet = np.abs(det)
    bound = fma32(adet, np.full_like(adet, F(1) + F(2.0 ** -20)), k3)
    a, b = fma32(x, s, kx), fma32(y, s, ky)
    e = fma32(-(x + y).astype(F), s, bound)
    f = (ed - adet).astype(F)
    return ~((np.signbit(a) | np.signbit(b) | np.signbit(e)) & np.signbit(f))


def exact_accepts(v0, e1, e2, o, d):
    """testTriangle with best.t = +inf: the reference's arithmetic in float64 (numpy has no FMA;
    a last-bit difference in u, v or t only matters within ~1e-16 of an edge)."""
    p = np.cross(d, e2)
    det = np.einsum("ij,ij->i", e1, p)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / det
        tv = o - v0
        u = np.einsum("ij,ij->i", tv, p) * inv
        q = np.cross(tv, e1)
        v = np.einsum("ij,ij->i", d, q) * inv
        t = np.einsum("ij,ij->i", e2, q) * inv
        reject = (np.abs(det) < EPS) | (u < 0) | (u > 1) | (v < 0) | (u + v > 1)
        return ~reject & (t > EPS)


def make_pairs(n, seed, scale):
    """(ray, triangle) pairs: a third of the rays aimed at a random point of their triangle (hits),
    a third at a point on an edge or a vertex (grazing), a third anywhere."""
    rng = np.random.default_rng(seed)
    v0 = rng.uniform(-scale, scale, (n, 3))
    e1 = rng.normal(size=(n, 3)) * rng.uniform(0.01, 1.0, (n, 1)) * scale
    e2 = rng.normal(size=(n, 3)) * rng.uniform(0.01, 1.0, (n, 1)) * scale
    o = rng.uniform(-scale, scale, (n, 3))
    a, b = rng.uniform(size=(2, n))
    flip = a + b > 1
    a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
    kind = rng.integers(0, 3, n)
    edge = kind == 1
    which = rng.integers(0, 4, n)
    a = np.where(edge & (which == 0), 0.0, a)            # on the v0-v2 edge
    b = np.where(edge & (which == 1), 0.0, b)            # on the v0-v1 edge
    b = np.where(edge & (which == 2), 1.0 - a, b)        # on the v1-v2 edge
    a = np.where(edge & (which == 3), 0.0, a)            # vertex v0 ...
    b = np.where(edge & (which == 3), 0.0, b)
    target = v0 + a[:, None] * e1 + b[:, None] * e2
    d = np.where((kind == 2)[:, None], rng.normal(size=(n, 3)), target - o)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return v0, e1, e2, o, d


@pytest.mark.parametrize("scale", [1.0, 30.0, 1e4])
def test_fp32_stage0_bounds_never_reject_an_exact_hit(scale):
    v0, e1, e2, o, d = make_pairs(400_000, seed=int(scale) + 3, scale=scale)
    origin_bound = float(np.linalg.norm(o, axis=1).max()) * (1 + 1e-6)
    keep_cmp, keep_sign, keep_cmp_t, keep_sign_t = stage0(v0, e1, e2, o, d, origin_bound)
    accepts = exact_accepts(v0, e1, e2, o, d)
    assert accepts.sum() > 100_000          # the aimed rays do hit
    assert (~keep_cmp).sum() > 50_000       # ... and the filter does reject the others
    keep_moment = stage0_moment(v0, e1, e2, o, d, origin_bound)
    assert (~keep_moment).sum() > 50_000
    for keep in (keep_cmp, keep_sign, keep_cmp_t, keep_sign_t, keep_moment):
        assert not (accepts & ~keep).any()  # soundness
    # the moment form is (slightly) the tighter filter: its rounding errors are about half
    assert keep_moment.sum() <= keep_sign.sum() * 1.001


@pytest.mark.parametrize("scale", [1.0, 30.0])
def test_sign_bit_formulation_takes_the_same_decisions(scale):
    v0, e1, e2, o, d = make_pairs(400_000, seed=int(scale) + 11, scale=scale)
    origin_bound = float(np.linalg.norm(o, axis=1).max()) * (1 + 1e-6)
    keep_cmp, keep_sign, keep_cmp_t, keep_sign_t = stage0(v0, e1, e2, o, d, origin_bound)
    assert np.array_equal(keep_cmp, keep_sign)
    assert np.array_equal(keep_cmp_t, keep_sign_t)


def test_padding_triangles_are_rejected_outright():
    """All-zero padding triangles carry Ed = 2Ex = 2Ey = -1, K3 = 2Et = 0 (buildFilterKernel): det32
    is +0, so f = Ed - |det32| = -1 and a = 0*s + 2Ex = -1 are both negative: rejected."""
    n = 8
    z = np.zeros((n, 3), dtype=F)
    det = np.zeros(n, dtype=F)
    ed = np.full(n, -1, dtype=F)
    kx = np.full(n, -1, dtype=F)
    s = np.ones(n, dtype=F)
    a = fma32(z[:, 0], s, kx)
    f = (ed - np.abs(det)).astype(F)
    assert (np.signbit(a) & np.signbit(f)).all()
    assert not ((np.abs(det) <= ed) | ~(z[:, 0] < -kx)).any()  # the comparison form agrees
