"""Multi-GPU paths (need >= 2 devices; skipped otherwise): the in-process ptb200_render_multi
(row-partitioned framebuffer, host-side gather) must equal the single-device render bit for bit
in keyed mode, and the pass-partitioned sequential mode must equal it to summation rounding."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def need_devices(capi, n):
    if capi.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")


def test_render_multi_keyed_equals_single_device(scenes, capi):
    need_devices(capi, 2)
    scene = scenes["cornell"]
    w, h, spp = 64, 50, 6
    cam = scene.camera(w, h)
    single, st1 = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4))
    for devices in ([0, 1], "all"):
        multi, stn = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4), devices=devices)
        assert np.array_equal(multi["sum"], single["sum"])
        assert np.array_equal(multi["n"], single["n"])
        assert stn["casts"] == st1["casts"] and stn["samples"] == st1["samples"]


def test_render_multi_sequential_partitions_passes(scenes, capi):
    need_devices(capi, 2)
    scene = scenes["cornell"]
    w, h, spp = 32, 24, 6
    cam = scene.camera(w, h)
    opts = capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL)
    single, st1 = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4), opts)
    multi, stn = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4), opts, devices=[0, 1])
    assert np.array_equal(multi["n"], single["n"]) and stn["casts"] == st1["casts"]
    # pass blocks are summed per device first, then across devices: association differs
    np.testing.assert_allclose(multi["sum"], single["sum"], rtol=1e-14, atol=1e-14)


def test_each_device_renders_the_same(scenes, capi):
    need_devices(capi, 2)
    scene = scenes["example1"]
    w, h = 40, 30
    cam = scene.camera(w, h)
    a, _ = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=9), capi.make_options(device=0))
    b, _ = capi.render(scene, cam, capi.make_params(w, h, spp=2, seed=9), capi.make_options(device=1))
    assert np.array_equal(a["sum"], b["sum"])


def test_render_multi_forwards_progress_on_the_calling_thread(scenes, capi):
    """ptb200_render_multi_progress: every device renders a slice of passes, the callback then sees
    the WHOLE partial frame (all rows, n == passes so far) on the caller's thread; the final frame
    equals the single-device render bit for bit."""
    import threading
    need_devices(capi, 2)
    scene = scenes["cornell"]
    w, h, spp = 48, 35, 10
    cam = scene.camera(w, h)
    seen = []

    def progress(user, pixels, done, total):
        arr = np.ctypeslib.as_array((capi.C.c_uint8 * (w * h * 32)).from_address(pixels)).view(capi.PIXEL_DTYPE)
        seen.append((threading.get_ident(), done, total, int(arr["n"].min()), int(arr["n"].max())))
        return 0

    multi, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4),
                            capi.make_options(passes_per_batch=4), progress=progress, devices=[0, 1])
    assert [s[1] for s in seen] == [4, 8, 10] and all(s[2] == spp for s in seen)
    assert all(s[0] == threading.get_ident() for s in seen)
    assert all(s[3] == s[4] == s[1] for s in seen)
    single, st1 = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=4))
    assert np.array_equal(multi["sum"], single["sum"]) and np.array_equal(multi["n"], single["n"])
    assert st["casts"] == st1["casts"] and st["samples"] == st1["samples"]
