"""Builds and runs the C++ host-side tests (pt_three_ways_b200/host/host_tests.cpp): the
reference's host-level test cases (OBJ loader, ArrayOutput) restated against our host types,
scene recipes + loader against the fixtures the reference's own loader produced, the
SceneBuilder adaptor, and the no-CPU-fallback behaviour of render()."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "pt_three_ways_b200", "host")


def test_host_cpp_suite():
    if not os.path.exists(os.path.join(ROOT, "pt_three_ways_b200", "libptb200.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "pt_three_ways_b200", "csrc")], check=True,
                       capture_output=True)
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    res = subprocess.run([os.path.join(HOST, "host_tests"), "--scenes", "/root/reference/scenes",
                          "--fixtures", os.path.join(ROOT, "tests", "golden", "scenes")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert " 0 failures" in res.stdout


def test_cli_rejects_bad_arguments():
    exe = os.path.join(HOST, "pt_b200")
    if not os.path.exists(exe):
        subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 1 and "Missing output filename" in res.stderr
    res = subprocess.run([exe, "--way", "smallpt", "out.png"], capture_output=True, text=True)
    assert res.returncode == 1 and "Unknown way" in res.stderr
    res = subprocess.run([exe, "--scene", "nonesuch", "--scenes", "/nonexistent", "out.png"],
                         capture_output=True, text=True)
    assert res.returncode == 1 and "Unknown scene" in res.stderr


def test_cli_accepts_every_way_and_has_no_cpu_fallback(tmp_path):
    """--way dod|fp|oo are all accepted (main.cpp:345-366); without a CUDA device the render fails
    loudly instead of falling back to a CPU path."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "pt_three_ways_b200", "libptb200.so"))
    count = ctypes.c_int32(0)
    lib.ptb200_device_count(ctypes.byref(count))
    if count.value > 0:
        pytest.skip("a CUDA device is present")
    exe = os.path.join(HOST, "pt_b200")
    scene = os.path.join(ROOT, "tests", "golden", "scenes", "cornell.ptscene")
    for way in ("dod", "fp", "oo"):
        res = subprocess.run([exe, "--way", way, "--ptscene", scene, "--raw", "-w", "8", "-h", "6", "--spp", "1",
                              "--seed", "1", str(tmp_path / "out.raw")], capture_output=True, text=True)
        assert res.returncode == 1 and "no CPU fallback" in res.stderr, res.stderr
        assert "38 triangles and 1 spheres" in res.stdout
        assert not (tmp_path / "out.raw").exists()


def test_merge_mode_sums_raw_framebuffers(tmp_path):
    """pt_b200 --merge a.raw b.raw out.png (what the reference's raw_to_png tool is for,
    src/main/raw_to_png.cpp:39-59): the sum is saved as a valid PNG or, with --raw, as a raw file
    again; the reference's own raw file (tests/golden) is accepted as input."""
    import struct
    import zlib
    import numpy as np
    exe = os.path.join(HOST, "pt_b200")
    subprocess.run(["make", "-C", HOST], check=True, capture_output=True)
    golden = os.path.join(ROOT, "tests", "golden", "render_asis_cornell.raw")
    out = str(tmp_path / "merged.png")
    res = subprocess.run([exe, "--merge", golden, golden, out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert "Merged 2 framebuffers of 16x16: 7680 samples, 30 per pixel" in res.stdout
    data = open(out, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    w, h = struct.unpack(">II", data[16:24])
    assert (w, h) == (16, 16)
    # decode the single IDAT chunk and compare with ArrayOutput::pixelAt's formula
    pos, idat = 8, b""
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        if kind == b"IDAT":
            idat += data[pos + 8:pos + 8 + n]
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(16, 1 + 16 * 3)
    assert (rows[:, 0] == 0).all()
    from oracle import oracle_binding as ob
    sums, counts = ob.read_raw(golden)
    mean = sums / counts[..., None]  # doubling sums and counts leaves the mean unchanged
    want = np.round(np.clip(mean, 0, 1) ** (1 / 2.2) * 255).astype(np.uint8)
    assert np.abs(rows[:, 1:].reshape(16, 16, 3).astype(int) - want.astype(int)).max() <= 1
    # --raw: the merged framebuffer in the reference's raw format, sums and counts doubled
    raw_out = str(tmp_path / "merged.raw")
    res = subprocess.run([exe, "--merge", golden, golden, "--raw", raw_out], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    sums2, counts2 = ob.read_raw(raw_out)
    assert np.array_equal(sums2, sums + sums) and np.array_equal(counts2, counts * 2)
    res = subprocess.run([exe, "--merge", golden, str(tmp_path / "missing.raw"), out], capture_output=True, text=True)
    assert res.returncode == 1 and "Error" in res.stderr
