"""Builds and runs the C++ host-side tests (pt_three_ways_b200/host/host_tests.cpp): the
reference's host-level test cases (OBJ loader, ArrayOutput) restated against our host types,
scene recipes + loader against the fixtures the reference's own loader produced, the
SceneBuilder adaptor, and the no-CPU-fallback behaviour of render()."""
import os
import subprocess

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
    res = subprocess.run([exe, "--way", "oo", "out.png"], capture_output=True, text=True)
    assert res.returncode == 1 and "Unknown way" in res.stderr
    res = subprocess.run([exe, "--scene", "nonesuch", "--scenes", "/nonexistent", "out.png"],
                         capture_output=True, text=True)
    assert res.returncode == 1 and "Unknown scene" in res.stderr
