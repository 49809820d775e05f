#!/usr/bin/env python
"""Measures the keyed megakernel under each launch configuration (PTB200_KEYED_CONFIG) in
separate subprocesses, plus the sequential kernel; prints one JSON line per measurement.
Run on the GPU box:  python tools/sweep_configs.py [scene] [width] [height] [spp]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import json, os, sys
sys.path.insert(0, %(root)r)
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load(os.path.join(%(root)r, "tests/golden/scenes", %(scene)r + ".ptscene"))
w, h, spp, mode = %(w)d, %(h)d, %(spp)d, %(mode)d
ctx = capi.Context(0)
ctx.upload_scene(scene)
cam = scene.camera(w, h)
params = capi.make_params(w, h, spp=spp, seed=1)
opts = capi.make_options(rng_mode=mode)
ctx.render(cam, params, opts)
best = None
for _ in range(3):
    st = ctx.render(cam, params, opts)
    if best is None or st["sweep_kernel_ms"] < best["sweep_kernel_ms"]:
        best = st
ms = best["sweep_kernel_ms"]
print(json.dumps(dict(scene=%(scene)r, w=w, h=h, spp=spp, mode=mode,
    config=os.environ.get("PTB200_KEYED_CONFIG", "default"), ms=ms,
    msamples_s=best["samples"] / ms / 1e3, mcasts_s=best["casts"] / ms / 1e3,
    casts_per_sample=best["casts"] / best["samples"],
    logical_gbs=best["casts"] * scene.sweep_bytes() / ms / 1e6,
    fp64_tflops=best["casts"] * scene.sweep_flops() / ms / 1e9)))
"""


def run(scene, w, h, spp, mode, config):
    env = dict(os.environ, PTB200_KEYED_CONFIG=str(config))
    code = CHILD % dict(root=ROOT, scene=scene, w=w, h=h, spp=spp, mode=mode)
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(res.stdout.strip() or res.stderr.strip()[-400:], flush=True)


if __name__ == "__main__":
    scene = sys.argv[1] if len(sys.argv) > 1 else "cornell"
    w = int(sys.argv[2]) if len(sys.argv) > 2 else 640
    h = int(sys.argv[3]) if len(sys.argv) > 3 else 480
    spp = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    configs = [int(c) for c in os.environ.get("SWEEP_CONFIGS", "1,0,2,11,12,21,22").split(",")]
    for config in configs:
        run(scene, w, h, spp, 0, config)
    if os.environ.get("SWEEP_SEQUENTIAL", "1") == "1":
        run(scene, max(16, w // 8), max(12, h // 8), min(spp, 296), 1, 0)
