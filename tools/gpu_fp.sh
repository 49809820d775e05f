#!/bin/bash
# GPU session for the fp way (PTB200_RNG_MT19937_PER_PIXEL): parity tests, throughput next to the
# keyed policy, and a memcheck pass over a small render.  Usage: gpurun -- 'bash tools/gpu_fp.sh TAG'
TAG=${1:-fp}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
timeout 900 python -m pytest tests/test_gpu_fp_way.py -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_tests.log
cat > /tmp/fp_rates.py <<'PY'
import json, os, sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
cases = (("cornell", 640, 480, 64), ("cornell", 640, 480, 256), ("suzanne", 640, 480, 32), ("ce", 320, 180, 8))
if os.environ.get("PTB200_KEYED_CONFIG"):
    cases = cases[1:2]
for name, w, h, spp in cases:
    scene = scenefile.load(f"tests/golden/scenes/{name}.ptscene")
    ctx = capi.Context(0)
    ctx.upload_scene(scene)
    cam = scene.camera(w, h)
    row = {"scene": name, "w": w, "h": h, "spp": spp, "config": os.environ.get("PTB200_KEYED_CONFIG", "auto")}
    for label, mode in (("keyed", capi.RNG_KEYED_PHILOX), ("fp", capi.RNG_MT19937_PER_PIXEL)):
        best = 0.0
        for _ in range(3):
            st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=mode))
            best = max(best, st["samples"] / st["kernel_ms"] / 1e3)
        row[label + "_msamples_s"] = round(best, 2)
        row[label + "_casts_per_sample"] = round(st["casts"] / st["samples"], 3)
    print(json.dumps(row), flush=True)
    ctx.close()
PY
(timeout 300 python /tmp/fp_rates.py; PTB200_KEYED_CONFIG=4 timeout 120 python /tmp/fp_rates.py) 2>&1 | tee gpurun_out/${TAG}_rates.jsonl
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python - <<'PY' 2>&1 | tail -6 | tee gpurun_out/${TAG}_memcheck.log
import sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
for name, w, h in (("cornell", 24, 18), ("ce", 8, 6)):
    scene = scenefile.load(f"tests/golden/scenes/{name}.ptscene")
    px, st = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=2, seed=1),
                         capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    print(name, st["casts"], float(px["sum"].mean()))
PY
echo "memcheck exit: $?"
