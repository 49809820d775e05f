#!/bin/bash
# Verification session of the round's final build (default = sign-bit stage 0, variant 6): smoke,
# every GPU test, fp-way rates per instantiation, both bench arms, the ncu launch list and one full
# capture of the megakernel with the DEFAULT configuration.
# Usage: gpurun -- 'bash tools/gpu_r1t.sh r1t'
TAG=${1:-r1t}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1 | tee $OUT/${TAG}_gpu.txt
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee $OUT/${TAG}_tests.log
echo "== fp way rate per instantiation"
timeout 150 python - <<'PY' 2>&1 | tee $OUT/${TAG}_fp_rates.jsonl
import json, os, subprocess, sys
code = """
import json, os, sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
ctx = capi.Context(0)
ctx.upload_scene(scene)
w, h, spp = 640, 480, 64
cam = scene.camera(w, h)
best = 0.0
for _ in range(3):
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    best = max(best, st["samples"] / st["kernel_ms"] / 1e3)
print(json.dumps({"mode": "fp", "config": os.environ.get("PTB200_KEYED_CONFIG", "auto"), "msamples_s": round(best, 2)}))
"""
for config in ("24", "5", "6", "26"):
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PTB200_KEYED_CONFIG=config),
                         capture_output=True, text=True)
    print(res.stdout.strip() or res.stderr.strip()[-300:], flush=True)
PY
echo "== bench (ours)"
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 700 $OUT/bench_${TAG}.json; tail -2 $OUT/bench_${TAG}.err
echo "== bench (reference arm)"
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; tail -c 500 $OUT/bench_ref_${TAG}.json
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -2 $OUT/launches_${TAG}.csv
echo "== ncu full (BENCH_SPP=16)"
BENCH_SPP=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f -o $OUT/prof_keyed_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -1 $OUT/ncu_full_${TAG}.log
