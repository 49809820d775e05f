#!/bin/bash
# Session r1u: register-pressure changes in the megakernel (strata sums and output slot in shared
# memory, per-CTA cast counter, lane flags in one register, shallow stacks in registers).
# Every GPU test, a configuration sweep, both defaults' bench lines, launch list + full capture.
# Usage: gpurun -- 'bash tools/gpu_r1u.sh r1u'
TAG=${1:-r1u}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1 | tee $OUT/${TAG}_gpu.txt
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee $OUT/${TAG}_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=26,6,46,56,24 SWEEP_SEQUENTIAL=0 timeout 200 python tools/sweep_configs.py cornell 640 480 32 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== sweep suzanne / ce"
SWEEP_CONFIGS=6,3 SWEEP_SEQUENTIAL=0 timeout 100 python tools/sweep_configs.py suzanne 640 480 4 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
SWEEP_CONFIGS=6,3 SWEEP_SEQUENTIAL=0 timeout 100 python tools/sweep_configs.py ce 320 180 1 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== bench (ours)"
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 300 $OUT/bench_${TAG}.json; tail -2 $OUT/bench_${TAG}.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_%s.json" % __import__("os").environ.get("TAG", "r1u")))
print("VALUE", d["value"], "E2E", d["e2e"]["value"], "FP", d["fp_way_mode"]["value"], "OO", d["oo_way_mode"]["value"])
PY
echo "== ncu launch list"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -2 $OUT/launches_${TAG}.csv
echo "== ncu full (BENCH_SPP=16)"
BENCH_SPP=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f -o $OUT/prof_keyed_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -1 $OUT/ncu_full_${TAG}.log
