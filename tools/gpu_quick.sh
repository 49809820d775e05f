#!/bin/bash
TAG=r2l
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_tests.log
echo "== bench"; timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; head -c 250 $OUT/bench_${TAG}.json; echo; tail -2 $OUT/bench_${TAG}.err
echo "== sweep"; SWEEP_CONFIGS=128,207 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-160
SWEEP_CONFIGS=207 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py suzanne 640 480 16 2>&1 | cut -c1-160
