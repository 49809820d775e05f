#!/bin/bash
# Round-2 session f: stage 0 out of the constant bank (sweep variant 9) vs shared memory (7).
TAG=r2f
OUT=gpurun_out
mkdir -p $OUT
echo "== config identity + parity subset"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "configuration or render_matches or random_scene" 2>&1 | tail -4 | tee $OUT/${TAG}_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=127,129,149,169,109 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== ncu full (variant 9, 256x3)"
PTB200_KEYED_CONFIG=129 BENCH_SPP=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPathKernel -c 1 -f -o $OUT/prof_subpath_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -2 $OUT/ncu_full_${TAG}.log
