#!/bin/bash
# Round-2 session g: constant-bank stage 0 with a rolled loop (variant 8) vs shared memory (7) vs unrolled (9).
TAG=r2g
OUT=gpurun_out
mkdir -p $OUT
echo "== config identity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "configuration" 2>&1 | tail -3 | tee $OUT/${TAG}_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=127,128,148,147,129 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== ncu full (variant 8, 256x3)"
PTB200_KEYED_CONFIG=128 BENCH_SPP=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPathKernel -c 1 -f -o $OUT/prof_subpath_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -2 $OUT/ncu_full_${TAG}.log
