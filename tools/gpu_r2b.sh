#!/bin/bash
# Round-2 session a: the three-kernel keyed pipeline (pt_split.cu) against the one-kernel form.
TAG=r2b
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=26,126,127,146,147,136,137,106,107 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== sweep suzanne"
SWEEP_CONFIGS=106,107,126,127 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py suzanne 640 480 8 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
echo "== sweep ce"
SWEEP_CONFIGS=106,107 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py ce 640 360 1 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== ncu full (sub-path kernel; BENCH_SPP=16 keeps the replays short)"
BENCH_SPP=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPathKernel -c 1 -f -o $OUT/prof_subpath_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -3 $OUT/ncu_full_${TAG}.log; ls -la $OUT | tail -5
