#!/bin/bash
# Round-2 session h: two sub-paths per lane (configs 2x7) on the sweep-bound scenes; variant 8 shapes on Cornell.
TAG=r2h
OUT=gpurun_out
mkdir -p $OUT
echo "== config identity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "configuration" 2>&1 | tail -3 | tee $OUT/${TAG}_tests.log
echo "== sweep suzanne"
SWEEP_CONFIGS=107,207,217,227 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py suzanne 640 480 16 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
echo "== sweep ce"
SWEEP_CONFIGS=107,207,217,227 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py ce 1280 720 2 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== sweep cornell"
SWEEP_CONFIGS=128,168,188,207 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
