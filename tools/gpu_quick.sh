#!/bin/bash
# Short single-GPU check (about two minutes of box time): the parity tests that cover every kernel
# form, then the keyed pipeline's rate on the three BASELINE scenes and the exact-stream rates.
# Usage (from the repo root, under gpurun):  bash tools/gpu_quick.sh
# (Same-session A/Bs of compile-time variants were run with this script's ancestor: build the variants
#  with -DPT_INLINE_LEVEL=n / -DPT_CONSTANTS_IN_BANK=n into separate .so files, copy each over
#  pt_three_ways_b200/libptb200.so in turn and run the sweeps — profiles/r2x_constants_ab.txt,
#  profiles/r2y_inline_ab.txt.)
OUT=gpurun_out
mkdir -p $OUT
echo "== parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "render_matches_oracle or lane_group or configuration" 2>&1 | tail -2
echo "== keyed pipeline"
SWEEP_CONFIGS=128 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-170
SWEEP_CONFIGS=217 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py suzanne 640 480 16 2>&1 | cut -c1-170
SWEEP_CONFIGS=217 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py ce 1280 720 2 2>&1 | cut -c1-170
echo "== exact stream"; timeout 300 python tools/sequential_rates.py cornell 160 120 4096 0 2>&1 | cut -c1-170
