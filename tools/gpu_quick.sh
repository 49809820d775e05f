#!/bin/bash
# Scratch A/B session.
OUT=gpurun_out
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_at_size.py -m gpu -x -q 2>&1 | tail -2
echo "== cornell 128 (fan groups)"; SWEEP_CONFIGS=128 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-170
echo "== cornell 128 (PTB200_NO_FAN_GROUPS=1)"; PTB200_NO_FAN_GROUPS=1 SWEEP_CONFIGS=128 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-170
