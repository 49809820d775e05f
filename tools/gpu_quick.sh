#!/bin/bash
# Scratch A/B session: config identity + one sweep line per configuration.
OUT=gpurun_out
mkdir -p $OUT
echo "== config identity"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "configuration" 2>&1 | tail -2
SWEEP_CONFIGS=${SUZANNE_CONFIGS:-207,217,227} SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py suzanne 640 480 16 2>&1 | cut -c1-170
SWEEP_CONFIGS=${CE_CONFIGS:-207,217} SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py ce 1280 720 2 2>&1 | cut -c1-170
SWEEP_CONFIGS=${CORNELL_CONFIGS:-128,168} SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-170
