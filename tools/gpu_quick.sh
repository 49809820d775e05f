#!/bin/bash
# Scratch session for a quick A/B on the GPU box: parity tests + one sweep line per configuration.
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q 2>&1 | tail -2
SWEEP_CONFIGS=${SWEEP_CONFIGS:-24,4,3} SWEEP_SEQUENTIAL=1 python tools/sweep_configs.py cornell 640 480 32
