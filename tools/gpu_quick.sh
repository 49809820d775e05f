#!/bin/bash
# Scratch session: lane groups of the exact-stream policies.
OUT=gpurun_out
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_oo_way.py -m gpu -x -q -k "lane_group or sequential or oo_way" 2>&1 | tail -4 | tee $OUT/r2p_seq_tests.log
echo "== rates"; timeout 900 python tools/sequential_rates.py cornell 160 120 256,4096 0,32,16,8,4 2>&1 | tee $OUT/r2p_sequential_rates.jsonl
