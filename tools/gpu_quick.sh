#!/bin/bash
# Scratch session: lane groups of the exact-stream policies (lazy twist + shuffle argmin).
OUT=gpurun_out
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_oo_way.py -m gpu -x -q -k "lane_group or sequential or oo_way or intersect" 2>&1 | tail -4 | tee $OUT/r2r_seq_tests.log
: > $OUT/r2r_sequential_rates.jsonl
for u in 1 2; do
echo "== rates unroll $u"; PTB200_SEQUENTIAL_UNROLL=$u timeout 900 python tools/sequential_rates.py cornell 160 120 4096 16,8,4 2>&1 | sed "s/^{/{\"unroll\": $u, /" | tee -a $OUT/r2r_sequential_rates.jsonl
done
echo "== rates 16384 passes"; timeout 900 python tools/sequential_rates.py cornell 80 60 16384 16,8,4 2>&1 | tee -a $OUT/r2r_sequential_rates.jsonl
echo "== rates 256/1024 passes"; timeout 900 python tools/sequential_rates.py cornell 160 120 256,1024 32,16,8 2>&1 | tee -a $OUT/r2r_sequential_rates.jsonl
