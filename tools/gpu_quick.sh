#!/bin/bash
# Scratch session: full capture of the exact-stream kernel with 16 lanes per pass at 4096 passes.
OUT=gpurun_out
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:renderSequential -c 1 -f -o $OUT/prof_seq_g16_r2s python tools/sequential_rates.py cornell 16 12 4096 16 > $OUT/ncu_seq_g16.log 2>&1
tail -2 $OUT/ncu_seq_g16.log
