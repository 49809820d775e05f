#!/bin/bash
# Round-2 session i: full ncu captures of the sub-path kernel on suzanne (one and two sub-paths per lane).
TAG=r2i
OUT=gpurun_out
mkdir -p $OUT
for CFG in 107 217; do
  PTB200_KEYED_CONFIG=$CFG BENCH_SPP=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPath -c 1 -f -o $OUT/prof_suzanne_${CFG}_${TAG} python bench.py --config 2 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_suzanne_${CFG}_${TAG}.log 2>&1
  tail -1 $OUT/ncu_suzanne_${CFG}_${TAG}.log
done
