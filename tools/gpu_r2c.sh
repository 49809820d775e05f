#!/bin/bash
# Round-2 session c: new parity-at-size tests, the reworked bench line, occupancy shapes, DRAM traffic.
TAG=r2c
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
echo "== bench (ours)"; timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 2500 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
echo "== sweep cornell (launch shapes)"
SWEEP_CONFIGS=127,167,177,187,137 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== ncu: launch list + DRAM bytes of one full-size step"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 135 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -3 $OUT/launches_${TAG}.csv
echo "== bench --config 2 (suzanne 640x480 @256)"; timeout 900 python bench.py --config 2 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_config2_${TAG}.json 2> $OUT/bench_config2_${TAG}.err; head -c 600 $OUT/bench_config2_${TAG}.json; echo
echo "== bench --config 3 (ce 1280x720, BENCH_SPP=32)"; BENCH_SPP=32 timeout 900 python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/bench_config3_${TAG}.json 2> $OUT/bench_config3_${TAG}.err; head -c 600 $OUT/bench_config3_${TAG}.json; echo
ls -la $OUT | tail -8
