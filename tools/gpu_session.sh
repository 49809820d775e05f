#!/bin/bash
# One single-GPU evidence session: parity tests, the bench line (both arms), the launch list with DRAM
# bytes of one full-size step, full ncu captures of the sub-path kernel, the other BASELINE workloads,
# the exact-stream policies' rates.
# Usage (from the repo root, under gpurun):  bash tools/gpu_session.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_tests.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
echo "== bench (ours)"; timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; head -c 300 $OUT/bench_${TAG}.json; echo; tail -2 $OUT/bench_${TAG}.err
echo "== bench (reference)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; head -c 300 $OUT/bench_ref_${TAG}.json; echo
echo "== ncu: launch list + DRAM bytes of one full-size step"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 135 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -1 $OUT/launches_${TAG}.csv | cut -c1-200
echo "== ncu full: sub-path kernel, Cornell (BENCH_SPP=16 keeps the replays short)"
BENCH_SPP=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPath -c 1 -f -o $OUT/prof_subpath_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -1 $OUT/ncu_full_${TAG}.log
echo "== bench --config 2 (suzanne 640x480 @256)"; timeout 900 python bench.py --config 2 --steps 2 --warmup 1 > $OUT/bench_config2_${TAG}.json 2> $OUT/bench_config2_${TAG}.err; head -c 300 $OUT/bench_config2_${TAG}.json; echo
echo "== bench --config 3 (ce 1280x720, BENCH_SPP=64)"; BENCH_SPP=64 timeout 900 python bench.py --config 3 --steps 1 --warmup 0 > $OUT/bench_config3_${TAG}.json 2> $OUT/bench_config3_${TAG}.err; head -c 300 $OUT/bench_config3_${TAG}.json; echo
echo "== ncu full: sub-path kernel, suzanne and ce"
BENCH_SPP=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:subPath -c 1 -f -o $OUT/prof_suzanne_${TAG} python bench.py --config 2 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_suzanne_${TAG}.log 2>&1
# (ce capture: see r2t; the dual kernel is unchanged since)
echo "== exact-stream policies: rates against passes and lanes per pass"
timeout 600 python tools/sequential_rates.py cornell 160 120 256,4096 0,32,16,8 2>&1 | tee $OUT/${TAG}_sequential_rates.jsonl
SEQUENTIAL_MODES=oo timeout 600 python tools/sequential_rates.py cornell 160 120 4096,8192 0 2>&1 | tee -a $OUT/${TAG}_sequential_rates.jsonl
timeout 600 python tools/sequential_rates.py cornell 160 120 8192 0 2>&1 | tee -a $OUT/${TAG}_sequential_rates.jsonl
ls $OUT | grep ${TAG} | tr '\n' ' '
