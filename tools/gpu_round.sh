#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), launch list and one full ncu capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench (ours)"; timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 3000 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
echo "== bench (reference)"; timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; tail -c 1500 $OUT/bench_ref_${TAG}.json
echo "== sweep"; SWEEP_CONFIGS=24,23,4,34 timeout 900 python tools/sweep_configs.py cornell 640 480 32 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
SWEEP_CONFIGS=3,4,23 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py suzanne 640 480 4 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
SWEEP_CONFIGS=3,4,23 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py ce 320 180 1 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -5 $OUT/launches_${TAG}.csv
echo "== ncu full (keyed megakernel; BENCH_SPP=16 keeps the ~45 replays short)"
BENCH_SPP=16 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f -o $OUT/prof_keyed_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -3 $OUT/ncu_full_${TAG}.log; ls -la $OUT
