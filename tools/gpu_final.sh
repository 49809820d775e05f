#!/bin/bash
# Final verification of a build: what the driver runs at round end, plus the JSON-line check.
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; python __graft_entry__.py smoke 2>&1 | tail -3
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench ours"; timeout 900 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; wc -l $OUT/bench_${TAG}.json; tail -c 600 $OUT/bench_${TAG}.json; echo
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref_${TAG}.json 2> $OUT/bench_ref_${TAG}.err; wc -l $OUT/bench_ref_${TAG}.json; tail -c 300 $OUT/bench_ref_${TAG}.json; echo
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
echo "== ncu full (BENCH_SPP=16)"
BENCH_SPP=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f -o $OUT/prof_keyed_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -1 $OUT/ncu_full_${TAG}.log
