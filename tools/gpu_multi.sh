#!/bin/bash
# Multi-GPU session (under gpurun --gpus N): multi-device tests + bench at N ranks, both arms.
N=${1:-2}
TAG=${2:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
for n in 1 $N; do
  echo "== bench N=$n"
  PORT=$((29500 + n))
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_n${n}_${TAG}.json 2> $OUT/bench_n${n}_${TAG}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $n --steps 3 --warmup 3 > $OUT/bench_n${n}_${TAG}.json 2> $OUT/bench_n${n}_${TAG}.err
  fi
  tail -c 1200 $OUT/bench_n${n}_${TAG}.json; tail -5 $OUT/bench_n${n}_${TAG}.err
done
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29600 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>&1 | tail -c 600
