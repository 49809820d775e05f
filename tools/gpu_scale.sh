#!/bin/bash
# Scaling session on an 8-GPU box: bench at N=1,2,4,8 (weak scaling), our arm only.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -2
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/scale_n${n}_${TAG}.json 2> $OUT/scale_n${n}_${TAG}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --steps 3 --warmup 3 > $OUT/scale_n${n}_${TAG}.json 2> $OUT/scale_n${n}_${TAG}.err
  fi
  python - <<PY
import json
try:
    b = json.load(open("$OUT/scale_n${n}_${TAG}.json"))
    print("N=$n value", round(b["value"], 2), "e2e", round(b["e2e"]["value"], 2), "ms/step", round(b["ms_per_step"], 1), "launches", b["gpu_launches"])
except Exception as e:
    print("N=$n FAILED", e); print(open("$OUT/scale_n${n}_${TAG}.err").read()[-1500:])
PY
done
