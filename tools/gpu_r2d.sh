#!/bin/bash
# Round-2 session d (2 GPUs): multi-device parity, the 2-rank bench line, ticket-ahead prefetch + unroll shapes.
TAG=r2d
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_gpu.txt
echo "== multi-device parity"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fp_way.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_multi_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=127,128,148,108,147 SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py cornell 640 480 64 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== bench N=1"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_n1_${TAG}.json 2> $OUT/bench_n1_${TAG}.err; head -c 400 $OUT/bench_n1_${TAG}.json; echo; tail -2 $OUT/bench_n1_${TAG}.err
echo "== bench N=2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/bench_n2_${TAG}.json 2> $OUT/bench_n2_${TAG}.err; head -c 400 $OUT/bench_n2_${TAG}.json; echo; tail -3 $OUT/bench_n2_${TAG}.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.load(open(f"gpurun_out/bench_n{n}_r2d.json"))
        print(n, "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "weak", d.get("weak_scaling"))
    except Exception as e:
        print(n, "failed", e)
PY
