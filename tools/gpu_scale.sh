#!/bin/bash
# 8-GPU session: strong scaling of the bench workload at N = 1, 2, 4, 8 (the driver's launch lines),
# configs[4] at N = 8 (inside the N = 8 bench line), multi-device parity.
# Usage (under gpurun --gpus 8):  bash tools/gpu_scale.sh <tag>
TAG=${1:-r2e}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_gpus.txt
echo "== multi-device parity"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_multi_tests.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/scale_n1_${TAG}.json 2> $OUT/scale_n1_${TAG}.err
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 5 --warmup 3 > $OUT/scale_n${N}_${TAG}.json 2> $OUT/scale_n${N}_${TAG}.err
  tail -2 $OUT/scale_n${N}_${TAG}.err | cut -c1-300
done
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_n{n}_{tag}.json"))
    except Exception as e:
        print(n, "failed", e); continue
    base = base or d
    print(f"N={n}: value {d['value']:.1f} ({d['value'] / base['value'] / n:.4f}), e2e {d['e2e']['value']:.1f} "
          f"({d['e2e']['value'] / base['e2e']['value'] / n:.4f}), weak {d.get('weak_scaling', {}).get('value')}, clocks {d['clocks']['sm_mhz']}")
    if "config4" in d:
        print("  config4:", json.dumps(d["config4"]))
PY
