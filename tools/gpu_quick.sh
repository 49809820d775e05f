#!/bin/bash
# Two-GPU session: multi-device parity with the final kernels (the sequential policies share the passes
# out between devices, each device picks its own lanes per pass) and the 2-rank bench line.
OUT=gpurun_out
mkdir -p $OUT
echo "== multi-device parity"; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fp_way.py -m gpu -q 2>&1 | tail -3 | tee $OUT/r2u_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/scale_n2_r2u.json 2> $OUT/scale_n2_r2u.err
head -c 400 $OUT/scale_n2_r2u.json; echo; tail -2 $OUT/scale_n2_r2u.err | cut -c1-300
