#!/bin/bash
# Scratch A/B: how much of the out-of-line arithmetic to inline.  L = level in the three-kernel pipeline
# (1 IEEE sqrt/div, 2 + sin/cos, 3 + Philox, 4 + cone sampling), K = the same in pt_kernels.cu
# (fp-way megakernel, exact-stream kernel).
OUT=gpurun_out
mkdir -p $OUT
cp pt_three_ways_b200/libptb200.so /tmp/libptb200_keep.so
for v in L1K0 L2K0 L3K0 L4K0; do
  cp pt_three_ways_b200/variants/libptb200_$v.so pt_three_ways_b200/libptb200.so
  echo "== variant $v"
  SWEEP_CONFIGS=128 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py cornell 640 480 64 2>&1 | cut -c1-140
  SWEEP_CONFIGS=217 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py suzanne 640 480 16 2>&1 | cut -c1-140
  SWEEP_CONFIGS=217 SWEEP_SEQUENTIAL=0 timeout 300 python tools/sweep_configs.py ce 1280 720 2 2>&1 | cut -c1-140
done 2>&1 | tee $OUT/r2y_inline_ab.txt
for v in L1K0 L1K1 L1K2; do
  cp pt_three_ways_b200/variants/libptb200_$v.so pt_three_ways_b200/libptb200.so
  echo "== variant $v (pt_kernels.cu)"
  timeout 300 python tools/sequential_rates.py cornell 160 120 4096 0 2>&1 | cut -c1-150
  timeout 300 python - <<'PY'
import sys; sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
ctx = capi.Context(0); ctx.upload_scene(scene)
cam = scene.camera(640, 480); params = capi.make_params(640, 480, spp=64, seed=1)
opts = capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL)
ctx.render(cam, params, opts)
best = min(ctx.render(cam, params, opts)["sweep_kernel_ms"] for _ in range(3))
print("fp way Msamples/s", 640 * 480 * 64 / best / 1e3)
PY
done 2>&1 | tee -a $OUT/r2y_inline_ab.txt
cp /tmp/libptb200_keep.so pt_three_ways_b200/libptb200.so
