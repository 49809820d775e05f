#!/bin/bash
# One full ncu capture of the fp way's megakernel (CornellBox 640x480, 16 passes).
TAG=${1:-fp}
mkdir -p gpurun_out
cat > /tmp/fp_render.py <<'PY'
import sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
px, st = capi.render(scene, scene.camera(640, 480), capi.make_params(640, 480, spp=16, seed=1),
                     capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
print(st)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f \
  -o gpurun_out/prof_keyed_${TAG} python /tmp/fp_render.py > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -3 gpurun_out/ncu_full_${TAG}.log; ls -la gpurun_out | tail -5
