#!/bin/bash
# compute-sanitizer passes over small renders of both kernels and every sweep variant.
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/san_render.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from pt_three_ways_b200 import capi, scenefile
for name, w, h in (("cornell", 24, 18), ("ce", 8, 6), ("bbc-owl", 16, 12)):
    scene = scenefile.load(f"tests/golden/scenes/{name}.ptscene")
    for mode in (capi.RNG_KEYED_PHILOX, capi.RNG_MT19937_SEQUENTIAL):
        px, st = capi.render(scene, scene.camera(w, h), capi.make_params(w, h, spp=2, seed=3), capi.make_options(rng_mode=mode))
        print(name, mode, st["casts"], float(px["sum"].sum()))
    rays = np.random.default_rng(0).normal(size=(64, 6)); rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
    for sweep in (0, 1, 2, 3, 4):
        capi.intersect(scene, rays, sweep=sweep)
    capi.intersect(scene, rays, warp_cooperative=True)
PY
for tool in memcheck racecheck synccheck; do
  for cfg in auto 1 3 0; do
    echo "== compute-sanitizer --tool $tool PTB200_KEYED_CONFIG=$cfg"
    if [ "$cfg" = "auto" ]; then unset PTB200_KEYED_CONFIG; else export PTB200_KEYED_CONFIG=$cfg; fi
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_render.py 2>&1 | tail -4
  done
done | tee $OUT/sanitizer.log
