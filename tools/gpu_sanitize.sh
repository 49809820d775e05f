#!/bin/bash
# compute-sanitizer passes over small renders of every round-2 kernel: the three-kernel keyed pipeline
# (constant-bank stage 0, tile sweep, two sub-paths per lane, streamed tiles), the fp way's megakernel,
# the exact-stream kernel with 32 / 16 / 8 / 4 lanes per pass (both estimators), the intersect kernel.
# Usage (under gpurun):  bash tools/gpu_sanitize.sh <tag>
TAG=${1:-r2v}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/san_render.py <<'PY'
import os, sys
sys.path.insert(0, ".")
import numpy as np
from pt_three_ways_b200 import capi, scenefile
forced = os.environ.get("PTB200_KEYED_CONFIG")
for name, w, h, spp in (("cornell", 16, 12, 2), ("ce", 8, 6, 1), ("suzanne", 8, 6, 1), ("multi-sphere", 12, 9, 2)):
    scene = scenefile.load(f"tests/golden/scenes/{name}.ptscene")
    cam = scene.camera(w, h)
    px, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=3), capi.make_options(rng_mode=capi.RNG_KEYED_PHILOX))
    print(name, "keyed", st["casts"], float(px["sum"].sum()))
    if forced is None:  # the other policies do not depend on the keyed configuration
        px, st = capi.render(scene, cam, capi.make_params(w, h, spp=spp, seed=3), capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
        print(name, "fp way", st["casts"], float(px["sum"].sum()))
        for mode in (capi.RNG_MT19937_SEQUENTIAL, capi.RNG_MT19937_SEQUENTIAL_OO):
            for lanes in (32, 16, 8, 4):
                px, st = capi.render(scene, cam, capi.make_params(w, h, spp=5, seed=3),
                                     capi.make_options(rng_mode=mode, lanes_per_pass=lanes))
                print(name, "mode", mode, "lanes", lanes, st["casts"], float(px["sum"].sum()))
        rays = np.random.default_rng(0).normal(size=(64, 6)); rays[:, 3:] /= np.linalg.norm(rays[:, 3:], axis=1, keepdims=True)
        for sweep in (1, 6, 7):
            capi.intersect(scene, rays, sweep=sweep)
        capi.intersect(scene, rays, warp_cooperative=True)
PY
for tool in memcheck racecheck synccheck; do
  for cfg in auto 127 207 101; do
    echo "== compute-sanitizer --tool $tool PTB200_KEYED_CONFIG=$cfg"
    if [ "$cfg" = "auto" ]; then unset PTB200_KEYED_CONFIG; else export PTB200_KEYED_CONFIG=$cfg; fi
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_render.py > $OUT/san_one.log 2>&1
    echo "exit code $?"
    grep -c "keyed\|fp way\|lanes" $OUT/san_one.log | sed 's/^/renders completed: /'
    grep -v "^cornell\|^ce \|^suzanne\|^multi-sphere" $OUT/san_one.log | tail -8
  done
done | tee $OUT/${TAG}_sanitizer.log
