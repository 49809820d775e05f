#!/bin/bash
# Profiles the sequential (exact mt19937 stream) kernel and measures it at scale.
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
cat > /tmp/seq_render.py <<'PY'
import sys, json
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
w, h, spp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = capi.Context(0); ctx.upload_scene(scene)
st = ctx.render(scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL))
print(json.dumps(dict(w=w, h=h, spp=spp, ms=st["sweep_kernel_ms"], msamples_s=st["samples"]/st["sweep_kernel_ms"]/1e3, us_per_cast_per_pass=st["sweep_kernel_ms"]*1e3/(st["casts"]/spp))))
PY
echo "== sequential kernel, scaling in passes"
python /tmp/seq_render.py 80 60 32 | tee $OUT/seq_${TAG}.jsonl
python /tmp/seq_render.py 80 60 296 | tee -a $OUT/seq_${TAG}.jsonl
python /tmp/seq_render.py 80 60 1184 | tee -a $OUT/seq_${TAG}.jsonl
python /tmp/seq_render.py 80 60 4736 | tee -a $OUT/seq_${TAG}.jsonl
echo "== ncu full (sequential kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:renderSequential -c 1 -f -o $OUT/prof_seq_${TAG} python /tmp/seq_render.py 80 60 32 > $OUT/ncu_seq_${TAG}.log 2>&1
tail -2 $OUT/ncu_seq_${TAG}.log
