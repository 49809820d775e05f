python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q 2>&1 | tail -2
python /dev/stdin <<'PY'
import sys, json
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
ctx = capi.Context(0); ctx.upload_scene(scene)
for w, h, spp in ((80, 60, 32), (80, 60, 296), (160, 120, 256)):
    st = ctx.render(scene.camera(w, h), capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=capi.RNG_MT19937_SEQUENTIAL))
    print(json.dumps(dict(w=w, h=h, spp=spp, ms=st["sweep_kernel_ms"], msamples_s=st["samples"]/st["sweep_kernel_ms"]/1e3, us_per_cast_per_pass=st["sweep_kernel_ms"]*1e3/(st["casts"]/spp))))
PY
