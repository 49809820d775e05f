#!/bin/bash
# One SHORT GPU-box session (the round's last GPU minutes): every GPU test, the configuration sweep
# with the sign-bit stage-0 variants (5/6) next to the current defaults, one bench line, and the
# ncu launch list + full capture of the fastest Cornell configuration.  Results land in gpurun_out/
# step by step, most important first, so a clamped call still leaves evidence.
# Usage: gpurun -- 'bash tools/gpu_r1s.sh r1s'
TAG=${1:-r1s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1 | tee $OUT/${TAG}_gpu.txt
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -25 | tee $OUT/${TAG}_tests.log
echo "== sweep cornell"
SWEEP_CONFIGS=24,23,25,26,45,46,55,56,5,35,44 SWEEP_SEQUENTIAL=0 timeout 240 python tools/sweep_configs.py cornell 640 480 32 2>&1 | tee $OUT/sweep_cornell_${TAG}.jsonl
echo "== sweep suzanne / ce"
SWEEP_CONFIGS=3,6,26,46 SWEEP_SEQUENTIAL=0 timeout 120 python tools/sweep_configs.py suzanne 640 480 4 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
SWEEP_CONFIGS=3,6 SWEEP_SEQUENTIAL=0 timeout 120 python tools/sweep_configs.py ce 320 180 1 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== fp way rate (default vs sign-bit stage 0)"
timeout 120 python - <<'PY' 2>&1 | tee $OUT/${TAG}_fp_rates.jsonl
import json, os, subprocess, sys
code = """
import json, os, sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
ctx = capi.Context(0)
ctx.upload_scene(scene)
w, h, spp = 640, 480, 64
cam = scene.camera(w, h)
best = 0.0
for _ in range(3):
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=capi.RNG_MT19937_PER_PIXEL))
    best = max(best, st["samples"] / st["kernel_ms"] / 1e3)
print(json.dumps({"mode": "fp", "config": os.environ.get("PTB200_KEYED_CONFIG", "auto"), "msamples_s": round(best, 2)}))
"""
for config in ("24", "25", "45", "5"):
    res = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PTB200_KEYED_CONFIG=config),
                         capture_output=True, text=True)
    print(res.stdout.strip() or res.stderr.strip()[-300:], flush=True)
PY
echo "== oo way rate"
timeout 120 python - <<'PY' 2>&1 | tee $OUT/${TAG}_oo_rate.jsonl
import json, sys
sys.path.insert(0, ".")
from pt_three_ways_b200 import capi, scenefile
scene = scenefile.load("tests/golden/scenes/cornell.ptscene")
ctx = capi.Context(0)
ctx.upload_scene(scene)
w, h, spp = 160, 120, 256
cam = scene.camera(w, h)
for label, mode in (("dod_exact", capi.RNG_MT19937_SEQUENTIAL), ("oo", capi.RNG_MT19937_SEQUENTIAL_OO)):
    st = ctx.render(cam, capi.make_params(w, h, spp=spp, seed=1), capi.make_options(rng_mode=mode))
    print(json.dumps({"mode": label, "w": w, "h": h, "spp": spp, "msamples_s": st["samples"] / st["kernel_ms"] / 1e3,
                      "casts": st["casts"]}), flush=True)
PY
BEST=$(TAG=$TAG python - <<'PY'
import json, os
best, cfg = 0.0, "24"
for line in open("gpurun_out/sweep_cornell_%s.jsonl" % os.environ["TAG"]):
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        if d.get("msamples_s", 0.0) > best:
            best, cfg = d["msamples_s"], d["config"]
print(cfg)
PY
)
echo "best cornell config: $BEST" | tee $OUT/${TAG}_best.txt
echo "== bench (default config)"
timeout 300 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -c 1200 $OUT/bench_${TAG}.json; tail -2 $OUT/bench_${TAG}.err
echo "== bench (best config $BEST, no cpu baseline)"
PTB200_KEYED_CONFIG=$BEST timeout 200 python bench.py --no-cpu-baseline > $OUT/bench_best_${TAG}.json 2> $OUT/bench_best_${TAG}.err; tail -c 600 $OUT/bench_best_${TAG}.json
echo "== ncu launch list (best config)"
PTB200_KEYED_CONFIG=$BEST timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu_${TAG}.log 2>&1
tail -3 $OUT/launches_${TAG}.csv
echo "== ncu full (best config; BENCH_SPP=16)"
PTB200_KEYED_CONFIG=$BEST BENCH_SPP=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:renderKeyed -c 1 -f -o $OUT/prof_keyed_${TAG} python bench.py --steps 1 --warmup 0 --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
tail -1 $OUT/ncu_full_${TAG}.log; ls -la $OUT | tail -20
