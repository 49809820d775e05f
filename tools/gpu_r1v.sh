#!/bin/bash
# Last short session of the round (3.9 GPU-minutes left): what the one-big-CTA launch shapes 66/76 do
# for the sweep-bound scenes, and the parity tests that involve them (every megakernel configuration
# on suzanne, cornell and ce renders bit-identically; renders equal the oracle); smoke.  The rest of
# the tree is what sessions r1t/r1u ran in full.
# Usage: gpurun -- 'bash tools/gpu_r1v.sh r1v'
TAG=${1:-r1v}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1 | tee $OUT/${TAG}_gpu.txt
echo "== sweep suzanne / ce"
SWEEP_CONFIGS=6,66,76 SWEEP_SEQUENTIAL=0 timeout 60 python tools/sweep_configs.py suzanne 640 480 4 2>&1 | tee $OUT/sweep_suzanne_${TAG}.jsonl
SWEEP_CONFIGS=6,66,76 SWEEP_SEQUENTIAL=0 timeout 60 python tools/sweep_configs.py ce 320 180 1 2>&1 | tee $OUT/sweep_ce_${TAG}.jsonl
echo "== pytest -m gpu"
timeout 150 python -m pytest tests -m gpu -q -p no:cacheprovider -k "every_megakernel_configuration or render_matches_oracle or every_instantiation" 2>&1 | tail -15 | tee $OUT/${TAG}_tests.log
echo "== smoke"; timeout 60 python __graft_entry__.py smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.log
