#!/bin/bash
# Scratch A/B session.
OUT=gpurun_out
mkdir -p $OUT
true
for scene in "suzanne 640 480 32" "ce 1280 720 4" "cornell 640 480 64"; do
  set -- $scene
  cfgs="217,127"; [ "$1" = cornell ] && cfgs="128,127"
  echo "== $scene (fan groups)"; SWEEP_CONFIGS="$cfgs" SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py $scene 2>&1 | cut -c1-170
  echo "== $scene (PTB200_NO_FAN_GROUPS=1)"; PTB200_NO_FAN_GROUPS=1 SWEEP_CONFIGS="$cfgs" SWEEP_SEQUENTIAL=0 timeout 600 python tools/sweep_configs.py $scene 2>&1 | cut -c1-170
done
