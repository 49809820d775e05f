SWEEP_CONFIGS=24,26,23,25 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py cornell 640 480 32
SWEEP_CONFIGS=3,5,23,25 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py suzanne 640 480 4
SWEEP_CONFIGS=3,5 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py ce 320 180 1
python -m pytest tests/test_gpu_parity.py -q -k "adaptor or one_stage or intersect_matches" 2>&1 | tail -2
