python -m pytest tests/test_gpu_parity.py -x -q -k "ce or suzanne or intersect_matches" 2>&1 | tail -2
SWEEP_CONFIGS=3,23 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py ce 320 180 1
SWEEP_CONFIGS=3 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py suzanne 640 480 4
