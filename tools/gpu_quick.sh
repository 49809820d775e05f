SWEEP_CONFIGS=4,24,54,44 SWEEP_SEQUENTIAL=0 python tools/sweep_configs.py cornell 640 480 32
