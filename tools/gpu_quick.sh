python -m pytest tests/test_gpu_parity.py -x -q -k "random_scene or aperture" 2>&1 | tail -12
