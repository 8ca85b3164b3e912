python -m pytest tests/test_dist.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
