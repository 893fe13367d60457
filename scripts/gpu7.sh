python -m pytest tests/test_gpu_quicksurf.py -m gpu -x -q 2>&1 | tail -25
