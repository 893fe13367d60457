timeout 900 python -m pytest tests/test_gpu_mt.py tests/test_gpu_plugin.py -m gpu -x -q 2>&1 | tail -25
