timeout 420 python -m pytest tests/test_gpu_vector.py tests/test_gpu_plugin.py tests/test_gpu_quicksurf.py tests/test_gpu_golden.py -m gpu -q 2>&1 | tail -40
