#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_coarse_grids.py tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -4
timeout 100 python scripts/stager_check.py 2>&1 | tail -4
