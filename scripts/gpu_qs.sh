#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/quicksurf_tail.json
timeout 900 python -m pytest tests/test_gpu_quicksurf_ref.py tests/test_gpu_quicksurf.py -m gpu -q 2>&1 | tail -40
cat gpurun_out/quicksurf_tail.json
