#!/bin/bash
# memcheck over the suites that bin into slab-local cell layers (czBase/czCount)
mkdir -p gpurun_out
T="tests/test_gpu_slabs.py tests/test_gpu_vector.py tests/test_gpu_quicksurf.py tests/test_gpu_coarse_grids.py tests/test_gpu_general_supports.py"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q > gpurun_out/sanitize_memcheck_slabs.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_slabs.log | tail -3
