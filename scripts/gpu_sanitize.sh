#!/bin/bash
# On a B200 box: compute-sanitizer over the suites that exercise the kernels changed in round 2 (memcheck everywhere, racecheck on the
# shared-memory kernels).  Results in gpurun_out/sanitize_*.log; the summary lines go to profiles/.
mkdir -p gpurun_out
T="tests/test_gpu_indexed.py tests/test_gpu_quicksurf.py tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_golden.py tests/test_gpu_coarse_grids.py tests/test_gpu_general_supports.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $T -m gpu -x -q > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_indexed.py tests/test_gpu_quicksurf.py tests/test_gpu_golden.py tests/test_gpu_variants.py -m gpu -x -q > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/sanitize_racecheck.log | tail -3
