#!/bin/bash
TAG=${1:-s3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_general_supports.py tests/test_gpu_coarse_grids.py tests/test_gpu_quicksurf.py tests/test_gpu_golden.py tests/test_gpu_variants.py tests/test_gpu_vector.py tests/test_gpu_parity.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py tests/test_gpu_stream.py -m gpu -q 2>&1 | tail -60
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_$TAG.log 2>&1
tail -1 gpurun_out/bench_dev_$TAG.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$TAG', d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_$TAG.log
bash scripts/gpu_ncu_emit.sh $TAG ${2:-density_splat}
