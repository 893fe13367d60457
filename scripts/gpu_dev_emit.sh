#!/bin/bash
# dev: new emit kernel vs the round-1 kernel (MMS_EMIT_V4=1): parity tests, then device-resident bench lines of both
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_slabs.py tests/test_gpu_quicksurf.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15
for v in new v4; do
  if [ $v = v4 ]; then export MMS_EMIT_V4=1; else unset MMS_EMIT_V4; fi
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_$v.log 2>&1
  tail -1 gpurun_out/bench_dev_$v.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_$v.log
done
