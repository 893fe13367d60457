#!/bin/bash
# On a B200 box (gpurun -- 'bash scripts/gpu_check.sh'): the whole GPU test suite, then a short device-resident bench line.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev.log 2>&1
tail -1 gpurun_out/bench_dev.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev.log
