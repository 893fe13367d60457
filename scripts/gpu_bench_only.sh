#!/bin/bash
# dev: device-resident bench line only (optionally an ncu capture of kernel regex $2)
TAG=${1:-b}
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_$TAG.log 2>&1
tail -1 gpurun_out/bench_dev_$TAG.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$TAG', d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_$TAG.log
if [ -n "$2" ]; then bash scripts/gpu_ncu_emit.sh $TAG $2; fi
