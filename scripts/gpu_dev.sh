#!/bin/bash
# dev loop on a B200 box: whole GPU suite, V1/V2 splat bit-equality, device-resident bench line; optional ncu capture (TAG, kernel regex)
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python scripts/splat_ab.py 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/bench_dev_$TAG.log 2>&1
tail -1 gpurun_out/bench_dev_$TAG.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$TAG', d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_$TAG.log
if [ -n "$2" ]; then
rm -f gpurun_out/prof_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${3:-3} -c ${4:-1} -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/b_ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
fi
