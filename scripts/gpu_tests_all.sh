#!/bin/bash
# dev: the whole GPU suite (first failure stops), then a device-resident bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAILN:-40}
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_all.log 2>&1
tail -1 gpurun_out/bench_dev_all.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_all.log
