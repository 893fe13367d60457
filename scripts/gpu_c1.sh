#!/bin/bash
# dev: C1 (1 M particles -> 128^3) bench line and its launch list
mkdir -p gpurun_out
python bench.py --workload c1 --steps 50 --warmup 5 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/c1_dev.json
python -c "
import json; d=json.loads(open('gpurun_out/c1_dev.json').read()); print('C1', d['ms_per_step'], d['stages_ms'], d.get('gpu_launches'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/c1_launches.csv python bench.py --workload c1 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/c1_ncu.log 2>&1
tail -3 gpurun_out/c1_ncu.log | cut -c1-300
