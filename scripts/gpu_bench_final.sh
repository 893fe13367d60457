#!/bin/bash
# the default bench line + reference arm + launch list (refreshes profiles/r2_bench_n1.json, r2_bench_reference.json, r2_launches.csv)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -1 gpurun_out/r2_bench_n1.json | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu.log 2>&1
MMS_NO_SPECULATION=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('no speculation', d['ms_per_step'], d['stages_ms'])"
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('speculation', d['ms_per_step'], d['stages_ms'])"
