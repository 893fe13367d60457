#!/bin/bash
# round-2 re-entry baseline: whole GPU suite, device-resident bench line, one --set full capture of emit + splat + count (with source)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_dev_base.log 2>&1
tail -1 gpurun_out/bench_dev_base.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'])
" || tail -20 gpurun_out/bench_dev_base.log
rm -f gpurun_out/prof_base.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count' -s 9 -c 3 -o gpurun_out/prof_base python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu_base.log 2>&1
ls -la gpurun_out/*.ncu-rep
