#!/bin/bash
# dev: one --set full capture (with source) of mc_emit_kernel on the C2 bench workload. TAG names the output.
TAG=${1:-r2}
mkdir -p gpurun_out
rm -f gpurun_out/prof_$TAG.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${2:-mc_emit}" -s ${3:-3} -c ${4:-1} -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
