#!/bin/bash
# On a B200 box: the ncu launch list of a short bench run and one --set full capture of the three big kernels (TAG names the outputs).
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu.log 2>&1
rm -f gpurun_out/prof_$TAG.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count' -s 9 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep gpurun_out/launches_$TAG.csv
