#!/bin/bash
# On a B200 box (gpurun --timeout 900 -- 'bash scripts/gpu_evidence.sh TAG'): the evidence the round's profile summary is written from.
#   bench line (ours) + reference arm, ncu launch list of the same command, one ncu --set full capture of the three big kernels,
#   the other BASELINE workloads.  Everything lands in gpurun_out/ (scratch); copy what is to be judged into profiles/.
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -1 gpurun_out/bench_n1.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
rm -f gpurun_out/prof_$TAG.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count' -s 9 -c 3 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
for WL in c1 c3; do python bench.py --steps 5 --warmup 3 --no-cpu --workload $WL 2>/dev/null | tail -1 > gpurun_out/bench_$WL.json; cut -c1-400 gpurun_out/bench_$WL.json; done
python bench.py --workload c5 --frames 12 2>/dev/null | tail -1 > gpurun_out/bench_c5.json; cut -c1-300 gpurun_out/bench_c5.json
python bench.py --steps 3 --warmup 3 --no-cpu --algorithm mt 2>/dev/null | tail -1 > gpurun_out/bench_mt.json; cut -c1-300 gpurun_out/bench_mt.json
ls -la gpurun_out | tail -8
