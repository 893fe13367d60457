#!/bin/bash
# On a B200 box (gpurun --timeout 2400 -- 'bash scripts/gpu_evidence_r2.sh'): the evidence the round-2 profile summary is written from.
# Everything lands in gpurun_out/ (scratch); what is to be judged is copied into profiles/ afterwards.
T=r2
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
tail -1 gpurun_out/${T}_bench_n1.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_ref.err
tail -1 gpurun_out/${T}_bench_reference.json | cut -c1-200
# launch list of the same command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu.log 2>&1
# one --set full capture (with source) of the big kernels of the C2 frame, and of the Gaussian density kernel (C3)
rm -f gpurun_out/${T}_prof_c2.ncu-rep gpurun_out/${T}_prof_c3.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count|cell_order|bin_' -s 14 -c 6 -o gpurun_out/${T}_prof_c2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/b_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'density_gauss' -s 3 -c 1 -o gpurun_out/${T}_prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --workload c3 > gpurun_out/b_ncu3.log 2>&1
ncu --set full --clock-control none -k regex:'mcx_' -s 3 -c 3 -o gpurun_out/${T}_prof_indexed python scripts/ix_profile.py > gpurun_out/b_ncu4.log 2>&1
# the other BASELINE workloads
for WL in c1 c3; do python bench.py --steps 10 --warmup 3 --no-cpu --workload $WL 2>/dev/null | tail -1 > gpurun_out/${T}_bench_$WL.json; cut -c1-300 gpurun_out/${T}_bench_$WL.json; done
python bench.py --steps 5 --warmup 3 --no-cpu --radius 1.0 2>/dev/null | tail -1 > gpurun_out/${T}_bench_c2_radius1.json; cut -c1-300 gpurun_out/${T}_bench_c2_radius1.json
python bench.py --steps 3 --warmup 3 --no-cpu --algorithm mt 2>/dev/null | tail -1 > gpurun_out/${T}_bench_c2_marching_tets.json; cut -c1-300 gpurun_out/${T}_bench_c2_marching_tets.json
python bench.py --workload c5 --frames 100 2>gpurun_out/c5.err | tail -1 > gpurun_out/${T}_bench_c5_100frames.json; cut -c1-400 gpurun_out/${T}_bench_c5_100frames.json
python bench.py --workload c5 --frames 100 --indexed 2>gpurun_out/c5i.err | tail -1 > gpurun_out/${T}_bench_c5_100frames_indexed.json; cut -c1-400 gpurun_out/${T}_bench_c5_100frames_indexed.json
ls -la gpurun_out | tail -20
