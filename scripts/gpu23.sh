# round-1 final N=1 evidence: bench (ours + reference arm), launch list, ncu full capture of the three big kernels, other workloads
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -1 gpurun_out/bench_n1.json | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-200
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
rm -f gpurun_out/prof_r1i.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count' -s 9 -c 3 -o gpurun_out/prof_r1i python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
for WL in c1 c3; do python bench.py --steps 5 --warmup 3 --no-cpu --workload $WL 2>/dev/null | tail -1 > gpurun_out/bench_$WL.json; cut -c1-400 gpurun_out/bench_$WL.json; done
python bench.py --workload c5 --frames 12 2>/dev/null | tail -1 > gpurun_out/bench_c5.json; cut -c1-300 gpurun_out/bench_c5.json
ls -la gpurun_out | tail -8
