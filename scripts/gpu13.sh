mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -1 gpurun_out/bench_n1.json | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -1 gpurun_out/bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|density_splat|mc_count' -s 8 -c 3 -o gpurun_out/prof_r1f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -5
