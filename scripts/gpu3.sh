mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu.log 2>&1
