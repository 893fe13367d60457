mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'density_splat|mc_emit' -s 6 -c 2 -o gpurun_out/prof_r1b python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out | tail -3
