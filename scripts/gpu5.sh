mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'mc_emit' -s 0 -c 1 -o gpurun_out/prof_r1d python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu5.log 2>&1
