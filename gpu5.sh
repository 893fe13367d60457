mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|mc_count' -s 0 -c 2 -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu4.log 2>&1
