rm -f gpurun_out/prof_r1g.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'mc_emit|mc_count' -s 4 -c 2 -o gpurun_out/prof_r1g python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -3
