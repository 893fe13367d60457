rm -f gpurun_out/prof_r1h.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'density_splat' -s 4 -c 1 -o gpurun_out/prof_r1h python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/b_ncu2.log 2>&1
ls -la gpurun_out | tail -3
