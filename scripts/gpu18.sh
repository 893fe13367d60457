timeout 300 compute-sanitizer --tool memcheck python scripts/mc_small.py 28 24 20 > gpurun_out/san_tma.log 2>&1
grep -v "^$" gpurun_out/san_tma.log | head -40
