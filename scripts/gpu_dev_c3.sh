#!/bin/bash
# dev: QuickSurf (Gaussian mode) tests + the C3 bench line with the generic gather kernel and with the warp-patch kernel; optional ncu
TAG=${1:-c3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quicksurf.py tests/test_gpu_quicksurf_ref.py tests/test_gpu_slabs.py tests/test_gpu_plugin.py -m gpu -x -q 2>&1 | tail -6
for V in generic patch; do
  if [ $V = generic ]; then export MMS_GATHER_GENERIC=1; else unset MMS_GATHER_GENERIC; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload c3 2>/dev/null | tail -1 > gpurun_out/bench_c3_${TAG}_$V.json
  python -c "
import sys, json
d=json.loads(open('gpurun_out/bench_c3_${TAG}_$V.json').read()); print('$V', d['ms_per_step'], d['stages_ms'], d['roofline'])
"
done
if [ -n "$2" ]; then
rm -f gpurun_out/prof_$TAG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --workload c3 > gpurun_out/b_ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
fi
