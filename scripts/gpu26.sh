timeout 600 python -m pytest tests/test_gpu_quicksurf.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu --workload c3 2>/dev/null | tail -1 > gpurun_out/bench_c3.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3.json').read()); print('c3', d['ms_per_step'], d['stages_ms'])"
rm -f gpurun_out/prof_c3.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:'density_gather' -s 2 -c 1 -o gpurun_out/prof_c3 python bench.py --steps 1 --warmup 3 --no-cpu --workload c3 > gpurun_out/b_ncu3.log 2>&1
ls -la gpurun_out/prof_c3.ncu-rep
