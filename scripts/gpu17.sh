# small MC check in both loader modes, full GPU test suite, then a short bench (device + e2e arms)
python scripts/mc_small.py 28 24 20 2>&1 | tail -3
MMS_NO_TMA=1 python scripts/mc_small.py 70 33 21 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_dev.log 2>&1
tail -1 gpurun_out/bench_dev.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['config']['triangles'])
" || tail -20 gpurun_out/bench_dev.log
