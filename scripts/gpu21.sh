timeout 300 python scripts/splat_ab.py 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_dev.log 2>&1
tail -1 gpurun_out/bench_dev.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['config']['triangles'])
" || tail -20 gpurun_out/bench_dev.log
