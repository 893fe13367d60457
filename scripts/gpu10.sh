python -m pytest tests/test_gpu_quicksurf.py -m gpu -x -q 2>&1 | tail -4
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['config']['triangles'], d['e2e']['ms_per_step'])
"
