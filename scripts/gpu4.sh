python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -3 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['per_stage_frac'], d['e2e']['ms_per_step'])
"
