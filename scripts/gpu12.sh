python bench.py --workload c5 --frames 8 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['frames_per_s'], d['config']['latency_ms'], d['stages_ms'], d['config']['file_write_s'])
"
