for d in 0 1 2; do
MMS_DEBUG_MC=$d python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('debug', $d, 'mc ms', d['stages_ms']['mc'])
"
done
