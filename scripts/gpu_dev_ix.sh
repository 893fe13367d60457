#!/bin/bash
# dev: indexed-mesh tests, then the whole GPU suite, then the full default bench line (with e2e and e2e_indexed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_indexed.py -m gpu -x -q 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_ix.log 2>&1
tail -1 gpurun_out/bench_ix.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['stages_ms'], d['roofline']['frac']); print('e2e', d['e2e']); print('ix', d['e2e_indexed'])
" || tail -20 gpurun_out/bench_ix.log
