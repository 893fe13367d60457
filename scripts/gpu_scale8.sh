#!/bin/bash
# On an N-GPU box (gpurun --gpus N -- 'bash scripts/gpu_scale8.sh N'): the default bench line at N (weak C2 + the C4 strong-scaling figures inside)
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_scale_n$N.log 2>&1
grep -A12 "Traceback" gpurun_out/r2_scale_n$N.log | head -30
tail -1 gpurun_out/r2_scale_n$N.log > gpurun_out/r2_scale_n$N.json
python -c "
import sys, json
d=json.loads(open('gpurun_out/r2_scale_n$N.json').read()); print('N=$N', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e'].get('ms_per_step'), d['config']['triangles'])
c=d.get('c4')
if c: print('  c4', c['ms_per_step'], c['stages_ms'], c['e2e_mesh_on_device']['ms_per_step'], (c['e2e'] or {}).get('ms_per_step'))
"
