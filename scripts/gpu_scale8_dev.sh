#!/bin/bash
# N-GPU box: the device-resident bench line at N (weak C2 + C4 strong figures), no end-to-end legs (short)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_scale_dev_n$N.log 2>&1
grep -A12 "Traceback" gpurun_out/r2_scale_dev_n$N.log | head -30
tail -1 gpurun_out/r2_scale_dev_n$N.log > gpurun_out/r2_scale_dev_n$N.json
python -c "
import sys, json
d=json.loads(open('gpurun_out/r2_scale_dev_n$N.json').read()); print('N=$N', d['value'], d['ms_per_step'], d['stages_ms'], d['config']['triangles'])
c=d.get('c4')
if c: print('  c4', c['ms_per_step'], c['stages_ms'])
"
