#!/bin/bash
# On an N-GPU box (gpurun --gpus N -- 'bash scripts/gpu_scale2.sh N'): the 2-rank parity test (both exchange paths, all gather modes),
# then the default bench line at N (weak C2 + the C4 figures inside), fused and nccl exchange.
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab_group.py tests/test_gpu_plugin.py -m gpu -q 2>&1 | tail -30
cat gpurun_out/multi_gpu_check.log
for EX in fused nccl; do
  EXTRA=""; if [ $EX = nccl ]; then EXTRA="--no-c4 --no-e2e"; fi
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --exchange $EX $EXTRA > gpurun_out/scale_${EX}_n$N.log 2>&1
  grep -A12 "Traceback" gpurun_out/scale_${EX}_n$N.log | head -30
  tail -1 gpurun_out/scale_${EX}_n$N.log > gpurun_out/scale_${EX}_n$N.json
  python -c "
import sys, json
d=json.loads(open('gpurun_out/scale_${EX}_n$N.json').read()); print('$EX N=$N', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e'].get('ms_per_step'), d['config']['triangles'])
c=d.get('c4')
if c: print('  c4', c['ms_per_step'], c['stages_ms'], c['e2e_mesh_on_device']['ms_per_step'], (c['e2e'] or {}).get('ms_per_step'))
"
done
