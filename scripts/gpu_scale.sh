#!/bin/bash
# On an N-GPU box (gpurun --gpus N -- 'bash scripts/gpu_scale.sh N'): the 2-rank NCCL tests, then the C2 (weak) and C4 (strong) lines.
N=$1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | tail -4
for WL in c2 c4; do
  EXTRA=""; if [ $WL = c4 ] && [ $N -lt 8 ]; then EXTRA="--no-e2e"; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --workload $WL $EXTRA > gpurun_out/scale_${WL}_n$N.log 2>&1
  grep -A12 "Traceback" gpurun_out/scale_${WL}_n$N.log | head -30
  tail -1 gpurun_out/scale_${WL}_n$N.log > gpurun_out/scale_${WL}_n$N.json
  python -c "
import sys, json
d=json.loads(open('gpurun_out/scale_${WL}_n$N.json').read()); print('$WL N=$N', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e'].get('ms_per_step'), d['config']['triangles'], d['scaling'])
"
done
