#!/bin/bash
# 2-GPU box, short: the real-NCCL 2-rank parity test, then the device-resident weak-scaling line (no end-to-end legs, no C4)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab_group.py -m gpu -q 2>&1 | tail -5
cat gpurun_out/multi_gpu_check.log 2>/dev/null | tail -4
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-e2e --no-c4 > gpurun_out/scale_quick_n2.log 2>&1
grep -A12 "Traceback" gpurun_out/scale_quick_n2.log | head -20
tail -1 gpurun_out/scale_quick_n2.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('N=2', d['value'], d['ms_per_step'], d['stages_ms'])"
