python -m pytest tests/test_gpu_parity.py tests/test_gpu_quicksurf.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/n2_fused.log 2>&1
grep -A12 "Traceback" gpurun_out/n2_fused.log | head -30
tail -1 gpurun_out/n2_fused.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('fused', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['ms_per_step'], d['config']['triangles'])
"
