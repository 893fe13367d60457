python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/n2.log 2>&1
grep -A12 "Traceback" gpurun_out/n2.log | head -30
tail -1 gpurun_out/n2.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['ms_per_step'])
"
