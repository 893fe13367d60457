N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu > gpurun_out/scale_n$N.log 2>&1
grep -A12 "Traceback" gpurun_out/scale_n$N.log | head -30
tail -1 gpurun_out/scale_n$N.log > gpurun_out/scale_n$N.json
python -c "
import sys, json
d=json.loads(open('gpurun_out/scale_n$N.json').read()); print('N=$N', d['value'], d['ms_per_step'], d['stages_ms'], d['e2e']['ms_per_step'], d['config']['triangles'])
"
