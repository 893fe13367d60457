N=$1; WL=${2:-c2}
nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -i "numa\|socket\|^CPU(s)" | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --workload $WL > gpurun_out/scale_${WL}_n$N.log 2>&1
grep -A12 "Traceback" gpurun_out/scale_${WL}_n$N.log | head -30
tail -1 gpurun_out/scale_${WL}_n$N.log > gpurun_out/scale_${WL}_n$N.json
python -c "
import sys, json
d=json.loads(open('gpurun_out/scale_${WL}_n$N.json').read()); print('$WL N=$N', d['value'], d['ms_per_step'], d['e2e'].get('ms_per_step'), d['e2e_mesh_on_device'], d['config']['host_affinity_rank0'])
"
