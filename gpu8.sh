python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/n2.log 2>&1
grep -v "Warning\|OMP_NUM\|\*\*\*" gpurun_out/n2.log | grep -A12 "Traceback" | head -40
tail -1 gpurun_out/n2.log
