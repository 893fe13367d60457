timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_dev.log 2>&1
tail -1 gpurun_out/bench_dev.log | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['e2e_mesh_on_device']['ms_per_step'])
" || tail -20 gpurun_out/bench_dev.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --algorithm mt > gpurun_out/bench_mt.log 2>&1
tail -1 gpurun_out/bench_mt.log > gpurun_out/bench_mt.json; python -c "
import sys, json
d=json.loads(open('gpurun_out/bench_mt.json').read()); print('mt', d['value'], d['ms_per_step'], d['stages_ms'], d['config']['triangles'], d['e2e']['ms_per_step'])
" || tail -20 gpurun_out/bench_mt.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
