#!/bin/bash
# On a B200 box: what the driver runs at round end -- the GPU test suite, smoke(), and the default bench line (with e2e and CPU legs).
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 900 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
