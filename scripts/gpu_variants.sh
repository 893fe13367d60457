#!/bin/bash
# dev: A/B of library variants built into gpurun_variants/libmmsurf_<name>.so (device-resident C2 bench line each, twice)
mkdir -p gpurun_out
cp megamol_b200/libmmsurf.so /tmp/libmmsurf_orig.so
for f in gpurun_variants/libmmsurf_*.so; do
  n=$(basename $f .so); n=${n#libmmsurf_}
  cp $f megamol_b200/libmmsurf.so
  for rep in 1 2; do
    timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('$n', d['ms_per_step'], d['stages_ms']['bin'], d['stages_ms']['density'], d['stages_ms']['mc'], d['stages_ms']['mc_emit'], d['config']['triangles'])"
  done
done
cp /tmp/libmmsurf_orig.so megamol_b200/libmmsurf.so
