"""dev: the C2 frame with the indexed mesh, a few times (for ncu launch lists / captures of mcx_* kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from megamol_b200.slabs import SlabJob
w = bench.workload("c2")
job = SlabJob(w, 0, 1, 0, iso=bench.ISO, radius=bench.RADIUS)
job.surf.set_mesh_indexed(True)
for _ in range(3):
    job._compute(job.d_xyz.data_ptr(), job.n_local)
    job.surf.synchronize()
print(job.surf.timings())
