"""V1 vs V2 splat kernels: bit-equality of the volume (debug helper; run on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megamol_b200 as mm
from megamol_b200 import synth

def run(xyz, box, res, cyc, radius, v1):
    if v1: os.environ["MMS_SPLAT_V1"] = "1"
    else: os.environ.pop("MMS_SPLAT_V1", None)
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), box, res, cyc)
    s.set_params(mode=0, aggregator=0, normalize=0, sigma=1.0)
    s.push_particles([dict(vtx=xyz, vtx_type=1, count=len(xyz), global_radius=radius)])
    s.compute_density()
    v = s.get_density().copy()
    t = s.timings()
    s.close()
    return v, t

ok = True
for (n, res, cyc, radius) in [(200_000, (128, 128, 128), (True, True, True), 0.5), (50_000, (70, 45, 33), (False, True, False), 0.6),
                              (300_000, (96, 64, 40), (True, False, True), 0.45), (2_000_000, (256, 256, 256), (True, True, True), 0.5)]:
    L = float(np.float32(res[0] - 1) * np.float32(0.4563))
    box = (L, L * res[1] / res[0], L * res[2] / res[0])
    xyz = synth.uniform_box(n, 1.0) * np.asarray(box, np.float32)
    if n == 300_000: xyz = (xyz * np.float32(1.6) - np.float32(0.3) * np.asarray(box, np.float32)).astype(np.float32) # particles outside the box: binned through the wrap
    a, ta = run(xyz, box, res, cyc, radius, True)
    b, tb = run(xyz, box, res, cyc, radius, False)
    same = np.array_equal(a.view(np.uint32), b.view(np.uint32))
    ok &= same
    print(n, res, cyc, "equal" if same else f"DIFFERENT: {np.count_nonzero(a != b)} voxels, max abs {np.abs(a - b).max():.3e}", "sum", float(a.sum()),
          "density ms v1 %.3f v2 %.3f" % (ta["density"], tb["density"]), flush=True)
sys.exit(0 if ok else 1)
