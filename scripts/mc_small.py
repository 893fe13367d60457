"""Small marching-cubes run against the oracle (debug helper; run on a GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megamol_b200 as mm
from oracle import oracle_binding as ob

res = tuple(int(v) for v in (sys.argv[1:4] or (28, 24, 20)))
iso = 6.5
z, y, x = np.mgrid[0:res[2], 0:res[1], 0:res[0]].astype(np.float32)
vol = np.sqrt((x - 13.3) ** 2 + (y - 11.1) ** 2 + (z - 9.7) ** 2).astype(np.float32)
s = mm.Surf(0)
s.set_grid((0, 0, 0), tuple(float(r - 1) for r in res), res, (False,) * 3)
s.set_params(want_cell_tricounts=1)
s.set_density(vol)
s.extract_isosurface(iso)
counts = s.cell_tricounts()
pos, nrm = s.get_mesh()
o = ob.Oracle()
total, rc, _ = o.mc_count(vol, iso)
rpos, rnrm, _ = o.mc_emit(vol, (0, 0, 0), (1, 1, 1), iso)
print("res", res, "tris", pos.shape[0], "ref", total, "counts equal", np.array_equal(counts, rc))
if pos.shape == rpos.shape:
    print("max pos err", np.abs(pos - rpos).max(), "max nrm err", np.abs(nrm - rnrm).max())
    bad = np.argwhere(np.abs(pos - rpos).max(axis=(1, 2)) > 1e-4)[:5, 0]
    for b in bad:
        print("tri", b, pos[b].tolist(), rpos[b].tolist())
else:
    d = np.argwhere(counts != rc)
    print("differing cells", len(d), d[:10].tolist(), counts[counts != rc][:10], rc[counts != rc][:10])
