"""Shared test helpers: workloads as particle-list dicts + comparison utilities."""
import numpy as np

from megamol_b200 import synth

VERT_FLOAT_XYZ, VERT_FLOAT_XYZR, VERT_SHORT_XYZ, VERT_DOUBLE_XYZ = 1, 2, 3, 4
COL_NONE, COL_UINT8_RGB, COL_UINT8_RGBA, COL_FLOAT_RGB, COL_FLOAT_RGBA, COL_FLOAT_I, COL_USHORT_RGBA, COL_DOUBLE_I = range(8)

# Density tolerance (BASELINE.json: "density within 1e-5 relative (fp32)").  The relative error is taken against
# max(|ref|, DENSITY_FLOOR): below the floor (2.7e-5 of the kernel's peak value e^-1) a voxel holds nothing but the
# far tail exp(-1/(1-q^2)), q -> 1, where ONE ulp of q^2 already moves the reference's own value by more than 1e-5
# relative (d/dq^2 of the exponent is 1/(1-q^2)^2); the absolute error there is < 1e-10.
DENSITY_RTOL = 1e-5
DENSITY_FLOOR = 1e-5
# "vertex positions within 1e-4 of the cell size"
VERTEX_TOL_CELLS = 1e-4


def xyz_list(xyz, radius):
    xyz = np.ascontiguousarray(xyz, np.float32)
    return dict(vtx=xyz, vtx_type=VERT_FLOAT_XYZ, count=len(xyz), global_radius=float(radius))


def xyzr_list(xyzr):
    xyzr = np.ascontiguousarray(xyzr, np.float32)
    return dict(vtx=xyzr, vtx_type=VERT_FLOAT_XYZR, count=len(xyzr))


def uniform_case(n, box, radius, seed=synth.SEED + 1, outside=0.0):
    xyz = synth.uniform_box(n, box * (1 + 2 * outside), seed=seed) - np.float32(box * outside)
    return [xyz_list(xyz.astype(np.float32), radius)], (0.0, 0.0, 0.0), (box, box, box)


def density_close(gpu, ref):
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64)) / np.maximum(np.abs(ref.astype(np.float64)), DENSITY_FLOOR)
    return float(err.max()) if err.size else 0.0


def tie_mask_cells(vol, iso, tol_rel=DENSITY_RTOL):
    """Cells with a corner whose value is within the density tolerance of iso ("iso-value ties"):
    their classification may legitimately differ.  vol: (sz,sy,sx) -> mask (sz-1,sy-1,sx-1)."""
    near = np.abs(vol.astype(np.float64) - iso) <= tol_rel * max(abs(iso), DENSITY_FLOOR) + 1e-12
    m = np.zeros(tuple(s - 1 for s in vol.shape), bool)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                m |= near[dz:dz + m.shape[0], dy:dy + m.shape[1], dx:dx + m.shape[2]]
    return m


def voxel_index_of(pos, bmin, bext, res):
    """Grid-particle positions (ParticlesToDensity "outParticles", aggregator 2) -> linear voxel indices (x fastest)."""
    sd = np.asarray(bext, np.float64) / (np.asarray(res, np.float64) - 1.0)
    ijk = np.rint((np.asarray(pos, np.float64) - np.asarray(bmin, np.float64)) / sd).astype(np.int64)
    return ijk[:, 0] + res[0] * (ijk[:, 1] + res[1] * ijk[:, 2])


def vector_tail_mask(oracle, c):
    """Aggregator 2 divides sum(w d) by sum(w).  Where sum(w) is a handful of subnormal units, ONE unit of difference in a weight
    (expf implementations differ by an ulp) moves the quotient by percents: those voxels (sz,sy,sx) are only checked for being
    non-zero, not for their value."""
    lists = [{k: v for k, v in l.items() if k not in ("dir", "dir_stride")} for l in c["lists"]]
    w, _ = oracle.density_p2d(lists, c["bmin"], c["bext"], c["res"], c["cyclic"], sigma=c["sigma"], aggregator=0, normalize=False)
    return (w > 0) & (w < 1e-30)
