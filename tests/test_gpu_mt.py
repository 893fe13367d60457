"""GPU: the marching-tetrahedra compatibility mode (MMS_ISO_MARCHING_TETS) reproduces the reference IsoSurface module's output
bit for bit -- against the golden vectors recorded from the unmodified module and against the oracle on larger volumes."""
import os

import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import synth
from tests import golden_util as G

pytestmark = pytest.mark.gpu


def _mt_mesh(vol, bbox, iso, slab=None, counts=False):
    sz, sy, sx = vol.shape
    s = mm.Surf(0)
    try:
        s.set_grid(tuple(float(v) for v in bbox[:3]), tuple(float(np.float32(bbox[a + 3]) - np.float32(bbox[a])) for a in range(3)), (sx, sy, sz), (False,) * 3)
        s.set_params(want_cell_tricounts=int(counts))
        s.set_isosurface_mode(mm.api.ISO_MARCHING_TETS)
        if slab is not None:
            z0, nz, c0, cn = slab
            s.set_slab(z0, nz, c0, cn)
            s.set_density(np.ascontiguousarray(vol[z0:z0 + nz]))
        else:
            s.set_density(vol)
        s.extract_isosurface(iso)
        pos, nrm = s.get_mesh()
        tc = s.cell_tricounts() if counts else None
        return pos, nrm, tc
    finally:
        s.close()


@pytest.mark.parametrize("name", ["sigma_clipped", "aniso_mixedcyc"])
def test_golden_reference_meshes(name):
    z = np.load(os.path.join(G.GOLDEN, f"isosurface_mt_{name}.npz"))
    pos, nrm, _ = _mt_mesh(z["volume"], z["bbox"], float(z["iso"]))
    assert pos.shape[0] * 3 == z["pos"].shape[0]
    assert np.array_equal(pos.reshape(-1, 3), z["pos"]), "vertices differ from the reference IsoSurface module"
    assert np.array_equal(nrm.reshape(-1, 3), z["nrm"]), "normals differ from the reference IsoSurface module"


def test_against_oracle_large_and_sharded(oracle, surf):
    """a P2D volume of 96 x 70 x 45 voxels: whole and as three z-slabs (cell layers split, global frame)"""
    n, res = 40_000, (96, 70, 45)
    box = tuple(float(np.float32(r - 1) * np.float32(0.5)) for r in res)
    xyz = synth.uniform_box(n, 1.0) * np.asarray(box, np.float32)
    surf.clear_particles()
    surf.set_grid((0, 0, 0), box, res, (True, False, True))
    surf.set_params(mode=0, aggregator=0, normalize=0, sigma=1.0)
    surf.push_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.7)])
    surf.compute_density()
    vol = surf.get_density().copy()
    bbox = (0, 0, 0) + box
    iso = 0.35
    rpos, rnrm = oracle.mt_emit(vol, bbox, iso)
    pos, nrm, tc = _mt_mesh(vol, bbox, iso, counts=True)
    assert rpos.shape[0] > 50_000 and pos.shape == rpos.shape
    assert np.array_equal(pos, rpos) and np.array_equal(nrm, rnrm)
    assert int(tc.sum()) == pos.shape[0] and tc.max() <= 12
    parts = []
    for c0, c1 in ((0, 15), (15, 31), (31, 44)):
        p, q, _ = _mt_mesh(vol, bbox, iso, slab=(c0, c1 - c0 + 1, c0, c1 - c0))
        parts.append((p, q))
    assert np.array_equal(np.concatenate([p for p, _ in parts]), rpos)
    assert np.array_equal(np.concatenate([q for _, q in parts]), rnrm)
