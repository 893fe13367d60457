"""GPU: the opt-in INDEXED marching-cubes mesh (mms_set_mesh_indexed; CallTriMeshData's SetVertexData + SetTriangleData(uint32) form,
plugins/geometry_calls_gl/include/geometry_calls_gl/CallTriMeshDataGL.h:897-1000) against the default triangle soup, which is itself pinned to the oracle
(tests/test_gpu_parity.py, test_gpu_fullsize.py):

  * expanding the indices reproduces the soup BIT FOR BIT (positions and normals, triangle for triangle)
  * the vertex count is the number of crossed grid edges (counted on the host from the volume), every vertex is referenced, and no two
    vertices share an edge (positions are distinct)
  * grids whose x resolution makes node 32 of the last segment the grid's last node / puts a lone node into an extra segment"""
import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def crossed_edges(vol, iso):
    b = vol < np.float32(iso)
    return int((b[:, :, 1:] != b[:, :, :-1]).sum() + (b[:, 1:, :] != b[:, :-1, :]).sum() + (b[1:, :, :] != b[:-1, :, :]).sum())


def soup_and_indexed(surf, iso):
    surf.set_mesh_indexed(False)
    surf.extract_isosurface(iso)
    pos, nrm = surf.get_mesh()
    surf.set_mesh_indexed(True)
    surf.extract_isosurface(iso)
    vpos, vnrm, idx = surf.get_mesh_indexed()
    with pytest.raises(mm.MmsError):
        surf.get_mesh()  # the soup getter refuses an indexed mesh
    surf.set_mesh_indexed(False)
    return pos, nrm, vpos, vnrm, idx


def check(vol, iso, pos, nrm, vpos, vnrm, idx):
    nt = pos.shape[0]
    assert idx.shape == (nt, 3) and idx.dtype == np.uint32
    nv = vpos.shape[0]
    assert nv == crossed_edges(vol, iso)
    assert int(idx.max()) < nv
    assert np.array_equal(vpos[idx].view(np.uint32), pos.view(np.uint32)), "expanded positions must equal the soup bit for bit"
    assert np.array_equal(vnrm[idx].view(np.uint32), nrm.view(np.uint32)), "expanded normals must equal the soup bit for bit"
    used = np.zeros(nv, bool)
    used[idx.reshape(-1)] = True
    assert used.all(), "every crossed edge belongs to a triangle"
    # (two crossed edges that meet in a grid node whose value IS the iso value share that point: only without such ties are positions distinct)
    if not (vol == np.float32(iso)).any():
        assert np.unique(vpos, axis=0).shape[0] == nv, "one vertex per crossed edge"


@pytest.mark.parametrize("res", [(64, 48, 40), (33, 20, 17), (65, 33, 20), (70, 41, 37), (31, 9, 5)], ids=lambda r: "x".join(map(str, r)))
def test_indexed_mesh_expands_to_the_soup_noise(surf, res):
    """white noise: about half of all grid edges are crossed, every marching-cubes case occurs"""
    vol = synth.uniform(4242, 0, res[0] * res[1] * res[2], 0).reshape(res[2], res[1], res[0]).astype(np.float32)
    surf.clear_particles()
    surf.set_grid((0.5, -1.0, 2.0), tuple(float(r - 1) * 0.37 for r in res), res, (False,) * 3)
    surf.set_params(mode=0, want_cell_tricounts=0)
    surf.set_density(vol)
    out = soup_and_indexed(surf, 0.5)
    assert out[0].shape[0] > vol.size
    check(vol, 0.5, *out)


def test_indexed_mesh_of_a_particle_density(surf):
    lists, bmin, bext = H.uniform_case(40000, 30.0, 0.7)
    res = (96, 80, 72)
    surf.clear_particles()
    surf.set_grid(bmin, bext, res, (True, True, True))
    surf.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0, want_home_voxels=0, want_cell_tricounts=0)
    surf.push_particles(lists)
    surf.compute_density()
    vol = surf.get_density().copy()
    out = soup_and_indexed(surf, 0.3)
    assert out[0].shape[0] > 10000
    check(vol, 0.3, *out)
    # ~6 corners share a vertex on a closed surface: the indexed mesh is much smaller
    assert out[2].shape[0] * 24 + out[4].shape[0] * 12 < 0.5 * out[0].shape[0] * 72


def test_indexed_mesh_needs_the_whole_volume():
    s = mm.Surf(0)
    try:
        s.set_grid((0, 0, 0), (15, 15, 15), (16, 16, 16), (False,) * 3)
        s.set_slab(0, 9, 0, 8)
        s.set_params(mode=0)
        s.set_density(np.zeros((9, 16, 16), np.float32))
        s.set_mesh_indexed(True)
        with pytest.raises(mm.MmsError):
            s.extract_isosurface(0.5)
    finally:
        s.close()


def test_empty_indexed_mesh(surf):
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (15, 15, 15), (16, 16, 16), (False,) * 3)
    surf.set_params(mode=0)
    surf.set_density(np.ones((16, 16, 16), np.float32))
    surf.set_mesh_indexed(True)
    surf.extract_isosurface(0.5)
    vpos, vnrm, idx = surf.get_mesh_indexed()
    surf.set_mesh_indexed(False)
    assert vpos.shape == (0, 3) and idx.shape == (0, 3)
