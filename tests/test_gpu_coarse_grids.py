"""GPU: coarse volume grids -- the modules' DEFAULT parameters (sizex = sizey = sizez = 16, cyclic, normalize;
ParticlesToDensity.cpp:126-140) on realistic particle counts, and sort cells holding far more than 65535 records.

With voxels much larger than the kernel almost no particle has a grid node inside its support: the binning drops those
(bin.cuh supportHasNode) and what is left is small.  With voxels AND kernels large every cell is crowded: the canonical
in-cell order then comes from the merge sort (bin.cuh cell_sort_big_kernel) instead of the all-pairs ranking."""
import numpy as np
import pytest

from megamol_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _density(surf, lists, box, res, cyclic, normalize, aggregator=0, sigma=1.0):
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (box,) * 3, res, cyclic)
    surf.set_params(mode=0, aggregator=aggregator, normalize=int(normalize), defer_normalize=0, sigma=sigma)
    surf.push_particles(lists)
    surf.compute_density()
    return surf.get_density().copy()


def test_module_defaults_on_a_large_frame(surf, oracle):
    """16^3, cyclic, normalize, r = 0.5 on 4 M LJ-fluid-like particles: 62 k particles per sort cell before the cull"""
    n = 4_000_000
    xyz, box = synth.lj_fluid(n)
    lists = [H.xyz_list(xyz, 0.5)]
    res, cyc = (16, 16, 16), (True, True, True)
    gpu = _density(surf, lists, box, res, cyc, normalize=True)
    ref, _ = oracle.density_p2d(lists, (0, 0, 0), (box,) * 3, res, cyc, sigma=1.0, normalize=True)
    assert ref.max() == 1.0 and np.count_nonzero(ref) > 100
    # (the far tail of the bump -- dis within 0.5 % of eps -- underflows on the SFU path and stays a denormal on the CPU: the non-zero
    #  patterns may differ there, the values agree within the floor of helpers.density_close)
    assert H.density_close(gpu, ref) < H.DENSITY_RTOL
    assert np.count_nonzero(gpu) >= 0.99 * np.count_nonzero(ref)
    surf.extract_isosurface(0.5)   # and the isosurface stage accepts the result
    pos, _ = surf.get_mesh()
    assert pos.shape[0] > 0


@pytest.mark.parametrize("radius,aggregator", [(1.5, 0), (3.2, 0), (1.5, 1)], ids=["box3", "box7", "intensity"])
def test_crowded_cells(surf, oracle, radius, aggregator):
    """16^3 grid with kernels as wide as the voxels: every sort cell holds thousands of records (up to 75 k)"""
    n, box, res = 600_000, 16.0, (16, 16, 16)
    xyz = synth.uniform_box(n, box)
    if aggregator == 1:
        inten = (synth.uniform(77, 0, n, 0) + np.float32(0.25)).astype(np.float32)
        inter = np.concatenate([xyz, inten[:, None]], axis=1).astype(np.float32)
        lists = [dict(vtx=inter, vtx_type=H.VERT_FLOAT_XYZ, vtx_stride=16, count=n, global_radius=radius,
                      col=inter.ctypes.data + 12, col_type=H.COL_FLOAT_I, col_stride=16)]
    else:
        lists = [H.xyz_list(xyz, radius)]
    cyc = (True, False, True)
    a = _density(surf, lists, box, res, cyc, normalize=False, aggregator=aggregator)
    b = _density(surf, lists, box, res, cyc, normalize=False, aggregator=aggregator)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "the merge-sorted in-cell order must be reproducible"
    ref, _ = oracle.density_p2d(lists, (0, 0, 0), (box,) * 3, res, cyc, sigma=1.0, aggregator=aggregator)
    # every voxel sums thousands of positive terms: the fp32 summation order (the reference's own depends on its thread count) shows at
    # sqrt(terms) * 2^-24; 5e-5 relative still pins every term
    assert H.density_close(a, ref) < 5e-5
