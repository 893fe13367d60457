"""GPU suite: the single-process slab group of the C ABI (mms_slabs_*: one context per device, fused halo push over peer memory, range
combined by peer reads) against one GPU -- volume and mesh bit for bit.  The multi-device cases need >= 2 GPUs (skipped on a single-GPU
box); a group may also name ONE device several times: z-chunks on a single GPU, the way a volume of 2^32 voxels or more (more than one
context indexes) is computed on one device -- the reference's QuickSurf chunks its volume in z for the same reason
(CUDAQuickSurf.cu:1050-1126, 1406-1447)."""
import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cyclic", [(True, True, True), (False, True, False)], ids=["cyclic", "open-z"])
@pytest.mark.parametrize("ndev", [2, 4])
def test_slab_group_equals_one_gpu(ndev, cyclic):
    import torch
    if torch.cuda.device_count() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    n, res, box, radius, iso = 400_000, (96, 80, 72), (40.0, 33.0, 30.0), 0.6, 0.35
    xyz = synth.uniform_box(n, 1.0, seed=321) * np.asarray(box, np.float32)
    lists = [dict(vtx=xyz, vtx_type=1, count=n, global_radius=radius)]
    one = mm.Surf(0)
    one.set_grid((0, 0, 0), box, res, cyclic)
    one.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    one.push_particles(lists)
    one.compute_density()
    ref = one.get_density().copy()
    rng = one.density_range()
    one.extract_isosurface(iso)
    rpos, rnrm = one.get_mesh()
    rpos, rnrm = rpos.copy(), rnrm.copy()
    one.close()
    g = mm.SurfGroup(list(range(ndev)))
    try:
        g.set_grid((0, 0, 0), box, res, cyclic)
        g.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
        for _ in range(2):   # second frame: counter parity, buffer re-use
            g.clear_particles()
            g.push_particles(lists)
            g.compute_density()
            vol = g.get_density()
            g.extract_isosurface(iso)
            pos, nrm = g.get_mesh()
            assert np.array_equal(vol.view(np.uint32), ref.view(np.uint32)), "volume differs from one GPU"
            assert pos.shape == rpos.shape and np.array_equal(pos, rpos) and np.array_equal(nrm, rnrm), "mesh differs from one GPU"
        assert rpos.shape[0] > 10000
    finally:
        g.close()


@pytest.mark.parametrize("nchunks", [2, 3])
def test_z_chunks_on_one_device_equal_one_context(nchunks):
    n, res, box, radius, iso = 300_000, (96, 80, 75), (40.0, 33.0, 31.0), 0.6, 0.35
    xyz = synth.uniform_box(n, 1.0, seed=77) * np.asarray(box, np.float32)
    lists = [dict(vtx=xyz, vtx_type=1, count=n, global_radius=radius)]
    one = mm.Surf(0)
    one.set_grid((0, 0, 0), box, res, (True, True, True))
    one.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    one.push_particles(lists)
    one.compute_density()
    ref = one.get_density().copy()
    one.extract_isosurface(iso)
    rpos, rnrm = one.get_mesh()
    rpos, rnrm = rpos.copy(), rnrm.copy()
    one.close()
    g = mm.SurfGroup([0] * nchunks)
    try:
        g.set_grid((0, 0, 0), box, res, (True, True, True))
        g.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
        for _ in range(2):
            g.clear_particles()
            g.push_particles(lists)
            g.compute_density()
            vol = g.get_density()
            g.extract_isosurface(iso)
            pos, nrm = g.get_mesh()
            assert np.array_equal(vol.view(np.uint32), ref.view(np.uint32)), "volume differs from one context"
            assert pos.shape == rpos.shape and np.array_equal(pos, rpos) and np.array_equal(nrm, rnrm), "mesh differs from one context"
        assert rpos.shape[0] > 10000
    finally:
        g.close()


@pytest.mark.parametrize("colour", [False, True], ids=["density", "colour"])
@pytest.mark.parametrize("nchunks", [2, 3])
def test_gaussian_mode_on_z_chunks_equals_one_context(nchunks, colour):
    """The QuickSurf-Gaussian mode (per-atom radii, non-periodic grid, wide supports: density_gauss_kernel) through the slab group:
    the halo bound comes from gausslim * radscale * max radius, each chunk bins only the cell layers that reach it, and with the
    colour volume on the halo records carry their RGBA -- density, colour volume and (coloured) mesh bit-identical to one context."""
    from megamol_b200 import quicksurf
    from tests import helpers as H
    n = 6000
    data, _, _ = synth.protein_like(n, seed=9, nballs=6, extent=40.0)
    data = np.ascontiguousarray(data)  # x y z r | R G B A, stride 32
    lists = [dict(vtx=data, vtx_type=H.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=data.ctypes.data + 16, col_type=H.COL_FLOAT_RGBA,
                  col_stride=32)]
    radscale, spacing, iso = 1.0, 0.8, 0.5
    org, ext, res = quicksurf.grid_from_particles(data[:, :3], data[:, 3], radscale, spacing)
    gl = quicksurf.GAUSSLIM[1]
    one = mm.Surf(0)
    one.set_grid(org, ext, res, (False,) * 3)
    one.set_params(mode=1, aggregator=0, normalize=0, radscale=radscale, gausslim=gl, colour=int(colour))
    one.push_particles(lists)
    one.compute_density()
    ref, refrgb = one.get_density(with_rgb=True)
    one.extract_isosurface(iso)
    if colour:
        rpos, rnrm, rcol = (a.copy() for a in one.get_mesh(colours=True))
    else:
        rpos, rnrm = (a.copy() for a in one.get_mesh())
        rcol = None
    one.close()
    assert float(ref.max()) > 1.0 and rpos.shape[0] > 1000 and (refrgb is not None) == colour
    g = mm.SurfGroup([0] * nchunks)
    try:
        g.set_grid(org, ext, res, (False,) * 3)
        g.set_params(mode=1, aggregator=0, normalize=0, radscale=radscale, gausslim=gl, colour=int(colour))
        for _ in range(2):
            g.clear_particles()
            g.push_particles(lists)
            g.compute_density()
            vol = g.get_density()
            rgb = g.get_colour_volume()
            g.extract_isosurface(iso)
            pos, nrm = g.get_mesh()
            col = g.get_mesh_colours(pos.shape[0] * 3)
            assert np.array_equal(vol.view(np.uint32), ref.view(np.uint32)), "volume differs from one context"
            assert pos.shape == rpos.shape and np.array_equal(pos, rpos) and np.array_equal(nrm, rnrm), "mesh differs from one context"
            if colour:
                assert np.array_equal(rgb.view(np.uint32), refrgb.view(np.uint32)), "colour volume differs from one context"
                assert np.array_equal(col.reshape(-1), rcol.reshape(-1)), "mesh colours differ from one context"
            else:
                assert rgb is None and col is None
    finally:
        g.close()
