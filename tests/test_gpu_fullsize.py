"""GPU suite at BASELINE.json's full single-GPU size (config 2: 10 M LJ-fluid-like particles -> 512^3 -> marching cubes, the bench
workload): the CUDA path against the oracle on the WHOLE problem -- the oracle's z-partitioned threads finish the density in a few
seconds -- plus the size-independent properties (idempotence, decomposition invariance) where a direct comparison would need the
8.4 GB mesh on the host."""
import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

N, RES, RADIUS, ISO = 10_000_000, (512, 512, 512), 0.5, 0.5


@pytest.fixture(scope="module")
def c2():
    xyz, box = synth.lj_fluid(N)
    return dict(xyz=xyz, box=box, lists=[dict(vtx=xyz, vtx_type=1, count=N, global_radius=RADIUS)])


def test_c2_full_size_against_the_oracle(c2, oracle):
    box, lists = c2["box"], c2["lists"]
    s = mm.Surf(0)
    try:
        s.set_grid((0, 0, 0), (box,) * 3, RES, (True,) * 3)
        s.set_params(mode=0, aggregator=0, normalize=1, defer_normalize=0, sigma=1.0, want_home_voxels=1, want_cell_tricounts=1)
        s.push_particles(lists)
        s.compute_density()
        # (1) binning: every home voxel, bit-exact
        assert np.array_equal(s.home_voxels(), oracle.home_voxels(lists, (0, 0, 0), (box,) * 3, RES, (1, 1, 1)))
        # (2) density: the whole normalised volume within 1e-5 relative
        vol = s.get_density()
        mn, mx = s.density_range()
        ref, (rmn, rmx) = oracle.density_p2d(lists, (0, 0, 0), (box,) * 3, RES, (1, 1, 1), sigma=1.0, normalize=True)
        assert abs(mx - rmx) <= 1e-5 * rmx and mn == rmn == 0.0
        floor = H.DENSITY_FLOOR / rmx  # the floor of helpers.density_close, in normalised units
        worst = 0.0
        for z in range(0, RES[2], 64):
            g, r = vol[z:z + 64].astype(np.float64), ref[z:z + 64].astype(np.float64)
            worst = max(worst, float((np.abs(g - r) / np.maximum(np.abs(r), floor)).max()))
        assert worst < H.DENSITY_RTOL, worst
        # (3) marching cubes: per-cell triangle counts on the same volume, bit-exact over all 133 M cells
        s.extract_isosurface(ISO)
        counts = s.cell_tricounts()
        total, ref_counts, _ = oracle.mc_count(vol, ISO)
        assert np.array_equal(counts, ref_counts)
        assert s.count_isosurface(ISO) == total and total > 100_000_000
        del ref, ref_counts, counts
        # (4) vertices and normals of twelve cell layers in the middle of the volume (the slab API recomputes exactly these planes;
        #     one extra plane on each side makes the gradients those of the whole volume)
        z0, nz, cz0, cnz = 199, 15, 200, 12
        s.set_slab(z0, nz, cz0, cnz)
        s.set_params(want_home_voxels=0, want_cell_tricounts=0, defer_normalize=1)
        s.clear_particles()
        s.push_particles(lists)
        s.compute_density()
        s.normalize(mn, mx)
        sub = s.get_density()
        assert np.array_equal(sub.view(np.uint32), vol[z0:z0 + nz].view(np.uint32)), "slab planes differ from the whole volume's"
        s.extract_isosurface(ISO)
        pos, nrm = s.get_mesh()
        sd = np.array([np.float32(box) / np.float32(r - 1) for r in RES], np.float32)
        _, layer_counts, _ = oracle.mc_count(sub, ISO)
        skip = int(layer_counts[:cz0 - z0].sum(dtype=np.int64))
        take = int(layer_counts[cz0 - z0:cz0 - z0 + cnz].sum(dtype=np.int64))
        rpos, rnrm, _ = oracle.mc_emit(sub, (0, 0, 0), sd, ISO, z_offset=z0)
        assert pos.shape[0] == take > 1_000_000
        rpos, rnrm = rpos[skip:skip + take], rnrm[skip:skip + take]
        assert np.abs((pos - rpos) / sd.astype(np.float64)).max() <= H.VERTEX_TOL_CELLS
        assert np.abs(nrm - rnrm).max() < 1e-4
    finally:
        s.close()


def test_c2_full_size_is_reproducible(c2):
    """Idempotence at full size without moving the mesh: two runs give the same volume bits and the same triangle count, and the
    device-side mesh of the second run equals the first one's (compared on the GPU)."""
    import torch
    box, lists = c2["box"], c2["lists"]
    s = mm.Surf(0)
    try:
        sums = []
        keep = None
        for _ in range(2):
            s.clear_particles()
            s.set_grid((0, 0, 0), (box,) * 3, RES, (True,) * 3)
            s.set_params(mode=0, aggregator=0, normalize=1, defer_normalize=0, sigma=1.0)
            s.push_particles(lists)
            s.compute_density()
            vol = s.get_density(copy=False)
            n = s.count_isosurface(ISO)
            pos = torch.empty(n * 9, device="cuda")
            nrm = torch.empty(n * 9, device="cuda")
            s.emit_isosurface(pos.data_ptr(), nrm.data_ptr(), None, 0)
            s.synchronize()
            sums.append((int(vol.view(np.uint32).sum(dtype=np.uint64)), n))
            if keep is None:
                keep = (pos, nrm)
            else:
                assert torch.equal(keep[0], pos) and torch.equal(keep[1], nrm)
        assert sums[0] == sums[1]
    finally:
        s.close()


def test_count_kernel_with_rebalanced_layers(oracle):
    """mc_count_kernel picks the number of cell layers a block marches so that the grid fills whole waves of resident blocks
    (McGeo::countLayers).  A 1024 x 1024 x 161 volume has 128 block columns x 10 default z-blocks = 1280 blocks for 1184 resident
    ones on a B200, so the kernel runs with 18 instead of 16 layers per block: every per-cell triangle count equals the oracle's."""
    res = (1024, 1024, 161)
    z, y, x = np.ogrid[0:res[2], 0:res[1], 0:res[0]]
    vol = (np.sin(0.05 * x) + np.sin(0.043 * y) + np.sin(0.061 * z)).astype(np.float32)
    vol += (synth.uniform(4242, 0, 4096, 0).astype(np.float32) * np.float32(0.02))[(x * 7 + y * 13 + z * 29) % 4096]
    s = mm.Surf(0)
    try:
        s.set_grid((0, 0, 0), tuple(float(r - 1) for r in res), res, (False,) * 3)
        s.set_params(want_cell_tricounts=1)
        s.set_density(vol)
        s.extract_isosurface(1.1)
        counts = s.cell_tricounts()
        total = s.count_isosurface(1.1)
        ref_total, ref_counts, _ = oracle.mc_count(vol, 1.1)
        assert total == ref_total and total > 1_000_000
        assert np.array_equal(counts, ref_counts)
    finally:
        s.close()
