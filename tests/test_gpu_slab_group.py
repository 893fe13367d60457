"""GPU suite, needs >= 2 GPUs (skipped on a single-GPU box): the single-process slab group of the C ABI (mms_slabs_*: one context per
device, fused halo push over peer memory, range combined by peer reads) against one GPU -- volume and mesh bit for bit."""
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
