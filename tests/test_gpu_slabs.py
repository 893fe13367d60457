"""GPU suite: z-slab sharding through the C ABI (mms_set_slab).  The slabs are computed by separate contexts (as separate
ranks would) from the routed particle subsets; density, range and mesh must equal the unsharded result BIT FOR BIT --
the summation order (colour phase, canonical in-cell order) does not depend on the decomposition."""
import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import slabs, synth

pytestmark = pytest.mark.gpu


def run_unsharded(xyz, box, res, radius, cyclic, iso):
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), box, res, cyclic)
    s.set_params(mode=0, aggregator=0, normalize=1, defer_normalize=1, sigma=1.0)
    s.push_particles([dict(vtx=xyz, vtx_type=1, count=len(xyz), global_radius=radius)])
    s.compute_density()
    mn, mx = s.density_range()
    s.normalize(mn, mx)
    vol = s.get_density()
    s.extract_isosurface(iso)
    pos, nrm = s.get_mesh()
    s.close()
    return vol, (mn, mx), pos, nrm


@pytest.mark.parametrize("world,cyclic", [(2, True), (3, False), (4, True), (8, True)])
def test_slabs_bit_identical_to_unsharded(world, cyclic):
    n, res, radius, iso = 60000, (48, 40, 72), 0.55, 0.3
    box = (24.0, 20.0, 36.0)
    xyz = synth.uniform_box(n, 1.0, seed=909) * np.array(box, np.float32)
    cyc = (cyclic,) * 3
    vol, (mn, mx), pos, nrm = run_unsharded(xyz, box, res, radius, cyc, iso)
    plan = slabs.plan_slabs(res[2], world)
    sdz = np.float32(box[2]) / np.float32(res[2] - 1)
    Z, f = slabs.home_and_filter_z(xyz[:, 2], np.full(n, radius, np.float32), 0.0, sdz, np)
    masks = slabs.destination_masks(Z, f, plan, res[2], cyclic, np)
    ctxs, ranges = [], []
    for g, sl in enumerate(plan):
        s = mm.Surf(0)
        s.set_grid((0, 0, 0), box, res, cyc)
        s.set_slab(sl["z0"], sl["nz"], sl["cell_z0"], sl["cell_nz"])
        s.set_params(mode=0, aggregator=0, normalize=1, defer_normalize=1, sigma=1.0)
        part = np.ascontiguousarray(xyz[masks[g]])   # what the all-to-all would deliver, in global particle order
        s.push_particles([dict(vtx=part, vtx_type=1, count=len(part), global_radius=radius)])
        s.compute_density()
        ranges.append(s.density_range())
        ctxs.append(s)
    gmn, gmx = min(r[0] for r in ranges), max(r[1] for r in ranges)   # the one all-reduce of two floats
    assert (gmn, gmx) == (mn, mx)
    all_pos, all_nrm = [], []
    for sl, s in zip(plan, ctxs):
        s.normalize(gmn, gmx)
        v = s.get_density()
        assert np.array_equal(v.view(np.uint32), vol[sl["z0"]:sl["z0"] + sl["nz"]].view(np.uint32)), "slab density differs"
        s.extract_isosurface(iso)
        p, q = s.get_mesh()
        all_pos.append(p)
        all_nrm.append(q)
        s.close()
    gpos, gnrm = np.concatenate(all_pos), np.concatenate(all_nrm)
    assert gpos.shape == pos.shape and np.array_equal(gpos, pos) and np.array_equal(gnrm, nrm)


@pytest.mark.parametrize("cyclic", [False, True])
def test_routing_kernel_matches_host_logic(cyclic):
    """mms_route_particles (stable partition by destination slab incl. halo copies) against the array restatement."""
    import torch
    n, res, radius, world = 50000, (32, 32, 96), 0.9, 5
    box = (16.0, 16.0, 48.0)
    xyz = synth.uniform_box(n, 1.0, seed=77) * np.array(box, np.float32)
    xyz[:100, 2] = np.float32(box[2]) + np.float32(0.3)   # homes beyond the last plane
    plan = slabs.plan_slabs(res[2], world)
    sdz = np.float32(box[2]) / np.float32(res[2] - 1)
    Z, f = slabs.home_and_filter_z(xyz[:, 2], np.full(n, radius, np.float32), 0.0, sdz, np)
    masks = slabs.destination_masks(Z, f, plan, res[2], cyclic, np)
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), box, res, (cyclic,) * 3)
    s.set_params(mode=0, aggregator=0, normalize=0, sigma=1.0)
    d = torch.from_numpy(xyz).cuda()
    cap = 2 * n
    send = torch.zeros((cap, 3), device="cuda")
    counts = s.route_particles(d.data_ptr(), n, plan, send.data_ptr(), cap, global_radius=radius)
    s.synchronize()
    assert counts == [int(m.sum()) for m in masks]
    got = send.cpu().numpy()
    off = 0
    for g, m in enumerate(masks):
        assert np.array_equal(got[off:off + counts[g]], xyz[m]), g   # original order inside every destination group
        off += counts[g]
    s.close()
