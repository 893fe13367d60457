"""CPU suite: the multi-GPU host logic (z-slab planning, halo routing, all-to-all-v exchange, mesh gather) under
torch.distributed/gloo with world_size 2.  The per-slab compute is done by the CPU oracle here (the GPU library cannot run
in this container); what is under test is that sharding + exchange + gather reproduce the unsharded result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megamol_b200 import slabs, synth


def test_plan_slabs_covers_all_cells():
    for sz in (2, 3, 17, 64, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            plan = slabs.plan_slabs(sz, world)
            cells = []
            for s in plan:
                cells += list(range(s["cell_z0"], s["cell_z0"] + s["cell_nz"]))
                assert 0 <= s["z0"] and s["z0"] + s["nz"] <= sz and s["nz"] >= 1
                if s["cell_nz"]:
                    # the planes its cells need (+1 halo plane on interior sides for the gradients) are inside the slab
                    assert s["z0"] <= max(s["cell_z0"] - 1, 0)
                    assert s["z0"] + s["nz"] - 1 >= min(s["cell_z0"] + s["cell_nz"] + 1, sz - 1)
            assert cells == list(range(sz - 1))


@pytest.mark.parametrize("cyclic", [False, True])
def test_destination_masks_match_bruteforce(cyclic):
    rng = np.random.default_rng(3)
    sz, world = 40, 4
    plan = slabs.plan_slabs(sz, world)
    z = (rng.random(5000).astype(np.float32) * 44 - 2).astype(np.float32)
    r = (rng.random(5000).astype(np.float32) * 2.5).astype(np.float32)
    sdz = np.float32(40.0 / 39.0)
    Z, f = slabs.home_and_filter_z(z, r, 0.0, sdz, np)
    masks = slabs.destination_masks(Z, f, plan, sz, cyclic, np)
    for g, s in enumerate(plan):
        want = np.zeros(len(z), bool)
        for i in range(len(z)):
            for h in range(Z[i] - f[i], Z[i] + f[i] + 1):
                t = h % sz if cyclic else h
                if s["z0"] <= t < s["z0"] + s["nz"]:
                    want[i] = True
        assert np.array_equal(masks[g], want), g


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cyclic, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle_binding as ob
        o = ob.Oracle()
        n, box, res, radius, iso = 6000, 12.0, (20, 18, 30), 0.9, 0.35
        xyz_all = synth.uniform_box(n, box, seed=404)
        i0, i1 = n * rank // world, n * (rank + 1) // world
        mine = torch.from_numpy(xyz_all[i0:i1].copy())
        plan = slabs.plan_slabs(res[2], world)
        me = plan[rank]
        sdz = float(np.float32(box) / np.float32(res[2] - 1))
        recv = slabs.route_and_exchange(mine, radius, plan, res[2], 0.0, sdz, cyclic)
        part = recv.numpy()
        lists = [dict(vtx=part, vtx_type=1, count=len(part), global_radius=radius)]
        cyc = (cyclic,) * 3
        vol, (mn, mx) = o.density_p2d(lists, (0, 0, 0), (box,) * 3, res, cyc, z0=me["z0"], nz=me["nz"])
        # global range with ONE all-reduce of two floats (instead of the reference's whole-volume MPI_Allreduce)
        t = torch.tensor([-mn, mx])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gmn, gmx = -float(t[0]), float(t[1])
        vol = o.normalize(vol, gmn, gmx)
        # marching cubes on this rank's cell layers: planes [cell_z0, cell_z0+cell_nz] plus one gradient halo plane each side
        a = me["cell_z0"] - me["z0"]
        lo = max(a - 1, 0)
        hi = min(a + me["cell_nz"] + 1, me["nz"] - 1)
        sub = vol[lo:hi + 1]
        sd = np.array([box / np.float32(r - 1) for r in res], np.float32)
        pos, nrm, _ = o.mc_emit(sub, (0, 0, 0), sd, iso, z_offset=me["z0"] + lo)
        # keep only triangles of my own cell layers (the halo layers belong to the neighbours)
        _, counts, _ = o.mc_count(sub, iso)
        per_layer = counts.reshape(counts.shape[0], -1).sum(1)
        first = int(per_layer[:a - lo].sum())
        mine_n = int(per_layer[a - lo:a - lo + me["cell_nz"]].sum())
        pos, nrm = pos[first:first + mine_n], nrm[first:first + mine_n]
        gpos, cnts = slabs.gather_rows_to_root(torch.from_numpy(pos.reshape(-1, 9).copy()), rank, world)
        gnrm, _ = slabs.gather_rows_to_root(torch.from_numpy(nrm.reshape(-1, 9).copy()), rank, world)
        gvol, _ = slabs.gather_rows_to_root(torch.from_numpy(vol[a:a + me["cell_nz"] + (1 if rank == world - 1 else 0)].reshape(-1, res[0] * res[1]).copy()), rank, world)
        if rank == 0:
            full, (fmn, fmx) = o.density_p2d([dict(vtx=xyz_all, vtx_type=1, count=n, global_radius=radius)], (0, 0, 0), (box,) * 3, res, cyc,
                                             normalize=True)
            fpos, fnrm, _ = o.mc_emit(full, (0, 0, 0), sd, iso)
            ok_vol = np.array_equal(gvol.numpy().reshape(full.shape).view(np.uint32), full.view(np.uint32))
            ok_rng = (gmn, gmx) == (fmn, fmx)
            ok_pos = gpos.shape[0] == fpos.shape[0] and np.array_equal(gpos.numpy().reshape(-1, 3, 3), fpos)
            ok_nrm = np.array_equal(gnrm.numpy().reshape(-1, 3, 3), fnrm)
            q.put((ok_vol, ok_rng, ok_pos, ok_nrm, int(fpos.shape[0]), cnts))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cyclic", [False, True])
def test_sharded_pipeline_equals_unsharded_gloo(cyclic):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cyclic, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_vol, ok_rng, ok_pos, ok_nrm, ntri, cnts = q.get(timeout=5)
    assert ntri > 1000 and sum(cnts) == ntri
    assert ok_vol, "sharded density must be bit-identical to the unsharded one"
    assert ok_rng and ok_pos and ok_nrm
