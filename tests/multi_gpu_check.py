"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py spawns it): the z-slab pipeline with real NCCL exchange, device-side range
all-reduce, both halo-exchange paths (fused push over CUDA IPC / NCCL all-to-all-v) and every mesh-gather mode must reproduce the
single-GPU result bit for bit."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import megamol_b200 as mm  # noqa: E402
from megamol_b200 import slabs, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = dict(name="check", n=150_000, res=(64, 48, 40), kind="uniform", box=32.0)
    iso, radius = 0.35, 0.7
    ok = True
    for gather, exchange in (("host", "fused"), ("nccl", "fused"), ("fused", "fused"), ("host", "nccl")):
        job = slabs.SlabJob(w, rank, world, local, iso=iso, radius=radius, gather=gather, exchange=exchange)
        for _ in range(2):
            job.step_device()
        # volume: every rank's own cell planes (+ the last plane on the last rank)
        vol = job.surf.get_density()
        me = job.me
        a = me["cell_z0"] - me["z0"]
        mine = vol[a:a + me["cell_nz"] + (1 if rank == world - 1 else 0)]
        gvol, _ = slabs.gather_rows_to_root(torch.from_numpy(mine.reshape(mine.shape[0], -1).copy()).cuda(), rank, world)
        if gather == "host":
            pos, nrm = job.surf.get_mesh()
            gp, _ = slabs.gather_rows_to_root(torch.from_numpy(pos.reshape(-1, 9).copy()).cuda(), rank, world)
            gn, _ = slabs.gather_rows_to_root(torch.from_numpy(nrm.reshape(-1, 9).copy()).cuda(), rank, world)
        else:
            gp, gn = (job._gpos, job._gnrm) if rank == 0 else (None, None)
        if rank == 0:
            n_total = job.n_total
            xyz = synth.uniform_box(n_total, 1.0) * np.asarray(job.box, np.float32)
            s = mm.Surf(local)
            s.set_grid((0, 0, 0), job.box, job.res, job.cyclic)
            s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
            s.push_particles([dict(vtx=xyz, vtx_type=1, count=n_total, global_radius=radius)])
            s.compute_density()
            ref = s.get_density()
            s.extract_isosurface(iso)
            rpos, rnrm = s.get_mesh()
            s.close()
            v = gvol.cpu().numpy().reshape(ref.shape)
            ntri = rpos.shape[0]
            p = gp.view(-1)[:ntri * 9].cpu().numpy().reshape(-1, 3, 3)
            q = gn.view(-1)[:ntri * 9].cpu().numpy().reshape(-1, 3, 3)
            ev, ep, en = np.array_equal(v.view(np.uint32), ref.view(np.uint32)), np.array_equal(p, rpos), np.array_equal(q, rnrm)
            counts = job.tri_counts() if gather == "host" else job.last["tri_counts"]
            total = sum(counts)
            good = ev and ep and en and total == ntri and ntri > 10000
            if not good:
                print(f"   volume equal {ev} (max abs diff {np.abs(v - ref).max():.3e}, differing planes {np.unique(np.argwhere(v != ref)[:, 0])[:12]}), "
                      f"pos equal {ep}, nrm equal {en}, counts {counts} total {total} vs {ntri}", flush=True)
            print(f"[multi_gpu_check] exchange={exchange} gather={gather} world={world} triangles={ntri} exchange_ms={job.last.get('exchange_ms', 0):.3f} "
                  f"-> {'OK' if good else 'MISMATCH'}", flush=True)
            ok = ok and good
        job.close_keep_group() if hasattr(job, "close_keep_group") else None
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
