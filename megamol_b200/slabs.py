"""Multi-GPU driver: z-slab sharding with halo, one process per GPU.

Rank g owns marching-cubes cell layers [c_g, c_{g+1}), c_g = floor(g * (sz-1) / G), and computes the density planes
those cells need: [c_g - 1, c_{g+1} + 1] clipped to the grid (one extra plane on each interior side so that the
gradient normals equal the unsharded ones).  A particle is needed by every slab its support box [Z-f, Z+f] touches
(periodic images in cyclic mode).  The ONLY data-path exchanges are

  1. particles:  all-to-all-v of 16-byte xyzr records (each rank starts with a contiguous chunk of the frame);
                 receivers concatenate in source-rank order, so the global particle order is preserved,
  2. normalise:  one all-reduce of (min, max)  -- instead of the reference's whole-volume MPI_Allreduce
                 (plugins/datatools/src/MPIVolumeAggregator.cpp:129),
  3. mesh:       all-gather of triangle counts + send/recv of the per-slab vertex arrays to rank 0.

torch.distributed is plumbing only (NCCL on the GPUs, gloo in the CPU tests); the compute is libmmsurf.
The planning functions are pure and are what tests/test_slabs_cpu.py exercises under gloo with world_size 2.
"""
from __future__ import annotations

import numpy as np


def plan_slabs(sz: int, world: int):
    """-> list of dicts(cell_z0, cell_nz, z0, nz) per rank."""
    out = []
    ncell = sz - 1
    for g in range(world):
        c0 = (g * ncell) // world
        c1 = ((g + 1) * ncell) // world
        p0 = max(c0 - 1, 0)
        p1 = min(c1 + 1, sz - 1)  # inclusive
        if c1 <= c0:  # more ranks than cell layers: this rank idles on an (empty) one-plane slab
            p0, p1 = min(c0, sz - 1), min(c0, sz - 1)
        out.append(dict(cell_z0=c0, cell_nz=c1 - c0, z0=p0, nz=p1 - p0 + 1))
    return out


def home_and_filter_z(z, r, zmin, sdz, xp):
    """Bit-exact fp32 restatement of static_cast<int>((z - minOS)/sliceDist) and (int)ceil(rad/sliceDist)
    (ParticlesToDensity.cpp:567,575) with array ops (xp = numpy or torch)."""
    if xp is np:
        Z = np.trunc((z.astype(np.float32) - np.float32(zmin)) / np.float32(sdz)).astype(np.int64)
        f = np.ceil(r.astype(np.float32) / np.float32(sdz)).astype(np.int64)
    else:
        import torch
        Z = torch.trunc((z - zmin) / sdz).to(torch.int64)
        f = torch.ceil(r / sdz).to(torch.int64)
    return Z, f


def destination_masks(Z, f, slabs, sz: int, cyclic_z: bool, xp):
    """mask[g][i] = particle i is needed by slab g."""
    masks = []
    if cyclic_z:
        Zw = Z % sz  # floor-mod for numpy and torch
    for s in slabs:
        lo, hi = s["z0"], s["z0"] + s["nz"] - 1
        if not cyclic_z:
            m = (Z + f >= lo) & (Z - f <= hi)
        else:
            a, b = Zw - f, Zw + f
            m = ((b >= lo) & (a <= hi)) | ((b - sz >= lo) & (a - sz <= hi)) | ((b + sz >= lo) & (a + sz <= hi)) | (2 * f + 1 >= sz)
        masks.append(m)
    return masks


def route_and_exchange(xyz, radius, slabs, sz, zmin, sdz, cyclic_z):
    """All-to-all-v of particle records by destination slab (torch tensors on any device; NCCL or gloo).
    xyz: [n, k] float32 rows whose column 2 is z; radius: float (global radius) or [n] tensor.
    Returns the rows this rank needs (own slab + halo), concatenated in SOURCE-RANK order = global particle order."""
    import torch
    import torch.distributed as dist
    r = radius if torch.is_tensor(radius) else torch.full((xyz.shape[0],), float(radius), device=xyz.device, dtype=torch.float32)
    Z, f = home_and_filter_z(xyz[:, 2], r, zmin, sdz, torch)
    masks = destination_masks(Z, f, slabs, sz, cyclic_z, torch)
    parts = [xyz[m] for m in masks]
    send_counts = torch.tensor([p.shape[0] for p in parts], device=xyz.device, dtype=torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    send = torch.cat(parts, 0).contiguous()
    recv = torch.empty((sum(rc), xyz.shape[1]), device=xyz.device, dtype=xyz.dtype)
    dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
    return recv


def gather_rows_to_root(local, rank, world, out=None):
    """all-gather of row counts, then every rank's [n_g, k] tensor travels to rank 0 (send/recv), concatenated in rank
    (= slab = cell-linear) order.  Returns (tensor on rank 0 | None, counts)."""
    import torch
    import torch.distributed as dist
    cnt = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    allc = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = [int(c.item()) for c in allc]
    if rank == 0:
        total = sum(counts)
        if out is None or out.shape[0] < total:
            out = torch.empty((total,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        out[:counts[0]].copy_(local)
        off = counts[0]
        ops = []
        for g in range(1, world):
            if counts[g]:
                ops.append(dist.P2POp(dist.irecv, out[off:off + counts[g]], g))
            off += counts[g]
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return out, counts
    if local.shape[0]:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), 0)]):
            req.wait()
    return None, counts


class SlabJob:
    """The bench/test driver: generates this rank's chunk of the synthetic frame and runs full steps."""

    def __init__(self, w, rank, world, local, iso, radius, sigma=1.0, cyclic=(True, True, True), normalize=True, gather="host",
                 exchange="fused"):
        import torch
        import megamol_b200 as mm
        from megamol_b200 import synth
        self.torch = torch
        self.w, self.rank, self.world, self.iso, self.radius = w, rank, world, iso, radius
        self.local = local
        # where the per-slab meshes end up:
        #   "host"  (default) they stay in their slab's HBM; the triangle counts are all-gathered so that every rank knows its offset
        #           in the frame's mesh, and (e2e) every rank copies its slab over ITS OWN PCIe link to that offset of the host mesh
        #           -- CallTriMeshData is a host-memory contract, the mesh never has to sit on one GPU
        #   "fused" every rank's mc_emit_kernel stores straight into rank 0's buffers over NVLink (CUDA IPC): emit + gather in one kernel
        #   "nccl"  emit locally, then NCCL send/recv to rank 0 (the baseline the fused kernel is compared with)
        self.gather = gather
        # how the halo particles travel:
        #   "fused" (default) every rank's halo_push_kernel appends what the other slabs need straight to THEIR receive buffers (CUDA IPC
        #           mappings, stores and counter atomics over NVLink); the only collective is a one-word all-reduce that orders "all pushes
        #           done" on the stream.  No host synchronisation, no count matrix, no all-to-all.
        #   "nccl"  the round-1 path, kept as the baseline: routing kernels -> host sync for the counts -> count matrix all-gather ->
        #           all-to-all-v
        self.exchange = exchange
        self._root = dict(gen=0, cap=0, pos=None, nrm=None)   # rank 0: owned buffers; others: IPC mappings
        self.dev = torch.device("cuda", local)
        self.normalize = normalize
        self.cyclic = cyclic
        # weak scaling: the lattice (and the grid) grows along z with the number of ranks
        n1 = w["n"]
        res = list(w["res"])
        self.protein = w["kind"] == "protein"
        if self.protein:
            if world != 1:
                raise SystemExit("the protein / QuickSurf workload (C3) is a single-GPU configuration")
            from megamol_b200 import quicksurf
            data, _, _ = synth.protein_like(n1)
            self.n_total = n1
            org, ext, _ = quicksurf.grid_from_particles(data[:, :3], data[:, 3], 1.0, 1.0)
            # fixed 512^3 grid: gridspacing = padded extent / 512
            self.origin = tuple(float(v) for v in org)
            pad_ext = ext + 1.0
            self.box = tuple(float(v) for v in pad_ext)
            # useful (voxel, atom) pairs, SURVEY 8d: sum over the atoms of (4/3) pi (cut-off / voxel size)^3, cut-off = gausslim * radscale * r
            h = np.asarray(pad_ext, np.float64) / (np.asarray(res, np.float64) - 1.0)
            self.gauss_pairs = float(np.sum((4.0 / 3.0) * np.pi * (3.0 * data[:, 3].astype(np.float64)) ** 3) / float(np.prod(h)))
        elif w["kind"] == "lj" and w.get("fixed_total"):
            # strong-scaling configuration (BASELINE configs[3]): the frame is fixed, the ranks share it
            Lc = int(w["lattice"])
            self.n_total = n1
            a = np.float32(1.0794)
            self.lat = (Lc, Lc, -(-self.n_total // (Lc * Lc)))
            self.box = (float(np.float32(Lc) * a),) * 3
        elif w["kind"] == "lj":
            L1 = int(np.ceil(n1 ** (1 / 3) - 1e-9))
            self.n_total = n1 * world
            a = np.float32(1.0794)
            self.lat = (L1, L1, -(-self.n_total // (L1 * L1)))
            self.box = (float(np.float32(L1) * a), float(np.float32(L1) * a), float(np.float32(self.lat[2]) * a))
            res[2] = res[2] * world
        else:
            self.n_total = n1 * world
            self.box = (w["box"], w["box"], w["box"] * world)
            res[2] = res[2] * world
        self.res = tuple(res)
        i0 = (self.n_total * rank) // world
        i1 = (self.n_total * (rank + 1)) // world
        if self.protein:
            xyz = data
        elif w["kind"] == "lj":
            xyz = self._lj_chunk(i0, i1)
        else:
            xyz = synth.uniform_box(self.n_total, 1.0, i0=i0, i1=i1) * np.asarray(self.box, np.float32)
        self.n_local = i1 - i0
        # host copy in pinned memory (what an MMPLD reader would fill), device copy for the resident arm
        self.h_xyz = torch.empty((self.n_local, xyz.shape[1]), dtype=torch.float32, pin_memory=True)
        self.h_xyz.numpy()[:] = xyz
        self.d_xyz = self.h_xyz.to(self.dev)
        self.slabs = plan_slabs(self.res[2], world)
        self.me = self.slabs[rank]
        self.surf = mm.Surf(local)
        if world > 1:  # one stream for the library's kernels and torch's collectives: ordering without host synchronisation
            self.surf.set_stream(torch.cuda.current_stream().cuda_stream)
        if self.protein:
            self.cyclic = cyclic = (False, False, False)
            self.normalize = False
            self.surf.set_grid(self.origin, self.box, self.res, cyclic)
            self.surf.set_params(mode=1, aggregator=0, normalize=0, radscale=1.0, gausslim=3.0, colour=1)
        else:
            self.surf.set_grid((0, 0, 0), self.box, self.res, cyclic)
            if world > 1:
                self.surf.set_slab(self.me["z0"], self.me["nz"], self.me["cell_z0"], self.me["cell_nz"])
            self.surf.set_params(mode=0, aggregator=0, normalize=int(normalize), defer_normalize=int(world > 1), sigma=sigma)
        self.sdz = float(np.float32(self.box[2]) / np.float32(self.res[2] - 1))
        if world > 1 and self.exchange == "fused":
            self._halo_setup()
        self._keep = []
        self.last = {}
        self.gather_ms = 0.0
        self.exchange_ms = 0.0

    def _lj_chunk(self, i0, i1):
        from megamol_b200 import synth
        Lx, Ly, Lz = self.lat
        idx = np.arange(i0, i1, dtype=np.int64)
        lat = (idx % Lx, (idx // Lx) % Ly, idx // (Lx * Ly))
        out = np.empty((i1 - i0, 3), np.float32)
        a = np.float32(1.0794)
        for k in range(3):
            u = synth.uniform(synth.SEED + 2, i0, i1, k)
            out[:, k] = (lat[k].astype(np.float32) + np.float32(0.5) + (u * np.float32(2) - np.float32(1)) * np.float32(0.15)) * a
        return out

    # ---- plumbing ---------------------------------------------------------------------------------------------
    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K steps between barriers; device time (CUDA events on the current torch stream AND the library stream are
        both drained by the trailing synchronize); returns max over ranks in ms."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self.surf.timer_start()
        for _ in range(steps):
            fn()
        lib_ms = self.surf.timer_stop()
        e1.record()
        torch.cuda.synchronize()
        ms = max(lib_ms, e0.elapsed_time(e1))
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        self.barrier()
        return ms

    def _halo_setup(self):
        """Once: every rank's receive buffer and counter block, exported over CUDA IPC and mapped by all the others."""
        import ctypes as C
        import torch.distributed as dist
        L = self.surf.L
        # a slab receives at most what its neighbours' halos hold; the frames here are z-ordered, so a rank's own share bounds it amply.
        # (An overflow is detected on the device and reported by the next host-synchronising call.)
        self._halo_cap = -(-self.n_total // self.world) + 1 + (1 << 20)  # the same on every rank (the buffer halves are addressed with it)
        buf, ctr = self.surf.halo_buffers(self._halo_cap)
        mine = []
        for ptr in (buf, ctr):
            h = (C.c_ubyte * 64)()
            if L.mms_ipc_export(self.local, C.c_void_p(ptr), h):
                raise RuntimeError("cudaIpcGetMemHandle failed")
            mine.append(bytes(h))
        allh = [None] * self.world
        dist.all_gather_object(allh, (mine[0], mine[1], self._halo_cap))
        self._peer_bufs, self._peer_ctrs, self._peer_open = [], [], []
        caps = []
        for g, (hb, hc, cap) in enumerate(allh):
            caps.append(cap)
            if g == self.rank:
                self._peer_bufs.append(buf)
                self._peer_ctrs.append(ctr)
                continue
            ptrs = []
            for raw in (hb, hc):
                hh = (C.c_ubyte * 64).from_buffer_copy(raw)
                p = C.c_void_p()
                if L.mms_ipc_open(self.local, hh, C.byref(p)):
                    raise RuntimeError("cudaIpcOpenMemHandle failed (is peer access available between the GPUs?)")
                ptrs.append(p.value)
                self._peer_open.append(p.value)
            self._peer_bufs.append(ptrs[0])
            self._peer_ctrs.append(ptrs[1])
        self._peer_cap = min(caps)
        self._tick = self.torch.zeros(1, device=self.dev, dtype=self.torch.float32)
        import os
        self._halo_allreduce = bool(os.environ.get("MMS_HALO_ALLREDUCE"))

    def _halo(self):
        """Fused halo exchange of the lists pushed so far: one push kernel per list and a signal kernel; the stream then waits on a flag
        in this GPU's memory until every other rank's records have landed (no collective, no host round trip); the received records join the
        frame as one more list (its length stays on the device).  MMS_HALO_ALLREDUCE=1: the earlier one-word all-reduce as that point."""
        self.surf.halo_push(self.slabs, self.rank, self._peer_bufs, self._peer_ctrs, self._peer_cap)
        if self._halo_allreduce:
            import torch.distributed as dist
            dist.all_reduce(self._tick)
        else:
            self.surf.halo_wait(self.world - 1)
        self.surf.halo_receive(self.radius)

    def _exchange(self, xyz_dev):
        """Halo exchange.  The rank's own chunk stays where it is (the binning kernel drops what does not reach the slab); only the records
        OTHER slabs need are extracted (libmmsurf's routing kernels with the own slab switched off: mms_route_particles) and travel in one
        all-to-all-v.  The canonical in-cell order makes the result independent of the order in which lists are pushed, so the pieces need
        not be re-assembled in global particle order.  Host synchronisations: the routing counts, and the count matrix.
        Returns the received records as one [n, k] tensor."""
        torch = self.torch
        import torch.distributed as dist
        n = xyz_dev.shape[0]
        cap = n + 4096
        if getattr(self, "_send", None) is None or self._send.shape[0] < cap:
            self._send = torch.empty((cap, xyz_dev.shape[1]), device=self.dev, dtype=torch.float32)
        others = [dict(z0=1, nz=0) if g == self.rank else s for g, s in enumerate(self.slabs)]   # plane_lo > plane_hi: switched off
        sc = self.surf.route_particles(xyz_dev.data_ptr(), n, others, self._send.data_ptr(), cap, global_radius=self.radius)
        mine = torch.tensor(sc, device=self.dev, dtype=torch.int64)
        if getattr(self, "_cmat", None) is None:
            self._cmat = torch.empty((self.world, self.world), device=self.dev, dtype=torch.int64)
        dist.all_gather_into_tensor(self._cmat, mine)
        rc = self._cmat[:, self.rank].tolist()         # what every source sends to me
        nrecv = sum(rc)
        if getattr(self, "_recv", None) is None or self._recv.shape[0] < nrecv:
            self._recv = torch.empty((nrecv + nrecv // 4 + 4096, xyz_dev.shape[1]), device=self.dev, dtype=torch.float32)
        recv = self._recv[:nrecv]
        dist.all_to_all_single(recv, self._send[:sum(sc)], output_split_sizes=rc, input_split_sizes=sc)
        return recv

    def _more(self, xyz_dev):
        """-> (extra (pointer, count) pieces for _compute, number of received records or -1 where only the device knows it)"""
        if self.exchange == "fused":
            return (), -1
        recv = self._exchange(xyz_dev)
        return ((recv.data_ptr(), recv.shape[0]),), int(recv.shape[0])

    def _gather_mesh(self):
        """all-gather of triangle counts, then the per-slab vertex/normal arrays travel to rank 0 over NCCL."""
        torch = self.torch
        nverts, ppos, pnrm = self.surf.mesh_device()

        def view(ptr, n):
            if n == 0:
                return torch.empty((0, 3), device=self.dev, dtype=torch.float32)
            return _tensor_from_ptr(torch, ptr, n * 3, self.dev).view(n, 3)
        self._gpos, counts = gather_rows_to_root(view(ppos, nverts), self.rank, self.world, getattr(self, "_gpos", None))
        self._gnrm, _ = gather_rows_to_root(view(pnrm, nverts), self.rank, self.world, getattr(self, "_gnrm", None))
        self.last["tri_counts"] = [c // 3 for c in counts]
        self.last["gathered_verts"] = sum(counts)

    def _allgather_counts(self, ntris=None):
        """mesh stays sharded: all-gather of the per-slab triangle counts -> every rank knows its offset in the frame's mesh.
        The collective is only enqueued here; the host reads the counts when somebody asks for them (tri_counts()).
        ntris given (the count is known, the emit kernel not yet launched): the collective goes to a side stream and runs UNDER the
        emit kernel; _join_counts() makes the main stream wait for it."""
        torch = self.torch
        import torch.distributed as dist
        if getattr(self, "_tcounts", None) is None:
            self._tcounts = torch.empty((self.world,), device=self.dev, dtype=torch.int64)
            self._side = torch.cuda.Stream(device=self.dev)
        self.last.pop("tri_counts", None)
        self.last["gathered_verts"] = 0
        if ntris is None:
            nverts, _, _ = self.surf.mesh_device()
            cnt = torch.tensor([nverts // 3], device=self.dev, dtype=torch.int64)
            dist.all_gather_into_tensor(self._tcounts, cnt)
            return
        with torch.cuda.stream(self._side):
            cnt = torch.tensor([int(ntris)], device=self.dev, dtype=torch.int64)
            dist.all_gather_into_tensor(self._tcounts, cnt)
        self._counts_pending = True

    def _join_counts(self):
        if getattr(self, "_counts_pending", False):
            self.torch.cuda.current_stream().wait_stream(self._side)
            self._counts_pending = False

    def tri_counts(self):
        """per-slab triangle counts of the last step (host list; synchronises if they are still on the device)"""
        if "tri_counts" not in self.last:
            counts = self._tcounts.tolist()
            self.last["tri_counts"] = counts
            self.last["tri_offset"] = sum(counts[:self.rank])
        return self.last["tri_counts"]

    # ---- fused emit + gather -------------------------------------------------------------------------------------
    def _emit_to_root(self):
        """Marching-cubes emission and mesh gather in ONE kernel: every rank's mc_emit_kernel writes its slab's triangles
        at their final offset of rank 0's mesh buffers (CUDA-IPC mapping, stores travel over NVLink).  The only collectives
        left are the all-gather of the triangle counts and a 144-byte broadcast of (generation, IPC handles)."""
        import ctypes as C
        torch = self.torch
        import torch.distributed as dist
        L = self.surf.L
        T = self.surf.count_isosurface(self.iso)
        cnt = torch.tensor([T], device=self.dev, dtype=torch.int64)
        allc = [torch.empty_like(cnt) for _ in range(self.world)]
        dist.all_gather(allc, cnt)
        counts = [int(c.item()) for c in allc]
        total, first = sum(counts), sum(counts[:self.rank])
        R = self._root
        msg = torch.zeros(144, dtype=torch.uint8, device=self.dev)
        if self.rank == 0:
            need = total * 36
            if R["cap"] < need:
                # the other ranks still map the current buffers (cudaIpcOpenMemHandle) and close those mappings only when they see
                # the new generation below: freeing an exported allocation before its importers have closed it is undefined
                # behaviour, so the buffers retire for one generation and are freed at the NEXT regrow (or at close, after a barrier)
                for ptr in R.get("retired", []):
                    L.mms_device_free(self.local, ptr)
                R["retired"] = [R[k] for k in ("pos", "nrm") if R[k]]
                cap = need + need // 8 + 256
                hb = bytearray(144)
                for i, k in enumerate(("pos", "nrm")):
                    p = C.c_void_p()
                    if L.mms_device_alloc(self.local, cap, C.byref(p)):
                        raise MemoryError(f"gather buffer of {cap} bytes")
                    R[k] = p.value
                    h = (C.c_ubyte * 64)()
                    if L.mms_ipc_export(self.local, p, h):
                        raise RuntimeError("cudaIpcGetMemHandle failed")
                    hb[16 + 64 * i:16 + 64 * (i + 1)] = bytes(h)
                R["cap"], R["gen"] = cap, R["gen"] + 1
                hb[0:8] = int(R["gen"]).to_bytes(8, "little")
                hb[8:16] = int(cap).to_bytes(8, "little")
                R["msg"] = bytes(hb)
            msg = torch.frombuffer(bytearray(R["msg"]), dtype=torch.uint8).to(self.dev)
        dist.broadcast(msg, 0)
        if self.rank != 0:
            raw = bytes(msg.cpu().numpy())
            gen = int.from_bytes(raw[0:8], "little")
            if gen != R["gen"]:
                for k in ("pos", "nrm"):
                    if R[k]:
                        L.mms_ipc_close(self.local, R[k])
                for i, k in enumerate(("pos", "nrm")):
                    h = (C.c_ubyte * 64).from_buffer_copy(raw[16 + 64 * i:16 + 64 * (i + 1)])
                    p = C.c_void_p()
                    if L.mms_ipc_open(self.local, h, C.byref(p)):
                        raise RuntimeError("cudaIpcOpenMemHandle failed (is peer access available between the GPUs?)")
                    R[k] = p.value
                R["gen"], R["cap"] = gen, int.from_bytes(raw[8:16], "little")
        self.surf.emit_isosurface(R["pos"], R["nrm"], None, first)
        self.surf.synchronize()
        dist.barrier()   # rank 0 may read the mesh once every rank's stores have landed
        self.last["tri_counts"] = counts
        self.last["gathered_verts"] = total * 3
        if self.rank == 0:
            self._gpos = _tensor_from_ptr(torch, R["pos"], total * 9, self.dev)
            self._gnrm = _tensor_from_ptr(torch, R["nrm"], total * 9, self.dev)

    # ---- steps ------------------------------------------------------------------------------------------------
    def _compute(self, xyz_ptr, n, extract=True, more=(), prefetch=False):
        """more: further (device pointer, count) pieces of the frame (the received halo); prefetch: start the volume's D2H copy before
        the isosurface kernels (end-to-end arm: the copy overlaps marching cubes)"""
        s = self.surf
        s.clear_particles()
        if self.protein:  # x y z r | R G B A interleaved, stride 32 (FLOAT_XYZR + FLOAT_RGBA)
            s.push_particles([dict(vtx=xyz_ptr, vtx_type=2, vtx_stride=32, count=n, col=xyz_ptr + 16, col_type=4, col_stride=32)])
        else:
            s.push_particles([dict(vtx=p, vtx_type=1, count=c, global_radius=self.radius) for p, c in ((xyz_ptr, n),) + tuple(more) if c > 0])
        if self.world > 1 and self.exchange == "fused":
            ev0, ev1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
            ev0.record()
            self._halo()
            ev1.record()
            self._halo_events = (ev0, ev1)
        s.compute_density()
        if self.world > 1 and self.normalize:
            # global range with ONE max-all-reduce of {-min, max}, in place on the library's device buffer: no host round trip
            import torch.distributed as dist
            ptr = s.density_range_device()
            t = _tensor_from_ptr(self.torch, ptr, 2, self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s.normalize_device(ptr)
        if prefetch:
            s.prefetch_density()
        if extract:
            if self.world > 1 and self.gather == "host" and not self.protein:
                # the library launches the emit kernel speculatively right behind the count (mms_extract_isosurface); the count's all-gather
                # goes to a side stream and runs under it (the side stream has nothing to wait for: the count comes from the host, and the
                # previous frame's readers of the gathered counts have been joined)
                s.extract_isosurface(self.iso)  # (returns once the count is on the host; the emit kernel is already running behind it)
                self._allgather_counts(s.mesh_device()[0] // 3)
            else:
                s.extract_isosurface(self.iso)

    def step_device(self):
        """inputs resident in HBM: (exchange) -> bin -> density -> (range all-reduce, normalise) -> MC -> (mesh gather)"""
        torch = self.torch
        if self.world == 1:
            self._compute(self.d_xyz.data_ptr(), self.n_local)
            self.surf.synchronize()
        else:
            # the library runs on torch's current stream (set_stream): everything below is stream-ordered, the host only waits where a
            # size decides an allocation (routing counts, count matrix, triangle count)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
            more, nrecv = self._more(self.d_xyz)
            ev[1].record()
            fused = self.gather == "fused"
            self._compute(self.d_xyz.data_ptr(), self.n_local, extract=not fused, more=more)
            ev[2].record()
            if fused:
                self._emit_to_root()
            elif self.gather == "nccl":
                self._gather_mesh()
            else:
                self._join_counts()
            ev[3].record()
            torch.cuda.current_stream().synchronize()
            self.last["exchange_ms"] = ev[0].elapsed_time(ev[1]) if self.exchange != "fused" else self._halo_events[0].elapsed_time(self._halo_events[1])
            self.last["gather_ms"] = ev[2].elapsed_time(ev[3])
            self.last["n_recv"] = nrecv
        self.last["n_in"] = self.n_local

    def step_e2e(self, mesh_to_host=True):
        """the same through HOST buffers: pinned H2D of this rank's chunk, D2H of its volume slab and of the mesh.
        mesh_to_host=False: the mesh stays in HBM (device-resident hand-off to a renderer); the particles still come from
        host memory and the volume still goes back."""
        torch = self.torch
        if self.world == 1:
            self._compute(self.h_xyz.data_ptr(), self.n_local, prefetch=True)
            self.surf.get_density(copy=False, with_rgb=self.protein)
            if mesh_to_host:
                self.surf.get_mesh(copy=False, colours=self.protein)
        elif not mesh_to_host:
            d = self.h_xyz.to(self.dev, non_blocking=True)
            more, _ = self._more(d)
            self._keep = [d]
            self._compute(d.data_ptr(), self.n_local, more=more, prefetch=True)
            self.surf.get_density(copy=False)
            self._join_counts()
            torch.cuda.current_stream().synchronize()
        else:
            d = self.h_xyz.to(self.dev, non_blocking=True)
            more, _ = self._more(d)
            self._keep = [d]
            fused = self.gather == "fused"
            self._compute(d.data_ptr(), self.n_local, extract=not fused, more=more, prefetch=True)
            self.surf.get_density(copy=False)
            if fused:
                self._emit_to_root()
            elif self.gather == "nccl":
                self._gather_mesh()
            else:
                self._join_counts()
                self.surf.get_mesh(copy=False)   # this rank's slab over this GPU's own PCIe link
            if self.rank == 0 and self.gather != "host":
                # D2H of the gathered mesh through a fixed 1 GiB pinned window (a consumer would map its own buffer; pinning
                # 8 x 8.4 GB on the host just for the measurement is not reasonable): the PCIe time is what is measured
                tot = self.last["gathered_verts"] * 3
                win = 256 * 1024 * 1024
                if getattr(self, "_hwin", None) is None:
                    self._hwin = torch.empty((win,), dtype=torch.float32, pin_memory=True)
                for src in (self._gpos, self._gnrm):
                    flat = src.view(-1)
                    for o in range(0, tot, win):
                        n = min(win, tot - o)
                        self._hwin[:n].copy_(flat[o:o + n], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def step_e2e_indexed(self):
        """single GPU, host buffers, the opt-in INDEXED mesh (one vertex per crossed edge + 3 x uint32 per triangle): pinned H2D of the
        particles, D2H of the volume and of the indexed mesh.  Returns the D2H bytes of the step."""
        assert self.world == 1 and not self.protein
        s = self.surf
        s.set_mesh_indexed(True)
        try:
            self._compute(self.h_xyz.data_ptr(), self.n_local, prefetch=True)
            s.get_density(copy=False)
            vpos, _, idx = s.get_mesh_indexed(copy=False)
        finally:
            s.set_mesh_indexed(False)
        self.last["indexed"] = (int(vpos.shape[0]), int(idx.shape[0]))
        return self.res[0] * self.res[1] * self.res[2] * 4 + vpos.shape[0] * 24 + idx.shape[0] * 12

    # ---- accounting -------------------------------------------------------------------------------------------
    def launches(self):
        return self.surf.launch_count()

    def stage_times(self):
        t = self.surf.timings()
        out = {k: round(v, 4) for k, v in t.items()}
        for k in ("exchange_ms", "gather_ms"):
            if k in self.last:
                out[k[:-3]] = round(self.last[k], 4)
        return out

    def local_tris(self):
        if self.world > 1 and self.gather in ("fused",):
            return int(self.last.get("tri_counts", [0] * self.world)[self.rank])
        n, _, _ = self.surf.mesh_device()
        return n // 3

    def totals(self):
        """(particles, voxels, triangles) of the whole job"""
        tris = self.local_tris()
        if self.world > 1:
            import torch.distributed as dist
            t = self.torch.tensor([tris], device=self.dev, dtype=self.torch.int64)
            dist.all_reduce(t)
            tris = int(t.item())
        self._tris_total = tris
        return self.n_total, self.res[0] * self.res[1] * self.res[2], tris

    def _local_alg_bytes(self):
        """Algorithmic HBM bytes of this rank's step: particles read (12 B xyz) + sorted records written and
        read (16 B each) + density written and read by MC (4 B each) + 72 B per triangle."""
        n = self.n_local
        v = self.res[0] * self.res[1] * self.me["nz"] if self.world > 1 else self.res[0] * self.res[1] * self.res[2]
        if self.protein:  # 32 B records in, 16 + 16 B sorted (xyzr + rgba), density + RGB3F volume, coloured mesh (108 B / triangle)
            return dict(bin=n * 32 + n * 32, density=n * 32 + v * 16, mc=v * 4 + self.local_tris() * 108, n=n, v=v)
        return dict(bin=n * 12 + n * 16, density=n * 16 + v * 4, mc=v * 4 + self.local_tris() * 72, n=n, v=v)

    def pipeline_bytes(self):
        b = self._local_alg_bytes()
        return (b["bin"] + b["density"] + b["mc"]) * self.world

    def roofline(self, stage, peak):
        """Roofline of the dominant KERNEL: algorithmic bytes of one launch / its CUDA-event time (library stream)."""
        b = self._local_alg_bytes()
        dens = "density_gauss_kernel" if self.protein else "density_splat_kernel"
        # mc_emit: reads the density once more, writes the mesh; mc_count reads the density once
        emit_bytes = b["mc"]
        cand = {dens: (stage["density"], b["density"]), "mc_emit_kernel": (stage.get("mc_emit", 0.0), emit_bytes)}
        if getattr(self.surf, "iso_mode", 0) == 1:
            cand["mt_emit_kernel"] = cand.pop("mc_emit_kernel")
        name = max(cand, key=lambda k: cand[k][0])
        ms, by = cand[name]
        ach = by / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        traffic = None
        try:
            import json, os
            with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")) as f:
                t = json.load(f)
            # (the committed captures are of the C2 frame per GPU -- the weak-scaled slabs run the same kernels on the same amount of
            #  work -- and of C3; the C2 frame uses the warp-tile splat kernel)
            if self.w.get("kind") == "lj" and not self.w.get("fixed_total") and abs(self.radius - 0.5) < 1e-6:
                traffic = t["c2"].get(name) or t["c2"].get(name.replace("density_splat_kernel", "density_splat3_kernel"))
            elif self.protein:
                traffic = t["c3"].get(name)
        except Exception:
            pass
        stages = {"bin (count+scan+scatter+order)": (stage["bin"], b["bin"]), dens: (stage["density"], b["density"]),
                  "marching cubes (count+scan+emit)": (stage["mc"], b["mc"] + b["v"] * 4)}
        per_stage = {k: (v[1] / (v[0] * 1e-3) / 1e9 / peak if v[0] > 0 else None) for k, v in stages.items()}
        if self.protein and name == dens:
            # The Gaussian density kernel is arithmetic-bound, not HBM-bound: per useful (voxel, atom) pair 3 sub + 3 FMA-class
            # for d^2, 1 compare, 1 mul + 1 ex2, 1 add, 3 FMA with colour = 13 FP32-pipe operations (19 flops, FMA = 2) and ONE SFU operation.
            # Peaks are nominal: 148 SMs x 128 FP32 lanes (x 2 flops per FMA) and 148 x 16 SFU lanes at the 1.965 GHz the bench runs at.
            pairs = getattr(self, "gauss_pairs", 0.0)
            fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
            tf = pairs * 19 / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            sfu = pairs / (ms * 1e-3) / (148 * 16 * 1.965e9) if ms > 0 else 0.0
            return {"bound": "fp32", "kernel": name, "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak, "traffic": traffic,
                    "peak_source": "nominal FP32 FMA peak (148 SMs x 128 lanes x 2 x 1.965 GHz); MEASURED_PEAKS.json holds no FP32 figure",
                    "algorithmic_flops": pairs * 19, "useful_pairs": pairs, "fp32_ops_per_pair": 13, "flops_per_pair": 19,
                    "issue_slot_frac": pairs * 13 / (ms * 1e-3) / (148 * 128 * 1.965e9) if ms > 0 else 0.0,
                    "sfu_frac": sfu, "hbm_frac": ach / peak, "algorithmic_bytes": by, "ms": ms, "per_stage_frac": per_stage}
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "algorithmic_bytes": by, "ms": ms, "per_stage_frac": per_stage}

    def h2d_bytes(self):
        return self.n_local * (32 if self.protein else 12) * self.world

    def d2h_bytes(self):
        v = self.res[0] * self.res[1] * self.res[2]
        if self.protein:
            return v * 16 + getattr(self, "_tris_total", 0) * 108
        return v * 4 + getattr(self, "_tris_total", 0) * 72

    def describe(self):
        how = {"fused": "marching-cubes kernels write straight into rank 0's mesh over NVLink (CUDA IPC): emit + gather fused",
               "nccl": "mesh gathered to rank 0 with NCCL send/recv",
               "host": "per-slab meshes stay in their GPU's HBM (counts all-gathered -> global offsets); e2e: every rank copies its slab "
                       "over its own PCIe link"}[self.gather]
        if self.w.get("fixed_total"):
            return (f"{self.w['name']} on {self.world} GPUs (strong scaling): {self.n_total} particles -> "
                    f"{self.res[0]}x{self.res[1]}x{self.res[2]}, z-slabs with halo, {self._exchange_how()}, {how}")
        return (f"{self.w['name']} weak-scaled x{self.world} along z: {self.n_total} particles -> "
                f"{self.res[0]}x{self.res[1]}x{self.res[2]}, z-slabs with halo, {self._exchange_how()}, {how}")

    def _exchange_how(self):
        return ("halo particles pushed straight into the neighbours' receive buffers over NVLink (fused kernel, CUDA IPC)" if self.exchange == "fused"
                else "particles exchanged with NCCL all-to-all-v")

    def close(self, destroy_group=True):
        R = self._root
        if self.world > 1:
            self.torch.cuda.synchronize()
            import torch.distributed as dist
            dist.barrier()
            for p in getattr(self, "_peer_open", []):  # halo receive buffers / counters of the other ranks
                self.surf.L.mms_ipc_close(self.local, p)
            self._peer_open = []
            if self.rank != 0:  # importers close their mappings first ...
                for k in ("pos", "nrm"):
                    if R[k]:
                        self.surf.L.mms_ipc_close(self.local, R[k])
                        R[k] = None
            dist.barrier()
            if self.rank == 0:  # ... then the exporter frees (current and retired buffers)
                for ptr in [R[k] for k in ("pos", "nrm") if R[k]] + R.get("retired", []):
                    self.surf.L.mms_device_free(self.local, ptr)
                R["pos"] = R["nrm"] = None
                R["retired"] = []
        self.surf.close()
        if self.world > 1 and destroy_group:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()

    def close_keep_group(self):
        self.close(destroy_group=False)


def _tensor_from_ptr(torch, ptr, nfloats, device):
    """Zero-copy float32 view of library-owned device memory (CUDA array interface)."""
    class _Wrap:
        pass
    w = _Wrap()
    w.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(w, device=device)
