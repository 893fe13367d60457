#!/usr/bin/env python
"""bench.py -- particles -> density volume -> isosurface throughput on N B200s (one process per GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): 10 M Lennard-Jones-fluid-like particles (jittered simple
cubic lattice, rho* = 0.795, FLOAT_XYZ + global radius 0.5) -> 512^3 ParticlesToDensity volume (bump kernel,
sigma 1, cyclic, normalize on) + marching-cubes isosurface at isoval 0.5 (the modules' default parameters).
For N > 1 the workload is weak-scaled: every rank owns a z-slab of 512 planes worth of particles of the
C4-style lattice (N = 8 -> 80 M particles, 512 x 512 x 4096 ... see slabs.py); `value` is the whole job.

One JSON line on stdout (rank 0).  `value` = device-resident throughput (inputs in HBM when the clock starts),
`e2e` = the same step through the C ABI with HOST buffers (pinned H2D of the particles, D2H of volume + mesh inside
the timed region).  `--impl reference` times the UNMODIFIED reference modules (oracle/_ref/libmmref.so: the
reference's ParticlesToDensity + IsoSurface translation units) on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from megamol_b200 import synth  # noqa: E402

METRIC = "Mparticles/s & Gvoxels/s density+isosurface"
ISO = 0.5
RADIUS = 0.5


def workload(name: str):
    if name == "c1":
        return dict(name="C1: 1M uniform-random spheres r=0.5 -> 128^3 P2D bump + MC", n=1_000_000, res=(128, 128, 128), kind="uniform", box=64.0)
    if name == "c2":
        return dict(name="C2: 10M LJ-fluid-like (jittered lattice, r=0.5, cyclic, normalize) -> 512^3 P2D bump + MC iso 0.5",
                    n=10_000_000, res=(512, 512, 512), kind="lj")
    if name == "c3":
        return dict(name="C3: 1M protein-like atoms (FLOAT_XYZR + FLOAT_RGBA, stride 32) -> 512^3 QuickSurf-Gaussian density + RGB volume "
                         "+ coloured MC surface (radscale 1, quality 2, iso 0.5)", n=1_000_000, res=(512, 512, 512), kind="protein")
    if name == "c4":
        return dict(name="C4: 100M LJ-fluid-like particles (465^3 lattice, r=0.5, cyclic, normalize) -> 1024^3 P2D bump + MC iso 0.5",
                    n=100_000_000, res=(1024, 1024, 1024), kind="lj", lattice=465, fixed_total=True)
    raise SystemExit(f"unknown workload {name}")


def make_particles(w, i0=0, i1=None):
    if w["kind"] == "uniform":
        xyz = synth.uniform_box(w["n"], w["box"], i0=i0, i1=i1)
        return xyz, w["box"]
    xyz, L = synth.lj_fluid(w["n"], i0=i0, i1=i1)
    return xyz, L


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for s in self.samples if len(s) >= 6 for k in range(4) if s[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_sample(threads=None, reps=1):
    """The reference's own CPU path on a bounded sample of the C2 workload: 54^3 = 157 464 lattice sites -> 128^3
    (same lattice spacing and voxel size as C2: 1/64 of the volume), ParticlesToDensity (OpenMP) + IsoSurface (serial),
    through the reference's call/slot API.  Returns dict(value Mparticles/s, gvoxels/s, ...)."""
    from oracle import ref_binding as rb
    n = 54 ** 3
    xyz, L = synth.lj_fluid(n)
    res = (128, 128, 128)
    # same voxel size as C2: box = 127 * (L_c2 / 511)
    if rb.available():
        h = rb.Harness()
        # torchrun exports OMP_NUM_THREADS=1 to its ranks: the reference arm always gets every host core it may use
        h.set_threads(threads or len(os.sched_getaffinity(0)) or os.cpu_count() or 1)
        cores = h.threads
        kind = "reference"
        best = None
        for _ in range(reps):
            h.set_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=RADIUS)], (0, 0, 0, L, L, L))
            h.set_p2d_params(res, cyclic=(True, True, True), normalize=True, sigma=1.0)
            t0 = time.perf_counter()
            _, meta = h.pull_volume(copy=False)
            m = h.pull_mesh(ISO, copy=False)
            dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, meta["ms"], m["ms"], m["nverts"] // 3)
        dt, ms_d, ms_i, ntri = best
        sample = f"{n} particles (54^3 sites of the C2 lattice) -> 128^3, P2D {ms_d:.0f} ms ({cores} OpenMP threads) + IsoSurface {ms_i:.0f} ms (serial, marching tets, {ntri} triangles)"
    else:
        from oracle import oracle_binding as ob
        o = ob.Oracle()
        cores = os.cpu_count() or 1
        kind = "port"
        t0 = time.perf_counter()
        vol, _ = o.density_p2d([dict(vtx=xyz, vtx_type=1, count=n, global_radius=RADIUS)], (0, 0, 0), (L, L, L), res, (1, 1, 1), normalize=True)
        ntri, _, _ = o.mc_count(vol, ISO)
        dt = time.perf_counter() - t0
        sample = f"{n} particles -> 128^3 with the oracle port (density OpenMP + MC classify), {ntri} triangles"
    return {"value": n / dt / 1e6, "unit": "Mparticles/s", "gvoxels_per_s": res[0] * res[1] * res[2] / dt / 1e9, "cores": cores,
            "kind": kind, "sample": sample, "seconds": dt}


def cpu_port_full():
    """The oracle's restatement of the path (not the reference's code: z-partitioned OpenMP density over ONE shared volume, marching-CUBES
    classification) on the FULL C2 workload, all host cores: what a tuned CPU implementation of the same algorithm does.  The vertex
    emission is left out (8.4 GB of host arrays), which flatters the CPU.  Informational, next to the reference-module baseline."""
    from oracle import oracle_binding as ob
    o = ob.Oracle()
    n = 10_000_000
    xyz, L = synth.lj_fluid(n)
    res = (512, 512, 512)
    t0 = time.perf_counter()
    vol, _ = o.density_p2d([dict(vtx=xyz, vtx_type=1, count=n, global_radius=RADIUS)], (0, 0, 0), (L, L, L), res, (1, 1, 1), normalize=True)
    t1 = time.perf_counter()
    ntri, _, _ = o.mc_count(vol, ISO)
    t2 = time.perf_counter()
    return {"value": n / (t2 - t0) / 1e6, "unit": "Mparticles/s", "gvoxels_per_s": 512 ** 3 / (t2 - t0) / 1e9, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"full C2 with the oracle port: density {1e3 * (t1 - t0):.0f} ms (OpenMP, z-partitioned) + marching-cubes classification "
                      f"{1e3 * (t2 - t1):.0f} ms (serial), {ntri} triangles counted, no vertex emission"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample()
    for _ in range(args.steps):
        vals.append(cpu_reference_sample())
    total_s = sum(v["seconds"] for v in vals)
    n = 54 ** 3
    value = n * len(vals) / total_s / 1e6
    cb = dict(vals[-1])
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mparticles/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_s / len(vals) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gvoxels_per_s": 128 ** 3 * len(vals) / total_s / 1e9,
            "config": {"workload": w["name"], "sample_per_step": cb["sample"]},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": "Mparticles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def pcie_probe(job, nbytes=1 << 30, reps=3):
    """Pinned-memory copy bandwidth of THIS box, all ranks at the same time (the e2e arm's roof): host -> device and device -> host,
    best of `reps`, one 1 GiB buffer per rank.  Returns GB/s per direction (this rank) -- the caller aggregates."""
    import torch
    n = nbytes // 4
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=job.dev)
    out = {}
    for name, dst, src in (("h2d", d, h), ("d2h", h, d)):
        best = 0.0
        for _ in range(reps):
            job.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        out[name] = best
    del h, d
    return out


def modules_e2e(w, steps=3):
    """The same end-to-end step through the DROP-IN MODULES (plugin/b200surf compiled against the reference's module/call runtime,
    oracle/_ref/libmmplug.so): a MultiParticleDataCall source -> ParticlesToDensityB200 -> IsoSurfaceB200 -> VolumetricDataCall +
    CallTriMeshData pulled like a renderer would (host pointers).  Adds the call walk, metadata and list handling of the modules to
    what `e2e` measures through the bare C ABI.  None where the harness library was not built."""
    from oracle import ref_binding as rb
    if not rb.available(rb.PLUG_LIB):
        return None
    xyz, L = synth.lj_fluid(w["n"])
    h = rb.Harness(rb.PLUG_LIB)
    lists = [dict(vtx=xyz, vtx_type=1, count=len(xyz), global_radius=RADIUS)]
    times = []
    for it in range(steps + 2):
        h.set_particles(lists, (0, 0, 0, L, L, L))   # new data hash: both modules recompute
        h.set_p2d_params(w["res"], cyclic=(True, True, True), normalize=True, sigma=1.0)
        t0 = time.perf_counter()
        h.pull_volume(copy=False)
        m = h.pull_mesh(ISO, copy=False)
        times.append(time.perf_counter() - t0)
    dt = float(np.median(times[2:]))
    h.close()
    return {"value": w["n"] / dt / 1e6, "unit": "Mparticles/s", "ms_per_step": dt * 1e3, "steps": steps, "triangles": m["nverts"] // 3,
            "note": "ParticlesToDensityB200 + IsoSurfaceB200 driven through the reference's Call/Slot runtime (libmmplug.so): host particle "
                    "array in, host volume + host mesh pointers out"}


def bind_to_gpu_numa_node(local):
    """Pin this rank's process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host memory is allocated: the pinned
    result buffers then live on the GPU's own NUMA node and N ranks do not funnel their D2H traffic into one socket's memory.
    Best effort (no NVML / no topology information -> no-op).  Returns a short description for the JSON line."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        masks = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (masks[i // 64] >> (i % 64)) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed and len(allowed) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, allowed)
            return f"cpus {allowed[0]}-{allowed[-1]} ({len(allowed)})"
        return "all cpus (no narrower GPU affinity reported)"
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def run_ours(args, w):
    import torch
    import megamol_b200 as mm
    from megamol_b200 import slabs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    job = slabs.SlabJob(w, rank, world, local, iso=ISO, radius=RADIUS, gather=args.gather, exchange=args.exchange)
    if args.algorithm == "mt":  # the reference IsoSurface's marching tetrahedra, bit for bit (compatibility mode; flat normals, ~2.7x the triangles)
        job.surf.set_isosurface_mode(mm.api.ISO_MARCHING_TETS)
    peak, peak_src = load_peaks()

    # ---- device-resident arm -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        job.step_device()
    job.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = job.launches()
    t_dev = job.timed(job.step_device, args.steps)  # max over ranks, ms for K steps
    launches = (job.launches() - launches0)
    stage = job.stage_times()
    # ---- end-to-end arm --------------------------------------------------------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    if args.no_e2e:
        t_e2e = float("nan")
    else:
        for _ in range(min(args.warmup, 3)):
            job.step_e2e()
        job.barrier()
        t_e2e = job.timed(job.step_e2e, e2e_steps)
    # the same with the mesh left in HBM (device-resident hand-off): host particles in, volume out
    t_e2e_dm = job.timed(lambda: job.step_e2e(mesh_to_host=False), e2e_steps) if args.gather == "host" else float("nan")
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=3)
    n_total, v_total, t_total = job.totals()
    # everything the record needs from the job (it may be replaced by the C4 job below)
    info = dict(h2d=job.h2d_bytes(), d2h=job.d2h_bytes(), roofline=job.roofline(stage, peak), describe=job.describe(), protein=job.protein,
                pipeline_bytes=job.pipeline_bytes())
    # single GPU, marching cubes: the opt-in indexed mesh through host buffers (about 28 instead of 72 bytes per triangle over PCIe)
    e2e_ix = None
    if world == 1 and not job.protein and args.algorithm == "mc" and not args.no_e2e:
        job.step_e2e_indexed()
        t_ix = job.timed(job.step_e2e_indexed, e2e_steps)
        nv_ix, nt_ix = job.last["indexed"]
        tm_ix = job.surf.timings()
        e2e_ix = {"value": job.n_total / (t_ix / e2e_steps * 1e-3) / 1e6, "unit": "Mparticles/s", "ms_per_step": t_ix / e2e_steps, "steps": e2e_steps,
                  "vertices": nv_ix, "triangles": nt_ix, "mc_ms": tm_ix.get("mc"), "mc_emit_ms": tm_ix.get("mc_emit"), "h2d_bytes_per_step": job.h2d_bytes(),
                  "d2h_bytes_per_step": job.res[0] * job.res[1] * job.res[2] * 4 + nv_ix * 24 + nt_ix * 12,
                  "note": "opt-in indexed mesh (mms_set_mesh_indexed / IsoSurfaceB200 'indexedMesh'): host particles in, host volume + "
                          "vertices (pos, nrm) + 32-bit indices out; the default contract stays the reference's triangle soup (e2e)"}
    # the e2e arm's own roof: pinned copies over PCIe, all ranks at once
    pc = pcie_probe(job) if not args.no_e2e else None
    if pc is not None and world > 1:
        import torch.distributed as dist
        t = torch.tensor([pc["h2d"], pc["d2h"]], device=job.dev, dtype=torch.float64)
        dist.all_reduce(t)
        pc = {"h2d": float(t[0].item()), "d2h": float(t[1].item())}
    # N > 1: the configuration the north-star target is quoted on (C4, strong scaling) inside the same record
    c4 = None
    if world > 1 and args.workload == "c2" and not args.no_c4:
        job.close_keep_group()
        job = None
        c4 = run_c4(args, rank, world, local)
    if rank != 0:
        if job is not None:
            job.close()
        else:
            import torch.distributed as dist
            dist.barrier()
            dist.destroy_process_group()
        return
    ms = t_dev / args.steps
    value = n_total / (ms * 1e-3) / 1e6
    ms_e = t_e2e / e2e_steps
    e2e_val = n_total / (ms_e * 1e-3) / 1e6
    e2e = {"value": e2e_val, "unit": "Mparticles/s", "ms_per_step": ms_e, "steps": e2e_steps,
           "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"]}
    if t_e2e_dm == t_e2e_dm:
        e2e_dm = {"value": n_total / (t_e2e_dm / e2e_steps * 1e-3) / 1e6, "unit": "Mparticles/s", "ms_per_step": t_e2e_dm / e2e_steps,
                  "note": "host particles in (pinned H2D), volume back to the host (D2H), mesh stays in HBM for a device-resident consumer",
                  "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": v_total * (16 if info["protein"] else 4)}
    else:
        e2e_dm = None
    if pc is not None and not args.no_e2e:
        # lower bound of the e2e step if it were nothing but its PCIe transfers at the measured copy bandwidth
        floor_ms = (info["h2d"] / (pc["h2d"] * 1e9) + info["d2h"] / (pc["d2h"] * 1e9)) * 1e3
        e2e["pcie"] = {"h2d_gbs": pc["h2d"], "d2h_gbs": pc["d2h"], "transfer_floor_ms": floor_ms, "frac": floor_ms / ms_e,
                       "how": "1 GiB pinned copies, all ranks at the same time, best of 3 (aggregate GB/s); frac = transfer floor / e2e step"}
    if args.no_e2e:
        e2e = {"value": None, "unit": "Mparticles/s", "skipped": "--no-e2e", "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"]}
    # roofline of the dominant kernel (largest share of the device step), algorithmic bytes per DESIGN.md
    rl = info["roofline"]
    rl.setdefault("peak_source", peak_src)
    line = {"metric": METRIC, "value": value, "unit": "Mparticles/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if w.get("fixed_total") else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "gvoxels_per_s": v_total / (ms * 1e-3) / 1e9,
            "config": {"workload": w["name"] if world == 1 else info["describe"], "particles": n_total, "voxels": v_total, "triangles": t_total,
                       "l2": "inputs (particles + volume + mesh) are larger than the 126 MB L2; no explicit flush",
                       "parallelism": f"z-slabs x{world}", "host_affinity_rank0": numa,
                       "isosurface": "marching cubes (default)" if args.algorithm == "mc" else "marching tetrahedra, reference-compatible mode"},
            "stages_ms": stage, "roofline": rl,
            "pipeline_hbm_frac": info["pipeline_bytes"] / (ms * 1e-3) / 1e9 / peak,
            "e2e": e2e, "e2e_mesh_on_device": e2e_dm, "e2e_indexed": e2e_ix, "gpu_launches": launches, "clocks": sampler.summary()}
    if not args.no_cpu:
        cb = cpu_reference_sample()
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_baseline"]["gvoxels_per_s"] = cb["gvoxels_per_s"]
        if world == 1 and args.workload == "c2":
            line["cpu_port"] = cpu_port_full()
    if c4 is not None:
        line["c4"] = c4
    if world == 1 and args.workload == "c2" and not args.no_e2e:
        job.close()
        job = None
        me = modules_e2e(w)
        if me is not None:
            line["e2e_modules"] = me
    print(json.dumps(line))
    if job is not None:
        job.close()
    elif world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_c4(args, rank, world, local):
    """BASELINE configs[3] / the north-star target: 100 M particles -> 1024^3, z-slabs over the ranks of this run (strong scaling)."""
    from megamol_b200 import slabs
    w = workload("c4")
    job = slabs.SlabJob(w, rank, world, local, iso=ISO, radius=RADIUS, gather=args.gather, exchange=args.exchange)
    for _ in range(2):
        job.step_device()
    t_dev = job.timed(job.step_device, 3) / 3
    stage = job.stage_times()
    job.step_e2e(mesh_to_host=False)
    t_dm = job.timed(lambda: job.step_e2e(mesh_to_host=False), 2) / 2
    t_full = None
    # the C4 mesh is 68 GB of host arrays: with the mesh on the host only where every rank's share (8.5 GB pinned) is moderate
    if not args.no_e2e and world >= 8 and os.environ.get("MMS_BENCH_C4_FULL_E2E", "1") != "0":
        job.step_e2e()
        t_full = job.timed(job.step_e2e, 2) / 2
    n_total, v_total, t_total = job.totals()
    out = {"workload": job.describe(), "particles": n_total, "voxels": v_total, "triangles": t_total, "ms_per_step": t_dev,
           "value": n_total / (t_dev * 1e-3) / 1e6, "unit": "Mparticles/s", "gvoxels_per_s": v_total / (t_dev * 1e-3) / 1e9, "stages_ms": stage,
           "e2e_mesh_on_device": {"ms_per_step": t_dm, "value": n_total / (t_dm * 1e-3) / 1e6,
                                  "note": "host particles in, 4.3 GB volume back to the host, mesh stays sharded in HBM"},
           "e2e": None if t_full is None else {"ms_per_step": t_full, "value": n_total / (t_full * 1e-3) / 1e6,
                                               "h2d_bytes_per_step": job.h2d_bytes(), "d2h_bytes_per_step": job.d2h_bytes()},
           "target": "north star: under 100 ms end to end on 8 GPUs"}
    job.close_keep_group()
    return out


def run_c5(args):
    """BASELINE config 5: MMPLD time series of 10 M particles per frame streamed through density + isosurface.
    The file holds `--frames` frames (default 12 = 1.4 GB; the configuration names 100) written to --tmpdir first."""
    import megamol_b200 as mm
    from megamol_b200 import mmpld, stream
    n, res = 10_000_000, (512, 512, 512)
    F = args.frames
    base, L = synth.lj_fluid(n)
    path = os.path.join(args.tmpdir, f"mmsurf_c5_{F}.mmpld")
    t0 = time.perf_counter()
    frames = []
    for f in range(F):   # seeded per-frame displacement (thermal jiggle), wrapped into the periodic box
        disp = np.stack([synth.uniform(synth.SEED + 50 + f, 0, n, k) for k in range(3)], 1)
        xyz = np.mod(base + (disp - np.float32(0.5)) * np.float32(0.2), np.float32(L)).astype(np.float32)
        frames.append((float(f), [dict(vtype=1, ctype=0, data=xyz, global_radius=RADIUS)]))
    mmpld.write_mmpld(path, frames, (0, 0, 0, L, L, L))
    del frames
    t_write = time.perf_counter() - t0
    rd = mmpld.Reader(path)
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), (L, L, L), res, (True, True, True))
    s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    if args.indexed:
        s.set_mesh_indexed(True)
    stream.stream_frames(s, rd, min(3, F), ISO, indexed=args.indexed)            # warm-up: allocations, page cache
    sampler = ClockSampler(0)
    sampler.start()
    l0 = s.launch_count()
    t0 = time.perf_counter()
    lat = stream.stream_frames(s, rd, F, ISO, indexed=args.indexed)
    s.synchronize()
    dt = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=3)
    if args.indexed:
        nvert, ntri = s.mesh_indexed_device()[:2]
        mesh_bytes = nvert * 24 + ntri * 12
    else:
        ntri = s.mesh_device()[0] // 3
        mesh_bytes = ntri * 72
    st = s.timings()
    line = {"metric": METRIC, "value": n * F / dt / 1e6, "unit": "Mparticles/s", "n_gpus": 1, "steps": F, "warmup": min(3, F),
            "ms_per_step": dt / F * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "frames_per_s": F / dt, "gvoxels_per_s": res[0] * res[1] * res[2] * F / dt / 1e9,
            "config": {"workload": f"C5: {F}-frame MMPLD v1.3 time series of 10M LJ-fluid-like particles -> 512^3 P2D bump + MC iso 0.5 per frame, "
                                   "streamed: pinned double-buffered reader, H2D of frame k+1 overlaps kernels + read-back of frame k",
                       "frames": F, "file_bytes": os.path.getsize(path), "file_write_s": t_write, "triangles_last_frame": ntri,
                       "latency_ms": {"median": float(np.median(lat)), "max": float(np.max(lat))}},
            "stages_ms": {k: round(v, 4) for k, v in st.items()},
            "e2e": {"value": n * F / dt / 1e6, "unit": "Mparticles/s", "h2d_bytes_per_step": n * 12, "d2h_bytes_per_step": res[0] * res[1] * res[2] * 4 + mesh_bytes},
            "mesh": "indexed (opt-in: one vertex per crossed grid edge + 32-bit indices)" if args.indexed else "triangle soup (the reference's contract)",
            "gpu_launches": s.launch_count() - l0, "clocks": sampler.summary()}
    print(json.dumps(line))
    s.close()
    rd.close()
    os.remove(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end arm (host buffers): the C4 mesh is 84 GB of pinned host memory")
    ap.add_argument("--gather", default="host", choices=["host", "fused", "nccl"],
                    help="multi-GPU: where the per-slab meshes go: host (default; stay sharded in HBM, counts all-gathered, e2e copies each slab over its "
                         "own PCIe link), fused (mc_emit stores into rank 0's mesh over NVLink), nccl (send/recv to rank 0)")
    ap.add_argument("--algorithm", default="mc", choices=["mc", "mt"],
                    help="isosurface triangulation: mc = marching cubes (default, the north-star path), mt = the reference module's marching tetrahedra")
    ap.add_argument("--frames", type=int, default=12, help="frames of the C5 time series")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="multi-GPU halo exchange: fused (default; one push kernel writes into the neighbours' receive buffers over NVLink/CUDA IPC, "
                         "no host synchronisation) or nccl (round-1 baseline: routing kernels + count matrix + all-to-all-v)")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the C4 (100 M particles -> 1024^3, strong scaling) figures in the record")
    ap.add_argument("--tmpdir", default="/tmp")
    ap.add_argument("--indexed", action="store_true", help="C5: read the opt-in indexed mesh back instead of the triangle soup")
    ap.add_argument("--radius", type=float, default=None, help="particle radius of the LJ workloads (default 0.5; SURVEY's C2 variant: 1.0, a 5^3 support)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "c5" and args.impl == "ours":
        return run_c5(args)
    w = workload("c2" if args.workload == "c5" else args.workload)
    if args.radius is not None:
        global RADIUS
        RADIUS = float(args.radius)
        w = dict(w, name=w["name"] + f" [radius {RADIUS:g}]")
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
