#!/usr/bin/env python
"""Generates the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE translation units
(oracle/_ref/libmmref.so = datatools::ParticlesToDensity + trisoup_gl::volumetrics::IsoSurface behind
oracle/ref_harness.cpp) in the dev container, where /root/reference exists.  The fixtures travel to the GPU box.

  p2d_case_*.npz     inputs + the reference's density volume at ONE OpenMP thread (accumulation = particle order),
                     its Min/MaxValues metadata, origin and slice distances
  home_voxel_kat.npz the reference's home voxels, read off the volume: with sigma = 4 and r < sliceDist the support box
                     is exactly home +- 1 and every voxel of it is non-zero, so a lone particle paints a 3x3x3 block
                     whose centre IS static_cast<int>((p - min)/sliceDist)  (ParticlesToDensity.cpp:563-579);
                     positions include exact multiples of the slice distance and their fp32 neighbours
  isosurface_ref.npz the reference IsoSurface (marching tetrahedra) vertex count / bbox of its output on one volume
                     (informational for the marching-CUBES path: different algorithm)
  isosurface_mt_*.npz the reference IsoSurface's complete output (vertices + normals) on two volumes: the golden vectors of the
                     marching-tetrahedra compatibility mode (oracle mmo_mt_emit, libmmsurf MMS_ISO_MARCHING_TETS)

  p2d_vec_*.npz      aggregator 2 (IVecToSingleCell_Volume) at ONE OpenMP thread: inputs incl. DIRDATA_FLOAT_XYZ, the 3-component volume,
                     Min/MaxValues, and what the module serves on "outParticles" (grid positions, directions, colours, global radius)
                     and "outInfo" (7-column table, names, ranges)

Run:  python oracle/tools/gen_golden.py            (everything)
      python oracle/tools/gen_golden.py --only mt  (the marching-tetrahedra fixtures only)
      python oracle/tools/gen_golden.py --only vec (the aggregator-2 fixtures only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from megamol_b200 import synth  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def p2d_cases():
    rng = np.random.default_rng(20071017)
    cases = []
    # name, n, res, box(min, ext), radius spec, sigma, cyclic, normalize, aggregator, vertex layout
    cases.append(dict(name="noncyc_outside", n=1500, res=(16, 16, 16), bmin=(0, 0, 0), bext=(8, 8, 8), r=0.9, sigma=1.0,
                      cyc=(0, 0, 0), norm=0, agg=0, outside=0.06))
    cases.append(dict(name="cyclic_default", n=1500, res=(16, 16, 16), bmin=(-4, -4, -4), bext=(8, 8, 8), r=0.9, sigma=1.0,
                      cyc=(1, 1, 1), norm=1, agg=0, outside=0.0))
    cases.append(dict(name="aniso_mixedcyc", n=4000, res=(24, 20, 17), bmin=(1, 2, 3), bext=(12, 9, 7), r=0.45, sigma=1.0,
                      cyc=(1, 0, 1), norm=0, agg=0, outside=0.0))
    cases.append(dict(name="sigma_clipped", n=2000, res=(18, 18, 18), bmin=(0, 0, 0), bext=(9, 9, 9), r=0.5, sigma=2.5,
                      cyc=(1, 1, 1), norm=0, agg=0, outside=0.0))
    cases.append(dict(name="xyzr_intensity", n=2500, res=(20, 20, 20), bmin=(0, 0, 0), bext=(10, 10, 10), r=None, sigma=1.0,
                      cyc=(1, 1, 1), norm=1, agg=1, outside=0.0))
    cases.append(dict(name="small_radius", n=3000, res=(32, 32, 32), bmin=(0, 0, 0), bext=(16, 16, 16), r=0.3, sigma=0.8,
                      cyc=(0, 1, 0), norm=0, agg=0, outside=0.03))
    for i, c in enumerate(cases):
        n = c["n"]
        ext = np.array(c["bext"], np.float32)
        mn = np.array(c["bmin"], np.float32)
        u = np.stack([synth.uniform(1000 + i, 0, n, k) for k in range(3)], 1)
        xyz = (mn + (u * (1 + 2 * c["outside"]) - c["outside"]) * ext).astype(np.float32)
        if c["r"] is None:
            rad = (0.25 + 0.8 * synth.uniform(2000 + i, 0, n, 0)).astype(np.float32)
            rad[::37] = 0.0
            inten = (synth.uniform(3000 + i, 0, n, 0) * 2 - 0.5).astype(np.float32)
            data = np.concatenate([xyz, rad[:, None], inten[:, None]], 1).astype(np.float32)  # x y z r I, stride 20
            c["data"] = data
        else:
            c["data"] = xyz
        yield c


def as_lists(c, module):
    d = c["data"]
    if c["r"] is None:
        return [dict(vtx=d, vtx_type=module.VERT_FLOAT_XYZR, vtx_stride=20, count=len(d), col=d.ctypes.data + 16,
                     col_type=module.COL_FLOAT_I, col_stride=20)]
    return [dict(vtx=d, vtx_type=module.VERT_FLOAT_XYZ, count=len(d), global_radius=c["r"])]


def gen_mt_fixtures(h):
    """complete reference IsoSurface meshes on the volumes of two P2D cases"""
    cases = {c["name"]: c for c in p2d_cases()}
    for name, iso in (("sigma_clipped", 0.2), ("aniso_mixedcyc", 0.5)):
        c = cases[name]
        mn, ext = c["bmin"], c["bext"]
        bbox = (mn[0], mn[1], mn[2], mn[0] + ext[0], mn[1] + ext[1], mn[2] + ext[2])
        h.set_particles(as_lists(c, rb), bbox)
        h.set_p2d_params(c["res"], cyclic=c["cyc"], normalize=False, sigma=c["sigma"])
        vol, _ = h.pull_volume()
        m = h.pull_mesh(iso)
        np.savez_compressed(os.path.join(OUT, f"isosurface_mt_{name}.npz"), volume=vol, iso=np.float32(iso), bbox=np.array(bbox, np.float32),
                            pos=m["pos"], nrm=m["nrm"])
        print("IsoSurface (marching tetrahedra) golden:", name, "iso", iso, m["nverts"] // 3, "triangles")


def vec_cases():
    """aggregator 2: (name, lists, keep-alive data, bbox min, extent, res, cyclic, normalize, sigma)"""
    n = 700
    xyz = (synth.uniform_box(n, 10.0, seed=4101)).astype(np.float32)
    d = (np.stack([synth.uniform(4102, 0, n, k) for k in range(3)], 1) * 2 - 1).astype(np.float32)
    yield dict(name="cyclic_tight", data=xyz, dirs=d, layout="xyz+dir", radius=0.6, bmin=(0, 0, 0), bext=(10, 10, 10), res=(40, 18, 16),
               cyc=(1, 1, 1), norm=1, sigma=1.0)
    n = 900
    buf = np.zeros((n, 7), np.float32)  # x y z r dx dy dz, stride 28
    u = np.stack([synth.uniform(4201, 0, n, k) for k in range(3)], 1)
    buf[:, :3] = np.array([1, 2, 3], np.float32) + (u * 1.1 - 0.05) * np.array([12, 9, 7], np.float32)  # some particles outside the box
    buf[:, 3] = 0.25 + 0.5 * synth.uniform(4202, 0, n, 0)
    buf[::29, 3] = 0.0
    buf[:, 4:7] = np.stack([synth.uniform(4203, 0, n, k) for k in range(3)], 1) * 3 - 1.5
    buf[::17, 4:7] = 0.0
    yield dict(name="noncyc_interleaved", data=buf, dirs=None, layout="xyzr|dir", radius=-1.0, bmin=(1, 2, 3), bext=(12, 9, 7), res=(40, 20, 18),
               cyc=(0, 0, 0), norm=0, sigma=0.9)


def vec_lists(c, module):
    d = c["data"]
    if c["layout"] == "xyz+dir":
        return [dict(vtx=d, vtx_type=module.VERT_FLOAT_XYZ, count=len(d), global_radius=c["radius"], dir=c["dirs"])]
    return [dict(vtx=d, vtx_type=module.VERT_FLOAT_XYZR, vtx_stride=28, count=len(d), dir=d.ctypes.data + 16, dir_stride=28)]


def gen_vec_fixtures(h):
    for c in vec_cases():
        mn, ext = c["bmin"], c["bext"]
        bbox = (mn[0], mn[1], mn[2], mn[0] + ext[0], mn[1] + ext[1], mn[2] + ext[2])
        h.set_particles(vec_lists(c, rb), bbox)
        h.set_p2d_params(c["res"], cyclic=c["cyc"], normalize=bool(c["norm"]), sigma=c["sigma"], aggregator=2)
        vol, meta = h.pull_volume(components=3)
        g = h.pull_grid_particles()
        t = h.pull_info()
        assert g["lists"] == 1 and g["vtx_type"] == 1 and g["col_type"] == 5 and g["dir_type"] == 1 and t["columns"] == 7
        np.savez_compressed(os.path.join(OUT, f"p2d_vec_{c['name']}.npz"), data=c["data"], dirs=c["dirs"] if c["dirs"] is not None else np.zeros(0, np.float32),
                            layout=c["layout"], radius=np.float32(c["radius"]), res=np.array(c["res"]), bmin=np.array(mn, np.float32),
                            bext=np.array(ext, np.float32), sigma=np.float32(c["sigma"]), cyclic=np.array(c["cyc"]), normalize=c["norm"],
                            volume=vol, minmax=np.array([meta["min"], meta["max"]]), grid_pos=g["pos"], grid_dir=g["dir"], grid_col=g["col"],
                            grid_radius=np.float32(g["global_radius"]), info=t["data"], info_names=np.array(t["names"]), info_ranges=t["ranges"])
        print("aggregator 2 golden:", c["name"], vol.shape, "min/max", meta["min"], meta["max"], g["count"], "grid particles")


def main():
    if "--only" in sys.argv:
        h = rb.Harness()
        h.set_threads(1)
        {"mt": gen_mt_fixtures, "vec": gen_vec_fixtures}[sys.argv[sys.argv.index("--only") + 1]](h)
        return
    os.makedirs(OUT, exist_ok=True)
    h = rb.Harness()
    h.set_threads(1)
    for c in p2d_cases():
        mn, ext = c["bmin"], c["bext"]
        bbox = (mn[0], mn[1], mn[2], mn[0] + ext[0], mn[1] + ext[1], mn[2] + ext[2])
        h.set_particles(as_lists(c, rb), bbox)
        h.set_p2d_params(c["res"], cyclic=c["cyc"], normalize=bool(c["norm"]), sigma=c["sigma"], aggregator=c["agg"])
        vol, meta = h.pull_volume()
        np.savez_compressed(os.path.join(OUT, f"p2d_case_{c['name']}.npz"), data=c["data"], res=np.array(c["res"]), bmin=np.array(mn, np.float32),
                            bext=np.array(ext, np.float32), radius=np.float32(-1 if c["r"] is None else c["r"]), sigma=np.float32(c["sigma"]),
                            cyclic=np.array(c["cyc"]), normalize=c["norm"], aggregator=c["agg"], volume=vol,
                            minmax=np.array([meta["min"], meta["max"]]), origin=np.array(meta["origin"], np.float32),
                            slicedist=np.array(meta["slicedist"], np.float32))
        print(c["name"], vol.shape, "sum", float(vol.sum(dtype=np.float64)), "min/max", meta["min"], meta["max"])

    # ---- home-voxel known-answer test ------------------------------------------------------------------------
    res = (64, 48, 40)
    bmin = np.array([-3.0, 0.5, 10.0], np.float32)
    bext = np.array([12.6, 9.4, 7.8], np.float32)
    # the reference derives the extent from the Cuboid it is given: Width() = Right - Left in fp32
    # (vislib Cuboid; ParticlesToDensity.cpp:423-425), so the effective extent is fl(fl(min+ext) - min)
    bext = ((bmin + bext).astype(np.float32) - bmin).astype(np.float32)
    sd = (bext / (np.array(res, np.float32) - np.float32(1))).astype(np.float32)
    # lattice of well separated particles (5 voxels apart) so that the 3x3x3 blocks never overlap; each gets a fractional
    # offset; a third of them sits EXACTLY on k*sd+min (as computed in fp32) or one ulp beside it
    pts = []
    rng = np.random.default_rng(7)
    for kz in range(2, res[2] - 2, 5):
        for ky in range(2, res[1] - 2, 5):
            for kx in range(2, res[0] - 2, 5):
                k = np.array([kx, ky, kz], np.float32)
                mode = rng.integers(0, 3)
                if mode == 0:
                    p = bmin + (k + rng.random(3).astype(np.float32) * np.float32(0.98) + np.float32(0.01)) * sd
                else:
                    p = (k * sd + bmin).astype(np.float32)
                    if mode == 2:
                        p = np.nextafter(p, np.float32(np.inf) * np.where(rng.random(3) < 0.5, -1, 1).astype(np.float32)).astype(np.float32)
                pts.append(p.astype(np.float32))
    pts = np.array(pts, np.float32)
    r = float(0.9 * sd.min())
    bbox = (bmin[0], bmin[1], bmin[2], bmin[0] + bext[0], bmin[1] + bext[1], bmin[2] + bext[2])
    h.set_particles([dict(vtx=pts, vtx_type=rb.VERT_FLOAT_XYZ, count=len(pts), global_radius=r)], bbox)
    h.set_p2d_params(res, cyclic=(0, 0, 0), normalize=False, sigma=4.0)
    vol, meta = h.pull_volume()
    nz = vol > 0
    # a voxel is a block centre iff all 27 neighbours are non-zero
    core = np.ones_like(nz)
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                core &= np.roll(nz, (dz, dy, dx), (0, 1, 2))
    centres = np.argwhere(core)[:, ::-1]  # (x, y, z)
    assert len(centres) == len(pts), (len(centres), len(pts))
    assert nz.sum() == 27 * len(pts)
    # match centres to particles: the home voxel is within one voxel of (p-min)/sd
    approx = np.floor((pts - bmin) / sd + 0.5).astype(int)
    home = np.zeros((len(pts), 3), np.int32)
    cs = {tuple(c) for c in centres}
    for i, a in enumerate(approx):
        cand = [(a[0] + dx, a[1] + dy, a[2] + dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1) if (a[0] + dx, a[1] + dy, a[2] + dz) in cs]
        assert len(cand) == 1, (i, cand)
        home[i] = cand[0]
    np.savez_compressed(os.path.join(OUT, "home_voxel_kat.npz"), points=pts, res=np.array(res), bmin=bmin, bext=bext, radius=np.float32(r),
                        home=home)
    print("home-voxel KAT:", len(pts), "particles;", int((home != np.trunc((pts - bmin) / sd)).any(1).sum()), "differ from a float64 guess")

    # ---- reference IsoSurface (marching tetrahedra), informational ----------------------------------------------
    c = next(iter(p2d_cases()))
    mn, ext = c["bmin"], c["bext"]
    bbox = (mn[0], mn[1], mn[2], mn[0] + ext[0], mn[1] + ext[1], mn[2] + ext[2])
    h.set_particles(as_lists(c, rb), bbox)
    h.set_p2d_params(c["res"], cyclic=c["cyc"], normalize=False, sigma=c["sigma"])
    vol, meta = h.pull_volume()
    m = h.pull_mesh(0.5)
    np.savez_compressed(os.path.join(OUT, "isosurface_ref.npz"), volume=vol, iso=np.float32(0.5), nverts=m["nverts"], ntris=m["ntris"],
                        pos_min=m["pos"].min(0), pos_max=m["pos"].max(0), bbox=np.array(bbox, np.float32))
    print("IsoSurface reference:", m["nverts"], "vertices, GetTriCount() =", m["ntris"])
    gen_mt_fixtures(h)
    gen_vec_fixtures(h)


if __name__ == "__main__":
    main()
