"""GPU parity: libmmsurf (through the C ABI) against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

from megamol_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def run_density(surf, lists, bmin, bext, res, cyclic, sigma=1.0, aggregator=0, normalize=False, home=False):
    surf.clear_particles()
    surf.set_grid(bmin, bext, res, cyclic)
    surf.set_params(mode=0, aggregator=aggregator, normalize=int(normalize), defer_normalize=0, sigma=sigma,
                    want_home_voxels=int(home), want_cell_tricounts=1)
    surf.push_particles(lists)
    surf.compute_density()
    return surf.get_density()


CASES = [
    # n, res, box, radius, sigma, cyclic, outside fraction
    (2000, (16, 16, 16), 8.0, 0.9, 1.0, (False, False, False), 0.05),
    (2000, (16, 16, 16), 8.0, 0.9, 1.0, (True, True, True), 0.0),
    (5000, (24, 20, 16), 8.0, 0.45, 1.0, (True, False, True), 0.0),
    (3000, (16, 16, 16), 8.0, 0.5, 2.5, (True, True, True), 0.0),
    (3000, (16, 16, 16), 8.0, 0.5, 2.5, (False, False, False), 0.1),
    (20000, (40, 37, 33), 20.0, 0.6, 1.0, (True, True, True), 0.0),
    (20000, (64, 48, 50), 20.0, 1.3, 1.0, (False, True, False), 0.02),
    (50000, (96, 96, 96), 48.0, 0.5, 1.0, (True, True, True), 0.0),
    # periodic, a quarter of the particles outside the box (binned through the wrap), on the warp-tile splat kernel (supports <= 3^3 voxels)
    (30000, (64, 48, 40), 24.0, 0.45, 1.0, (True, True, True), 0.15),
    (30000, (64, 48, 40), 24.0, 0.45, 1.0, (True, False, True), 0.15),
]


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_home_voxels_and_density(surf, oracle, case):
    n, res, box, radius, sigma, cyclic, outside = case
    lists, bmin, bext = H.uniform_case(n, box, radius, outside=outside)
    gpu = run_density(surf, lists, bmin, bext, res, cyclic, sigma=sigma, home=True)
    home = surf.home_voxels()
    ref_home = oracle.home_voxels(lists, bmin, bext, res, cyclic)
    assert np.array_equal(home, ref_home), "home voxels must be bit-exact"
    ref, (mn, mx) = oracle.density_p2d(lists, bmin, bext, res, cyclic, sigma=sigma)
    assert np.array_equal(gpu != 0, ref != 0) or H.density_close(gpu, ref) < H.DENSITY_RTOL
    err = H.density_close(gpu, ref)
    assert err < H.DENSITY_RTOL, f"density rel err {err}"
    gmn, gmx = surf.density_range()
    assert abs(gmn - mn) <= 1e-5 * max(abs(mn), 1e-5) and abs(gmx - mx) <= 1e-5 * abs(mx)


@pytest.mark.parametrize("ctype", ["USHORT_RGBA", "DOUBLE_I", "FLOAT_RGB"])
def test_intensity_from_the_remaining_colour_types(surf, oracle, ctype):
    """Aggregator 1 weights every particle with its colour-R accessor as a float (ParticlesToDensity.cpp:483,515;
    SimpleSphericalParticles.h:123-178): raw unsigned short, double narrowed to float, first float of an RGB triple.
    Also through QuickSurf's colour conversion (ushort / 255, intensity min-max normalised, QuickSurf.cpp:511-577)."""
    rng = np.random.default_rng(17)
    n = 5000
    xyz = synth.uniform_box(n, 12.0, seed=31)
    if ctype == "USHORT_RGBA":
        col = rng.integers(0, 65535, size=(n, 4), dtype=np.uint16)
        ct, stride = H.COL_USHORT_RGBA, 8
    elif ctype == "DOUBLE_I":
        col = (rng.random(n) * 5.0 + 0.25).astype(np.float64)
        ct, stride = H.COL_DOUBLE_I, 8
    else:
        col = rng.random((n, 3)).astype(np.float32)
        ct, stride = H.COL_FLOAT_RGB, 12
    col = np.ascontiguousarray(col)
    lists = [dict(vtx=xyz, vtx_type=H.VERT_FLOAT_XYZ, count=n, global_radius=0.9, col=col, col_type=ct, col_stride=stride)]
    res = (24, 24, 24)
    gpu = run_density(surf, lists, (0, 0, 0), (12, 12, 12), res, (True,) * 3, aggregator=1, normalize=False)
    ref, _ = oracle.density_p2d(lists, (0, 0, 0), (12, 12, 12), res, (True,) * 3, aggregator=1, normalize=False)
    assert ref.max() > 0
    err = np.abs(gpu.astype(np.float64) - ref) / np.maximum(np.abs(ref), H.DENSITY_FLOOR * max(1.0, float(ref.max())))
    assert err.max() < H.DENSITY_RTOL, err.max()
    # QuickSurf colour conversion of the same list (Gaussian mode, density-weighted RGB volume)
    data = np.concatenate([xyz, np.full((n, 1), 0.8, np.float32)], axis=1).astype(np.float32)
    ql = [dict(vtx=data, vtx_type=H.VERT_FLOAT_XYZR, vtx_stride=16, count=n, col=col, col_type=ct, col_stride=stride)]
    if ct == H.COL_DOUBLE_I:
        ql[0]["irange"] = (float(np.float32(col.min())), float(np.float32(col.max())))
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (12, 12, 12), res, (False,) * 3)
    surf.set_params(mode=1, aggregator=0, normalize=0, radscale=1.0, gausslim=2.0, colour=1)
    surf.push_particles(ql)
    surf.compute_density()
    vol, rgb = surf.get_density(with_rgb=True)
    sd = (np.float32(12.0) / np.float32(23.0)) * np.ones(3, np.float32)
    rvol, rrgb = oracle.density_gauss(ql, (0, 0, 0), sd, res, radscale=1.0, gausslim=2.0, colour=True)
    scale = float(rvol.max())
    assert (np.abs(vol.astype(np.float64) - rvol) / np.maximum(np.abs(rvol), 1e-5 * scale)).max() < 2e-5
    assert (np.abs(rgb.astype(np.float64) - rrgb) / np.maximum(np.abs(rrgb), 1e-5 * scale)).max() < 2e-5
    surf.set_params(mode=0, colour=0)


def test_density_deterministic(surf):
    lists, bmin, bext = H.uniform_case(30000, 24.0, 0.7)
    a = run_density(surf, lists, bmin, bext, (48, 48, 48), (True, True, True))
    for _ in range(3):
        b = run_density(surf, lists, bmin, bext, (48, 48, 48), (True, True, True))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "density must be bit-reproducible run to run"


def test_normalize_and_intensity(surf, oracle):
    rng = np.random.default_rng(5)
    n = 4000
    xyz = synth.uniform_box(n, 10.0)
    inten = rng.random(n).astype(np.float32) * 3 - 1
    inter = np.concatenate([xyz, inten[:, None]], axis=1).astype(np.float32)  # x y z I interleaved, stride 16
    lists = [dict(vtx=inter, vtx_type=H.VERT_FLOAT_XYZ, vtx_stride=16, count=n, global_radius=0.8,
                  col=inter.ctypes.data + 12, col_type=H.COL_FLOAT_I, col_stride=16)]
    res = (20, 20, 20)
    absl = [dict(lists[0], col=np.abs(inten), col_stride=4)]
    for agg, norm in ((1, False), (0, True), (1, True)):
        gpu = run_density(surf, lists, (0, 0, 0), (10, 10, 10), res, (True,) * 3, aggregator=agg, normalize=norm)
        ref, (mn, mx) = oracle.density_p2d(lists, (0, 0, 0), (10, 10, 10), res, (True,) * 3, aggregator=agg, normalize=norm)
        # signed weights cancel: the error of a sum is relative to the sum of the |terms|, not to the result
        mag, _ = oracle.density_p2d(absl if agg == 1 else lists, (0, 0, 0), (10, 10, 10), res, (True,) * 3, aggregator=agg)
        if norm:
            mag = mag / (mx - mn) + abs(mn) / (mx - mn)
        err = np.abs(gpu.astype(np.float64) - ref) / np.maximum(np.abs(mag), H.DENSITY_FLOOR)
        assert err.max() < H.DENSITY_RTOL, (agg, norm, err.max())


@pytest.mark.parametrize("vtype", ["xyzr", "double", "short", "unaligned"])
def test_vertex_types(surf, oracle, vtype):
    n = 3000
    xyz = synth.uniform_box(n, 12.0)
    res, cyc = (24, 24, 24), (True, True, True)
    bmin, bext = (0, 0, 0), (12, 12, 12)
    if vtype == "xyzr":
        r = (0.3 + 0.9 * synth.uniform(99, 0, n, 0)).astype(np.float32)
        r[::50] = 0.0  # rad == 0 particles are skipped
        lists = [H.xyzr_list(np.concatenate([xyz, r[:, None]], 1))]
    elif vtype == "double":
        lists = [dict(vtx=xyz.astype(np.float64), vtx_type=H.VERT_DOUBLE_XYZ, count=n, global_radius=0.7)]
    elif vtype == "short":
        q = (xyz * 5).astype(np.uint16)
        lists = [dict(vtx=q, vtx_type=H.VERT_SHORT_XYZ, count=n, global_radius=2.5)]
        bext = (60, 60, 60)
    else:  # xyz + uint8 rgb, stride 15: floats are not 4-byte aligned (legal MMPLD layout)
        raw = np.zeros((n, 15), np.uint8)
        raw[:, :12] = xyz.view(np.uint8).reshape(n, 12)
        raw[:, 12:] = 7
        lists = [dict(vtx=raw, vtx_type=H.VERT_FLOAT_XYZ, vtx_stride=15, count=n, global_radius=0.7,
                      col=raw.ctypes.data + 12, col_type=H.COL_UINT8_RGB, col_stride=15)]
    gpu = run_density(surf, lists, bmin, bext, res, cyc, home=True)
    assert np.array_equal(surf.home_voxels(), oracle.home_voxels(lists, bmin, bext, res, cyc))
    ref, _ = oracle.density_p2d(lists, bmin, bext, res, cyc)
    assert H.density_close(gpu, ref) < H.DENSITY_RTOL


def test_multiple_lists_and_empty(surf, oracle):
    a = synth.uniform_box(1500, 9.0, seed=11)
    b = synth.uniform_box(700, 9.0, seed=12)
    lists = [H.xyz_list(a, 0.6), dict(vtx=np.zeros((0, 3), np.float32), vtx_type=H.VERT_FLOAT_XYZ, count=0),
             H.xyz_list(b, 1.1)]
    res = (18, 18, 18)
    gpu = run_density(surf, lists, (0, 0, 0), (9, 9, 9), res, (False,) * 3)
    ref, _ = oracle.density_p2d([lists[0], lists[2]], (0, 0, 0), (9, 9, 9), res, (False,) * 3)
    assert H.density_close(gpu, ref) < H.DENSITY_RTOL
    # no particles at all -> zero volume, empty mesh
    surf.clear_particles()
    surf.compute_density()
    assert not surf.get_density().any()
    surf.extract_isosurface(0.5)
    pos, nrm = surf.get_mesh()
    assert pos.shape[0] == 0


MC_CASES = [(3000, (16, 16, 16), 8.0, 0.9, 0.5), (20000, (40, 37, 33), 20.0, 0.8, 0.3), (20000, (70, 33, 21), 20.0, 1.3, 1.0)]


@pytest.mark.parametrize("case", MC_CASES, ids=[f"m{i}" for i in range(len(MC_CASES))])
def test_marching_cubes(surf, oracle, case):
    n, res, box, radius, iso = case
    lists, bmin, bext = H.uniform_case(n, box, radius)
    gpu = run_density(surf, lists, bmin, bext, res, (False, False, False))
    surf.extract_isosurface(iso)
    counts = surf.cell_tricounts()
    pos, nrm = surf.get_mesh()
    # (1) classification on the SAME density array must be bit-exact
    total, ref_counts, _ = oracle.mc_count(gpu, iso)
    assert np.array_equal(counts, ref_counts)
    assert pos.shape[0] == total
    # (2) against the reference-order density: equal away from iso-value ties
    ref_vol, _ = oracle.density_p2d(lists, bmin, bext, res, (False, False, False))
    _, ref_counts2, _ = oracle.mc_count(ref_vol, iso)
    ties = H.tie_mask_cells(ref_vol, iso)
    assert np.array_equal(counts[~ties], ref_counts2[~ties])
    # (3) vertices / normals
    sd = [bext[a] / np.float32(res[a] - 1) for a in range(3)]
    rpos, rnrm, _ = oracle.mc_emit(gpu, bmin, np.array(sd, np.float32), iso)
    cell = np.array(sd, np.float64)
    assert np.abs((pos - rpos) / cell).max() <= H.VERTEX_TOL_CELLS
    assert np.abs(nrm - rnrm).max() < 1e-4
    ln = np.linalg.norm(nrm.reshape(-1, 3), axis=1)
    assert np.all((np.abs(ln - 1) < 1e-4) | (ln == 0))


def test_external_volume_isosurface(surf, oracle):
    """IsoSurface fed by a foreign VolumetricDataCall source: mms_set_density + extract."""
    z, y, x = np.mgrid[0:20, 0:24, 0:28].astype(np.float32)
    vol = np.sqrt((x - 13.3) ** 2 + (y - 11.1) ** 2 + (z - 9.7) ** 2).astype(np.float32)
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (27, 23, 19), (28, 24, 20), (False,) * 3)
    surf.set_params(want_cell_tricounts=1)
    surf.set_density(vol)
    surf.extract_isosurface(6.5)
    pos, nrm = surf.get_mesh()
    rpos, rnrm, _ = oracle.mc_emit(vol, (0, 0, 0), (1, 1, 1), 6.5)
    assert pos.shape == rpos.shape and pos.shape[0] > 100
    assert np.abs(pos - rpos).max() <= 1e-4
    # a sphere's iso-surface: normals point towards lower density = inwards here
    c = np.array([13.3, 11.1, 9.7])
    d = pos.reshape(-1, 3) - c
    cosang = -(d * nrm.reshape(-1, 3)).sum(1) / np.linalg.norm(d, axis=1)
    assert cosang.min() > 0.97


def test_noise_volume_isosurface_emits_in_windows(surf, oracle):
    """White noise crosses about half of all grid edges: far more than the emit kernel's compact vertex store holds per step (E_CAP),
    so the step is emitted in several windows of its crossing list.  Same mesh as the oracle, triangle for triangle."""
    res = (70, 41, 37)
    vol = synth.uniform(991, 0, res[0] * res[1] * res[2], 0).reshape(res[2], res[1], res[0]).astype(np.float32)
    surf.clear_particles()
    surf.set_grid((0, 0, 0), tuple(float(r - 1) for r in res), res, (False,) * 3)
    surf.set_params(want_cell_tricounts=1)
    surf.set_density(vol)
    surf.extract_isosurface(0.5)
    counts = surf.cell_tricounts()
    pos, nrm = surf.get_mesh()
    total, ref_counts, _ = oracle.mc_count(vol, 0.5)
    assert np.array_equal(counts, ref_counts) and pos.shape[0] == total and total > 2 * vol.size
    rpos, rnrm, _ = oracle.mc_emit(vol, (0, 0, 0), (1, 1, 1), 0.5)
    assert np.abs(pos - rpos).max() <= H.VERTEX_TOL_CELLS
    assert np.abs(nrm - rnrm).max() < 1e-4


def test_emit_into_caller_buffers(surf, oracle):
    """mms_count_isosurface + mms_emit_isosurface: the mesh lands at an offset of caller-owned device memory (the hook the
    z-slab driver uses to let every rank write into rank 0's mesh, and the hook for GL-interop vertex buffers)."""
    import torch
    lists, bmin, bext = H.uniform_case(8000, 10.0, 0.8)
    run_density(surf, lists, bmin, bext, (30, 30, 30), (False,) * 3)
    n = surf.count_isosurface(0.4)
    assert n > 500
    off = 37
    pos = torch.full(((n + off) * 9 + 5,), -7.0, device="cuda")
    nrm = torch.full(((n + off) * 9 + 5,), -7.0, device="cuda")
    surf.emit_isosurface(pos.data_ptr(), nrm.data_ptr(), None, off)
    surf.synchronize()
    with pytest.raises(Exception):
        surf.get_mesh()                      # the mesh lives in caller memory
    surf.extract_isosurface(0.4)
    rpos, rnrm = surf.get_mesh()
    p = pos.cpu().numpy()
    assert (p[:off * 9] == -7.0).all() and (p[(n + off) * 9:] == -7.0).all()
    assert np.array_equal(p[off * 9:(n + off) * 9].reshape(-1, 3, 3), rpos)
    assert np.array_equal(nrm.cpu().numpy()[off * 9:(n + off) * 9].reshape(-1, 3, 3), rnrm)
