"""GPU: the kernel variants the host picks between must be interchangeable bit for bit.

  * density_splat3_kernel (warp-owned sub-tiles, chosen for supports of at most 3x3x3 voxels) vs density_splat_kernel (coloured cells):
    equal up to the fp32 summation order
  * mc_emit_kernel with the TMA plane loader (x resolution a multiple of 4) vs the cp.async loader

The environment switches are read by libmmsurf on every call (debug knobs, not part of the C ABI)."""
import os

import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import synth

pytestmark = pytest.mark.gpu


def _run(xyz, box, res, cyclic, radius, iso, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        s = mm.Surf(0)
        s.set_grid((0, 0, 0), box, res, cyclic)
        s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
        s.push_particles([dict(vtx=xyz, vtx_type=1, count=len(xyz), global_radius=radius)])
        s.compute_density()
        vol = s.get_density().copy()
        s.extract_isosurface(iso)
        pos, nrm = s.get_mesh()
        pos, nrm = pos.copy(), nrm.copy()
        s.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return vol, pos, nrm


CASES = [(150_000, (128, 96, 72), (True, True, True), 0.5), (60_000, (72, 45, 33), (False, True, False), 0.6),
         (200_000, (96, 64, 40), (True, False, True), 0.45)]


@pytest.mark.parametrize("case", CASES, ids=[f"v{i}" for i in range(len(CASES))])
def test_variants_bit_identical(case):
    n, res, cyclic, radius = case
    sd = 0.4563
    box = tuple(float(np.float32(r - 1) * np.float32(sd)) for r in res)
    xyz = synth.uniform_box(n, 1.0) * np.asarray(box, np.float32)
    base = _run(xyz, box, res, cyclic, radius, 0.4, {})
    assert base[1].shape[0] > 1000
    for env in ({"MMS_NO_TMA": "1"}, {"MMS_EMIT_V4": "1"}):  # the isosurface variants: same bits
        other = _run(xyz, box, res, cyclic, radius, 0.4, env)
        assert np.array_equal(base[0].view(np.uint32), other[0].view(np.uint32)), f"density differs with {env}"
        assert base[1].shape == other[1].shape and np.array_equal(base[1], other[1]) and np.array_equal(base[2], other[2]), f"mesh differs with {env}"
    # the density kernels sum a voxel's contributions in different (each one fixed) orders: equal up to fp32 re-association
    for env in ({"MMS_SPLAT_V1": "1"},):
        other = _run(xyz, box, res, cyclic, radius, 0.4, env)
        a, b = base[0].astype(np.float64), other[0].astype(np.float64)
        assert np.array_equal(a == 0, b == 0), f"support differs with {env}"
        assert (np.abs(a - b) <= 4e-7 * np.maximum(np.abs(b), 1e-3)).all(), f"density differs with {env}"


def test_prefetched_volume_copy_equals_plain_copy():
    """mms_prefetch_density: the D2H copy that overlaps the isosurface kernels delivers the same volume as the plain copy"""
    n, res = 120_000, (96, 80, 64)
    box = tuple(float(np.float32(r - 1) * np.float32(0.4563)) for r in res)
    xyz = synth.uniform_box(n, 1.0) * np.asarray(box, np.float32)
    out = []
    for prefetch in (False, True):
        s = mm.Surf(0)
        s.set_grid((0, 0, 0), box, res, (True, True, True))
        s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
        for _ in range(2):  # second frame: the prefetch of frame 1 must not disturb frame 2
            s.clear_particles()
            s.push_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.5)])
            s.compute_density()
            if prefetch:
                s.prefetch_density()
            s.extract_isosurface(0.4)
            vol = s.get_density()
            pos, _ = s.get_mesh()
        out.append((vol.copy(), pos.copy()))
        s.close()
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32)) and np.array_equal(out[0][1], out[1][1])
    assert out[0][0].max() == 1.0 and out[0][1].shape[0] > 1000


def test_speculative_emit_regrows_and_matches_the_plain_path():
    """mms_extract_isosurface launches the emit kernel behind the count with the EXISTING mesh buffers' capacity as its limit; a frame
    whose mesh outgrows them is emitted again after growing them.  Sequence small -> large -> small -> large on one context, every mesh
    bit-identical to the one a fresh context produces without speculation (MMS_NO_SPECULATION, a debug knob)."""
    n, res = 150_000, (96, 80, 64)
    box = tuple(float(np.float32(r - 1) * np.float32(0.4563)) for r in res)
    xyz = synth.uniform_box(n, 1.0) * np.asarray(box, np.float32)

    def fresh(iso):
        os.environ["MMS_NO_SPECULATION"] = "1"
        try:
            s = mm.Surf(0)
            s.set_grid((0, 0, 0), box, res, (True, True, True))
            s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
            s.push_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.5)])
            s.compute_density()
            s.extract_isosurface(iso)
            pos, nrm = s.get_mesh()
            s.close()
            return pos.copy(), nrm.copy()
        finally:
            os.environ.pop("MMS_NO_SPECULATION", None)

    isos = (0.9, 0.3, 0.95, 0.2, 0.2)   # few triangles, many (outgrows the buffers), few (fits easily), more again, same again
    want = {iso: fresh(iso) for iso in set(isos)}
    assert want[0.2][0].shape[0] > 1.5 * want[0.3][0].shape[0] > 3 * want[0.9][0].shape[0] > 0
    s = mm.Surf(0)
    s.set_grid((0, 0, 0), box, res, (True, True, True))
    s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    s.push_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.5)])
    s.compute_density()
    for iso in isos:
        s.extract_isosurface(iso)
        pos, nrm = s.get_mesh()
        assert pos.shape == want[iso][0].shape, iso
        assert np.array_equal(pos, want[iso][0]) and np.array_equal(nrm, want[iso][1]), iso
    s.close()
