"""GPU suite: ParticlesToDensity aggregator 2 (IVecToSingleCell_Volume, ParticlesToDensity.cpp:493-508,629-667) through the C ABI
(mms_push_particles_dir / mms_get_vector_field) against the golden vectors of the UNMODIFIED reference module and the oracle."""
import os

import numpy as np
import pytest

from megamol_b200 import synth
from tests import golden_util as G
from tests import helpers as H

pytestmark = pytest.mark.gpu

# v = sum(w d)/sum(w): the weights carry ~1e-7 relative error and the two sums are taken in another order than the reference's, so
# the error of a component is bounded relative to the SCALE of the directions (signed d cancel), not to the component itself.
VEC_RTOL = 1e-5


def run(surf, c, normalize=None):
    surf.clear_particles()
    surf.set_grid(c["bmin"], c["bext"], c["res"], c["cyclic"])
    surf.set_params(mode=0, aggregator=2, normalize=c["normalize"] if normalize is None else normalize, defer_normalize=0, sigma=c["sigma"])
    surf.push_particles(c["lists"])
    surf.compute_density()
    return surf.get_vector_field() + (surf.density_range(),)


def dir_scale(c):
    l = c["lists"][0]
    if isinstance(l["dir"], np.ndarray):
        return float(np.abs(l["dir"]).max())
    return float(np.abs(c["keep"][0][:, 4:7]).max())


@pytest.mark.parametrize("path", G.vec_cases(), ids=lambda p: os.path.basename(p)[8:-4])
def test_vector_volume_vs_reference_golden(surf, oracle, path):
    c = G.load_vec(path)
    vec, mag, dirs, (mn, mx) = run(surf, c)
    _, omag, odirs, (omn, omx) = oracle.density_p2d_vector(c["lists"], c["bmin"], c["bext"], c["res"], c["cyclic"], sigma=c["sigma"], normalize=False)
    tail = H.vector_tail_mask(oracle, c)
    ref = c["volume"]
    scale = dir_scale(c) / ((omx - omn) if c["normalize"] else 1.0)
    err = np.abs(vec.astype(np.float64) - ref) / np.maximum(np.abs(ref), scale)
    assert err[~tail].max() < VEC_RTOL, err[~tail].max()
    # the same voxels carry a vector (the grid particles of "outParticles" are exactly these)
    assert np.array_equal(mag != 0, omag != 0)
    assert abs(mx - omx) <= VEC_RTOL * omx and abs(mn - omn) <= VEC_RTOL * max(omx, 1e-30)
    assert (np.abs(mag - omag) / np.maximum(omag, dir_scale(c)))[~tail].max() < VEC_RTOL
    # unit directions: compared where the magnitude is not itself a cancellation residue
    solid = (omag > 1e-3 * dir_scale(c)) & ~tail
    assert np.abs(dirs - odirs)[solid].max() < 1e-4
    assert np.all(dirs[omag == 0] == 0)
    # mms_get_density / the isosurface see the scalar |v| volume
    assert np.array_equal(surf.get_density(), mag)


def test_vector_field_slab_equals_whole_and_is_reproducible(surf):
    """Bit-identical for any z-slab decomposition and from run to run (fixed accumulation order, no atomics)."""
    c = G.load_vec(G.vec_cases()[0])
    whole = run(surf, c, normalize=0)
    again = run(surf, c, normalize=0)
    for a, b in zip(whole[:3], again[:3]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    sz = c["res"][2]
    for z0, nz in ((0, 7), (7, sz - 7)):
        surf.clear_particles()
        surf.set_grid(c["bmin"], c["bext"], c["res"], c["cyclic"])
        surf.set_slab(z0, nz, z0, max(nz - 1, 0))
        surf.set_params(mode=0, aggregator=2, normalize=0, defer_normalize=0, sigma=c["sigma"])
        surf.push_particles(c["lists"])
        surf.compute_density()
        part = surf.get_vector_field()
        for a, b in zip(whole[:3], part):
            assert np.array_equal(a[z0:z0 + nz].view(np.uint32), b.view(np.uint32))
    surf.set_grid(c["bmin"], c["bext"], c["res"], c["cyclic"])  # back to the whole volume for the tests that follow


def test_missing_direction_data_gives_a_zero_field(surf):
    """DIRDATA_NONE: the reference's accessors deliver 0 -> sum(w 0)/sum(w) = 0 everywhere, range 0..0."""
    n, box = 500, 8.0
    xyz = synth.uniform_box(n, box, seed=77)
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (box,) * 3, (40, 16, 16), (False,) * 3)
    surf.set_params(mode=0, aggregator=2, normalize=0, defer_normalize=0, sigma=1.0)
    surf.push_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.6)])
    surf.compute_density()
    vec, mag, dirs = surf.get_vector_field()
    assert not vec.any() and not mag.any() and not dirs.any()
    assert surf.density_range() == (0.0, 0.0)
    surf.set_params(aggregator=0)


def test_vector_field_needs_a_vector_compute(surf):
    import megamol_b200 as mm
    surf.set_params(mode=0, aggregator=0, sigma=1.0)
    n, box = 200, 8.0
    surf.clear_particles()
    surf.set_grid((0, 0, 0), (box,) * 3, (16, 16, 16), (False,) * 3)
    surf.push_particles([dict(vtx=synth.uniform_box(n, box, seed=78), vtx_type=1, count=n, global_radius=0.6)])
    surf.compute_density()
    with pytest.raises(mm.MmsError):
        surf.get_vector_field()  # the last compute was a scalar aggregator: there is no vector field


def test_mixed_lists_with_and_without_directions(surf, oracle):
    """Two lists in one call: FLOAT_XYZR with a strided direction array and a FLOAT_XYZ list with DIRDATA_NONE (its particles only add
    to the weights, pulling the average towards 0), the direction data handed over as a device pointer."""
    import torch
    n1, n2, box, res = 900, 600, 9.0, (40, 18, 20)
    a = np.zeros((n1, 4), np.float32)
    a[:, :3] = synth.uniform_box(n1, box, seed=501)
    a[:, 3] = 0.3 + 0.4 * synth.uniform(502, 0, n1, 0)
    d = np.zeros((n1, 5), np.float32)  # dx dy dz + two floats of padding: stride 20
    d[:, :3] = np.stack([synth.uniform(503, 0, n1, k) for k in range(3)], 1) * 2 - 1
    b = synth.uniform_box(n2, box, seed=504)
    lists = [dict(vtx=a, vtx_type=2, count=n1, dir=d, dir_stride=20), dict(vtx=b, vtx_type=1, count=n2, global_radius=0.5)]
    c = dict(lists=lists, bmin=(0, 0, 0), bext=(box,) * 3, res=res, cyclic=(1, 0, 1), sigma=0.9, normalize=0)
    ovec, omag, _, (omn, omx) = oracle.density_p2d_vector(lists, c["bmin"], c["bext"], res, c["cyclic"], sigma=0.9, normalize=False)
    dd = torch.from_numpy(d).cuda()
    gl = [dict(lists[0], dir=dd.data_ptr()), lists[1]]
    surf.clear_particles()
    surf.set_grid(c["bmin"], c["bext"], res, c["cyclic"])
    surf.set_params(mode=0, aggregator=2, normalize=0, defer_normalize=0, sigma=0.9)
    surf.push_particles(gl)
    surf.compute_density()
    vec, mag, _ = surf.get_vector_field()
    tail = H.vector_tail_mask(oracle, c)
    assert (np.abs(vec.astype(np.float64) - ovec) / np.maximum(np.abs(ovec), 1.0))[~tail].max() < VEC_RTOL
    assert np.array_equal(mag != 0, omag != 0)
    mn, mx = surf.density_range()
    assert abs(mx - omx) <= VEC_RTOL * omx and mn == omn == 0.0
    surf.set_params(aggregator=0, sigma=1.0)
