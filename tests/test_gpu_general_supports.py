"""GPU: the parameter combinations the reference computes and round 1 rejected (VERDICT r1 "what's missing" 3):

  * short periodic axes with wide supports (several periodic images of one particle reach a tile), down to a support box WIDER than the
    axis, where a voxel receives the same particle more than once (ParticlesToDensity.cpp:583-603),
  * sigma > 1 with supports wider than 8 voxels and with aggregator 2: the reference's integer box home +- ceil(rad/sliceDist) clips the
    kernel (:573-579),
  * the modules' default grid (16^3, all axes cyclic) with aggregator 2.
All against the oracle (bit-identical to the compiled reference on the golden cases, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

from megamol_b200 import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _scalar(surf, lists, box, res, cyc, sigma, aggregator=0):
    surf.clear_particles()
    surf.set_grid((0, 0, 0), box, res, cyc)
    surf.set_params(mode=0, aggregator=aggregator, normalize=0, defer_normalize=0, sigma=sigma)
    surf.push_particles(lists)
    surf.compute_density()
    return surf.get_density().copy()


CASES = [
    # name, n, res, box, radius, sigma, cyclic
    ("wide_on_short_axes", 1500, (16, 16, 16), (8.0, 8.0, 8.0), 1.6, 1.0, (True, True, True)),        # reach 3 of 16: images overlap a 32-voxel tile
    ("wider_than_the_axis", 600, (12, 40, 20), (6.0, 20.0, 10.0), 3.6, 1.0, (True, False, True)),    # 2f+1 = 17 > 12 on x
    ("sigma_gt1_wide", 800, (48, 40, 36), (24.0, 20.0, 18.0), 2.2, 2.5, (False, True, False)),       # eps = 5.5 = 11 voxels, box 5 voxels
    ("sigma_gt1_wide_cyclic", 800, (24, 24, 24), (12.0, 12.0, 12.0), 1.4, 4.0, (True, True, True)),  # eps 5.6 = 11 voxels of 24, box 3
    ("sigma_gt1_short_axis", 500, (10, 33, 17), (5.0, 16.0, 8.0), 1.1, 3.0, (True, True, True)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_scalar_density(surf, oracle, case):
    _, n, res, box, radius, sigma, cyc = case
    xyz = (synth.uniform_box(n, 1.0, seed=9100 + n) * (np.asarray(box, np.float32) * 1.1) - np.asarray(box, np.float32) * 0.05).astype(np.float32)
    lists = [H.xyz_list(xyz, radius)]
    gpu = _scalar(surf, lists, box, res, cyc, sigma)
    ref, _ = oracle.density_p2d(lists, (0, 0, 0), box, res, cyc, sigma=sigma)
    assert np.array_equal(gpu != 0, ref != 0) or H.density_close(gpu, ref) < H.DENSITY_RTOL
    assert H.density_close(gpu, ref) < 3e-5, H.density_close(gpu, ref)   # hundreds of terms per voxel: summation order
    again = _scalar(surf, lists, box, res, cyc, sigma)
    assert np.array_equal(gpu.view(np.uint32), again.view(np.uint32))


@pytest.mark.parametrize("sigma,res,cyc", [(1.0, (16, 16, 16), (True, True, True)), (2.0, (16, 16, 16), (True, True, True)),
                                           (1.5, (40, 18, 16), (True, False, True))], ids=["defaults", "defaults_sigma2", "sigma_gt1"])
def test_vector_field(surf, oracle, sigma, res, cyc):
    """aggregator 2 on the modules' default grid (all axes cyclic, 16 voxels: shorter than a gather tile + halo) and with sigma > 1"""
    n, box = 2500, (8.0, 8.0, 8.0)
    xyz = synth.uniform_box(n, 8.0, seed=4711)
    dirs = (np.stack([synth.uniform(4712, 0, n, k) for k in range(3)], 1) * 2 - 1).astype(np.float32)
    lists = [dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.7, dir=dirs)]
    surf.clear_particles()
    surf.set_grid((0, 0, 0), box, res, cyc)
    surf.set_params(mode=0, aggregator=2, normalize=0, defer_normalize=0, sigma=sigma)
    surf.push_particles(lists)
    surf.compute_density()
    vec, mag, _ = surf.get_vector_field()
    ovec, omag, _, (omn, omx) = oracle.density_p2d_vector(lists, (0, 0, 0), box, res, cyc, sigma=sigma, normalize=False)
    c = dict(lists=lists, bmin=(0, 0, 0), bext=box, res=res, cyclic=cyc, sigma=sigma)
    tail = H.vector_tail_mask(oracle, c)
    assert np.array_equal(mag != 0, omag != 0)
    assert (np.abs(vec.astype(np.float64) - ovec) / np.maximum(np.abs(ovec), 1.0))[~tail].max() < 1e-5
    mn, mx = surf.density_range()
    assert abs(mx - omx) <= 1e-5 * omx
    surf.set_params(aggregator=0, sigma=1.0)
