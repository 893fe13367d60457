"""CPU suite, dev container only (needs oracle/_ref/libmmref.so built from /root/reference): live comparison of the
oracle with the unmodified reference modules driven through their Call/Slot API."""
import numpy as np
import pytest

from megamol_b200 import synth

rb = pytest.importorskip("oracle.ref_binding")
pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref/libmmref.so not built (no /root/reference here)")


@pytest.fixture(scope="module")
def harness():
    return rb.Harness()


@pytest.mark.parametrize("threads", [1, 4])
@pytest.mark.parametrize("cyc", [False, True])
def test_density_live(harness, oracle, threads, cyc):
    n, box, res = 3000, 10.0, (20, 18, 16)
    xyz = synth.uniform_box(n, box * 1.06, seed=77) - np.float32(0.03 * box)
    lists = [dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.8)]
    harness.set_threads(threads)
    harness.set_particles(lists, (0, 0, 0, box, box, box))
    harness.set_p2d_params(res, cyclic=(cyc,) * 3, normalize=False, sigma=1.0)
    ref, meta = harness.pull_volume()
    vol, (mn, mx) = oracle.density_p2d(lists, (0, 0, 0), (box,) * 3, res, (cyc,) * 3)
    if threads == 1:
        assert np.array_equal(ref.view(np.uint32), vol.view(np.uint32))
    else:  # the reference's result depends on its thread count (per-thread volumes summed in thread order)
        assert np.abs(ref - vol).max() <= 1e-6 * max(1.0, float(vol.max()))
    assert np.array_equal(ref != 0, vol != 0)


def test_datahash_and_dirty_protocol(harness):
    """ParticlesToDensity recomputes iff frame / data hash / a parameter changed (ParticlesToDensity.cpp:233)."""
    xyz = synth.uniform_box(500, 4.0, seed=5)
    harness.set_threads(2)
    harness.set_particles([dict(vtx=xyz, vtx_type=1, count=500, global_radius=0.5)], (0, 0, 0, 4, 4, 4))
    harness.set_p2d_params((8, 8, 8))
    _, m1 = harness.pull_volume()
    _, m2 = harness.pull_volume()
    assert m1["datahash"] == m2["datahash"]
    harness.set_p2d_params((8, 8, 8), sigma=0.9)
    _, m3 = harness.pull_volume()
    assert m3["datahash"] == m2["datahash"] + 1


@pytest.mark.parametrize("iso", [0.15, 0.5, 1.1])
def test_marching_tetrahedra_live(harness, oracle, iso):
    """oracle mmo_mt_emit == the reference IsoSurface module, vertices and normals, bit for bit"""
    n, box, res = 3000, 10.0, (20, 18, 16)
    xyz = synth.uniform_box(n, box, seed=78)
    harness.set_threads(1)
    harness.set_particles([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.8)], (0, 0, 0, box, box, box))
    harness.set_p2d_params(res, cyclic=(True,) * 3, normalize=False, sigma=1.0)
    vol, _ = harness.pull_volume()
    m = harness.pull_mesh(iso)
    pos, nrm = oracle.mt_emit(vol, (0, 0, 0, box, box, box), iso)
    assert pos.shape[0] * 3 == m["nverts"] and m["nverts"] > 300
    assert np.array_equal(pos.reshape(-1, 3), m["pos"]) and np.array_equal(nrm.reshape(-1, 3), m["nrm"])
