"""GPU parity against the golden vectors of the UNMODIFIED reference translation units (tests/golden)."""
import os

import numpy as np
import pytest

from tests import golden_util as G
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", G.p2d_cases(), ids=lambda p: os.path.basename(p)[9:-4])
def test_density_vs_reference_golden(surf, path):
    c = G.load_p2d(path)
    surf.clear_particles()
    surf.set_grid(c["bmin"], c["bext"], c["res"], c["cyclic"])
    surf.set_params(mode=0, aggregator=c["aggregator"], normalize=c["normalize"], defer_normalize=0, sigma=c["sigma"])
    surf.push_particles(c["lists"])
    surf.compute_density()
    gpu = surf.get_density()
    ref = c["volume"]
    if c["aggregator"] == 1:   # signed weights cancel: error relative to the scale of the volume
        floor = max(H.DENSITY_FLOOR, 1e-2 * float(np.abs(ref).max()))
    else:
        floor = H.DENSITY_FLOOR
    err = np.abs(gpu.astype(np.float64) - ref) / np.maximum(np.abs(ref), floor)
    assert err.max() < H.DENSITY_RTOL, err.max()
    assert np.array_equal(gpu != 0, ref != 0) or c["aggregator"] == 1 or err.max() < H.DENSITY_RTOL
    mn, mx = surf.density_range()
    if not c["normalize"]:
        assert abs(mx - c["minmax"][1]) <= 1e-5 * abs(c["minmax"][1]) and abs(mn - c["minmax"][0]) <= 1e-5 * max(abs(c["minmax"][0]), 1e-5)


def test_home_voxels_vs_reference_kat(surf):
    z = np.load(os.path.join(G.GOLDEN, "home_voxel_kat.npz"))
    pts = np.ascontiguousarray(z["points"])
    surf.clear_particles()
    surf.set_grid(z["bmin"], z["bext"], z["res"], (False,) * 3)
    surf.set_params(mode=0, aggregator=0, normalize=0, sigma=4.0, want_home_voxels=1)
    surf.push_particles([dict(vtx=pts, vtx_type=1, count=len(pts), global_radius=float(z["radius"]))])
    surf.compute_density()
    assert np.array_equal(surf.home_voxels(), z["home"]), "home voxels must equal the reference's (read off its volume)"
    # and the painted pattern is the reference's: 27 non-zero voxels around every home voxel
    vol = surf.get_density()
    assert int((vol > 0).sum()) == 27 * len(pts)
    surf.set_params(sigma=1.0, want_home_voxels=0)
