"""QuickSurf grid set-up on the host (restating QuickSurf::calculateSurface, plugins/protein_cuda/src/QuickSurf.cpp:419-480
and :580-587): radius bounds -> padding -> padded box -> number of voxels from the grid spacing; quality -> gausslim."""
from __future__ import annotations

import math

import numpy as np

GAUSSLIM = {0: 2.0, 1: 2.5, 2: 3.0, 3: 4.0}  # quicksurf::quality (QuickSurf.cpp:580-587)


def grid_from_particles(xyz: np.ndarray, radii, radscale: float, gridspacing: float):
    """-> (origin(3) float32, extent(3) float32, res(3) int) such that node (i,j,k) = origin + (i,j,k)*gridspacing.
    gridpadding = max(1.5*radscale*r_max, ...) (QuickSurf.cpp:458-470); numVoxels = ceil(extent / gridspacing) (:476-478)."""
    rmax = float(np.max(radii))
    pad = np.float32(radscale * rmax * 1.5)
    mn = xyz.min(0).astype(np.float32) - pad
    mx = xyz.max(0).astype(np.float32) + pad
    res = [max(2, int(math.ceil(float(mx[a] - mn[a]) / gridspacing))) for a in range(3)]
    extent = np.array([(r - 1) * np.float32(gridspacing) for r in res], np.float32)
    return mn.astype(np.float32), extent, tuple(res)
