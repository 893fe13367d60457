"""Synthetic particle workloads: counter-based, reproducible on any machine and any rank.

u(i, k) = (splitmix64(seed * 0x9E3779B97F4A7C15 + 4*i + k) >> 40) * 2^-24  in [0, 1) as fp32.
Every generator takes a global index range [i0, i1) so that ranks can generate disjoint chunks of one data set.
The reference's own fixture (datatools::ParticleBoxGeneratorDataSource, seed 2007) uses std::mt19937 +
uniform_real_distribution, which is not reproducible across standard libraries; we keep its default seed.
"""
from __future__ import annotations

import numpy as np

SEED = 2007
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniform(seed: int, i0: int, i1: int, k: int) -> np.ndarray:
    """u(i, k) for i in [i0, i1) as float32 in [0, 1)."""
    idx = np.arange(i0, i1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(4) * idx + np.uint64(k)) & _M64
    return ((splitmix64(key) >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


def uniform_box(n: int, box: float, seed: int = SEED + 1, i0: int = 0, i1: int | None = None) -> np.ndarray:
    """C1: uniform-random positions in [0, box)^3, FLOAT_XYZ (n x 3 float32)."""
    i1 = n if i1 is None else i1
    out = np.empty((i1 - i0, 3), np.float32)
    for k in range(3):
        out[:, k] = uniform(seed, i0, i1, k) * np.float32(box)
    return out


def lj_fluid(n: int, seed: int = SEED + 2, i0: int = 0, i1: int | None = None, spacing: float = 1.0794,
             jitter: float = 0.15, radius: float | None = None):
    """C2/C4: jittered simple-cubic lattice ("LJ-fluid-like", rho* = 0.795): the first n sites of an L^3 lattice,
    L = ceil(n^(1/3)); jitter uniform +-jitter*spacing.  Returns (xyz or xyzr float32 array, box_length)."""
    i1 = n if i1 is None else i1
    L = int(np.ceil(n ** (1.0 / 3.0) - 1e-9))
    while L ** 3 < n:
        L += 1
    idx = np.arange(i0, i1, dtype=np.int64)
    ix, iy, iz = idx % L, (idx // L) % L, idx // (L * L)
    cols = 3 if radius is None else 4
    out = np.empty((i1 - i0, cols), np.float32)
    a = np.float32(spacing)
    for k, lat in enumerate((ix, iy, iz)):
        u = uniform(seed, i0, i1, k)
        out[:, k] = (lat.astype(np.float32) + np.float32(0.5) + (u * np.float32(2.0) - np.float32(1.0)) * np.float32(jitter)) * a
    if radius is not None:
        out[:, 3] = np.float32(radius)
    return out, float(np.float32(L) * a)


def droplet(n: int, box: float, seed: int = SEED + 6, i0: int = 0, i1: int | None = None, liquid_fraction: float = 0.9,
            radius_fraction: float = 0.36):
    """A liquid droplet in vapour: liquid_fraction of the particles uniformly inside a sphere of radius
    radius_fraction*box at the box centre (by rejection-free radial sampling), the rest uniform in the box."""
    i1 = n if i1 is None else i1
    u = [uniform(seed, i0, i1, k) for k in range(4)]
    out = np.empty((i1 - i0, 3), np.float32)
    liquid = u[3] < np.float32(liquid_fraction)
    # uniform in ball: direction from (u0,u1), radius ~ cbrt(u2)
    ct = u[0] * np.float32(2) - np.float32(1)
    st = np.sqrt(np.maximum(np.float32(0), np.float32(1) - ct * ct))
    ph = u[1] * np.float32(2 * np.pi)
    rr = np.cbrt(u[2]) * np.float32(radius_fraction * box)
    c = np.float32(box / 2)
    out[:, 0] = np.where(liquid, c + rr * st * np.cos(ph), u[0] * np.float32(box))
    out[:, 1] = np.where(liquid, c + rr * st * np.sin(ph), u[1] * np.float32(box))
    out[:, 2] = np.where(liquid, c + rr * ct, u[2] * np.float32(box))
    return out.astype(np.float32)


_ELEM_RADII = np.array([1.2, 1.52, 1.55, 1.7, 1.8], np.float32)          # H O N C S
_ELEM_CUM = np.cumsum(np.array([0.50, 0.10, 0.09, 0.30, 0.01], np.float64))
_ELEM_RGB = np.array([[1.0, 1.0, 1.0], [1.0, 0.05, 0.05], [0.2, 0.2, 1.0], [0.55, 0.55, 0.55], [1.0, 0.8, 0.2]], np.float32)


def protein_like(n: int, seed: int = SEED + 3, nballs: int = 60, extent: float = 230.0):
    """C3: points filling a union of overlapping balls along a seeded random walk; FLOAT_XYZR + FLOAT_RGBA interleaved
    (stride 32).  Returns (array n x 8 float32: x y z r R G B A, bbox_min(3), bbox_max(3))."""
    # ball centres: random walk with step ~ 0.55*ball radius, reflected into [margin, extent-margin]
    rb = extent / 9.0
    cu = [uniform(seed + 100, 0, nballs, k) for k in range(3)]
    centres = np.empty((nballs, 3), np.float64)
    p = np.array([extent / 2] * 3)
    for b in range(nballs):
        d = np.array([cu[0][b], cu[1][b], cu[2][b]], np.float64) * 2 - 1
        d /= max(np.linalg.norm(d), 1e-6)
        p = p + d * rb * 0.9
        p = np.clip(p, rb + 2.0, extent - rb - 2.0)
        centres[b] = p
    u = [uniform(seed, 0, n, k) for k in range(4)]
    ball = np.minimum((u[3] * np.float32(nballs)).astype(np.int64), nballs - 1)
    ct = u[0] * np.float32(2) - np.float32(1)
    st = np.sqrt(np.maximum(np.float32(0), np.float32(1) - ct * ct))
    ph = u[1] * np.float32(2 * np.pi)
    rr = np.cbrt(u[2]) * np.float32(rb)
    out = np.empty((n, 8), np.float32)
    out[:, 0] = centres[ball, 0].astype(np.float32) + rr * st * np.cos(ph)
    out[:, 1] = centres[ball, 1].astype(np.float32) + rr * st * np.sin(ph)
    out[:, 2] = centres[ball, 2].astype(np.float32) + rr * ct
    ue = uniform(seed + 7, 0, n, 0).astype(np.float64)
    elem = np.searchsorted(_ELEM_CUM, ue, side="right").clip(0, 4)
    out[:, 3] = _ELEM_RADII[elem]
    out[:, 4:7] = _ELEM_RGB[elem]
    out[:, 7] = 1.0
    return out, np.zeros(3, np.float32), np.full(3, extent, np.float32)
