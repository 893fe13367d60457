"""Time-series streaming (BASELINE config 5): MMPLD frames -> density + isosurface per frame, with the upload of frame k+1
(copy stream, second upload arena, prefetched by the reader's loader thread) overlapping the kernels and the read-back of
frame k.  The reference's equivalent is the AnimDataModule frame cache + its loader thread
(plugins/mmstd/include/mmstd/data/AnimDataModule.h:183-204) feeding ParticlesToDensity one frame per GetData."""
from __future__ import annotations

import time


def stream_frames(surf, reader, nframes, iso, on_result=None, first=0, fetch_volume=True, fetch_mesh=True, indexed=False):
    """Runs frames [first, first+nframes) through `surf` (grid / params already set).  on_result(k, volume, (pos, nrm)) sees
    library-owned host views valid until the next frame ((pos, nrm, idx) with indexed=True: the opt-in indexed mesh, 28 instead of
    72 bytes per triangle over PCIe; the caller switches the context with set_mesh_indexed).  Returns per-frame wall-clock latencies in ms."""
    lat = []
    n, lp, _ = reader.read_frame(first)
    if nframes > 1:
        reader.prefetch(first + 1)
    surf.clear_particles()
    surf.push_raw_lists(n, lp)
    for k in range(nframes):
        t0 = time.perf_counter()
        surf.compute_density()
        surf.extract_isosurface(iso)
        if k + 1 < nframes:
            n2, lp2, _ = reader.read_frame(first + k + 1)
            if k + 2 < nframes:
                reader.prefetch(first + k + 2)
            surf.clear_particles()          # flips to the other upload arena
            surf.push_raw_lists(n2, lp2)    # asynchronous H2D on the copy stream
        vol = surf.get_density(copy=False) if fetch_volume else None
        mesh = (surf.get_mesh_indexed(copy=False) if indexed else surf.get_mesh(copy=False)) if fetch_mesh else None
        if on_result:
            on_result(first + k, vol, mesh)
        lat.append((time.perf_counter() - t0) * 1e3)
    return lat
