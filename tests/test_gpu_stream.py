"""GPU suite: MMPLD time series streamed through the pinned reader and the double-buffered upload path; every frame must
equal a fresh, unpipelined computation of that frame (so no frame ever sees another frame's particles)."""
import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import mmpld, stream, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_streamed_frames_equal_isolated_frames(tmp_path, oracle):
    nfr, n, box, res, r, iso = 6, 20000, 16.0, (40, 40, 40), 0.6, 0.4
    frames, raws = [], []
    for f in range(nfr):
        xyz = synth.uniform_box(n + 100 * f, box, seed=300 + f)   # frames differ in size and content
        raws.append(xyz)
        frames.append((float(f), [dict(vtype=1, ctype=0, data=xyz, global_radius=r)]))
    path = str(tmp_path / "series.mmpld")
    mmpld.write_mmpld(path, frames, (0, 0, 0, box, box, box))
    rd = mmpld.Reader(path)
    s = mm.Surf(0)
    s.set_grid(rd.bbox[:3], [rd.bbox[3 + a] - rd.bbox[a] for a in range(3)], res, (True,) * 3)
    s.set_params(mode=0, aggregator=0, normalize=0, sigma=1.0)
    got = {}

    def grab(k, vol, mesh):
        got[k] = (vol.copy(), mesh[0].copy())
    lat = stream.stream_frames(s, rd, nfr, iso, on_result=grab)
    assert len(lat) == nfr and sorted(got) == list(range(nfr))
    for f in range(nfr):
        ref, _ = oracle.density_p2d([H.xyz_list(raws[f], r)], (0, 0, 0), (box,) * 3, res, (1, 1, 1))
        assert H.density_close(got[f][0], ref) < H.DENSITY_RTOL, f
        total, _, _ = oracle.mc_count(got[f][0], iso)
        assert got[f][1].shape[0] == total
    # frames pushed from host arrays through the same context, unpipelined, give bit-identical volumes
    for f in (0, nfr - 1):
        s.clear_particles()
        s.push_particles([H.xyz_list(raws[f], r)])
        s.compute_density()
        assert np.array_equal(s.get_density().view(np.uint32), got[f][0].view(np.uint32))
    s.close()
    rd.close()
