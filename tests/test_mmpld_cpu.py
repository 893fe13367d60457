"""CPU suite: MMPLD ingest (libmmsurf's pinned double-buffered reader) against files written by our writer, and the
writer against the reference's independent parser utils/MMPLD/mmpldinfo.py where the reference checkout exists."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from megamol_b200 import mmpld, synth

REF_INFO = "/root/reference/utils/MMPLD/mmpldinfo.py"


def make_file(tmp_path, version=103, nframes=3):
    frames = []
    keep = []
    for fi in range(nframes):
        n1, n2, n3 = 1000 + 10 * fi, 257, 64
        xyz = synth.uniform_box(n1, 10.0, seed=50 + fi)
        xyzr_rgba = np.zeros((n2, 8), np.float32)
        xyzr_rgba[:, :3] = synth.uniform_box(n2, 10.0, seed=60 + fi)
        xyzr_rgba[:, 3] = 0.3
        xyzr_rgba[:, 4:] = 0.5
        xyz_u8 = np.zeros((n3, 15), np.uint8)  # float xyz + uint8 rgb: stride 15 (unaligned floats)
        xyz_u8[:, :12] = synth.uniform_box(n3, 10.0, seed=70 + fi).view(np.uint8).reshape(n3, 12)
        xyz_u8[:, 12:] = (fi + 1) * 40
        xyz_i = np.concatenate([synth.uniform_box(n3, 10.0, seed=80 + fi), synth.uniform(81 + fi, 0, n3, 0)[:, None]], 1).astype(np.float32)
        lists = [dict(vtype=1, ctype=0, data=xyz, global_radius=0.4, global_rgb=(10, 20, 30)),
                 dict(vtype=2, ctype=5, data=xyzr_rgba),
                 dict(vtype=1, ctype=1, data=xyz_u8, global_radius=0.25),
                 dict(vtype=1, ctype=3, data=xyz_i, global_radius=0.2, irange=(0.0, 1.0))]
        frames.append((0.5 * fi, lists))
        keep.append(lists)
    path = str(tmp_path / f"series_v{version}.mmpld")
    mmpld.write_mmpld(path, frames, (0, 0, 0, 10, 10, 10), version=version)
    return path, keep


@pytest.mark.parametrize("version", [100, 102, 103])
def test_reader_roundtrip(tmp_path, version):
    path, keep = make_file(tmp_path, version)
    r = mmpld.Reader(path)
    assert r.frames == 3 and r.version == version and r.bbox == (0, 0, 0, 10, 10, 10)
    r.prefetch(1)
    for fi in (0, 1, 2, 1):
        n, lp, ts = r.read_frame(fi)
        if fi + 1 < r.frames:
            r.prefetch(fi + 1)
        assert n == 4
        assert ts == (0.5 * fi if version >= 102 else float(fi))
        want = keep[fi]
        # file colour codes -> in-memory enum: 0->0, 5->4 (FLOAT_RGBA), 1->1, 3->5 (FLOAT_I)
        assert [lp[i].col_type for i in range(4)] == [0, 4, 1, 5]
        assert [lp[i].vtx_type for i in range(4)] == [1, 2, 1, 1]
        assert [lp[i].vtx_stride for i in range(4)] == [12, 32, 15, 16]
        assert lp[0].vtx % 16 == 0, "first payload must be 16-byte aligned in the pinned buffer"
        assert abs(lp[0].global_radius - 0.4) < 1e-7 and tuple(lp[0].global_rgba) == (10, 20, 30, 255)
        assert abs(lp[1].global_radius - 0.05) < 1e-7   # XYZR lists carry no global radius (MMPLDDataSource.cpp:166)
        for i in range(4):
            raw = mmpld.Reader.list_as_numpy(lp[i])
            src = np.ascontiguousarray(want[i]["data"]).view(np.uint8).reshape(raw.shape)
            assert np.array_equal(raw, src)
            if lp[i].col:
                assert lp[i].col - lp[i].vtx == (12 if lp[i].vtx_type == 1 else 16)
    r.close()


def test_bad_files(tmp_path):
    p = tmp_path / "bad.mmpld"
    p.write_bytes(b"NOTMMPLD" + b"\0" * 64)
    with pytest.raises(Exception):
        mmpld.Reader(str(p))
    with pytest.raises(Exception):
        mmpld.Reader(str(tmp_path / "missing.mmpld"))
    path, _ = make_file(tmp_path)
    r = mmpld.Reader(path)
    with pytest.raises(Exception):
        r.read_frame(99)
    # truncated payload
    data = open(path, "rb").read()
    q = tmp_path / "trunc.mmpld"
    q.write_bytes(data[:len(data) - 500])
    r2 = mmpld.Reader(str(q))
    with pytest.raises(Exception):
        r2.read_frame(2)


@pytest.mark.skipif(not os.path.exists(REF_INFO), reason="reference checkout not present")
def test_writer_against_reference_mmpldinfo(tmp_path):
    path, keep = make_file(tmp_path, 103)
    out = subprocess.run([sys.executable, REF_INFO, "-vv", path], capture_output=True, text=True, timeout=60).stdout
    assert "3 time frames" in out and "4 particle lists per frame" in out and "1385 .. 1405 particles per frame" in out
    counts = [int(x) for x in re.findall(r"^\s+(\d+) particles?$", out, flags=re.M)]
    assert counts[:4] == [1000, 257, 64, 64]
    for needle in ("VERTDATA_FLOAT_XYZR, COLDATA_FLOAT_RGBA", "VERTDATA_FLOAT_XYZ, COLDATA_UINT8_RGB", "VERTDATA_FLOAT_XYZ, COLDATA_FLOAT_I",
                   "global color: (10, 20, 30, 255)", "15 bytes per particle", "intensity color range: [0.000000, 1.000000]"):
        assert needle in out, needle
