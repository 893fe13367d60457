"""GPU suite: the device-resident hand-off (mms_share_*).  The library keeps the volume and the mesh in exportable
device memory and hands out POSIX file descriptors; a consumer (here: this process through mms_share_open, and a second process that
gets the descriptor the way a renderer process would) maps them and must see exactly what the host contract (mms_get_density /
mms_get_mesh = the reference's VolumetricDataCall RAM variant and CallTriMeshData arrays) delivers."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import megamol_b200 as mm
from megamol_b200 import api, synth

pytestmark = pytest.mark.gpu


def _cudart():
    api.load_library()
    return C.CDLL("libcudart.so.12")


def _read(share, dtype=np.float32):
    """what a CUDA consumer does: map the allocation, copy the payload out, unmap"""
    L = api.load_library()
    p = C.c_void_p()
    assert L.mms_share_open(0, C.byref(share), C.byref(p)) == 0
    out = np.empty(share.bytes // np.dtype(dtype).itemsize, dtype)
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(out.ctypes.data, p, share.bytes, 2) == 0
    assert L.mms_share_close(0, p, C.byref(share)) == 0
    return out


def _frame(s, seed, n=300_000, res=(96, 80, 72), box=(40.0, 33.0, 30.0), colour=False):
    xyz = synth.uniform_box(n, 1.0, seed=seed) * np.asarray(box, np.float32)
    l = dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.6)
    if colour:
        l.update(col=np.random.default_rng(seed).random((n, 4), dtype=np.float32), col_type=4)
    s.set_grid((0, 0, 0), box, res, (True, True, True) if not colour else (False, False, False))
    if colour:
        s.set_params(mode=1, radscale=1.0, gausslim=2.0, colour=1)
    else:
        s.set_params(mode=0, aggregator=0, normalize=1, sigma=1.0)
    s.clear_particles()
    s.push_particles([l])
    s.compute_density()


def test_shared_volume_and_mesh_equal_the_host_contract():
    s = mm.Surf(0)
    s.share_enable()
    _frame(s, 5)
    vol = s.get_density().copy()
    s.extract_isosurface(0.35)
    pos, nrm = s.get_mesh()
    pos, nrm = pos.copy(), nrm.copy()
    sv, srgb = s.share_density()
    nverts, sp, sn, sc = s.share_mesh()
    try:
        assert srgb.fd == -1 and sc.fd == -1
        assert nverts * 3 == pos.size and sp.bytes == pos.nbytes and sp.alloc_bytes >= sp.bytes
        assert np.array_equal(_read(sv), vol.ravel())
        assert np.array_equal(_read(sp), pos.ravel())
        assert np.array_equal(_read(sn), nrm.ravel())
    finally:
        for x in (sv, sp, sn):
            os.close(x.fd)
    s.close()


def test_an_import_stays_valid_across_frames():
    """a renderer imports once per (re)allocation: the mapping made for frame 1 shows frame 2"""
    L = api.load_library()
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    s = mm.Surf(0)
    s.share_enable()
    _frame(s, 7)
    sv, _ = s.share_density()
    p = C.c_void_p()
    assert L.mms_share_open(0, C.byref(sv), C.byref(p)) == 0
    _frame(s, 8)
    sv2, _ = s.share_density()  # (synchronises)
    assert sv2.alloc_bytes == sv.alloc_bytes
    got = np.empty(sv.bytes // 4, np.float32)
    assert rt.cudaMemcpy(got.ctypes.data, p, sv.bytes, 2) == 0
    assert np.array_equal(got, s.get_density().ravel())
    assert L.mms_share_close(0, p, C.byref(sv)) == 0
    os.close(sv.fd), os.close(sv2.fd)
    s.close()


def test_colour_volume_and_colour_mesh_are_shared():
    s = mm.Surf(0)
    s.share_enable()
    _frame(s, 11, n=40_000, res=(64, 56, 48), box=(40.0, 33.0, 30.0), colour=True)
    s.extract_isosurface(0.5)
    pos, nrm, col = s.get_mesh(colours=True)
    col = col.copy()
    nverts, sp, sn, sc = s.share_mesh()
    sv, srgb = s.share_density()
    try:
        assert sc.fd >= 0 and srgb.fd >= 0 and srgb.bytes == 3 * sv.bytes
        assert np.array_equal(_read(sc), col.ravel())
    finally:
        for x in (sv, srgb, sp, sn, sc):
            os.close(x.fd)
    s.close()


_CONSUMER = r"""
import ctypes as C, sys, numpy as np
from megamol_b200 import api
fd, alloc, nbytes, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
L = api.load_library()
sh = api.MmsShare(fd, 0, alloc, 0, nbytes)
p = C.c_void_p()
assert L.mms_share_open(0, C.byref(sh), C.byref(p)) == 0, "open"
rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
a = np.empty(nbytes // 4, np.float32)
assert rt.cudaMemcpy(a.ctypes.data, p, nbytes, 2) == 0, "copy"
np.save(out, a)
"""


def test_another_process_maps_the_mesh(tmp_path):
    s = mm.Surf(0)
    s.share_enable()
    _frame(s, 13)
    s.extract_isosurface(0.35)
    pos, _ = s.get_mesh()
    pos = pos.copy()
    _, sp, sn, _ = s.share_mesh()
    out = str(tmp_path / "pos.npy")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(api.__file__))))
    r = subprocess.run([sys.executable, "-c", _CONSUMER, str(sp.fd), str(sp.alloc_bytes), str(sp.bytes), out], pass_fds=(sp.fd,), env=env,
        capture_output=True, text=True, timeout=300)
    os.close(sp.fd), os.close(sn.fd)
    assert r.returncode == 0, r.stderr[-2000:]
    assert np.array_equal(np.load(out), pos.ravel())
    s.close()


def test_without_share_enable_the_call_says_so():
    s = mm.Surf(0)
    _frame(s, 5, n=20_000, res=(32, 32, 32))
    with pytest.raises(mm.MmsError):
        s.share_density()
    s.close()
