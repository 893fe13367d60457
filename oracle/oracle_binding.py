"""ctypes binding of the CPU oracle (oracle/_build/libmmoracle.so, source oracle/mmoracle.cpp).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libmmoracle.so")

VERT_NONE, VERT_FLOAT_XYZ, VERT_FLOAT_XYZR, VERT_SHORT_XYZ, VERT_DOUBLE_XYZ = range(5)
(COL_NONE, COL_UINT8_RGB, COL_UINT8_RGBA, COL_FLOAT_RGB, COL_FLOAT_RGBA, COL_FLOAT_I, COL_USHORT_RGBA,
 COL_DOUBLE_I) = range(8)
VERT_SIZE = (0, 12, 16, 6, 24)
COL_SIZE = (0, 3, 4, 12, 16, 4, 8, 8)


class MmoList(C.Structure):
    _fields_ = [("vtx", C.c_void_p), ("col", C.c_void_p), ("count", C.c_uint64), ("vtx_type", C.c_int32),
                ("vtx_stride", C.c_uint32), ("col_type", C.c_int32), ("col_stride", C.c_uint32),
                ("global_radius", C.c_float), ("global_rgba", C.c_uint8 * 4), ("irange", C.c_float * 2)]


class MmoGrid(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("extent", C.c_float * 3), ("res", C.c_int32 * 3), ("cyclic", C.c_int32 * 3)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(_HERE, "mmoracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB


def pack_lists(lists, struct=MmoList):
    """lists: iterable of dicts {vtx: ndarray(raw), vtx_type, [vtx_stride], count, [col: ndarray|address], [col_type],
    [col_stride], [global_radius], [global_rgba], [irange]} -> (ctypes array, keep-alive list)."""
    arr = (struct * len(lists))()
    keep = []
    for i, l in enumerate(lists):
        vtx = l["vtx"]
        if isinstance(vtx, np.ndarray):
            vtx = np.ascontiguousarray(vtx)
            keep.append(vtx)
            arr[i].vtx = vtx.ctypes.data
        else:
            arr[i].vtx = int(vtx)
        arr[i].vtx_type = l["vtx_type"]
        arr[i].vtx_stride = l.get("vtx_stride", 0)
        arr[i].count = l["count"]
        col = l.get("col")
        if col is not None:
            if isinstance(col, np.ndarray):
                col = np.ascontiguousarray(col)
                keep.append(col)
                arr[i].col = col.ctypes.data
            else:
                arr[i].col = int(col)
        arr[i].col_type = l.get("col_type", COL_NONE)
        arr[i].col_stride = l.get("col_stride", 0)
        arr[i].global_radius = l.get("global_radius", 0.5)
        rgba = l.get("global_rgba", (255, 255, 255, 255))
        for k in range(4):
            arr[i].global_rgba[k] = rgba[k]
        ir = l.get("irange", (0.0, 1.0))
        arr[i].irange[0], arr[i].irange[1] = ir
    return arr, keep


def pack_dirs(lists):
    """Per-list direction data (DIRDATA_FLOAT_XYZ): optional keys `dir` (ndarray or address) and `dir_stride` of the list dicts
    -> (void* array, uint32 array, keep-alive list)."""
    ptrs = (C.c_void_p * max(len(lists), 1))()
    strides = (C.c_uint32 * max(len(lists), 1))()
    keep = []
    for i, l in enumerate(lists):
        d = l.get("dir")
        if d is None:
            continue
        if isinstance(d, np.ndarray):
            d = np.ascontiguousarray(d)
            keep.append(d)
            ptrs[i] = d.ctypes.data
        else:
            ptrs[i] = int(d)
        strides[i] = l.get("dir_stride", 0)
    return ptrs, strides, keep


def make_grid(bbox_min, bbox_extent, res, cyclic, struct=MmoGrid):
    g = struct()
    for a in range(3):
        g.min[a] = float(bbox_min[a])
        g.extent[a] = float(bbox_extent[a])
        g.res[a] = int(res[a])
        g.cyclic[a] = int(bool(cyclic[a]))
    return g


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.mmo_home_voxels.argtypes = [C.c_int, C.POINTER(MmoList), C.POINTER(MmoGrid), C.c_void_p]
        L.mmo_density_p2d.argtypes = [C.c_int, C.POINTER(MmoList), C.POINTER(MmoGrid), C.c_float, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mmo_normalize.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_float]
        L.mmo_density_gauss.argtypes = [C.c_int, C.POINTER(MmoList), C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.mmo_mc_count.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.mmo_mc_count.restype = C.c_int64
        L.mmo_mc_emit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mmo_mc_emit.restype = C.c_int64
        L.mmo_case_words.argtypes = [C.c_void_p]

    def home_voxels(self, lists, bbox_min, bbox_extent, res, cyclic=(0, 0, 0)):
        arr, keep = pack_lists(lists)
        g = make_grid(bbox_min, bbox_extent, res, cyclic)
        n = sum(int(l["count"]) for l in lists if l["vtx_type"] != VERT_NONE)
        out = np.empty((n, 3), np.int32)
        rc = self.lib.mmo_home_voxels(len(lists), arr, C.byref(g), out.ctypes.data)
        assert rc == 0
        return out

    def density_p2d(self, lists, bbox_min, bbox_extent, res, cyclic, sigma=1.0, aggregator=0, normalize=False,
                    z0=0, nz=None):
        arr, keep = pack_lists(lists)
        g = make_grid(bbox_min, bbox_extent, res, cyclic)
        nz = res[2] - z0 if nz is None else nz
        vol = np.empty((nz, res[1], res[0]), np.float32)
        mm = np.zeros(2, np.float32)
        rc = self.lib.mmo_density_p2d(len(lists), arr, C.byref(g), float(sigma), int(aggregator), int(normalize), int(z0),
                                      int(nz), vol.ctypes.data, mm.ctypes.data)
        if rc:
            raise RuntimeError(f"mmo_density_p2d rc={rc}")
        return vol, (float(mm[0]), float(mm[1]))

    def density_p2d_vector(self, lists, bbox_min, bbox_extent, res, cyclic, sigma=1.0, normalize=False):
        """Aggregator 2.  Returns (vec [sz,sy,sx,3], magnitude [sz,sy,sx], direction [sz,sy,sx,3], (minDens, maxDens))."""
        arr, keep = pack_lists(lists)
        ptrs, strides, keep2 = pack_dirs(lists)
        g = make_grid(bbox_min, bbox_extent, res, cyclic)
        shape = (res[2], res[1], res[0])
        vec = np.empty(shape + (3,), np.float32)
        mag = np.empty(shape, np.float32)
        dirs = np.empty(shape + (3,), np.float32)
        mm = np.zeros(2, np.float32)
        self.lib.mmo_density_p2d_vector.argtypes = [C.c_int, C.POINTER(MmoList), C.c_void_p, C.c_void_p, C.POINTER(MmoGrid), C.c_float,
                                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = self.lib.mmo_density_p2d_vector(len(lists), arr, ptrs, strides, C.byref(g), float(sigma), int(normalize),
                                             vec.ctypes.data, mag.ctypes.data, dirs.ctypes.data, mm.ctypes.data)
        if rc:
            raise RuntimeError(f"mmo_density_p2d_vector rc={rc}")
        return vec, mag, dirs, (float(mm[0]), float(mm[1]))

    def normalize(self, vol, mn, mx):
        self.lib.mmo_normalize(vol.ctypes.data, vol.size, float(mn), float(mx))
        return vol

    def density_gauss(self, lists, origin, spacing, res, radscale=1.0, gausslim=3.0, colour=False, z0=0, nz=None):
        arr, keep = pack_lists(lists)
        nz = res[2] - z0 if nz is None else nz
        o = np.asarray(origin, np.float32)
        s = np.asarray(spacing, np.float32)
        r = np.asarray(res, np.int32)
        vol = np.empty((nz, res[1], res[0]), np.float32)
        rgb = np.empty((nz, res[1], res[0], 3), np.float32) if colour else None
        rc = self.lib.mmo_density_gauss(len(lists), arr, o.ctypes.data, s.ctypes.data, r.ctypes.data, float(radscale),
                                        float(gausslim), int(z0), int(nz), vol.ctypes.data,
                                        rgb.ctypes.data if colour else None)
        if rc:
            raise RuntimeError(f"mmo_density_gauss rc={rc}")
        return vol, rgb

    def density_gauss_refset(self, xyzr, rgba, res, maxrad, radscale, gridspacing, isovalue, gausslim):
        """The reference QuickSurf's candidate set (no radial cut-off; mmo_density_gauss_refset).  xyzr relative to the grid origin."""
        xyzr = np.ascontiguousarray(xyzr, np.float32)
        r = np.asarray(res, np.int32)
        vol = np.empty((res[2], res[1], res[0]), np.float32)
        rgb = np.empty((res[2], res[1], res[0], 3), np.float32) if rgba is not None else None
        if rgba is not None:
            rgba = np.ascontiguousarray(rgba, np.float32)
        self.lib.mmo_density_gauss_refset.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                                      C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        rc = self.lib.mmo_density_gauss_refset(len(xyzr), xyzr.ctypes.data, rgba.ctypes.data if rgba is not None else None, r.ctypes.data,
                                               float(maxrad), float(radscale), float(gridspacing), float(isovalue), float(gausslim),
                                               vol.ctypes.data, rgb.ctypes.data if rgb is not None else None)
        if rc:
            raise RuntimeError(f"mmo_density_gauss_refset rc={rc}")
        return vol, rgb

    def mc_count(self, vol, iso, want_cubeidx=False):
        vol = np.ascontiguousarray(vol, np.float32)
        sz, sy, sx = vol.shape
        res = np.array([sx, sy, sz], np.int32)
        ncell = max(sx - 1, 0) * max(sy - 1, 0) * max(sz - 1, 0)
        cnt = np.zeros(ncell, np.uint8)
        cub = np.zeros(ncell, np.uint8) if want_cubeidx else None
        total = self.lib.mmo_mc_count(vol.ctypes.data, res.ctypes.data, float(iso), cnt.ctypes.data,
                                      cub.ctypes.data if want_cubeidx else None)
        shape = (max(sz - 1, 0), max(sy - 1, 0), max(sx - 1, 0))
        return int(total), cnt.reshape(shape), (cub.reshape(shape) if want_cubeidx else None)

    def mc_emit(self, vol, origin, sd, iso, rgb=None, z_offset=0, normals=True):
        vol = np.ascontiguousarray(vol, np.float32)
        sz, sy, sx = vol.shape
        res = np.array([sx, sy, sz], np.int32)
        total, _, _ = self.mc_count(vol, iso)
        o = np.asarray(origin, np.float32)
        s = np.asarray(sd, np.float32)
        pos = np.empty((total, 3, 3), np.float32)
        nrm = np.empty((total, 3, 3), np.float32) if normals else None
        col = np.empty((total, 3, 3), np.float32) if rgb is not None else None
        if rgb is not None:
            rgb = np.ascontiguousarray(rgb, np.float32)
        n = self.lib.mmo_mc_emit(vol.ctypes.data, rgb.ctypes.data if rgb is not None else None, res.ctypes.data,
                                 o.ctypes.data, s.ctypes.data, float(iso), int(z_offset), total, pos.ctypes.data,
                                 nrm.ctypes.data if normals else None, col.ctypes.data if col is not None else None)
        assert n == total, (n, total)
        return pos, nrm, col

    def mt_emit(self, vol, bbox, iso, normals=True):
        """The reference IsoSurface's marching-tetrahedra soup (mmo_mt_emit).  bbox = (left, bottom, back, right, top, front) of
        the object-space bounding box.  Returns (pos [T,3,3], nrm [T,3,3] | None)."""
        vol = np.ascontiguousarray(vol, np.float32)
        sz, sy, sx = vol.shape
        res = np.array([sx, sy, sz], np.int32)
        b = np.asarray(bbox, np.float32)
        bb = np.array([b[0], b[1], b[2], b[3] - b[0], b[4] - b[1], b[5] - b[2]], np.float32)  # Cuboid::Width() etc.: fp32 differences
        self.lib.mmo_mt_emit.restype = C.c_int64
        self.lib.mmo_mt_emit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
        total = self.lib.mmo_mt_emit(vol.ctypes.data, res.ctypes.data, bb.ctypes.data, float(iso), 0, None, None)
        pos = np.empty((total, 3, 3), np.float32)
        nrm = np.empty((total, 3, 3), np.float32) if normals else None
        n = self.lib.mmo_mt_emit(vol.ctypes.data, res.ctypes.data, bb.ctypes.data, float(iso), total, pos.ctypes.data,
                                 nrm.ctypes.data if normals else None)
        assert n == total, (n, total)
        return pos, nrm

    def case_words(self):
        w = np.empty(256, np.uint64)
        self.lib.mmo_case_words(w.ctypes.data)
        return w
