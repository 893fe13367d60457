"""ctypes binding of libmmsurf.so (include/mmsurf.h).  No CPU fallback: if the library or a CUDA device is
missing, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

VERT_NONE, VERT_FLOAT_XYZ, VERT_FLOAT_XYZR, VERT_SHORT_XYZ, VERT_DOUBLE_XYZ = range(5)
(COL_NONE, COL_UINT8_RGB, COL_UINT8_RGBA, COL_FLOAT_RGB, COL_FLOAT_RGBA, COL_FLOAT_I, COL_USHORT_RGBA,
 COL_DOUBLE_I) = range(8)
MODE_P2D_BUMP, MODE_QS_GAUSS = 0, 1
ISO_MARCHING_CUBES, ISO_MARCHING_TETS = 0, 1

EXPORTS = ["mms_create", "mms_destroy", "mms_last_error", "mms_set_grid", "mms_set_slab", "mms_set_params",
           "mms_clear_particles", "mms_push_particles", "mms_push_particles_dir", "mms_get_vector_field", "mms_get_vector_field_device", "mms_get_max_radius", "mms_compute_density", "mms_get_density_range", "mms_normalize", "mms_density_range_device", "mms_normalize_device", "mms_set_stream",
           "mms_get_density", "mms_prefetch_density", "mms_get_density_device", "mms_set_density", "mms_adopt_density", "mms_extract_isosurface", "mms_set_isosurface_mode", "mms_count_isosurface", "mms_emit_isosurface", "mms_device_alloc",
           "mms_device_free", "mms_route_particles", "mms_halo_buffers", "mms_halo_push", "mms_halo_receive", "mms_halo_wait", "mms_slabs_create", "mms_slabs_destroy", "mms_slabs_last_error", "mms_slabs_count", "mms_slabs_context", "mms_slabs_set_grid", "mms_slabs_set_params", "mms_slabs_clear_particles", "mms_slabs_push_particles", "mms_slabs_compute_density", "mms_slabs_get_density_range", "mms_slabs_get_density", "mms_slabs_adopt_density", "mms_slabs_extract_isosurface", "mms_slabs_get_mesh", "mms_slabs_get_colour_volume", "mms_slabs_get_mesh_colours", "mms_ipc_export", "mms_ipc_open", "mms_ipc_close", "mms_share_enable", "mms_share_density", "mms_share_mesh", "mms_share_open", "mms_share_close", "mms_get_mesh",
           "mms_get_mesh_device", "mms_set_mesh_indexed", "mms_get_mesh_indexed", "mms_get_mesh_indexed_device", "mms_get_home_voxels", "mms_get_cell_tricounts", "mms_get_timings", "mms_synchronize",
           "mms_timer_start", "mms_timer_stop", "mms_launch_count", "mms_alloc_pinned", "mms_free_pinned", "mms_version", "mms_mmpld_open", "mms_mmpld_close",
           "mms_mmpld_last_error", "mms_mmpld_info", "mms_mmpld_prefetch", "mms_mmpld_read_frame"]


class MmsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmmsurf error {code}: {msg}")
        self.code = code


class MmsConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("reserved", C.c_int32)]


class MmsList(C.Structure):
    _fields_ = [("vtx", C.c_void_p), ("col", C.c_void_p), ("count", C.c_uint64), ("vtx_type", C.c_int32),
                ("vtx_stride", C.c_uint32), ("col_type", C.c_int32), ("col_stride", C.c_uint32),
                ("global_radius", C.c_float), ("global_rgba", C.c_uint8 * 4), ("irange", C.c_float * 2)]


class MmsGrid(C.Structure):
    _fields_ = [("min", C.c_float * 3), ("extent", C.c_float * 3), ("res", C.c_int32 * 3), ("cyclic", C.c_int32 * 3)]


class MmsParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("aggregator", C.c_int32), ("normalize", C.c_int32), ("defer_normalize", C.c_int32),
                ("sigma", C.c_float), ("radscale", C.c_float), ("gausslim", C.c_float), ("colour", C.c_int32),
                ("want_home_voxels", C.c_int32), ("want_cell_tricounts", C.c_int32)]


class MmsShare(C.Structure):
    _fields_ = [("fd", C.c_int32), ("reserved", C.c_uint32), ("alloc_bytes", C.c_uint64), ("offset", C.c_uint64), ("bytes", C.c_uint64)]


class MmsTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h2d", "bin", "density", "normalize", "mc", "d2h_volume", "d2h_mesh", "mc_emit")]


def lib_path() -> str:
    return os.path.join(_HERE, "libmmsurf.so")


_LIB = None


def load_library():
    """Loads the in-tree libmmsurf.so.  Raises if it has not been built (python -m megamol_b200.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} is missing: build it with `python -m megamol_b200.build` (nvcc, sm_100a); "
                                "there is no fallback implementation")
    L = C.CDLL(p)
    vp = C.c_void_p
    L.mms_create.argtypes = [C.POINTER(vp), C.POINTER(MmsConfig)]
    L.mms_destroy.argtypes = [vp]
    L.mms_last_error.argtypes = [vp]
    L.mms_last_error.restype = C.c_char_p
    L.mms_set_grid.argtypes = [vp, C.POINTER(MmsGrid)]
    L.mms_set_slab.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.mms_set_params.argtypes = [vp, C.POINTER(MmsParams)]
    L.mms_clear_particles.argtypes = [vp]
    L.mms_push_particles.argtypes = [vp, C.c_int32, C.POINTER(MmsList)]
    L.mms_push_particles_dir.argtypes = [vp, C.c_int32, C.POINTER(MmsList), C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.mms_get_vector_field.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_get_vector_field_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_get_max_radius.argtypes = [vp, C.POINTER(C.c_float)]
    L.mms_compute_density.argtypes = [vp]
    L.mms_get_density_range.argtypes = [vp, C.POINTER(C.c_float)]
    L.mms_normalize.argtypes = [vp, C.c_float, C.c_float]
    L.mms_density_range_device.argtypes = [vp, C.POINTER(vp)]
    L.mms_normalize_device.argtypes = [vp, vp]
    L.mms_set_stream.argtypes = [vp, vp]
    L.mms_get_density.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mms_prefetch_density.argtypes = [vp]
    L.mms_get_density_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mms_set_density.argtypes = [vp, vp]
    L.mms_adopt_density.argtypes = [vp, vp]
    L.mms_set_isosurface_mode.argtypes = [vp, C.c_int32]
    L.mms_extract_isosurface.argtypes = [vp, C.c_float]
    L.mms_count_isosurface.argtypes = [vp, C.c_float, C.POINTER(C.c_uint64)]
    L.mms_emit_isosurface.argtypes = [vp, vp, vp, vp, C.c_uint64]
    L.mms_route_particles.argtypes = [vp, C.POINTER(MmsList), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, C.c_uint64,
                                      C.POINTER(C.c_uint64)]
    L.mms_device_alloc.argtypes = [C.c_int32, C.c_size_t, C.POINTER(vp)]
    L.mms_device_free.argtypes = [C.c_int32, vp]
    L.mms_halo_buffers.argtypes = [vp, C.c_uint64, C.POINTER(vp), C.POINTER(vp)]
    L.mms_halo_push.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(vp), C.POINTER(vp), C.c_uint64]
    L.mms_halo_receive.argtypes = [vp, C.c_float]
    L.mms_slabs_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int32), C.c_int32]
    L.mms_slabs_destroy.argtypes = [vp]
    L.mms_slabs_last_error.argtypes = [vp]
    L.mms_slabs_last_error.restype = C.c_char_p
    L.mms_slabs_count.argtypes = [vp]
    L.mms_slabs_context.argtypes = [vp, C.c_int32]
    L.mms_slabs_context.restype = vp
    L.mms_slabs_set_grid.argtypes = [vp, C.POINTER(MmsGrid)]
    L.mms_slabs_set_params.argtypes = [vp, C.POINTER(MmsParams)]
    L.mms_slabs_clear_particles.argtypes = [vp]
    L.mms_slabs_push_particles.argtypes = [vp, C.c_int32, C.POINTER(MmsList)]
    L.mms_slabs_compute_density.argtypes = [vp]
    L.mms_slabs_get_density_range.argtypes = [vp, C.POINTER(C.c_float)]
    L.mms_slabs_get_density.argtypes = [vp, C.POINTER(vp)]
    L.mms_slabs_adopt_density.argtypes = [vp, vp]
    L.mms_slabs_extract_isosurface.argtypes = [vp, C.c_float]
    L.mms_slabs_get_mesh.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp)]
    L.mms_slabs_get_colour_volume.argtypes = [vp, C.POINTER(vp)]
    L.mms_slabs_get_mesh_colours.argtypes = [vp, C.POINTER(vp)]
    L.mms_ipc_export.argtypes = [C.c_int32, vp, C.POINTER(C.c_ubyte)]
    L.mms_ipc_open.argtypes = [C.c_int32, C.POINTER(C.c_ubyte), C.POINTER(vp)]
    L.mms_ipc_close.argtypes = [C.c_int32, vp]
    sp = C.POINTER(MmsShare)
    L.mms_share_enable.argtypes = [vp, C.c_int32]
    L.mms_share_density.argtypes = [vp, sp, sp]
    L.mms_share_mesh.argtypes = [vp, C.POINTER(C.c_uint64), sp, sp, sp]
    L.mms_share_open.argtypes = [C.c_int32, sp, C.POINTER(vp)]
    L.mms_share_close.argtypes = [C.c_int32, vp, sp]
    L.mms_get_mesh.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_get_mesh_device.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_set_mesh_indexed.argtypes = [vp, C.c_int32]
    L.mms_get_mesh_indexed.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_get_mesh_indexed_device.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.mms_get_home_voxels.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.mms_get_cell_tricounts.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    L.mms_get_timings.argtypes = [vp, C.POINTER(MmsTimings)]
    L.mms_synchronize.argtypes = [vp]
    L.mms_timer_start.argtypes = [vp]
    L.mms_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    L.mms_launch_count.argtypes = [vp]
    L.mms_launch_count.restype = C.c_uint64
    L.mms_alloc_pinned.argtypes = [C.c_size_t]
    L.mms_alloc_pinned.restype = vp
    L.mms_free_pinned.argtypes = [vp]
    L.mms_mmpld_open.argtypes = [C.POINTER(vp), C.c_char_p]
    L.mms_mmpld_close.argtypes = [vp]
    L.mms_mmpld_last_error.argtypes = [vp]
    L.mms_mmpld_last_error.restype = C.c_char_p
    L.mms_mmpld_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.mms_mmpld_prefetch.argtypes = [vp, C.c_uint32]
    L.mms_mmpld_read_frame.argtypes = [vp, C.c_uint32, C.POINTER(C.c_int32), C.POINTER(C.POINTER(MmsList)), C.POINTER(C.c_float)]
    _LIB = L
    return L


def _np_view(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.empty(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class Surf:
    """One libmmsurf context (= one module instance on one GPU)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        self.h = C.c_void_p()
        cfg = MmsConfig(device, 0)
        rc = self.L.mms_create(C.byref(self.h), C.byref(cfg))
        if rc:
            raise MmsError(rc, self.L.mms_last_error(None).decode())
        self.res = None
        self.z0 = 0
        self.nz = 0
        self.cell_z0 = 0
        self.cell_nz = 0
        self._keep = []
        self.params = MmsParams(MODE_P2D_BUMP, 0, 1, 0, 1.0, 1.0, 3.0, 0, 0, 0)

    def close(self):
        if self.h:
            self.L.mms_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc:
            raise MmsError(rc, self.L.mms_last_error(self.h).decode())

    def set_grid(self, bbox_min, bbox_extent, res, cyclic=(True, True, True)):
        g = MmsGrid()
        for a in range(3):
            g.min[a] = float(bbox_min[a])
            g.extent[a] = float(bbox_extent[a])
            g.res[a] = int(res[a])
            g.cyclic[a] = int(bool(cyclic[a]))
        self._chk(self.L.mms_set_grid(self.h, C.byref(g)))
        self.res = tuple(int(r) for r in res)
        self.z0, self.nz = 0, self.res[2]
        self.cell_z0, self.cell_nz = 0, self.res[2] - 1

    def set_slab(self, z0, nz, cell_z0, cell_nz):
        self._chk(self.L.mms_set_slab(self.h, int(z0), int(nz), int(cell_z0), int(cell_nz)))
        self.z0, self.nz, self.cell_z0, self.cell_nz = int(z0), int(nz), int(cell_z0), int(cell_nz)

    def set_params(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.params, k):
                raise KeyError(k)
            setattr(self.params, k, v)
        self._chk(self.L.mms_set_params(self.h, C.byref(self.params)))

    def clear_particles(self):
        self._chk(self.L.mms_clear_particles(self.h))
        self._keep = []

    def push_raw_lists(self, nlists, lists_ptr):
        """lists_ptr: ctypes POINTER(MmsList) as returned by the MMPLD reader (no Python-side repacking)."""
        self._chk(self.L.mms_push_particles(self.h, int(nlists), lists_ptr))

    def push_particles(self, lists):
        """lists: dicts {vtx: ndarray | int address (host or device), vtx_type, count, [vtx_stride], [col], [col_type],
        [col_stride], [global_radius], [global_rgba], [irange], [dir: ndarray | address of DIRDATA_FLOAT_XYZ], [dir_stride]}"""
        arr = (MmsList * len(lists))()
        dirs = (C.c_void_p * max(len(lists), 1))()
        dstrides = (C.c_uint32 * max(len(lists), 1))()
        have_dirs = False
        for i, l in enumerate(lists):
            for key, fld in (("vtx", "vtx"), ("col", "col")):
                v = l.get(key)
                if v is None:
                    continue
                if isinstance(v, np.ndarray):
                    v = np.ascontiguousarray(v)
                    self._keep.append(v)
                    setattr(arr[i], fld, v.ctypes.data)
                else:
                    setattr(arr[i], fld, int(v))
            arr[i].vtx_type = l["vtx_type"]
            arr[i].vtx_stride = l.get("vtx_stride", 0)
            arr[i].count = l["count"]
            arr[i].col_type = l.get("col_type", COL_NONE)
            arr[i].col_stride = l.get("col_stride", 0)
            arr[i].global_radius = l.get("global_radius", 0.5)
            rgba = l.get("global_rgba", (255, 255, 255, 255))
            for k in range(4):
                arr[i].global_rgba[k] = rgba[k]
            ir = l.get("irange", (0.0, 1.0))
            arr[i].irange[0], arr[i].irange[1] = ir
            d = l.get("dir")
            if d is not None:
                if isinstance(d, np.ndarray):
                    d = np.ascontiguousarray(d)
                    self._keep.append(d)
                    d = d.ctypes.data
                dirs[i] = int(d)
                dstrides[i] = int(l.get("dir_stride", 0))
                have_dirs = True
        if have_dirs:
            self._chk(self.L.mms_push_particles_dir(self.h, len(lists), arr, dirs, dstrides))
        else:
            self._chk(self.L.mms_push_particles(self.h, len(lists), arr))

    def max_radius(self) -> float:
        r = C.c_float()
        self._chk(self.L.mms_get_max_radius(self.h, C.byref(r)))
        return float(r.value)

    def compute_density(self):
        self._chk(self.L.mms_compute_density(self.h))

    def density_range(self):
        mm = (C.c_float * 2)()
        self._chk(self.L.mms_get_density_range(self.h, mm))
        return float(mm[0]), float(mm[1])

    def prefetch_density(self):
        """start the D2H copy of the volume now (overlaps a following extract_isosurface); get_density() then only waits"""
        self._chk(self.L.mms_prefetch_density(self.h))

    def density_range_device(self) -> int:
        p = C.c_void_p()
        self._chk(self.L.mms_density_range_device(self.h, C.byref(p)))
        return p.value

    def normalize_device(self, ptr):
        self._chk(self.L.mms_normalize_device(self.h, int(ptr)))

    def set_stream(self, cuda_stream):
        """cuda_stream: a cudaStream_t handle as int; 0 (torch's default stream) is passed as cudaStreamLegacy (0x1) because NULL
        means 'back to the context's own stream' in the C ABI; None restores the context's own stream."""
        if cuda_stream is None:
            self._chk(self.L.mms_set_stream(self.h, None))
        else:
            self._chk(self.L.mms_set_stream(self.h, int(cuda_stream) or 1))

    def normalize(self, mn, mx):
        self._chk(self.L.mms_normalize(self.h, float(mn), float(mx)))

    def get_density(self, copy=True, with_rgb=False):
        p = C.c_void_p()
        q = C.c_void_p()
        self._chk(self.L.mms_get_density(self.h, C.byref(p), C.byref(q) if with_rgb else None))
        v = _np_view(p.value, (self.nz, self.res[1], self.res[0]), np.float32)
        v = v.copy() if copy else v
        if not with_rgb:
            return v
        c = _np_view(q.value, (self.nz, self.res[1], self.res[0], 3), np.float32) if q.value else None
        return v, (c.copy() if (copy and c is not None) else c)

    def get_vector_field(self):
        """Aggregator 2: (vec [nz,sy,sx,3] as handed to VolumetricDataCall, magnitude [nz,sy,sx] un-normalised, direction [nz,sy,sx,3])."""
        p, q, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_get_vector_field(self.h, C.byref(p), C.byref(q), C.byref(r)))
        shape = (self.nz, self.res[1], self.res[0])
        return (_np_view(p.value, shape + (3,), np.float32).copy(), _np_view(q.value, shape, np.float32).copy(),
                _np_view(r.value, shape + (3,), np.float32).copy())

    def density_device_ptr(self):
        p = C.c_void_p()
        q = C.c_void_p()
        self._chk(self.L.mms_get_density_device(self.h, C.byref(p), C.byref(q)))
        return p.value

    def set_density(self, vol):
        if isinstance(vol, np.ndarray):
            vol = np.ascontiguousarray(vol, np.float32)
            self._keep.append(vol)
            self._chk(self.L.mms_set_density(self.h, vol.ctypes.data))
            self._chk(self.L.mms_synchronize(self.h))
        else:
            self._chk(self.L.mms_set_density(self.h, int(vol)))

    def adopt_density(self, producer: "Surf"):
        """device-resident hand-off: use the producer context's (colour) volume by reference, with this context's own mesh buffers"""
        self._chk(self.L.mms_adopt_density(self.h, producer.h))
        self.res, self.nz = producer.res, producer.nz

    def extract_isosurface(self, iso):
        self._chk(self.L.mms_extract_isosurface(self.h, float(iso)))

    # ---- device-resident hand-off (mms_share_*): exportable volume / mesh buffers, file-descriptor handles ----
    def share_enable(self, on=True):
        self._chk(self.L.mms_share_enable(self.h, 1 if on else 0))

    def share_density(self):
        """-> (volume share, colour/vector share); the caller owns (closes) the descriptors"""
        v, r = MmsShare(), MmsShare()
        self._chk(self.L.mms_share_density(self.h, C.byref(v), C.byref(r)))
        return v, r

    def share_mesh(self):
        """-> (vertex count, positions share, normals share, colours share)"""
        n = C.c_uint64()
        p, q, c = MmsShare(), MmsShare(), MmsShare()
        self._chk(self.L.mms_share_mesh(self.h, C.byref(n), C.byref(p), C.byref(q), C.byref(c)))
        return n.value, p, q, c

    def route_particles(self, ptr, count, slabs, send_ptr, capacity, vtx_type=VERT_FLOAT_XYZ, stride=0, global_radius=0.5):
        """Stable partition of a device-resident list by destination slab (mms_route_particles) -> per-slab record counts."""
        l = MmsList()
        l.vtx, l.count, l.vtx_type, l.vtx_stride, l.global_radius = int(ptr), int(count), vtx_type, stride, global_radius
        n = len(slabs)
        lo = (C.c_int32 * n)(*[s["z0"] for s in slabs])
        hi = (C.c_int32 * n)(*[s["z0"] + s["nz"] - 1 for s in slabs])
        cnt = (C.c_uint64 * n)()
        self._chk(self.L.mms_route_particles(self.h, C.byref(l), n, lo, hi, int(send_ptr), int(capacity), cnt))
        return [int(c) for c in cnt]

    def halo_buffers(self, capacity):
        """this context's halo receive buffer and counter block (device addresses), see include/mmsurf.h"""
        b, c = C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_halo_buffers(self.h, int(capacity), C.byref(b), C.byref(c)))
        return b.value, c.value

    def halo_push(self, slabs, mine, peer_bufs, peer_counters, capacity):
        """append what the other slabs need from the pushed lists to THEIR receive buffers (one kernel per list, no synchronisation)"""
        n = len(slabs)
        lo = (C.c_int32 * n)(*[s["z0"] for s in slabs])
        hi = (C.c_int32 * n)(*[s["z0"] + s["nz"] - 1 for s in slabs])
        pb = (C.c_void_p * n)(*[C.c_void_p(p) for p in peer_bufs])
        pc = (C.c_void_p * n)(*[C.c_void_p(p) for p in peer_counters])
        self._chk(self.L.mms_halo_push(self.h, n, int(mine), lo, hi, pb, pc, int(capacity)))

    def halo_wait(self, npeers):
        """stream-ordered: the context's stream waits until `npeers` slabs have signalled that their pushes of this frame have landed"""
        self._chk(self.L.mms_halo_wait(self.h, int(npeers)))

    def halo_receive(self, radius_bound):
        self._chk(self.L.mms_halo_receive(self.h, float(radius_bound)))

    def set_isosurface_mode(self, mode):
        """0 = marching cubes (default), 1 = the reference IsoSurface's marching tetrahedra, bit for bit (ISO_MARCHING_TETS)"""
        self._chk(self.L.mms_set_isosurface_mode(self.h, int(mode)))
        self.iso_mode = int(mode)

    def count_isosurface(self, iso) -> int:
        n = C.c_uint64()
        self._chk(self.L.mms_count_isosurface(self.h, float(iso), C.byref(n)))
        return int(n.value)

    def emit_isosurface(self, pos=None, nrm=None, col=None, first_triangle=0):
        """pos/nrm/col: device addresses (ints) of caller-owned buffers, or None for the library's own buffers."""
        self._chk(self.L.mms_emit_isosurface(self.h, pos, nrm, col, int(first_triangle)))

    def get_mesh(self, copy=True, normals=True, colours=False):
        n = C.c_uint64()
        p, q, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_get_mesh(self.h, C.byref(n), C.byref(p), C.byref(q) if normals else None, C.byref(r) if colours else None))
        nt = n.value // 3
        pos = _np_view(p.value, (nt, 3, 3), np.float32)
        nrm = _np_view(q.value, (nt, 3, 3), np.float32) if normals else None
        col = _np_view(r.value, (nt, 3, 3), np.float32) if (colours and r.value) else None
        if copy:
            pos = pos.copy()
            nrm = nrm.copy() if nrm is not None else None
            col = col.copy() if col is not None else None
        return (pos, nrm, col) if colours else (pos, nrm)

    def set_mesh_indexed(self, on=True):
        """opt-in indexed mesh (one vertex per crossed grid edge + 3 x uint32 per triangle) instead of the reference's triangle soup"""
        self._chk(self.L.mms_set_mesh_indexed(self.h, int(bool(on))))

    def get_mesh_indexed(self, copy=True):
        """-> (pos [nv, 3] f32, nrm [nv, 3] f32, idx [nt, 3] u32); host arrays (views of the library's pinned buffers unless copy)"""
        nv, nt = C.c_uint64(), C.c_uint64()
        p, q, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_get_mesh_indexed(self.h, C.byref(nv), C.byref(nt), C.byref(p), C.byref(q), C.byref(r)))
        pos = _np_view(p.value, (nv.value, 3), np.float32)
        nrm = _np_view(q.value, (nv.value, 3), np.float32)
        idx = _np_view(r.value, (nt.value, 3), np.uint32)
        return (pos.copy(), nrm.copy(), idx.copy()) if copy else (pos, nrm, idx)

    def mesh_indexed_device(self):
        """-> (nverts, ntris, pos, nrm, idx) device addresses"""
        nv, nt = C.c_uint64(), C.c_uint64()
        p, q, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_get_mesh_indexed_device(self.h, C.byref(nv), C.byref(nt), C.byref(p), C.byref(q), C.byref(r)))
        return nv.value, nt.value, p.value, q.value, r.value

    def mesh_device(self):
        n = C.c_uint64()
        p, q, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_get_mesh_device(self.h, C.byref(n), C.byref(p), C.byref(q), C.byref(r)))
        return n.value, p.value, q.value

    def home_voxels(self):
        p = C.c_void_p()
        n = C.c_uint64()
        self._chk(self.L.mms_get_home_voxels(self.h, C.byref(p), C.byref(n)))
        return _np_view(p.value, (n.value, 3), np.int32).copy()

    def cell_tricounts(self):
        p = C.c_void_p()
        n = C.c_uint64()
        self._chk(self.L.mms_get_cell_tricounts(self.h, C.byref(p), C.byref(n)))
        return _np_view(p.value, (self.cell_nz, self.res[1] - 1, self.res[0] - 1), np.uint8).copy()

    def timings(self):
        t = MmsTimings()
        self._chk(self.L.mms_get_timings(self.h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in MmsTimings._fields_}

    def timer_start(self):
        self._chk(self.L.mms_timer_start(self.h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self._chk(self.L.mms_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def synchronize(self):
        self._chk(self.L.mms_synchronize(self.h))

    def launch_count(self):
        return int(self.L.mms_launch_count(self.h))


class SurfGroup:
    """Several GPUs behind one handle inside one process (mms_slabs_*, include/mmsurf.h): z-slabs with halo, fused halo push over peer
    memory, global range by peer reads.  Host lists in, whole host volume / whole host mesh out; bit-identical to one GPU."""

    def __init__(self, devices):
        self.L = load_library()
        self.h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        rc = self.L.mms_slabs_create(C.byref(self.h), devs, len(devices))
        if rc:
            raise MmsError(rc, self.L.mms_slabs_last_error(None).decode())
        self.params = MmsParams(mode=0, aggregator=0, normalize=1, defer_normalize=0, sigma=1.0, radscale=1.0, gausslim=3.0)
        self.res = None
        self._keep = []

    def _chk(self, rc):
        if rc:
            raise MmsError(rc, self.L.mms_slabs_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.mms_slabs_destroy(self.h)
            self.h = C.c_void_p()

    def set_grid(self, bbox_min, bbox_extent, res, cyclic=(True, True, True)):
        g = MmsGrid()
        for a in range(3):
            g.min[a], g.extent[a], g.res[a], g.cyclic[a] = float(bbox_min[a]), float(bbox_extent[a]), int(res[a]), int(bool(cyclic[a]))
        self._chk(self.L.mms_slabs_set_grid(self.h, C.byref(g)))
        self.res = tuple(int(r) for r in res)

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._chk(self.L.mms_slabs_set_params(self.h, C.byref(self.params)))

    def clear_particles(self):
        self._chk(self.L.mms_slabs_clear_particles(self.h))
        self._keep = []

    def push_particles(self, lists):
        arr = (MmsList * len(lists))()
        for i, l in enumerate(lists):
            v = np.ascontiguousarray(l["vtx"])
            self._keep.append(v)
            arr[i].vtx, arr[i].vtx_type, arr[i].vtx_stride, arr[i].count = v.ctypes.data, l["vtx_type"], l.get("vtx_stride", 0), l["count"]
            arr[i].global_radius = l.get("global_radius", 0.5)
            col = l.get("col")
            if col is not None:  # ndarray or the address of colour data interleaved with the (kept) vertex array
                if isinstance(col, np.ndarray):
                    col = np.ascontiguousarray(col)
                    self._keep.append(col)
                    col = col.ctypes.data
                arr[i].col, arr[i].col_type, arr[i].col_stride = int(col), l.get("col_type", COL_NONE), l.get("col_stride", 0)
            rgba = l.get("global_rgba", (255, 255, 255, 255))
            for k in range(4):
                arr[i].global_rgba[k] = rgba[k]
            ir = l.get("irange", (0.0, 1.0))
            arr[i].irange[0], arr[i].irange[1] = ir
        self._chk(self.L.mms_slabs_push_particles(self.h, len(lists), arr))

    def compute_density(self):
        self._chk(self.L.mms_slabs_compute_density(self.h))

    def density_range(self):
        mm = (C.c_float * 2)()
        self._chk(self.L.mms_slabs_get_density_range(self.h, mm))
        return float(mm[0]), float(mm[1])

    def get_density(self):
        p = C.c_void_p()
        self._chk(self.L.mms_slabs_get_density(self.h, C.byref(p)))
        return _np_view(p.value, (self.res[2], self.res[1], self.res[0]), np.float32).copy()

    def get_colour_volume(self):
        """QuickSurf colour mode: the density-weighted RGB volume (sz, sy, sx, 3), or None"""
        p = C.c_void_p()
        self._chk(self.L.mms_slabs_get_colour_volume(self.h, C.byref(p)))
        if not p.value:
            return None
        return _np_view(p.value, (self.res[2], self.res[1], self.res[0], 3), np.float32).copy()

    def get_mesh_colours(self, nverts):
        p = C.c_void_p()
        self._chk(self.L.mms_slabs_get_mesh_colours(self.h, C.byref(p)))
        if not p.value:
            return None
        return _np_view(p.value, (nverts // 3, 3, 3), np.float32).copy()

    def extract_isosurface(self, iso):
        self._chk(self.L.mms_slabs_extract_isosurface(self.h, float(iso)))

    def get_mesh(self):
        n, p, q = C.c_uint64(), C.c_void_p(), C.c_void_p()
        self._chk(self.L.mms_slabs_get_mesh(self.h, C.byref(n), C.byref(p), C.byref(q)))
        nt = n.value // 3
        if nt == 0:
            return np.zeros((0, 3, 3), np.float32), np.zeros((0, 3, 3), np.float32)
        return _np_view(p.value, (nt, 3, 3), np.float32).copy(), _np_view(q.value, (nt, 3, 3), np.float32).copy()
