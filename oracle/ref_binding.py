"""ctypes binding of oracle/_ref/libmmref.so (the UNMODIFIED reference modules behind oracle/ref_harness.cpp)
and of oracle/_ref/libmmplug.so (our drop-in modules in the same graph).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libmmref.so")
PLUG_LIB = os.path.join(_HERE, "_ref", "libmmplug.so")

# SimpleSphericalParticles.h:27-50
VERT_NONE, VERT_FLOAT_XYZ, VERT_FLOAT_XYZR, VERT_SHORT_XYZ, VERT_DOUBLE_XYZ = range(5)
(COL_NONE, COL_UINT8_RGB, COL_UINT8_RGBA, COL_FLOAT_RGB, COL_FLOAT_RGBA, COL_FLOAT_I, COL_USHORT_RGBA,
 COL_DOUBLE_I) = range(8)


class MmhList(C.Structure):
    _fields_ = [("vtx", C.c_void_p), ("col", C.c_void_p), ("count", C.c_uint64), ("vtx_type", C.c_int32),
                ("vtx_stride", C.c_uint32), ("col_type", C.c_int32), ("col_stride", C.c_uint32),
                ("global_radius", C.c_float), ("global_rgba", C.c_uint8 * 4), ("irange", C.c_float * 2)]


def available(path: str = REF_LIB) -> bool:
    return os.path.exists(path)


class Harness:
    """source -> ParticlesToDensity -> IsoSurface -> sink, driven through the reference's Call/Slot API."""

    def __init__(self, lib_path: str = REF_LIB, preload: str | None = None, molecule: bool = False):
        """molecule=True: the density module is fed by a MolecularDataCall source (set_molecule) instead of particle lists."""
        if preload:
            C.CDLL(preload, mode=C.RTLD_GLOBAL)
        self.lib = C.CDLL(lib_path)
        L = self.lib
        L.mmh_create.restype = C.c_void_p
        L.mmh_create_molecule.restype = C.c_void_p
        L.mmh_set_molecule.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]
        L.mmh_destroy.argtypes = [C.c_void_p]
        L.mmh_set_particles.argtypes = [C.c_void_p, C.c_int, C.POINTER(MmhList), C.POINTER(C.c_float), C.c_uint]
        L.mmh_set_p2d_params.argtypes = [C.c_void_p] + [C.c_int] * 8 + [C.c_float, C.c_int]
        L.mmh_set_param_int.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.mmh_set_param_float.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_float]
        L.mmh_set_param_string.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_char_p]
        L.mmh_pull_volume.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double)]
        L.mmh_pull_mesh.argtypes = [C.c_void_p, C.c_uint, C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_double)]
        L.mmh_copy_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        if hasattr(L, "mmh_copy_indices"):
            L.mmh_copy_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.mmh_pull_mesh2.argtypes = L.mmh_pull_mesh.argtypes
        L.mmh_select_mesh.argtypes = [C.c_void_p, C.c_int]
        L.mmh_mc_tables.argtypes = [C.c_void_p] * 5
        L.mmh_set_directions.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint]
        L.mmh_pull_grid_particles.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.c_void_p, C.c_void_p,
                                              C.c_void_p]
        L.mmh_pull_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_void_p, C.c_void_p]
        self.h = L.mmh_create_molecule() if molecule else L.mmh_create()
        if not self.h:
            raise RuntimeError("mmh_create failed" + (": the density module does not accept a MolecularDataCall" if molecule else ""))
        self._keep = []
        self.res = (16, 16, 16)
        self.frame = 0

    def close(self):
        if self.h:
            self.lib.mmh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self) -> int:
        return self.lib.mmh_threads()

    def set_threads(self, n: int):
        self.lib.mmh_set_threads(int(n))

    def set_particles(self, lists, bbox, frame_id: int = 0):
        """lists: iterable of dicts with keys vtx (np array, raw bytes), vtx_type, vtx_stride, optional col/col_type/
        col_stride, global_radius, global_rgba, irange, dir (np array of DIRDATA_FLOAT_XYZ) / dir_stride.
        bbox = (minx,miny,minz,maxx,maxy,maxz)."""
        arr = (MmhList * len(lists))()
        self._keep = []
        for i, l in enumerate(lists):
            vtx = np.ascontiguousarray(l["vtx"])
            self._keep.append(vtx)
            arr[i].vtx = vtx.ctypes.data
            arr[i].vtx_type = l["vtx_type"]
            arr[i].vtx_stride = l.get("vtx_stride", 0)
            arr[i].count = l["count"]
            col = l.get("col")
            if col is not None:
                col = np.ascontiguousarray(col) if not isinstance(col, int) else col
                if isinstance(col, int):
                    arr[i].col = col
                else:
                    self._keep.append(col)
                    arr[i].col = col.ctypes.data
            arr[i].col_type = l.get("col_type", COL_NONE)
            arr[i].col_stride = l.get("col_stride", 0)
            arr[i].global_radius = l.get("global_radius", 0.5)
            rgba = l.get("global_rgba", (255, 255, 255, 255))
            for k in range(4):
                arr[i].global_rgba[k] = rgba[k]
            ir = l.get("irange", (0.0, 1.0))
            arr[i].irange[0], arr[i].irange[1] = ir
        bb = (C.c_float * 6)(*[float(b) for b in bbox])
        self.frame = frame_id
        rc = self.lib.mmh_set_particles(self.h, len(lists), arr, bb, frame_id)
        if rc:
            raise RuntimeError(f"mmh_set_particles rc={rc}")
        for i, l in enumerate(lists):
            d = l.get("dir")
            if d is None:
                continue
            if isinstance(d, np.ndarray):
                d = np.ascontiguousarray(d)
                self._keep.append(d)
                d = d.ctypes.data
            if self.lib.mmh_set_directions(self.h, i, int(d), int(l.get("dir_stride", 0))):
                raise RuntimeError("mmh_set_directions failed")

    def set_molecule(self, pos, type_idx, radii, rgb, bbox):
        """pos [n,3] float32, type_idx [n] uint32, radii [t] float32, rgb [t,3] uint8; bbox = (minx,miny,minz,maxx,maxy,maxz)."""
        pos = np.ascontiguousarray(pos, np.float32)
        type_idx = np.ascontiguousarray(type_idx, np.uint32)
        radii = np.ascontiguousarray(radii, np.float32)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        bb = (C.c_float * 6)(*[float(b) for b in bbox])
        rc = self.lib.mmh_set_molecule(self.h, len(pos), pos.ctypes.data, type_idx.ctypes.data, len(radii), radii.ctypes.data, rgb.ctypes.data, bb)
        if rc:
            raise RuntimeError(f"mmh_set_molecule rc={rc}")

    def set_molecule_structure(self, bfactor=None, atoms_per_residue=0, residues_per_molecule=0, molecules_per_chain=0):
        """optional B-factors [n] and a synthetic structure of equal-sized residues / molecules / chains (after set_molecule)"""
        bf = None if bfactor is None else np.ascontiguousarray(bfactor, np.float32)
        self.lib.mmh_set_molecule_structure.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_uint, C.c_uint]
        if self.lib.mmh_set_molecule_structure(self.h, None if bf is None else bf.ctypes.data, int(atoms_per_residue), int(residues_per_molecule),
                                               int(molecules_per_chain)):
            raise RuntimeError("mmh_set_molecule_structure failed")

    def molecule_colour_table(self, natoms, mode0, mode1, weight, grad):
        """the unmodified reference's ProteinColor::MakeWeightedColorTable for the molecule source's atoms -> [n, 3] float32"""
        g = np.ascontiguousarray(grad, np.float32).reshape(9)
        out = np.empty((natoms, 3), np.float32)
        self.lib.mmh_molecule_colour_table.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        n = self.lib.mmh_molecule_colour_table(self.h, int(mode0), int(mode1), float(weight), g.ctypes.data, out.ctypes.data)
        if n != natoms:
            raise RuntimeError(f"mmh_molecule_colour_table returned {n} atoms")
        return out

    def set_p2d_params(self, res, cyclic=(True, True, True), normalize=True, sigma=1.0, aggregator=0,
                       for_surface=False):
        self.res = tuple(int(r) for r in res)
        rc = self.lib.mmh_set_p2d_params(self.h, aggregator, *self.res, *[int(bool(c)) for c in cyclic],
                                         int(bool(normalize)), float(sigma), int(bool(for_surface)))
        if rc:
            raise RuntimeError(f"mmh_set_p2d_params rc={rc}")

    def set_param(self, module: int, name: str, value):
        if isinstance(value, str):
            if self.lib.mmh_set_param_string(self.h, module, name.encode(), value.encode()):
                raise KeyError(name)
            return
        if isinstance(value, float):
            rc = self.lib.mmh_set_param_float(self.h, module, name.encode(), value)
        else:
            rc = self.lib.mmh_set_param_int(self.h, module, name.encode(), int(value))
        if rc:
            raise RuntimeError(f"no parameter {name!r} on module {module}")

    def pull_volume(self, copy: bool = True, components: int = 1):
        """components = 3 for aggregator 2 (the buffer must be sized before the call; the metadata confirms it)."""
        sx, sy, sz = self.res
        info = (C.c_uint64 * 5)()
        mm = (C.c_double * 2)()
        org = (C.c_float * 3)()
        sd = (C.c_float * 3)()
        ms = C.c_double()
        shape = (sz, sy, sx) if components == 1 else (sz, sy, sx, components)
        out = np.empty(shape, dtype=np.float32) if copy else None
        rc = self.lib.mmh_pull_volume(self.h, self.frame, out.ctypes.data if copy else None, info, mm, org, sd,
                                      C.byref(ms))
        if rc:
            raise RuntimeError(f"mmh_pull_volume rc={rc}")
        meta = {"resolution": tuple(info[:3]), "components": info[3], "datahash": info[4], "min": mm[0], "max": mm[1],
                "origin": tuple(org), "slicedist": tuple(sd), "ms": ms.value}
        if copy and info[3] != components:
            raise RuntimeError(f"volume has {info[3]} components, buffer was sized for {components}")
        return out, meta

    def pull_grid_particles(self):
        """'outParticles' of ParticlesToDensity (aggregator 2): dict(lists, count, vtx_type, col_type, dir_type, datahash,
        global_radius, pos [n,3], dir [n,3], col [n])."""
        info = (C.c_uint64 * 6)()
        rad = C.c_float()
        rc = self.lib.mmh_pull_grid_particles(self.h, self.frame, info, C.byref(rad), None, None, None)
        if rc:
            raise RuntimeError(f"mmh_pull_grid_particles rc={rc}")
        n = int(info[1])
        pos = np.zeros((n, 3), np.float32)
        dirs = np.zeros((n, 3), np.float32)
        col = np.zeros(n, np.float32)
        if n:
            rc = self.lib.mmh_pull_grid_particles(self.h, self.frame, info, C.byref(rad), pos.ctypes.data, dirs.ctypes.data,
                                                  col.ctypes.data)
            if rc:
                raise RuntimeError(f"mmh_pull_grid_particles rc={rc}")
        return {"lists": int(info[0]), "count": n, "vtx_type": int(info[2]), "col_type": int(info[3]), "dir_type": int(info[4]),
                "datahash": int(info[5]), "global_radius": rad.value, "pos": pos, "dir": dirs, "col": col}

    def pull_info(self):
        """'outInfo' of ParticlesToDensity (aggregator 2): dict(columns, rows, datahash, names, ranges [c,2], data [rows,c])."""
        dims = (C.c_uint64 * 3)()
        rc = self.lib.mmh_pull_info(self.h, dims, None, None, None)
        if rc:
            raise RuntimeError(f"mmh_pull_info rc={rc}")
        cols, rows = int(dims[0]), int(dims[1])
        data = np.zeros((rows, cols), np.float32)
        names = C.create_string_buffer(32 * max(cols, 1))
        ranges = np.zeros((cols, 2), np.float32)
        rc = self.lib.mmh_pull_info(self.h, dims, data.ctypes.data if rows * cols else None, names, ranges.ctypes.data if cols else None)
        if rc:
            raise RuntimeError(f"mmh_pull_info rc={rc}")
        nm = [names.raw[32 * c:32 * c + 32].split(b"\0")[0].decode() for c in range(cols)]
        return {"columns": cols, "rows": rows, "datahash": int(dims[2]), "names": nm, "ranges": ranges, "data": data}

    def reread_mesh(self, which: int, nverts: int):
        """the mesh isosurface module `which` handed out last, read AGAIN through the pointers it left in its CallTriMeshData"""
        if self.lib.mmh_select_mesh(self.h, int(which)):
            raise RuntimeError("no mesh has been pulled from that module")
        pos = np.empty((nverts, 3), np.float32)
        nrm = np.empty((nverts, 3), np.float32)
        if self.lib.mmh_copy_mesh(self.h, pos.ctypes.data, nrm.ctypes.data, None):
            raise RuntimeError("mmh_copy_mesh failed")
        return pos, nrm

    def pull_mesh(self, isoval: float, copy: bool = True, colours: bool = False, which: int = 0):
        nv = C.c_uint64()
        nt = C.c_uint64()
        ms = C.c_double()
        fn = self.lib.mmh_pull_mesh2 if which else self.lib.mmh_pull_mesh
        rc = fn(self.h, self.frame, float(isoval), C.byref(nv), C.byref(nt), C.byref(ms))
        if rc:
            raise RuntimeError(f"mmh_pull_mesh rc={rc}")
        res = {"nverts": nv.value, "ntris": nt.value, "ms": ms.value}
        if copy and nv.value:
            pos = np.empty((nv.value, 3), np.float32)
            nrm = np.empty((nv.value, 3), np.float32)
            col = np.empty((nv.value, 3), np.float32) if colours else None
            rc = self.lib.mmh_copy_mesh(self.h, pos.ctypes.data, nrm.ctypes.data, col.ctypes.data if colours else None)
            if rc:
                raise RuntimeError(f"mmh_copy_mesh rc={rc}")
            res.update(pos=pos, nrm=nrm, col=col)
        if copy and nt.value:  # an indexed mesh (the reference's own IsoSurface hands out 0 triangles: an unindexed soup)
            idx = np.empty((nt.value, 3), np.uint32)
            rc = self.lib.mmh_copy_indices(self.h, idx.ctypes.data)
            if rc:
                raise RuntimeError(f"mmh_copy_indices rc={rc}")
            res.update(idx=idx)
        return res

    def share_density(self):
        """B200 modules, 'memoryLocation' = VRAM: dict(fd, alloc_bytes, offset, bytes, memloc) of the volume's device allocation"""
        out = (C.c_int64 * 5)()
        self.lib.mmh_share_density.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        if self.lib.mmh_share_density(self.h, out):
            raise RuntimeError("the density module has no device-resident volume to share")
        return dict(fd=out[0], alloc_bytes=out[1], offset=out[2], bytes=out[3], memloc=out[4])

    def share_mesh(self, which: int = 0):
        """B200 modules, 'deviceMesh' on: (nverts, positions share, normals share) as dicts(fd, alloc_bytes, offset, bytes)"""
        out = (C.c_int64 * 9)()
        self.lib.mmh_share_mesh.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        if self.lib.mmh_share_mesh(self.h, int(which), out):
            raise RuntimeError("the isosurface module has no device-resident mesh to share")
        mk = lambda o: dict(fd=out[o], alloc_bytes=out[o + 1], offset=out[o + 2], bytes=out[o + 3])
        return out[0], mk(1), mk(5)

    def mc_tables(self):
        tri = np.empty((256, 16), np.int32)
        cnt = np.empty(256, np.uint8)
        ef = np.empty(256, np.uint32)
        vo = np.empty((8, 3), np.uint32)
        ec = np.empty((12, 2), np.uint32)
        self.lib.mmh_mc_tables(tri.ctypes.data, cnt.ctypes.data, ef.ctypes.data, vo.ctypes.data, ec.ctypes.data)
        return {"tri": tri, "count": cnt, "edgeflags": ef, "vertoff": vo, "edgeconn": ec}
