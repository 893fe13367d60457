"""ctypes binding of oracle/_ref/libmmrefqs.so: the reference's UNMODIFIED QuickSurf density kernels (CUDA) behind oracle/ref_qs_harness.cu.
TEST INFRASTRUCTURE -- needs a GPU; built by oracle/Makefile.ref in the dev container and shipped to the GPU box with the snapshot."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libmmrefqs.so")


def available() -> bool:
    return os.path.exists(LIB)


def density(xyzr, rgba, res, maxrad, radscale, gridspacing, isovalue, gausslim):
    """xyzr: [n,4] positions RELATIVE TO THE GRID ORIGIN + radius; rgba: [n,4] or None.  -> (density [sz,sy,sx], rgb [sz,sy,sx,3] | None,
    acceleration-grid size)"""
    L = C.CDLL(LIB)
    L.mmq_density.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                              C.c_void_p, C.c_void_p, C.c_void_p]
    xyzr = np.ascontiguousarray(xyzr, np.float32)
    nv = np.asarray(res, np.int32)
    vol = np.empty((res[2], res[1], res[0]), np.float32)
    rgb = None
    if rgba is not None:
        rgba = np.ascontiguousarray(rgba, np.float32)
        rgb = np.empty((res[2], res[1], res[0], 3), np.float32)
    accel = np.zeros(3, np.int32)
    rc = L.mmq_density(len(xyzr), xyzr.ctypes.data, rgba.ctypes.data if rgba is not None else None, nv.ctypes.data, float(maxrad),
                       float(radscale), float(gridspacing), float(isovalue), float(gausslim), vol.ctypes.data,
                       rgb.ctypes.data if rgb is not None else None, accel.ctypes.data)
    if rc:
        raise RuntimeError(f"mmq_density rc={rc}")
    return vol, rgb, tuple(int(a) for a in accel)
