"""MMPLD files: a writer for synthetic time series (tests, bench config C5) and the binding of libmmsurf's pinned,
double-buffered frame reader.  Format: plugins/moldyn/src/io/MMPLDDataSource.cpp:61-217 (frame), :375-401 (header);
utils/MMPLD/mmpldinfo.py is an independent parser of the same format."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from . import api

VERT_SIZE = (0, 12, 16, 6, 24)
# FILE colour codes (differ from the in-memory enum): 0 none, 1 u8 rgb, 2 u8 rgba, 3 float I, 4 float rgb, 5 float rgba,
# 6 ushort rgba, 7 double I
FILE_COL_SIZE = (0, 3, 4, 4, 12, 16, 8, 8)


def write_mmpld(path, frames, bbox, clipbox=None, version=103):
    """frames: list of frames; a frame = (timestamp, [list, ...]); a list = dict(vtype, ctype (FILE code), data (n x stride
    uint8/any array reinterpreted as raw bytes), [global_radius], [global_rgb], [irange], [bbox])."""
    clipbox = bbox if clipbox is None else clipbox
    blobs = []
    for ts, lists in frames:
        b = bytearray()
        if version >= 102:
            b += struct.pack("<f", float(ts))
        b += struct.pack("<I", len(lists))
        for l in lists:
            vt, ct = l["vtype"], l.get("ctype", 0)
            raw = np.ascontiguousarray(l["data"]).view(np.uint8).reshape(-1)
            stride = VERT_SIZE[vt] + FILE_COL_SIZE[ct]
            n = len(raw) // stride if stride else 0
            assert stride == 0 or len(raw) == n * stride
            b += struct.pack("<BB", vt, ct)
            if vt in (1, 3, 4):
                b += struct.pack("<f", float(l.get("global_radius", 0.5)))
            if ct == 0:
                rgb = l.get("global_rgb", (192, 192, 192))
                b += struct.pack("<BBBB", rgb[0], rgb[1], rgb[2], 255)
            elif ct in (3, 7):
                ir = l.get("irange", (0.0, 1.0))
                b += struct.pack("<ff", float(ir[0]), float(ir[1]))
            b += struct.pack("<Q", n)
            if version >= 103:
                bb = l.get("bbox", bbox)
                b += struct.pack("<6f", *[float(v) for v in bb])
            b += raw.tobytes()
        blobs.append(bytes(b))
    header = 6 + 2 + 4 + 24 + 24 + 8 * (len(frames) + 1)
    offs = [header]
    for bl in blobs:
        offs.append(offs[-1] + len(bl))
    with open(path, "wb") as f:
        f.write(b"MMPLD\x00")
        f.write(struct.pack("<H", version))
        f.write(struct.pack("<I", len(frames)))
        f.write(struct.pack("<6f", *[float(v) for v in bbox]))
        f.write(struct.pack("<6f", *[float(v) for v in clipbox]))
        f.write(struct.pack(f"<{len(offs)}Q", *offs))
        for bl in blobs:
            f.write(bl)
    return offs


class Reader:
    """Binding of mms_mmpld_* (pinned double-buffered reader)."""

    def __init__(self, path):
        self.L = api.load_library()
        self.h = C.c_void_p()
        rc = self.L.mms_mmpld_open(C.byref(self.h), str(path).encode())
        if rc:
            msg = self.L.mms_mmpld_last_error(self.h).decode() if self.h else "open failed"
            if self.h:
                self.L.mms_mmpld_close(self.h)
                self.h = C.c_void_p()
            raise api.MmsError(rc, msg)
        fr, ver = C.c_uint32(), C.c_uint32()
        bb, cb = (C.c_float * 6)(), (C.c_float * 6)()
        self.L.mms_mmpld_info(self.h, C.byref(fr), C.byref(ver), bb, cb)
        self.frames, self.version = fr.value, ver.value
        self.bbox, self.clipbox = tuple(bb), tuple(cb)

    def close(self):
        if self.h:
            self.L.mms_mmpld_close(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def prefetch(self, frame):
        rc = self.L.mms_mmpld_prefetch(self.h, int(frame))
        if rc:
            raise api.MmsError(rc, self.L.mms_mmpld_last_error(self.h).decode())

    def read_frame(self, frame):
        """-> (nlists, POINTER(MmsList), timestamp); the pointer can go straight into Surf.push_raw_lists."""
        n = C.c_int32()
        lp = C.POINTER(api.MmsList)()
        ts = C.c_float()
        rc = self.L.mms_mmpld_read_frame(self.h, int(frame), C.byref(n), C.byref(lp), C.byref(ts))
        if rc:
            raise api.MmsError(rc, self.L.mms_mmpld_last_error(self.h).decode())
        return n.value, lp, ts.value

    @staticmethod
    def list_as_numpy(l):
        """raw bytes of one list as (count, stride) uint8 view (for tests)."""
        n, stride = int(l.count), int(l.vtx_stride)
        buf = (C.c_char * (n * stride)).from_address(l.vtx)
        return np.frombuffer(buf, np.uint8).reshape(n, stride)
