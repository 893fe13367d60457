"""Loading the golden fixtures written by oracle/tools/gen_golden.py (reference TU outputs)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def p2d_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "p2d_case_*.npz")))


def load_p2d(path):
    z = np.load(path)
    data = np.ascontiguousarray(z["data"])
    if float(z["radius"]) < 0:  # x y z r I interleaved, stride 20
        lists = [dict(vtx=data, vtx_type=2, vtx_stride=20, count=len(data), col=data.ctypes.data + 16, col_type=5, col_stride=20)]
    else:
        lists = [dict(vtx=data, vtx_type=1, count=len(data), global_radius=float(z["radius"]))]
    return dict(lists=lists, res=tuple(int(r) for r in z["res"]), bmin=tuple(float(v) for v in z["bmin"]),
                bext=tuple(float(v) for v in z["bext"]), sigma=float(z["sigma"]), cyclic=tuple(int(c) for c in z["cyclic"]),
                normalize=int(z["normalize"]), aggregator=int(z["aggregator"]), volume=z["volume"],
                minmax=(float(z["minmax"][0]), float(z["minmax"][1])), slicedist=z["slicedist"], origin=z["origin"], keep=data)


def vec_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "p2d_vec_*.npz")))


def load_vec(path):
    """Aggregator-2 fixture -> dict; lists carry the direction data under `dir` / `dir_stride`."""
    z = np.load(path)
    data = np.ascontiguousarray(z["data"])
    if str(z["layout"]) == "xyz+dir":
        dirs = np.ascontiguousarray(z["dirs"])
        lists = [dict(vtx=data, vtx_type=1, count=len(data), global_radius=float(z["radius"]), dir=dirs)]
    else:  # x y z r dx dy dz interleaved, stride 28
        dirs = None
        lists = [dict(vtx=data, vtx_type=2, vtx_stride=28, count=len(data), dir=data.ctypes.data + 16, dir_stride=28)]
    out = {k: z[k] for k in ("volume", "grid_pos", "grid_dir", "grid_col", "info", "info_ranges")}
    out.update(lists=lists, res=tuple(int(r) for r in z["res"]), bmin=tuple(float(v) for v in z["bmin"]),
               bext=tuple(float(v) for v in z["bext"]), sigma=float(z["sigma"]), cyclic=tuple(int(c) for c in z["cyclic"]),
               normalize=int(z["normalize"]), minmax=(float(z["minmax"][0]), float(z["minmax"][1])), grid_radius=float(z["grid_radius"]),
               info_names=[str(n) for n in z["info_names"]], keep=(data, dirs))
    return out
