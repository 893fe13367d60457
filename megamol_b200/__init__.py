"""megamol_b200 -- B200-native particle -> density volume -> isosurface path of MegaMol.

The product is libmmsurf.so (CUDA kernels for sm_100a behind the C ABI of include/mmsurf.h) plus the MegaMol
modules in plugin/b200surf.  This package is the thin Python host layer used by the tests and by bench.py:
ctypes binding (api.py), synthetic workloads (synth.py) and the multi-GPU z-slab driver (slabs.py).
"""
from .api import Surf, SurfGroup, MmsError, lib_path, load_library  # noqa: F401
