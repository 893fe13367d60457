"""CPU suite: the C-ABI library loads and exports every symbol include/mmsurf.h declares; without a GPU the only
compute entry point that may be touched, mms_create, must fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    text = open(os.path.join(ROOT, "include", "mmsurf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mms_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from megamol_b200 import api, build
    lib = build.build()  # nvcc cross-compiles for sm_100a without a GPU
    L = ctypes.CDLL(lib)
    names = declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mmsurf.h but not exported by libmmsurf.so"
    assert sorted(api.EXPORTS) == names, "megamol_b200.api.EXPORTS is out of sync with include/mmsurf.h"
    assert L.mms_version() == 1


def test_struct_layouts_match_header():
    from megamol_b200 import api
    assert ctypes.sizeof(api.MmsList) == 56 and ctypes.sizeof(api.MmsGrid) == 48
    assert ctypes.sizeof(api.MmsParams) == 40 and ctypes.sizeof(api.MmsTimings) == 32
    assert api.MmsList.count.offset == 16 and api.MmsList.global_radius.offset == 40


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import megamol_b200 as mm
    with pytest.raises(mm.MmsError) as e:
        mm.Surf(0)
    assert "no CUDA device" in str(e.value) or "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The product path (megamol_b200/, plugin/, include/) must not import, link or load anything under oracle/."""
    bad = []
    for base in ("megamol_b200", "plugin", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".txt", ".cmake")):
                    text = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"oracle[/.]|mmoracle|ref_binding|oracle_binding|libmmref", text):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
