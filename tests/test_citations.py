"""CPU suite (dev container only): every `file:line` citation of the reference in the headers, docs and kernels must point at
an existing reference file with at least that many lines -- the citations are how parity claims get checked, so they must not rot.
Skipped where the reference checkout does not exist (the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
SOURCES = ["include/mmsurf.h", "DESIGN.md", "INTEGRATION.md", "README.md", "oracle/mmoracle.cpp", "oracle/ref_harness.cpp",
           "megamol_b200/csrc/common.cuh", "megamol_b200/csrc/bin.cuh", "megamol_b200/csrc/density.cuh", "megamol_b200/csrc/mc.cuh",
           "megamol_b200/csrc/mt.cuh", "megamol_b200/csrc/mmsurf.cu", "megamol_b200/csrc/mmpld.cpp",
           "plugin/b200surf/src/ParticlesToDensityB200.cpp", "plugin/b200surf/src/ParticlesToDensityB200.h",
           "plugin/b200surf/src/IsoSurfaceB200.cpp", "plugin/b200surf/src/IsoSurfaceB200.h"]
CITE = re.compile(r"([A-Za-z0-9_./]+\.(?:cpp|h|cu|py|md|cmake|txt)):(\d+)(?:-(\d+))?((?:,\s*\d+(?:-\d+)?)*)")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "plugins")), reason="no reference checkout here")


def _index():
    by_name = {}
    for dp, dn, files in os.walk(REF):
        dn[:] = [d for d in dn if d not in (".git", "externals")]
        for f in files:
            by_name.setdefault(f, []).append(os.path.join(dp, f))
    return by_name


def _resolve(path, by_name):
    full = os.path.join(REF, path)
    if os.path.isfile(full):
        return full
    cands = [c for c in by_name.get(os.path.basename(path), []) if c.endswith("/" + path) or "/" not in path]
    return cands[0] if len(cands) >= 1 else None


def test_reference_citations_resolve():
    by_name = _index()
    own = {os.path.basename(s) for s in SOURCES} | {"mmsurf.h", "bench.py", "slabs.py", "api.py", "stream.py", "gen_golden.py", "probe.cu"}
    lengths, bad, checked = {}, [], 0
    for src in SOURCES:
        text = open(os.path.join(ROOT, src), errors="ignore").read()
        for m in CITE.finditer(text):
            path = m.group(1)
            if os.path.basename(path) in own or path.startswith(("tests/", "profiles/", "scripts/", "oracle/", "megamol_b200/", "csrc/")):
                continue  # citations of this repository's own files
            target = _resolve(path, by_name)
            if target is None:
                if os.path.basename(path) in by_name or "/" in path:
                    bad.append(f"{src}: {path} does not resolve")
                continue
            if target not in lengths:
                with open(target, errors="ignore") as f:
                    lengths[target] = sum(1 for _ in f)
            nums = [int(m.group(2))] + ([int(m.group(3))] if m.group(3) else []) + [int(x) for x in re.findall(r"\d+", m.group(4) or "")]
            checked += 1
            if max(nums) > lengths[target]:
                bad.append(f"{src}: {path}:{max(nums)} is beyond the file's {lengths[target]} lines")
    assert checked > 100, checked
    assert not bad, "\n".join(bad[:40])
