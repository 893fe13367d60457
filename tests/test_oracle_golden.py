"""CPU suite: the oracle against the golden vectors produced by the UNMODIFIED reference translation units
(tests/golden/*, generator oracle/tools/gen_golden.py).  This is what pins the oracle."""
import os

import numpy as np
import pytest

from tests import golden_util as G


@pytest.mark.parametrize("path", G.p2d_cases(), ids=lambda p: os.path.basename(p)[9:-4])
def test_density_bit_identical_to_reference(oracle, path):
    c = G.load_p2d(path)
    vol, (mn, mx) = oracle.density_p2d(c["lists"], c["bmin"], c["bext"], c["res"], c["cyclic"], sigma=c["sigma"],
                                       aggregator=c["aggregator"], normalize=c["normalize"])
    # reference at one OpenMP thread accumulates in particle order, exactly like the oracle: every bit must match
    assert np.array_equal(vol.view(np.uint32), c["volume"].view(np.uint32))
    if not c["normalize"]:
        assert (mn, mx) == c["minmax"]
    else:
        assert c["minmax"] == (0.0, 1.0)
    sd = np.array(c["bext"], np.float32) / (np.array(c["res"], np.float32) - np.float32(1))
    assert np.array_equal(sd, c["slicedist"])


@pytest.mark.parametrize("path", G.vec_cases(), ids=lambda p: os.path.basename(p)[8:-4])
def test_vector_field_bit_identical_to_reference(oracle, path):
    """Aggregator 2 (IVecToSingleCell_Volume): the 3-component volume bit for bit, and the module's "outParticles" / "outInfo" payloads
    rebuilt from the oracle's magnitude / direction volumes (the order of equal magnitudes is the sort's business: compared per voxel)."""
    from tests import helpers as H
    c = G.load_vec(path)
    vec, mag, dirs, (mn, mx) = oracle.density_p2d_vector(c["lists"], c["bmin"], c["bext"], c["res"], c["cyclic"], sigma=c["sigma"],
                                                         normalize=c["normalize"])
    assert np.array_equal(vec.view(np.uint32), c["volume"].view(np.uint32))
    assert c["minmax"] == ((0.0, 1.0) if c["normalize"] else (mn, mx))
    # grid particles: exactly the voxels with a non-zero magnitude, sorted by magnitude (descending)
    idx = H.voxel_index_of(c["grid_pos"], c["bmin"], c["bext"], c["res"])
    assert len(np.unique(idx)) == len(idx) == int(np.count_nonzero(mag))
    assert np.array_equal(np.sort(idx), np.flatnonzero(mag.ravel()))
    m = mag.ravel()[idx]
    assert np.all(np.diff(m) <= 0)
    assert np.array_equal(c["grid_dir"], dirs.reshape(-1, 3)[idx])
    col = ((m - np.float32(mn)) / (np.float32(mx) - np.float32(mn))).astype(np.float32)
    assert np.array_equal(c["grid_col"], col)
    sd = np.array(c["bext"], np.float32) / (np.array(c["res"], np.float32) - np.float32(1))
    ijk = np.stack([idx % c["res"][0], (idx // c["res"][0]) % c["res"][1], idx // (c["res"][0] * c["res"][1])], 1).astype(np.float32)
    assert np.array_equal(c["grid_pos"], (np.array(c["bmin"], np.float32) + sd * ijk).astype(np.float32))
    assert abs(c["grid_radius"] - c["bext"][0] / c["res"][0] / 5.0) < 1e-7
    # table: position, direction, magnitude (normalised like the colours if "normalize")
    assert c["info_names"] == ["PositionX", "PositionY", "PositionZ", "VelocityX", "VelocityY", "VelocityZ", "VelocityMag"]
    assert np.array_equal(c["info"][:, :3], c["grid_pos"]) and np.array_equal(c["info"][:, 3:6], c["grid_dir"])
    assert np.array_equal(c["info"][:, 6], col if c["normalize"] else m)


def test_vector_field_of_a_uniform_direction_is_that_direction(oracle):
    """Property of aggregator 2: sum(w d0)/sum(w) = d0 wherever any particle reaches, whatever the weights; 0 elsewhere."""
    from megamol_b200 import synth
    n, box, res = 800, 8.0, (20, 18, 16)
    xyz = synth.uniform_box(n, box, seed=61)
    d0 = np.array([0.3, -1.2, 0.7], np.float32)
    lists = [dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.45, dir=np.tile(d0, (n, 1)))]
    vec, mag, dirs, (mn, mx) = oracle.density_p2d_vector(lists, (0, 0, 0), (box,) * 3, res, (1, 0, 1), sigma=1.0, normalize=False)
    w, _ = oracle.density_p2d([dict(vtx=xyz, vtx_type=1, count=n, global_radius=0.45)], (0, 0, 0), (box,) * 3, res, (1, 0, 1))
    hit = w > 1e-30
    assert hit.any() and (~hit).any()
    assert np.abs(vec[hit] - d0).max() < 1e-5 and not vec[w == 0].any()
    # (a voxel whose weights are a few subnormal units carries percents of rounding error, in the reference too: the range is only
    #  checked over the voxels with normal weights)
    assert abs(float(mag[hit].max()) - float(np.linalg.norm(d0))) < 1e-5 and mn == 0.0 and mx >= float(mag[hit].max())
    assert np.abs(dirs[hit] - d0 / np.linalg.norm(d0)).max() < 1e-5


def test_home_voxels_match_reference_kat(oracle):
    z = np.load(os.path.join(G.GOLDEN, "home_voxel_kat.npz"))
    pts = np.ascontiguousarray(z["points"])
    lists = [dict(vtx=pts, vtx_type=1, count=len(pts), global_radius=float(z["radius"]))]
    home = oracle.home_voxels(lists, z["bmin"], z["bext"], z["res"])
    assert np.array_equal(home, z["home"])
    # the KAT is not trivial: a float64 evaluation of (p-min)/sd disagrees for a good part of the particles
    sd = z["bext"].astype(np.float64) / (z["res"] - 1)
    naive = np.trunc((pts.astype(np.float64) - z["bmin"]) / sd).astype(np.int32)
    assert (naive != z["home"]).any()


def test_mc_tables_equal_reference(oracle):
    z = np.load(os.path.join(G.GOLDEN, "mc_tables_ref.npz"))
    words = oracle.case_words()
    assert np.array_equal(words, z["words"])
    tri, cnt = z["tri"].astype(np.int64), z["count"]
    for c in range(256):
        w = int(words[c])
        assert (w & 15) == int(cnt[c])
        for k in range(15):
            e = (w >> (4 + 4 * k)) & 15
            assert e == (int(tri[c, k]) if tri[c, k] >= 0 else 0)
    # the product's copy of the table is the same text
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    a = open(os.path.join(here, "oracle", "mc_case_words.inc")).read()
    b = open(os.path.join(here, "megamol_b200", "csrc", "mc_case_words.inc")).read()
    assert a == b
    # classic table facts: 820 triangles over the 256 cases, empty at both ends, at most 5 per cell
    assert int(cnt.sum()) == 820 and cnt[0] == 0 and cnt[255] == 0 and cnt.max() == 5


def test_marching_cubes_oracle_properties(oracle):
    """Closed surface of a sphere: watertight (every edge shared by exactly two triangles), outward normals,
    vertices on the iso-level of the trilinear field along grid edges."""
    z, y, x = np.mgrid[0:24, 0:26, 0:28].astype(np.float32)
    vol = (10.0 - np.sqrt((x - 13.2) ** 2 + (y - 12.1) ** 2 + (z - 11.7) ** 2)).astype(np.float32)
    pos, nrm, _ = oracle.mc_emit(vol, (0, 0, 0), (1, 1, 1), 2.5)
    total, counts, cub = oracle.mc_count(vol, 2.5, want_cubeidx=True)
    assert pos.shape[0] == total == counts.sum()
    v = pos.reshape(-1, 3)
    keys = [tuple(np.round(p * 4096).astype(np.int64)) for p in v]
    from collections import Counter
    edges = Counter()
    for t in range(total):
        a, b, c = keys[3 * t], keys[3 * t + 1], keys[3 * t + 2]
        for e in ((a, b), (b, c), (c, a)):
            if e[0] != e[1]:
                edges[tuple(sorted(e))] += 1
    assert set(edges.values()) == {2}
    r = np.linalg.norm(v - np.array([13.2, 12.1, 11.7]), axis=1)
    assert np.abs(r - 7.5).max() < 0.05
    # density falls outwards, so -grad points outwards
    out = (v - np.array([13.2, 12.1, 11.7])) / r[:, None]
    assert (out * nrm.reshape(-1, 3)).sum(1).min() > 0.98


def test_reference_isosurface_is_a_different_algorithm(oracle):
    """Informational pin (SURVEY 8c iv): the reference's IsoSurface is marching TETRAHEDRA; on the same volume our
    marching cubes yields fewer triangles, the same bounding box to within a cell."""
    z = np.load(os.path.join(G.GOLDEN, "isosurface_ref.npz"))
    vol, iso = z["volume"], float(z["iso"])
    total, _, _ = oracle.mc_count(vol, iso)
    ref_tris = int(z["nverts"]) // 3
    assert int(z["ntris"]) == 0            # the reference never fills its index buffer (IsoSurface.cpp:180)
    assert 0 < total < ref_tris < 3 * total


@pytest.mark.parametrize("name", ["sigma_clipped", "aniso_mixedcyc"])
def test_marching_tetrahedra_equals_the_reference_isosurface(oracle, name):
    """Golden vectors = the complete output of the unmodified trisoup_gl IsoSurface module (vertices + normals): the oracle's
    restatement of buildMesh / makeTet / interpolate (IsoSurface.cpp:229-309, 430-465, 606-735) reproduces them bit for bit."""
    z = np.load(os.path.join(G.GOLDEN, f"isosurface_mt_{name}.npz"))
    pos, nrm = oracle.mt_emit(z["volume"], z["bbox"], float(z["iso"]))
    assert pos.shape[0] * 3 == z["pos"].shape[0] and pos.shape[0] > 1000
    assert np.array_equal(pos.reshape(-1, 3), z["pos"])
    assert np.array_equal(nrm.reshape(-1, 3), z["nrm"])
