"""GPU suite: the MegaMol drop-in modules (plugin/b200surf, compiled against the reference headers into
oracle/_ref/libmmplug.so) next to the UNMODIFIED reference modules (oracle/_ref/libmmref.so), both driven through the
reference's own Module / CallerSlot / CalleeSlot / Call machinery by the same harness code."""
import numpy as np
import pytest

from megamol_b200 import synth
from tests import helpers as H

rb = pytest.importorskip("oracle.ref_binding")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (rb.available(rb.REF_LIB) and rb.available(rb.PLUG_LIB)), reason="oracle/_ref libraries not built")]


@pytest.fixture(scope="module")
def ref():
    h = rb.Harness(rb.REF_LIB)
    h.set_threads(4)
    return h


@pytest.fixture(scope="module")
def plug():
    return rb.Harness(rb.PLUG_LIB)


def feed(h, lists, bbox, res, **kw):
    h.set_particles(lists, bbox)
    h.set_p2d_params(res, **kw)


@pytest.mark.parametrize("cyc,norm", [(True, True), (False, False), (True, False)])
def test_volume_call_matches_reference(ref, plug, cyc, norm):
    n, box, res = 6000, 12.0, (28, 24, 20)
    xyz = synth.uniform_box(n, box, seed=321)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=0.8)]
    outs = []
    for h in (ref, plug):
        feed(h, lists, (0, 0, 0, box, box, box), res, cyclic=(cyc,) * 3, normalize=norm, sigma=1.0)
        outs.append(h.pull_volume())
    (rv, rm), (pv, pm) = outs
    for k in ("resolution", "components", "origin", "slicedist"):
        assert rm[k] == pm[k], k
    assert H.density_close(pv, rv) < H.DENSITY_RTOL
    assert abs(pm["min"] - rm["min"]) <= 1e-5 and abs(pm["max"] - rm["max"]) <= 1e-5 * max(1.0, abs(rm["max"]))


def test_interleaved_xyzr_rgba_and_intensity(ref, plug):
    """MMPLD-style interleaved list (x y z r | R G B A floats, stride 32), aggregator 1 uses the colour's R channel."""
    n = 3000
    buf = np.zeros((n, 8), np.float32)
    buf[:, :3] = synth.uniform_box(n, 9.0, seed=99)
    buf[:, 3] = 0.4 + 0.5 * synth.uniform(5, 0, n, 0)
    buf[:, 4] = 0.2 + synth.uniform(6, 0, n, 0)
    buf[:, 7] = 1.0
    lists = [dict(vtx=buf, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=buf.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA,
                  col_stride=32)]
    vols = []
    for h in (ref, plug):
        feed(h, lists, (0, 0, 0, 9, 9, 9), (18, 18, 18), cyclic=(True,) * 3, normalize=False, sigma=1.0, aggregator=1)
        vols.append(h.pull_volume()[0])
    assert H.density_close(vols[1], vols[0]) < H.DENSITY_RTOL


def test_datahash_protocol_matches_reference(ref, plug):
    xyz = synth.uniform_box(800, 4.0, seed=5)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=800, global_radius=0.5)]
    deltas = []
    for h in (ref, plug):
        feed(h, lists, (0, 0, 0, 4, 4, 4), (8, 8, 8))
        _, a = h.pull_volume()
        _, b = h.pull_volume()                       # nothing changed -> no recompute, same hash
        h.set_p2d_params((8, 8, 8), sigma=0.9)       # parameter dirty -> recompute, hash + 1
        _, c = h.pull_volume()
        h.set_particles(lists, (0, 0, 0, 4, 4, 4))   # new input data hash -> recompute, hash + 1
        _, d = h.pull_volume()
        deltas.append((b["datahash"] - a["datahash"], c["datahash"] - b["datahash"], d["datahash"] - c["datahash"]))
    assert deltas[0] == deltas[1] == (0, 1, 1)


def test_mesh_call_contract(ref, plug, oracle):
    """CallTriMeshData from IsoSurfaceB200: same contract as the reference (one unindexed soup, GetTriCount()==0), vertices =
    marching cubes of the very volume the VolumetricDataCall delivers."""
    n, box, res = 5000, 10.0, (24, 24, 24)
    xyz = synth.uniform_box(n, box, seed=11)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=1.0)]
    feed(plug, lists, (0, 0, 0, box, box, box), res, cyclic=(False,) * 3, normalize=True)
    vol, meta = plug.pull_volume()
    m = plug.pull_mesh(0.5)
    feed(ref, lists, (0, 0, 0, box, box, box), res, cyclic=(False,) * 3, normalize=True)
    ref.pull_volume()
    rmesh = ref.pull_mesh(0.5, copy=False)
    assert m["ntris"] == rmesh["ntris"] == 0 and m["nverts"] % 3 == 0 and m["nverts"] > 0
    total, _, _ = oracle.mc_count(vol, 0.5)
    assert m["nverts"] == 3 * total
    pos, nrm, _ = oracle.mc_emit(vol, meta["origin"], np.array(meta["slicedist"], np.float32), 0.5)
    cell = np.array(meta["slicedist"], np.float64)
    assert np.abs((m["pos"].reshape(-1, 3, 3) - pos) / cell).max() <= H.VERTEX_TOL_CELLS
    assert np.abs(m["nrm"].reshape(-1, 3, 3) - nrm).max() < 1e-4
    # a second pull without changes must not recompute and must hand out the same data
    m2 = plug.pull_mesh(0.5)
    assert m2["nverts"] == m["nverts"] and np.array_equal(m2["pos"], m["pos"])
    # changing isoval recomputes
    m3 = plug.pull_mesh(0.3)
    assert m3["nverts"] != m["nverts"]


def test_indexed_mesh_parameter(plug):
    """IsoSurfaceB200 'indexedMesh': CallTriMeshData's indexed form (SetVertexData + SetTriangleData with 32-bit indices); expanding the
    indices gives the module's default soup bit for bit, and switching back restores the reference's contract (0 triangles)."""
    n, box, res = 6000, 10.0, (40, 28, 24)
    xyz = synth.uniform_box(n, box, seed=13)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=0.9)]
    feed(plug, lists, (0, 0, 0, box, box, box), res, cyclic=(False,) * 3, normalize=True)
    plug.pull_volume()
    soup = plug.pull_mesh(0.4)
    assert soup["ntris"] == 0 and soup["nverts"] > 3000
    plug.set_param(1, "indexedMesh", 1)
    try:
        m = plug.pull_mesh(0.4)
        assert m["ntris"] * 3 == soup["nverts"] and 0 < m["nverts"] < soup["nverts"] // 3
        assert int(m["idx"].max()) < m["nverts"]
        assert np.array_equal(m["pos"][m["idx"].reshape(-1)].view(np.uint32), soup["pos"].view(np.uint32))
        assert np.array_equal(m["nrm"][m["idx"].reshape(-1)].view(np.uint32), soup["nrm"].view(np.uint32))
    finally:
        plug.set_param(1, "indexedMesh", 0)
    back = plug.pull_mesh(0.4)
    assert back["ntris"] == 0 and np.array_equal(back["pos"], soup["pos"])


def test_two_isosurface_modules_share_one_density_module(plug):
    """Two IsoSurfaceB200 behind one ParticlesToDensityB200 (different iso values): each keeps its mesh in its own context
    (mms_adopt_density takes only the device volume from the producer).  The mesh module 0 handed out stays intact -- its pointers are
    read again -- after module 1 has extracted another surface from the same producer, and after the producer has recomputed."""
    n, box, res = 6000, 10.0, (32, 28, 24)
    xyz = synth.uniform_box(n, box, seed=23)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=1.0)]
    feed(plug, lists, (0, 0, 0, box, box, box), res, cyclic=(True,) * 3, normalize=True)
    plug.pull_volume(copy=False)
    a = plug.pull_mesh(0.3, which=0)
    b = plug.pull_mesh(0.6, which=1)
    assert a["nverts"] > 0 and b["nverts"] > 0 and a["nverts"] != b["nverts"]
    pos, nrm = plug.reread_mesh(0, a["nverts"])
    assert np.array_equal(pos, a["pos"]) and np.array_equal(nrm, a["nrm"]), "module 1 overwrote the mesh module 0 had handed out"
    # the producer recomputes (new particles) and only module 1 pulls again: module 0's pointers still hold its old mesh
    xyz2 = synth.uniform_box(n, box, seed=24)
    plug.set_particles([dict(vtx=xyz2, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=1.0)], (0, 0, 0, box, box, box))
    plug.pull_volume(copy=False)
    b2 = plug.pull_mesh(0.6, which=1)
    assert b2["nverts"] != b["nverts"] or not np.array_equal(b2["pos"], b["pos"])
    pos, nrm = plug.reread_mesh(0, a["nverts"])
    assert np.array_equal(pos, a["pos"]) and np.array_equal(nrm, a["nrm"])


def _read_share(sh):
    """a CUDA consumer: import the descriptor, map, copy out, unmap"""
    import ctypes as C
    import os
    from megamol_b200 import api
    L = api.load_library()
    share = api.MmsShare(int(sh["fd"]), 0, int(sh["alloc_bytes"]), int(sh["offset"]), int(sh["bytes"]))
    p = C.c_void_p()
    assert L.mms_share_open(0, C.byref(share), C.byref(p)) == 0
    out = np.empty(share.bytes // 4, np.float32)
    rt = C.CDLL("libcudart.so.12")
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    assert rt.cudaMemcpy(out.ctypes.data, p, share.bytes, 2) == 0
    assert L.mms_share_close(0, p, C.byref(share)) == 0
    os.close(share.fd)
    return out


def test_device_resident_hand_off_through_the_modules():
    """'memoryLocation' = VRAM on the density module (MemLoc VRAM, GetData() = nullptr, no host copy) and 'deviceMesh'
    on the isosurface module (CallTriMeshData without objects): a consumer that imports the modules' device memory sees exactly the
    volume and the soup the host contract delivers."""
    n, box, res = 8000, 10.0, (40, 36, 32)
    xyz = synth.uniform_box(n, box, seed=29)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=1.0)]
    host = rb.Harness(rb.PLUG_LIB)
    feed(host, lists, (0, 0, 0, box, box, box), res, cyclic=(True,) * 3, normalize=True)
    vol, _ = host.pull_volume()
    mesh = host.pull_mesh(0.4)
    dev = rb.Harness(rb.PLUG_LIB)
    feed(dev, lists, (0, 0, 0, box, box, box), res, cyclic=(True,) * 3, normalize=True)
    dev.set_param(0, "memoryLocation", 0)  # geocalls::MemoryLocation::VRAM
    dev.set_param(1, "deviceMesh", 1)
    _, meta = dev.pull_volume(copy=False)
    assert meta["resolution"] == res
    with pytest.raises(RuntimeError):  # GetData() is nullptr for a VRAM volume (VolumetricDataCall.h:136-144)
        dev.pull_volume(copy=True)
    sv = dev.share_density()
    assert sv["memloc"] == 0
    assert np.array_equal(_read_share(sv), vol.ravel())
    m = dev.pull_mesh(0.4, copy=False)
    assert m["nverts"] == 0  # no host object
    nverts, sp, sn = dev.share_mesh()
    assert nverts == mesh["nverts"]
    assert np.array_equal(_read_share(sp), mesh["pos"].ravel())
    assert np.array_equal(_read_share(sn), mesh["nrm"].ravel())
    # back to the host contract on the same modules
    dev.set_param(0, "memoryLocation", 1)
    dev.set_param(1, "deviceMesh", 0)
    vol2, _ = dev.pull_volume()
    assert np.array_equal(vol2, vol)
    m2 = dev.pull_mesh(0.4)
    assert np.array_equal(m2["pos"], mesh["pos"])
    host.close(), dev.close()


def test_modules_on_two_devices(plug):
    """The modules' `devices` parameter: ParticlesToDensityB200 computes the volume in z-slabs on two GPUs behind one handle (mms_slabs_*),
    IsoSurfaceB200 adopts the slabs' device volumes into its own group; volume and mesh equal the single-device modules' bit for bit."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, box, res = 200_000, 20.0, (64, 56, 48)
    xyz = synth.uniform_box(n, box, seed=29)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=0.6)]
    outs = []
    for devices in ("", "0,1"):
        plug.set_param(0, "devices", devices)
        feed(plug, lists, (0, 0, 0, box, box, box), res, cyclic=(True,) * 3, normalize=True)
        vol, meta = plug.pull_volume()
        m = plug.pull_mesh(0.4)
        outs.append((vol.copy(), meta, m))
    plug.set_param(0, "devices", "")
    (v1, m1, s1), (v2, m2, s2) = outs
    assert np.array_equal(v1.view(np.uint32), v2.view(np.uint32))
    assert (m1["min"], m1["max"]) == (m2["min"], m2["max"]) and m1["resolution"] == m2["resolution"] and m1["slicedist"] == m2["slicedist"]
    assert s1["nverts"] == s2["nverts"] > 1000 and np.array_equal(s1["pos"], s2["pos"]) and np.array_equal(s1["nrm"], s2["nrm"])


def test_for_surface_reconstruction_bbox(ref, plug):
    xyz = synth.uniform_box(2000, 6.0, seed=17)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=2000, global_radius=0.7)]
    outs = []
    for h in (ref, plug):
        feed(h, lists, (0, 0, 0, 6, 5, 4), (16, 16, 16), cyclic=(False,) * 3, normalize=False, for_surface=True)
        outs.append(h.pull_volume(copy=False)[1])
    assert outs[0]["resolution"] == outs[1]["resolution"]
    assert np.allclose(outs[0]["origin"], outs[1]["origin"]) and np.allclose(outs[0]["slicedist"], outs[1]["slicedist"])


def test_quicksurf_mode_through_the_modules(plug, oracle):
    """The drop-in modules' extra `mode` parameter: QuickSurf-Gaussian density + colour volume -> coloured CallTriMeshData."""
    n = 3000
    data, _, _ = synth.protein_like(n, seed=3, nballs=5, extent=36.0)
    data = np.ascontiguousarray(data)
    lists = [dict(vtx=data, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=data.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA,
                  col_stride=32)]
    res = (46, 46, 46)
    feed(plug, lists, (0, 0, 0, 36, 36, 36), res, cyclic=(True,) * 3, normalize=True)
    plug.set_param(0, "mode", 1)
    plug.set_param(0, "quicksurf::quality", 1)
    plug.set_param(0, "quicksurf::colour", 1)
    try:
        vol, meta = plug.pull_volume()
        sd = np.array(meta["slicedist"], np.float32)
        rvol, rrgb = oracle.density_gauss(lists, meta["origin"], sd, res, radscale=1.0, gausslim=2.5, colour=True)
        assert (np.abs(vol - rvol) / np.maximum(rvol, 1e-5 * rvol.max())).max() < 2e-5
        m = plug.pull_mesh(0.5, colours=True)
        assert m["nverts"] > 1000 and m["col"] is not None
        rpos, rnrm, rcol = oracle.mc_emit(vol, meta["origin"], sd, 0.5, rgb=rrgb)
        assert m["nverts"] == 3 * rpos.shape[0]
        assert np.abs(m["pos"].reshape(-1, 3, 3) - rpos).max() <= 1e-4 * float(sd.max())
        assert np.abs(m["col"].reshape(-1, 3, 3) - rcol).max() < 2e-4
    finally:
        plug.set_param(0, "mode", 0)
        plug.set_param(0, "quicksurf::colour", 0)


def test_quicksurf_mode_on_z_chunks_through_the_modules(plug):
    """`devices` with the QuickSurf mode: the slab group (here the one device named twice = two z-chunks, which runs on a single-GPU box)
    computes the Gaussian density and its colour volume with coloured halo records, IsoSurfaceB200 adopts the slabs and hands out the
    coloured mesh -- volume, positions, normals and colours equal the single-context modules' bit for bit."""
    n = 3000
    data, _, _ = synth.protein_like(n, seed=3, nballs=5, extent=36.0)
    data = np.ascontiguousarray(data)
    lists = [dict(vtx=data, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=data.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA,
                  col_stride=32)]
    res = (46, 46, 46)
    plug.set_param(0, "mode", 1)
    plug.set_param(0, "quicksurf::quality", 1)
    plug.set_param(0, "quicksurf::colour", 1)
    outs = []
    try:
        for devices in ("", "0,0"):
            plug.set_param(0, "devices", devices)
            feed(plug, lists, (0, 0, 0, 36, 36, 36), res, cyclic=(False,) * 3, normalize=False)
            vol, _ = plug.pull_volume()
            m = plug.pull_mesh(0.5, colours=True)
            outs.append((vol.copy(), {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in m.items()}))
    finally:
        plug.set_param(0, "devices", "")
        plug.set_param(0, "mode", 0)
        plug.set_param(0, "quicksurf::colour", 0)
    (v1, s1), (v2, s2) = outs
    assert np.array_equal(v1.view(np.uint32), v2.view(np.uint32))
    assert s1["nverts"] == s2["nverts"] > 1000 and s1["col"] is not None and s2["col"] is not None
    assert np.array_equal(s1["pos"], s2["pos"]) and np.array_equal(s1["nrm"], s2["nrm"]) and np.array_equal(s1["col"], s2["col"])


def test_marching_tetrahedra_algorithm_matches_the_reference_module(ref, plug, oracle):
    """IsoSurfaceB200 with algorithm = MarchingTetrahedra: on the volume its VolumetricDataCall delivers, the mesh equals the oracle's
    restatement of the reference IsoSurface bit for bit (which itself equals the unmodified module bit for bit: tests/test_oracle_golden.py,
    tests/test_oracle_vs_reference.py); next to the reference graph (whose density differs in the last bits: summation order) the two
    soups agree triangle for triangle up to iso-value ties."""
    n, box, res = 5000, 10.0, (24, 22, 20)
    xyz = synth.uniform_box(n, box, seed=12)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=1.0)]
    bbox = (0, 0, 0, box, box, box)
    feed(plug, lists, bbox, res, cyclic=(False,) * 3, normalize=True)
    plug.set_param(1, "algorithm", 1)
    try:
        vol, _ = plug.pull_volume()
        m = plug.pull_mesh(0.5)
        pos, nrm = oracle.mt_emit(vol, bbox, 0.5)
        assert m["ntris"] == 0 and m["nverts"] == 3 * pos.shape[0] and pos.shape[0] > 5000
        assert np.array_equal(m["pos"], pos.reshape(-1, 3)) and np.array_equal(m["nrm"], nrm.reshape(-1, 3))
        feed(ref, lists, bbox, res, cyclic=(False,) * 3, normalize=True)
        ref.pull_volume()
        r = ref.pull_mesh(0.5)
        assert abs(r["nverts"] - m["nverts"]) <= 0.002 * r["nverts"]
        if r["nverts"] == m["nverts"]:
            assert np.abs(r["pos"] - m["pos"]).max() <= 1e-3 * box / res[0]
    finally:
        plug.set_param(1, "algorithm", 0)


def test_molecule_colouring_modes_of_the_reference_quicksurf(plug):
    """The reference QuickSurf module colours a MolecularDataCall with two protein_calls::ProteinColor modes blended by a weight
    (plugins/protein_cuda/src/QuickSurf.cpp:62-109, 281-320, 596-616; parameters color::coloringMode0/1, color::colorWeighting,
    color::min/mid/maxGradColor).  ParticlesToDensityB200 has the same parameters and calls the same (unmodified) ProteinColor code: for
    several mode pairs the density-weighted colour volume equals, bit for bit, the one of the equivalent FLOAT_XYZR + FLOAT_RGBA particle
    list whose colours are the reference's table for that pair."""
    n = 3000
    data, _, _ = synth.protein_like(n, seed=9, nballs=5, extent=36.0)
    radii = np.array([1.2, 1.52, 1.55, 1.7, 1.8], np.float32)
    rgb = np.array([[255, 255, 255], [255, 13, 13], [48, 80, 248], [144, 144, 144], [255, 255, 48]], np.uint8)
    tidx = (np.arange(n) * 7 % 5).astype(np.uint32)
    pos = np.ascontiguousarray(data[:, :3])
    bfac = (np.random.default_rng(3).random(n) * 80.0 + 5.0).astype(np.float32)
    bbox, res = (0, 0, 0, 36, 36, 36), (40, 40, 40)
    ELEMENT, RAINBOW, BFACTOR, CHAIN, MOLECULE, RESIDUE = 0, 2, 3, 6, 7, 8   # protein_calls::ProteinColor::ColoringMode
    grad = np.array([[0x14, 0x64, 0x96], [0xf0, 0xf0, 0xf0], [0xae, 0x3b, 0x32]], np.float32) / np.float32(255.0)  # the parameters' defaults

    mol = rb.Harness(rb.PLUG_LIB, molecule=True)
    mol.set_molecule(pos, tidx, radii, rgb, bbox)
    mol.set_molecule_structure(bfac, atoms_per_residue=9, residues_per_molecule=25, molecules_per_chain=3)

    def setup(h):
        h.set_p2d_params(res, cyclic=(False,) * 3, normalize=False)
        h.set_param(0, "mode", 1)
        h.set_param(0, "quicksurf::quality", 1)
        h.set_param(0, "quicksurf::colour", 1)

    seen = []
    try:
        for m0, m1, w in ((CHAIN, ELEMENT, 0.5), (ELEMENT, ELEMENT, 0.5), (BFACTOR, RAINBOW, 0.25), (MOLECULE, RESIDUE, 0.7), (RAINBOW, CHAIN, 1.0)):
            setup(mol)
            mol.set_param(0, "color::coloringMode0", m0)
            mol.set_param(0, "color::coloringMode1", m1)
            mol.set_param(0, "color::colorWeighting", float(w))
            mvol, _ = mol.pull_volume()
            table = mol.molecule_colour_table(n, m0, m1, w, grad)
            assert table.min() >= 0.0 and table.max() <= 1.0 + 1e-6
            seen.append(table)
            buf = np.zeros((n, 8), np.float32)
            buf[:, :3], buf[:, 3], buf[:, 4:7], buf[:, 7] = pos, radii[tidx], table, 1.0
            lists = [dict(vtx=buf, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=buf.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA, col_stride=32)]
            plug.set_particles(lists, bbox)
            setup(plug)
            pvol, _ = plug.pull_volume()
            assert np.array_equal(mvol, pvol) and mvol.max() > 0.5
            mm_, pm = mol.pull_mesh(0.5, colours=True), plug.pull_mesh(0.5, colours=True)
            assert mm_["nverts"] == pm["nverts"] > 1000
            assert np.array_equal(mm_["col"], pm["col"]), (m0, m1, w)   # the coloured mesh carries the colour volume
    finally:
        plug.set_param(0, "mode", 0)
        plug.set_param(0, "quicksurf::colour", 0)
        mol.close()
    # the modes really differ (chain / b-factor / rainbow tables are not the element colours)
    assert not np.allclose(seen[0], seen[1]) and not np.allclose(seen[2], seen[1]) and not np.allclose(seen[3], seen[4])


@pytest.mark.parametrize("cyc,norm", [(True, True), (False, False)])
def test_vector_aggregator_matches_the_reference_module(ref, plug, oracle, cyc, norm):
    """aggregator = IVecToSingleCell_Volume: the 3-component VolumetricDataCall, the grid particles on "outParticles" and the table on
    "outInfo" of ParticlesToDensityB200 next to the unmodified ParticlesToDensity (ParticlesToDensity.cpp:249-378, 629-727)."""
    n, box, res = 1500, 10.0, (40, 20, 18)
    xyz = synth.uniform_box(n, box, seed=4301)
    d = (np.stack([synth.uniform(4302, 0, n, k) for k in range(3)], 1) * 2 - 1).astype(np.float32)
    lists = [dict(vtx=xyz, vtx_type=rb.VERT_FLOAT_XYZ, count=n, global_radius=0.6, dir=d)]
    bbox = (0, 0, 0, box, box, box)
    out = []
    for h in (ref, plug):
        feed(h, lists, bbox, res, cyclic=(cyc,) * 3, normalize=norm, sigma=1.0, aggregator=2)
        vol, meta = h.pull_volume(components=3)
        out.append((vol, meta, h.pull_grid_particles(), h.pull_info()))
    (rv, rm, rg, rt), (pv, pm, pg, pt) = out
    for k in ("resolution", "components", "origin", "slicedist"):
        assert rm[k] == pm[k], k
    assert pm["components"] == 3
    assert abs(pm["min"] - rm["min"]) <= 1e-5 and abs(pm["max"] - rm["max"]) <= 1e-5 * max(1.0, abs(rm["max"]))
    _, omag, _, (omn, omx) = oracle.density_p2d_vector(lists, bbox[:3], (box,) * 3, res, (cyc,) * 3, sigma=1.0, normalize=False)
    c = dict(lists=lists, bmin=bbox[:3], bext=(box,) * 3, res=res, cyclic=(cyc,) * 3, sigma=1.0)
    tail = H.vector_tail_mask(oracle, c)
    scale = 1.0 / ((omx - omn) if norm else 1.0)
    assert (np.abs(pv.astype(np.float64) - rv) / np.maximum(np.abs(rv), scale))[~tail].max() < 1e-5
    # grid particles: same voxels, sorted by magnitude, same payload per voxel
    for k in ("lists", "count", "vtx_type", "col_type", "dir_type"):
        assert rg[k] == pg[k], k
    assert abs(rg["global_radius"] - pg["global_radius"]) < 1e-7
    ri, pi = H.voxel_index_of(rg["pos"], bbox[:3], (box,) * 3, res), H.voxel_index_of(pg["pos"], bbox[:3], (box,) * 3, res)
    assert np.array_equal(np.sort(ri), np.sort(pi))
    assert np.all(np.diff(pg["col"]) <= 0)
    ro, po = np.argsort(ri), np.argsort(pi)
    assert np.array_equal(rg["pos"][ro], pg["pos"][po])
    keep = ~tail.ravel()[ri[ro]]
    assert np.abs(rg["col"][ro] - pg["col"][po])[keep].max() < 1e-5
    solid = keep & (omag.ravel()[ri[ro]] > 1e-3)
    assert np.abs(rg["dir"][ro] - pg["dir"][po])[solid].max() < 1e-4
    # table
    assert rt["names"] == pt["names"] and rt["columns"] == pt["columns"] == 7 and rt["rows"] == pt["rows"] == pg["count"]
    assert np.allclose(rt["ranges"], pt["ranges"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(pt["data"][:, :3], pg["pos"]) and np.array_equal(pt["data"][:, 3:6], pg["dir"])
    assert np.abs(rt["data"][ro][:, 6] - pt["data"][po][:, 6])[keep].max() < 1e-5 * max(1.0, omx)
    # back to a scalar aggregator: no grid particles, empty table
    feed(plug, lists, bbox, res, cyclic=(cyc,) * 3, normalize=norm, aggregator=0)
    plug.pull_volume()
    assert plug.pull_grid_particles()["lists"] == 0 and plug.pull_info()["rows"] == 0


def test_molecular_data_call_input(plug, oracle):
    """ParticlesToDensityB200.inData connected to a protein_calls::MolecularDataCall (the second input of the reference's QuickSurf
    module, QuickSurf.cpp:326-404): one sphere per atom, radius and colour of its atom type -> same volume, same coloured mesh as
    the equivalent FLOAT_XYZR + FLOAT_RGBA particle list through the MultiParticleDataCall path, bit for bit."""
    n = 2500
    data, _, _ = synth.protein_like(n, seed=5, nballs=5, extent=36.0)
    radii = np.array([1.2, 1.52, 1.55, 1.7, 1.8], np.float32)
    rgb = np.array([[255, 255, 255], [255, 13, 13], [48, 80, 248], [144, 144, 144], [255, 255, 48]], np.uint8)
    tidx = (np.arange(n) * 7 % 5).astype(np.uint32)
    pos = np.ascontiguousarray(data[:, :3])
    buf = np.zeros((n, 8), np.float32)
    buf[:, :3] = pos
    buf[:, 3] = radii[tidx]
    buf[:, 4:7] = rgb[tidx].astype(np.float32) / np.float32(255.0)
    buf[:, 7] = 1.0
    lists = [dict(vtx=buf, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=buf.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA, col_stride=32)]
    bbox, res = (0, 0, 0, 36, 36, 36), (46, 44, 42)

    def pull(h):
        h.set_p2d_params(res, cyclic=(False,) * 3, normalize=False)
        h.set_param(0, "mode", 1)
        h.set_param(0, "quicksurf::quality", 1)
        h.set_param(0, "quicksurf::colour", 1)
        vol, meta = h.pull_volume()
        return vol, meta, h.pull_mesh(0.5, colours=True)

    mol = rb.Harness(rb.PLUG_LIB, molecule=True)
    mol.set_molecule(pos, tidx, radii, rgb, bbox)
    mvol, mmeta, mmesh = pull(mol)
    plug.set_particles(lists, bbox)
    try:
        pvol, pmeta, pmesh = pull(plug)
    finally:
        plug.set_param(0, "mode", 0)
        plug.set_param(0, "quicksurf::colour", 0)
    assert mmeta["resolution"] == pmeta["resolution"] and mmeta["origin"] == pmeta["origin"]
    assert np.array_equal(mvol, pvol) and mvol.max() > 0.5
    assert mmesh["nverts"] == pmesh["nverts"] > 1000
    for k in ("pos", "nrm", "col"):
        assert np.array_equal(mmesh[k], pmesh[k]), k
    # and it is the right volume: the oracle's Gaussian density of those spheres
    sd = np.array(mmeta["slicedist"], np.float32)
    rvol, _ = oracle.density_gauss(lists, mmeta["origin"], sd, res, radscale=1.0, gausslim=2.5, colour=False)
    assert (np.abs(mvol - rvol) / np.maximum(rvol, 1e-5 * rvol.max())).max() < 2e-5
    # bump mode through the same input
    mol.set_param(0, "mode", 0)
    mol.set_p2d_params(res, cyclic=(False,) * 3, normalize=False, sigma=1.0)
    bvol, _ = mol.pull_volume()
    rb_vol, _ = oracle.density_p2d(lists, bbox[:3], (36.0,) * 3, res, (0, 0, 0), sigma=1.0)
    assert H.density_close(bvol, rb_vol) < H.DENSITY_RTOL
    mol.close()


def test_quicksurf_own_grid_setup(plug, oracle):
    """quicksurf::gridSpacing > 0: the module sets the grid up like QuickSurf::calculateSurface (QuickSurf.cpp:456-480): bounding box
    padded by max(1.5 radscale r_max, 0.4 sqrt(4/3 pi pad^3)), ceil(extent / spacing) voxels of exactly that spacing."""
    n = 2000
    data, _, _ = synth.protein_like(n, seed=9, nballs=4, extent=30.0)
    data = np.ascontiguousarray(data)
    lists = [dict(vtx=data, vtx_type=rb.VERT_FLOAT_XYZR, vtx_stride=32, count=n, col=data.ctypes.data + 16, col_type=rb.COL_FLOAT_RGBA,
                  col_stride=32)]
    bbox = (1.0, 2.0, 3.0, 31.0, 30.0, 29.0)
    spacing, radscale = np.float32(0.8), np.float32(1.25)
    feed(plug, lists, bbox, (16, 16, 16), cyclic=(False,) * 3, normalize=False)
    plug.set_param(0, "mode", 1)
    plug.set_param(0, "quicksurf::quality", 0)
    plug.set_param(0, "quicksurf::radiusScale", float(radscale))
    plug.set_param(0, "quicksurf::gridSpacing", float(spacing))
    try:
        rmax = np.float32(data[:, 3].max())
        pad = np.float32(radscale * rmax * np.float32(1.5))
        padrad = np.float32(0.4 * np.sqrt(4.0 / 3.0 * float(np.float32(np.pi)) * float(pad) ** 3))
        pad = max(pad, padrad)
        lo = np.array(bbox[:3], np.float32) - pad
        hi = np.array(bbox[3:], np.float32) + pad
        res = tuple(int(np.ceil(np.float32(hi[a] - lo[a]) / spacing)) for a in range(3))
        plug.res = res  # the harness sizes its copy buffer from this
        vol, meta = plug.pull_volume()
        assert meta["resolution"] == res
        assert np.allclose(meta["origin"], lo, rtol=0, atol=1e-5)
        assert np.allclose(meta["slicedist"], [spacing] * 3, rtol=1e-6)
        rvol, _ = oracle.density_gauss(lists, meta["origin"], np.array(meta["slicedist"], np.float32), res, radscale=float(radscale), gausslim=2.0)
        assert (np.abs(vol - rvol) / np.maximum(rvol, 1e-5 * rvol.max())).max() < 2e-5
        m = plug.pull_mesh(0.5)
        assert m["nverts"] > 1000
        assert m["pos"].min(0).min() >= lo.min() - 1e-3 and (m["pos"].max(0) <= hi + spacing).all()
    finally:
        plug.set_param(0, "mode", 0)
        plug.set_param(0, "quicksurf::gridSpacing", 0.0)
        plug.set_param(0, "quicksurf::radiusScale", 1.0)
        plug.set_param(0, "quicksurf::quality", 2)
