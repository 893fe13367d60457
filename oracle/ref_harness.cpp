// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Drives MegaMol modules through the reference's real Module / CallerSlot / CalleeSlot / Call
// machinery, the way MegaMolGraph::add_module / add_call do (core/src/MegaMolGraph.cpp:567-790),
// and exposes the result over a small C interface that the pytest suite and bench.py load with
// ctypes.  It is compiled twice by oracle/Makefile.ref:
//   * oracle/_ref/libmmref.so       P2D = datatools::ParticlesToDensity (plugins/datatools/src/
//                                   ParticlesToDensity.cpp), ISO = trisoup_gl::volumetrics::IsoSurface
//                                   (plugins/trisoup_gl/src/volumetrics/IsoSurface.cpp): the UNMODIFIED
//                                   reference translation units == the parity anchor and CPU baseline.
//   * oracle/_ref/libmmplug.so      -DMMH_B200: the same graph with our drop-in modules
//                                   (plugin/b200surf/src) in place of the reference ones.
// The file includes reference headers but copies no reference code.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <omp.h>
#include <unistd.h>

#include "datatools/table/TableDataCall.h"
#include "geometry_calls/MultiParticleDataCall.h"
#include "geometry_calls/VolumetricDataCall.h"
#include "geometry_calls_gl/CallTriMeshDataGL.h"
#include "mmcore/Call.h"
#include "mmcore/CalleeSlot.h"
#include "mmcore/CallerSlot.h"
#include "mmcore/Module.h"
#include "mmcore/RootModuleNamespace.h"
#include "mmcore/factories/CallAutoDescription.h"
#include "mmcore/param/BoolParam.h"
#include "mmcore/param/EnumParam.h"
#include "mmcore/param/FloatParam.h"
#include "mmcore/param/IntParam.h"
#include "mmcore/param/StringParam.h"
#include "mmcore/param/ParamSlot.h"
#include "mmcore/utility/log/Log.h"
#include "protein_calls/MolecularDataCall.h"
#include "protein_calls/ProteinColor.h"
#include "trisoup/volumetrics/MarchingCubeTables.h"

#ifdef MMH_B200
#include "IsoSurfaceB200.h"
#include "ParticlesToDensityB200.h"
using P2DModule = megamol::b200surf::ParticlesToDensityB200;
using IsoModule = megamol::b200surf::IsoSurfaceB200;
#else
#include "ParticlesToDensity.h"
#include "volumetrics/IsoSurface.h"
using P2DModule = megamol::datatools::ParticlesToDensity;
using IsoModule = megamol::trisoup_gl::volumetrics::IsoSurface;
#endif

using namespace megamol;

extern "C" {
/** One particle list as a MultiParticleDataCall carries it (enums numerically equal to
 *  SimpleSphericalParticles::VertexDataType / ColourDataType, SimpleSphericalParticles.h:27-50). */
struct mmh_list {
    const void* vtx;
    const void* col;
    uint64_t count;
    int32_t vtx_type;
    uint32_t vtx_stride;
    int32_t col_type;
    uint32_t col_stride;
    float global_radius;
    uint8_t global_rgba[4];
    float irange[2];
};
}

namespace {

/** Answers MultiParticleDataCall GetData/GetExtent from caller-supplied arrays (stands in for MMPLDDataSource). */
class ParticleSource : public core::Module {
public:
    ParticleSource() : outSlot("outData", "particles") {
        outSlot.SetCallback(geocalls::MultiParticleDataCall::ClassName(), "GetData", &ParticleSource::getData);
        outSlot.SetCallback(geocalls::MultiParticleDataCall::ClassName(), "GetExtent", &ParticleSource::getExtent);
        MakeSlotAvailable(&outSlot);
    }
    ~ParticleSource() override { Release(); }
    std::vector<mmh_list> lists;
    std::vector<const void*> dirs;      // per list: DIRDATA_FLOAT_XYZ pointer or nullptr (DIRDATA_NONE)
    std::vector<unsigned> dirStrides;
    float bbox[6] = {0, 0, 0, 1, 1, 1};
    unsigned frameCount = 1;
    unsigned frameID = 0;
    size_t hash = 1;

protected:
    bool create() override { return true; }
    void release() override {}

private:
    bool getExtent(core::Call& c) {
        auto* m = dynamic_cast<geocalls::MultiParticleDataCall*>(&c);
        if (!m) return false;
        m->SetFrameCount(frameCount);
        m->AccessBoundingBoxes().Clear();
        m->AccessBoundingBoxes().SetObjectSpaceBBox(bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]);
        m->AccessBoundingBoxes().SetObjectSpaceClipBox(bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]);
        m->SetFrameID(frameID);
        m->SetDataHash(hash);
        m->SetUnlocker(nullptr);
        return true;
    }
    bool getData(core::Call& c) {
        auto* m = dynamic_cast<geocalls::MultiParticleDataCall*>(&c);
        if (!m) return false;
        m->SetFrameID(frameID);
        m->SetDataHash(hash);
        m->SetParticleListCount(static_cast<unsigned>(lists.size()));
        for (size_t i = 0; i < lists.size(); ++i) {
            auto& p = m->AccessParticles(static_cast<unsigned>(i));
            const auto& l = lists[i];
            p.SetCount(l.count);
            p.SetGlobalRadius(l.global_radius);
            p.SetGlobalColour(l.global_rgba[0], l.global_rgba[1], l.global_rgba[2], l.global_rgba[3]);
            p.SetColourMapIndexValues(l.irange[0], l.irange[1]);
            p.SetVertexData(static_cast<geocalls::SimpleSphericalParticles::VertexDataType>(l.vtx_type), l.vtx,
                l.vtx_stride);
            p.SetColourData(static_cast<geocalls::SimpleSphericalParticles::ColourDataType>(l.col_type), l.col,
                l.col_stride);
            if (i < dirs.size() && dirs[i] != nullptr)
                p.SetDirData(geocalls::SimpleSphericalParticles::DIRDATA_FLOAT_XYZ, dirs[i], dirStrides[i]);
        }
        m->SetUnlocker(nullptr);
        return true;
    }
    core::CalleeSlot outSlot;
};

/** Answers MolecularDataCall GetData/GetExtent (stands in for PDBLoader): atoms with per-type radius and colour. */
class MoleculeSource : public core::Module {
public:
    MoleculeSource() : outSlot("outData", "molecule") {
        using MDC = protein_calls::MolecularDataCall;
        outSlot.SetCallback(MDC::ClassName(), MDC::FunctionName(MDC::CallForGetData), &MoleculeSource::getData);
        outSlot.SetCallback(MDC::ClassName(), MDC::FunctionName(MDC::CallForGetExtent), &MoleculeSource::getExtent);
        MakeSlotAvailable(&outSlot);
    }
    ~MoleculeSource() override { Release(); }
    std::vector<float> pos;
    std::vector<unsigned> typeIdx;
    std::vector<protein_calls::MolecularDataCall::AtomType> types;
    std::vector<float> bfactor; // optional (B-factor colouring)
    // optional synthetic structure (chain / molecule / residue colouring): equal-sized residues, molecules, chains
    std::vector<protein_calls::MolecularDataCall::Residue> residues;
    std::vector<const protein_calls::MolecularDataCall::Residue*> residuePtrs;
    std::vector<protein_calls::MolecularDataCall::Molecule> molecules;
    std::vector<protein_calls::MolecularDataCall::Chain> chains;
    std::vector<vislib::StringA> residueTypeNames;
    float bbox[6] = {0, 0, 0, 1, 1, 1};
    size_t hash = 1;

    /** What getData puts into a call (also used to fill a local call for the expected colour table). */
    void fill(protein_calls::MolecularDataCall& m) {
        m.SetAtoms(static_cast<unsigned>(typeIdx.size()), static_cast<unsigned>(types.size()), typeIdx.data(), pos.data(), types.data(),
            nullptr, bfactor.empty() ? nullptr : bfactor.data(), nullptr, nullptr);
        if (!bfactor.empty()) m.SetBFactorRange(*std::min_element(bfactor.begin(), bfactor.end()), *std::max_element(bfactor.begin(), bfactor.end()));
        if (!chains.empty()) {
            m.SetResidueTypeNames(static_cast<unsigned>(residueTypeNames.size()), residueTypeNames.data());
            m.SetResidues(static_cast<unsigned>(residuePtrs.size()), residuePtrs.data());
            m.SetMolecules(static_cast<unsigned>(molecules.size()), molecules.data());
            m.SetChains(static_cast<unsigned>(chains.size()), chains.data());
        }
    }

protected:
    bool create() override { return true; }
    void release() override {}

private:
    bool getExtent(core::Call& c) {
        auto* m = dynamic_cast<protein_calls::MolecularDataCall*>(&c);
        if (!m) return false;
        m->SetFrameCount(1);
        m->AccessBoundingBoxes().Clear();
        m->AccessBoundingBoxes().SetObjectSpaceBBox(bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]);
        m->AccessBoundingBoxes().SetObjectSpaceClipBox(bbox[0], bbox[1], bbox[2], bbox[3], bbox[4], bbox[5]);
        m->SetDataHash(hash);
        m->SetUnlocker(nullptr);
        return true;
    }
    bool getData(core::Call& c) {
        auto* m = dynamic_cast<protein_calls::MolecularDataCall*>(&c);
        if (!m) return false;
        m->SetDataHash(hash);
        fill(*m);
        m->SetUnlocker(nullptr);
        return true;
    }
    core::CalleeSlot outSlot;
};

/** The consumer end: what a renderer would be. */
class Sink : public core::Module {
public:
    Sink() : volSlot("inVolume", "volume"), meshSlot("inMesh", "mesh"), mesh2Slot("inMesh2", "mesh of a second isosurface module"),
             gridSlot("inGrid", "grid particles"), infoSlot("inInfo", "table") {
        gridSlot.SetCompatibleCall<geocalls::MultiParticleDataCallDescription>();
        MakeSlotAvailable(&gridSlot);
        infoSlot.SetCompatibleCall<datatools::table::TableDataCallDescription>();
        MakeSlotAvailable(&infoSlot);
        volSlot.SetCompatibleCall<geocalls::VolumetricDataCallDescription>();
        MakeSlotAvailable(&volSlot);
        meshSlot.SetCompatibleCall<geocalls_gl::CallTriMeshDataGLDescription>();
        MakeSlotAvailable(&meshSlot);
        mesh2Slot.SetCompatibleCall<geocalls_gl::CallTriMeshDataGLDescription>();
        MakeSlotAvailable(&mesh2Slot);
    }
    ~Sink() override { Release(); }
    core::CallerSlot volSlot, meshSlot, mesh2Slot, gridSlot, infoSlot;

protected:
    bool create() override { return true; }
    void release() override {}
};

template<class Desc>
bool connect(core::Module& from, const char* fromSlot, core::Module& to, const char* toSlot,
    std::vector<std::unique_ptr<core::Call>>& keep) {
    auto desc = std::make_shared<Desc>();
    auto* caller = dynamic_cast<core::CallerSlot*>(from.FindSlot(fromSlot));
    auto* callee = dynamic_cast<core::CalleeSlot*>(to.FindSlot(toSlot));
    if (!caller || !callee) return false;
    core::Call* call = desc->CreateCall();
    keep.emplace_back(call);
    if (!callee->ConnectCall(call, desc)) return false;
    return caller->ConnectCall(call);
}

struct Harness {
    std::shared_ptr<core::RootModuleNamespace> root = std::make_shared<core::RootModuleNamespace>();
    std::shared_ptr<ParticleSource> src = std::make_shared<ParticleSource>();
    std::shared_ptr<MoleculeSource> mol = std::make_shared<MoleculeSource>();
    std::shared_ptr<P2DModule> p2d = std::make_shared<P2DModule>();
    std::shared_ptr<IsoModule> iso = std::make_shared<IsoModule>();
    std::shared_ptr<IsoModule> iso2 = std::make_shared<IsoModule>(); // a second consumer of the same density module
    std::shared_ptr<Sink> sink = std::make_shared<Sink>();
    std::vector<std::unique_ptr<core::Call>> calls;
    const geocalls_gl::CallTriMeshDataGL::Mesh* mesh = nullptr;
    const geocalls_gl::CallTriMeshDataGL::Mesh* meshOf[2] = {nullptr, nullptr};
    bool ok = false;

    /** molecule: feed the density module from the MolecularDataCall source instead of the particle source. */
    explicit Harness(bool molecule = false) {
        core::utility::log::Log::DefaultLog.SetLevel(core::utility::log::Log::log_level::error);
        core::utility::log::Log::DefaultLog.SetEchoLevel(core::utility::log::Log::log_level::error);
        src->setName("src");
        p2d->setName("p2d");
        iso->setName("iso");
        iso2->setName("iso2");
        sink->setName("sink");
        mol->setName("mol");
        root->AddChild(mol);
        root->AddChild(src);
        root->AddChild(p2d);
        root->AddChild(iso);
        root->AddChild(iso2);
        root->AddChild(sink);
        ok = src->Create() && mol->Create() && p2d->Create() && iso->Create() && iso2->Create() && sink->Create();
        if (molecule) ok = ok && connect<protein_calls::MolecularDataCallDescription>(*p2d, "inData", *mol, "outData", calls);
        else ok = ok && connect<geocalls::MultiParticleDataCallDescription>(*p2d, "inData", *src, "outData", calls);
        ok = ok && connect<geocalls::VolumetricDataCallDescription>(*iso, "inData", *p2d, "outData", calls);
        ok = ok && connect<geocalls::VolumetricDataCallDescription>(*sink, "inVolume", *p2d, "outData", calls);
        ok = ok && connect<geocalls_gl::CallTriMeshDataGLDescription>(*sink, "inMesh", *iso, "outData", calls);
        ok = ok && connect<geocalls::VolumetricDataCallDescription>(*iso2, "inData", *p2d, "outData", calls);
        ok = ok && connect<geocalls_gl::CallTriMeshDataGLDescription>(*sink, "inMesh2", *iso2, "outData", calls);
        ok = ok && connect<geocalls::MultiParticleDataCallDescription>(*sink, "inGrid", *p2d, "outParticles", calls);
        ok = ok && connect<datatools::table::TableDataCallDescription>(*sink, "inInfo", *p2d, "outInfo", calls);
    }
    ~Harness() {
        // callers first, so that no slot is left pointing at a destroyed call
        sink->volSlot.ConnectCall(nullptr);
        sink->meshSlot.ConnectCall(nullptr);
        sink->mesh2Slot.ConnectCall(nullptr);
        sink->gridSlot.ConnectCall(nullptr);
        sink->infoSlot.ConnectCall(nullptr);
    }
};

template<class P, class V>
bool setParam(core::Module& m, const char* name, V v) {
    auto* s = dynamic_cast<core::param::ParamSlot*>(m.FindSlot(name));
    if (!s) return false;
    auto* p = s->Param<P>();
    if (!p) return false;
    p->SetValue(v);
    return true;
}

double nowMs() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

extern "C" {

/** 0 = reference modules, 1 = B200 drop-in modules. */
int mmh_flavour() {
#ifdef MMH_B200
    return 1;
#else
    return 0;
#endif
}

int mmh_threads() { return omp_get_max_threads(); }
void mmh_set_threads(int n) { omp_set_num_threads(n); }

void* mmh_create() {
    auto* h = new Harness();
    if (!h->ok) {
        delete h;
        return nullptr;
    }
    return h;
}

/** The same graph fed by a MolecularDataCall source.  NULL where the density module's inData does not accept that call
 *  (the reference ParticlesToDensity: MultiParticleDataCall only, ParticlesToDensity.cpp:149-150). */
void* mmh_create_molecule() {
    auto* h = new Harness(true);
    if (!h->ok) {
        delete h;
        return nullptr;
    }
    return h;
}

void mmh_destroy(void* hv) { delete static_cast<Harness*>(hv); }

/** Atoms of the molecule source: positions (3 floats per atom), type index per atom, per type a radius and an RGB byte triple. */
int mmh_set_molecule(void* hv, unsigned natoms, const float* pos, const unsigned* type_idx, unsigned ntypes, const float* radii,
    const uint8_t* rgb, const float bbox[6]) {
    auto* h = static_cast<Harness*>(hv);
    auto& m = *h->mol;
    m.pos.assign(pos, pos + 3 * static_cast<size_t>(natoms));
    m.typeIdx.assign(type_idx, type_idx + natoms);
    m.types.clear();
    for (unsigned t = 0; t < ntypes; ++t)
        m.types.emplace_back(vislib::StringA("X"), radii[t], rgb[3 * t], rgb[3 * t + 1], rgb[3 * t + 2]);
    std::memcpy(m.bbox, bbox, sizeof(float) * 6);
    ++m.hash;
    return 0;
}

/** Optional extras of the molecule source: B-factors (NULL: none) and a synthetic structure of equal-sized residues / molecules / chains
 *  (atoms_per_residue == 0: none).  Call after mmh_set_molecule. */
int mmh_set_molecule_structure(void* hv, const float* bfactor, unsigned atoms_per_residue, unsigned residues_per_molecule, unsigned molecules_per_chain) {
    auto* h = static_cast<Harness*>(hv);
    auto& m = *h->mol;
    using MDC = protein_calls::MolecularDataCall;
    const unsigned n = static_cast<unsigned>(m.typeIdx.size());
    m.bfactor.clear();
    if (bfactor) m.bfactor.assign(bfactor, bfactor + n);
    m.residues.clear(), m.residuePtrs.clear(), m.molecules.clear(), m.chains.clear();
    m.residueTypeNames = {vislib::StringA("ALA"), vislib::StringA("GLY"), vislib::StringA("SOL"), vislib::StringA("LYS"), vislib::StringA("TOL")};
    if (atoms_per_residue && residues_per_molecule && molecules_per_chain && n) {
        const unsigned nres = (n + atoms_per_residue - 1) / atoms_per_residue;
        const unsigned nmol = (nres + residues_per_molecule - 1) / residues_per_molecule;
        const unsigned nchain = (nmol + molecules_per_chain - 1) / molecules_per_chain;
        m.residues.reserve(nres);
        for (unsigned r = 0; r < nres; ++r) {
            const unsigned first = r * atoms_per_residue, cnt = std::min(atoms_per_residue, n - first);
            m.residues.emplace_back(first, cnt, vislib::math::Cuboid<float>(0, 0, 0, 1, 1, 1), r % 5u, static_cast<int>(r / residues_per_molecule), r);
        }
        for (auto& r : m.residues) m.residuePtrs.push_back(&r);
        for (unsigned k = 0; k < nmol; ++k) {
            const unsigned first = k * residues_per_molecule, cnt = std::min(residues_per_molecule, nres - first);
            m.molecules.emplace_back(first, cnt, static_cast<int>(k / molecules_per_chain));
        }
        for (unsigned c = 0; c < nchain; ++c) {
            const unsigned first = c * molecules_per_chain, cnt = std::min(molecules_per_chain, nmol - first);
            m.chains.emplace_back(first, cnt, static_cast<char>('A' + c % 26));
        }
    }
    ++m.hash;
    return 0;
}

/** The colour table the UNMODIFIED reference code (protein_calls::ProteinColor::MakeWeightedColorTable, what QuickSurf.cpp:596-616 calls)
 *  makes for the molecule source's atoms: the expectation the drop-in module's colours are checked against.  grad = min, mid, max gradient
 *  colours (3 x RGB); out = 3 floats per atom. */
int mmh_molecule_colour_table(void* hv, int mode0, int mode1, float weight, const float grad[9], float* out) {
    auto* h = static_cast<Harness*>(hv);
    using protein_calls::ProteinColor;
    protein_calls::MolecularDataCall call;
    h->mol->fill(call);
    std::vector<glm::vec3> table, fileTable, rainbow;
    ProteinColor::ReadColorTableFromFile(std::string("colors.txt"), fileTable);
    ProteinColor::MakeRainbowColorTable(100, rainbow);
    const std::vector<glm::vec3> lookup = {glm::make_vec3(grad), glm::make_vec3(grad + 3), glm::make_vec3(grad + 6)};
    ProteinColor::MakeWeightedColorTable(call, static_cast<ProteinColor::ColoringMode>(mode0), static_cast<ProteinColor::ColoringMode>(mode1), weight,
        1.0 - weight, table, lookup, fileTable, rainbow, nullptr, nullptr, true);
    for (size_t i = 0; i < table.size(); ++i) out[3 * i] = table[i].r, out[3 * i + 1] = table[i].g, out[3 * i + 2] = table[i].b;
    return static_cast<int>(table.size());
}

/** Replaces the source's particle lists; bumps the data hash so that consumers recompute. */
int mmh_set_particles(void* hv, int nlists, const mmh_list* lists, const float bbox[6], unsigned frame_id) {
    auto* h = static_cast<Harness*>(hv);
    h->src->lists.assign(lists, lists + nlists);
    h->src->dirs.assign(nlists, nullptr);
    h->src->dirStrides.assign(nlists, 0u);
    std::memcpy(h->src->bbox, bbox, sizeof(float) * 6);
    h->src->frameID = frame_id;
    h->src->frameCount = frame_id + 1;
    ++h->src->hash;
    return 0;
}

/** Direction data (DIRDATA_FLOAT_XYZ) of list `list` of the last mmh_set_particles; consumed by aggregator 2. */
int mmh_set_directions(void* hv, int list, const void* dir, unsigned stride) {
    auto* h = static_cast<Harness*>(hv);
    if (list < 0 || static_cast<size_t>(list) >= h->src->dirs.size()) return -1;
    h->src->dirs[list] = dir;
    h->src->dirStrides[list] = stride;
    ++h->src->hash;
    return 0;
}

/** Parameters of ParticlesToDensity by their slot names (ParticlesToDensity.cpp:59-152). */
int mmh_set_p2d_params(void* hv, int aggregator, int sx, int sy, int sz, int cyclx, int cycly, int cyclz,
    int normalize, float sigma, int for_surface) {
    auto* h = static_cast<Harness*>(hv);
    using namespace core::param;
    bool ok = setParam<EnumParam>(*h->p2d, "aggregator", aggregator);
    ok = ok && setParam<IntParam>(*h->p2d, "sizex", sx) && setParam<IntParam>(*h->p2d, "sizey", sy) &&
         setParam<IntParam>(*h->p2d, "sizez", sz);
    ok = ok && setParam<BoolParam>(*h->p2d, "cyclX", cyclx != 0) && setParam<BoolParam>(*h->p2d, "cyclY", cycly != 0) &&
         setParam<BoolParam>(*h->p2d, "cyclZ", cyclz != 0);
    ok = ok && setParam<BoolParam>(*h->p2d, "normalize", normalize != 0) && setParam<FloatParam>(*h->p2d, "sigma", sigma);
    ok = ok && setParam<BoolParam>(*h->p2d, "forSurfaceReconstruction", for_surface != 0);
    return ok ? 0 : -1;
}

/** Generic parameter setters (for the extra parameters of the B200 modules). */
int mmh_set_param_int(void* hv, int module, const char* name, int v) {
    auto* h = static_cast<Harness*>(hv);
    core::Module& m = module == 0 ? static_cast<core::Module&>(*h->p2d) : static_cast<core::Module&>(*h->iso);
    if (setParam<core::param::IntParam>(m, name, v)) return 0;
    if (setParam<core::param::EnumParam>(m, name, v)) return 0;
    if (setParam<core::param::BoolParam>(m, name, v != 0)) return 0;
    return -1;
}
int mmh_set_param_string(void* hv, int module, const char* name, const char* v) {
    auto* h = static_cast<Harness*>(hv);
    core::Module& m = module == 0 ? static_cast<core::Module&>(*h->p2d) : static_cast<core::Module&>(*h->iso);
    auto* s = dynamic_cast<core::param::ParamSlot*>(m.FindSlot(name));
    if (!s) return -1;
    auto* p = s->Param<core::param::StringParam>();
    if (!p) return -1;
    p->SetValue(v);
    return 0;
}
int mmh_set_param_float(void* hv, int module, const char* name, float v) {
    auto* h = static_cast<Harness*>(hv);
    core::Module& m = module == 0 ? static_cast<core::Module&>(*h->p2d) : static_cast<core::Module&>(*h->iso);
    return setParam<core::param::FloatParam>(m, name, v) ? 0 : -1;
}

/**
 * Pulls the volume like a consumer would: VolumetricDataCall GetExtents(0), GetMetadata(2), GetData(1)
 * (the order IsoSurface.cpp:114-117 uses).  out_vol may be NULL (timing only).
 * info[0..2] = resolution, info[3] = components, info[4] = data hash; minmax = metadata Min/MaxValues[0].
 */
int mmh_pull_volume(void* hv, unsigned frame_id, float* out_vol, uint64_t info[5], double minmax[2],
    float origin[3], float slicedist[3], double* ms) {
    auto* h = static_cast<Harness*>(hv);
    auto* v = h->sink->volSlot.CallAs<geocalls::VolumetricDataCall>();
    if (!v) return -1;
    v->SetFrameID(frame_id, true);
    const double t0 = nowMs();
    if (!(*v)(geocalls::VolumetricDataCall::IDX_GET_EXTENTS)) return -2;
    if (!(*v)(geocalls::VolumetricDataCall::IDX_GET_METADATA)) return -3;
    if (!(*v)(geocalls::VolumetricDataCall::IDX_GET_DATA)) return -4;
    if (ms) *ms = nowMs() - t0;
    const auto* md = v->GetMetadata();
    if (!md) return -5;
    for (int i = 0; i < 3; ++i) {
        info[i] = md->Resolution[i];
        origin[i] = md->Origin[i];
        slicedist[i] = md->SliceDists[i] ? md->SliceDists[i][0] : 0.0f;
    }
    info[3] = md->Components;
    info[4] = v->DataHash();
    minmax[0] = md->MinValues ? md->MinValues[0] : 0.0;
    minmax[1] = md->MaxValues ? md->MaxValues[0] : 0.0;
    const void* data = v->GetData();
    if (out_vol) {
        if (!data) return -6;
        std::memcpy(out_vol, data, sizeof(float) * md->Resolution[0] * md->Resolution[1] * md->Resolution[2] * md->Components);
    }
    return 0;
}

/**
 * Pulls "outParticles" like a sphere/arrow renderer would: MultiParticleDataCall GetExtent(1) then GetData(0).
 * info = {list count, particle count of list 0, vertex type, colour type, direction type, data hash}; NULL outputs are skipped,
 * pos/dir receive 3 floats per particle, col one float (COLDATA_FLOAT_I).
 */
int mmh_pull_grid_particles(void* hv, unsigned frame_id, uint64_t info[6], float* global_radius, float* pos, float* dir, float* col) {
    auto* h = static_cast<Harness*>(hv);
    auto* g = h->sink->gridSlot.CallAs<geocalls::MultiParticleDataCall>();
    if (!g) return -1;
    g->SetFrameID(frame_id, true);
    if (!(*g)(1)) return -2;
    if (!(*g)(0)) return -3;
    for (int i = 0; i < 6; ++i) info[i] = 0;
    info[0] = g->GetParticleListCount();
    info[5] = g->DataHash();
    if (info[0] == 0) return 0;
    const auto& p = g->AccessParticles(0);
    info[1] = p.GetCount();
    info[2] = p.GetVertexDataType();
    info[3] = p.GetColourDataType();
    info[4] = p.GetDirDataType();
    if (global_radius) *global_radius = p.GetGlobalRadius();
    const size_t n = p.GetCount();
    auto gather = [n](float* dst, const void* src, unsigned stride, int comps) {
        for (size_t i = 0; i < n; ++i) std::memcpy(dst + i * comps, static_cast<const char*>(src) + i * stride, sizeof(float) * comps);
    };
    if (n > 0) {
        if (pos && p.GetVertexData()) gather(pos, p.GetVertexData(), p.GetVertexDataStride() ? p.GetVertexDataStride() : 12, 3);
        if (dir && p.GetDirData()) gather(dir, p.GetDirData(), p.GetDirDataStride() ? p.GetDirDataStride() : 12, 3);
        if (col && p.GetColourData()) gather(col, p.GetColourData(), p.GetColourDataStride() ? p.GetColourDataStride() : 4, 1);
    }
    return 0;
}

/**
 * Pulls "outInfo": TableDataCall GetHash(1) then GetData(0).  dims = {columns, rows, data hash}; data (rows x columns floats, row
 * major), names (columns x 32 chars) and ranges (columns x {min, max}) may be NULL.
 */
int mmh_pull_info(void* hv, uint64_t dims[3], float* data, char* names, float* ranges) {
    auto* h = static_cast<Harness*>(hv);
    auto* t = h->sink->infoSlot.CallAs<datatools::table::TableDataCall>();
    if (!t) return -1;
    if (!(*t)(1)) return -2;
    if (!(*t)(0)) return -3;
    dims[0] = t->GetColumnsCount();
    dims[1] = t->GetRowsCount();
    dims[2] = t->DataHash();
    if (data && t->GetData() && dims[0] * dims[1] > 0) std::memcpy(data, t->GetData(), sizeof(float) * dims[0] * dims[1]);
    for (size_t c = 0; c < dims[0]; ++c) {
        const auto& ci = t->GetColumnsInfos()[c];
        if (names) std::snprintf(names + 32 * c, 32, "%s", ci.Name().c_str());
        if (ranges) ranges[2 * c] = ci.MinimumValue(), ranges[2 * c + 1] = ci.MaximumValue();
    }
    return 0;
}

/** Pulls the mesh like TriSoupRenderer would: CallTriMeshData GetExtent(1) then GetData(0). */
static int pullMeshOf(Harness* h, int which, unsigned frame_id, float isoval, uint64_t* nverts, uint64_t* ntris, double* ms) {
    auto& module = which == 0 ? h->iso : h->iso2;
    if (!setParam<core::param::FloatParam>(*module, "isoval", isoval)) return -1;
    auto* t = (which == 0 ? h->sink->meshSlot : h->sink->mesh2Slot).CallAs<geocalls_gl::CallTriMeshDataGL>();
    if (!t) return -2;
    t->SetFrameID(frame_id, true);
    const double t0 = nowMs();
    if (!(*t)(1)) return -3;
    if (!(*t)(0)) return -4;
    if (ms) *ms = nowMs() - t0;
    h->meshOf[which] = nullptr;
    *nverts = 0;
    *ntris = 0;
    if (t->Count() >= 1 && t->Objects()) {
        h->meshOf[which] = &t->Objects()[0];
        *nverts = h->meshOf[which]->GetVertexCount();
        *ntris = h->meshOf[which]->GetTriCount();
    }
    h->mesh = h->meshOf[which];
    return 0;
}

int mmh_pull_mesh(void* hv, unsigned frame_id, float isoval, uint64_t* nverts, uint64_t* ntris, double* ms) {
    return pullMeshOf(static_cast<Harness*>(hv), 0, frame_id, isoval, nverts, ntris, ms);
}
/** The same through the SECOND isosurface module connected to the same density module (its own isoval). */
int mmh_pull_mesh2(void* hv, unsigned frame_id, float isoval, uint64_t* nverts, uint64_t* ntris, double* ms) {
    return pullMeshOf(static_cast<Harness*>(hv), 1, frame_id, isoval, nverts, ntris, ms);
}
/** Which mesh mmh_copy_mesh reads: the one last handed out by isosurface module 0 or 1 (pointers as the module left them). */
int mmh_select_mesh(void* hv, int which) {
    auto* h = static_cast<Harness*>(hv);
    if (which < 0 || which > 1 || !h->meshOf[which]) return -1;
    h->mesh = h->meshOf[which];
    return 0;
}

/** Copies the last pulled mesh (float positions / normals / colours, 3 per vertex); NULL outputs are skipped. */
int mmh_copy_mesh(void* hv, float* pos, float* nrm, float* col) {
    auto* h = static_cast<Harness*>(hv);
    if (!h->mesh) return -1;
    using Mesh = geocalls_gl::CallTriMeshDataGL::Mesh;
    const size_t n = h->mesh->GetVertexCount();
    if (pos) {
        if (h->mesh->GetVertexDataType() != Mesh::DT_FLOAT) return -2;
        std::memcpy(pos, h->mesh->GetVertexPointerFloat(), n * 3 * sizeof(float));
    }
    if (nrm) {
        if (h->mesh->GetNormalDataType() != Mesh::DT_FLOAT) return -3;
        std::memcpy(nrm, h->mesh->GetNormalPointerFloat(), n * 3 * sizeof(float));
    }
    if (col) {
        if (h->mesh->GetColourDataType() != Mesh::DT_FLOAT) return -4;
        std::memcpy(col, h->mesh->GetColourPointerFloat(), n * 3 * sizeof(float));
    }
    return 0;
}

/** Triangle indices of the last pulled mesh (3 x uint32 per triangle; an unindexed soup has 0 triangles, IsoSurface.cpp:171-181). */
int mmh_copy_indices(void* hv, unsigned* idx) {
    auto* h = static_cast<Harness*>(hv);
    if (!h->mesh) return -1;
    using Mesh = geocalls_gl::CallTriMeshDataGL::Mesh;
    const size_t n = h->mesh->GetTriCount();
    if (n == 0) return 0;
    if (h->mesh->GetTriDataType() != Mesh::DT_UINT32) return -2;
    std::memcpy(idx, h->mesh->GetTriIndexPointerUInt32(), n * 3 * sizeof(unsigned));
    return 0;
}

#ifdef MMH_B200
/** Device-resident hand-off of the B200 modules: out[0..4] = volume share {fd, alloc_bytes, offset, bytes, MemLoc of the metadata}. */
int mmh_share_density(void* hv, int64_t out[5]) {
    auto* h = static_cast<Harness*>(hv);
    mms_share v{}, rgb{};
    if (!h->p2d->ShareDensity(&v, &rgb)) return -1;
    if (rgb.fd >= 0) close(rgb.fd);
    auto* call = h->sink->volSlot.CallAs<geocalls::VolumetricDataCall>();
    out[0] = v.fd, out[1] = static_cast<int64_t>(v.alloc_bytes), out[2] = static_cast<int64_t>(v.offset), out[3] = static_cast<int64_t>(v.bytes);
    out[4] = (call && call->GetMetadata()) ? static_cast<int64_t>(call->GetMetadata()->MemLoc) : -1;
    return 0;
}
/** out = {nverts, then for positions and normals: fd, alloc_bytes, offset, bytes}. */
int mmh_share_mesh(void* hv, int which, int64_t out[9]) {
    auto* h = static_cast<Harness*>(hv);
    mms_share p{}, n{}, c{};
    uint64_t nv = 0;
    if (!(which ? h->iso2 : h->iso)->ShareMesh(&nv, &p, &n, &c)) return -1;
    if (c.fd >= 0) close(c.fd);
    out[0] = static_cast<int64_t>(nv);
    out[1] = p.fd, out[2] = static_cast<int64_t>(p.alloc_bytes), out[3] = static_cast<int64_t>(p.offset), out[4] = static_cast<int64_t>(p.bytes);
    out[5] = n.fd, out[6] = static_cast<int64_t>(n.alloc_bytes), out[7] = static_cast<int64_t>(n.offset), out[8] = static_cast<int64_t>(n.bytes);
    return 0;
}
#endif

/** The reference's marching-cubes tables (plugins/trisoup/src/volumetrics/MarchingCubeTables.cpp:11-285). */
void mmh_mc_tables(int32_t tri[256 * 16], uint8_t count[256], uint32_t edgeflags[256], uint32_t vertoff[8 * 3],
    uint32_t edgeconn[12 * 2]) {
    using T = trisoup::volumetrics::MarchingCubeTables;
    for (int i = 0; i < 256; ++i) {
        for (int j = 0; j < 16; ++j) tri[i * 16 + j] = T::a2iTriangleConnectionTable[i][j];
        count[i] = T::a2ucTriangleConnectionCount[i];
        edgeflags[i] = T::aiCubeEdgeFlags[i];
    }
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 3; ++j) vertoff[i * 3 + j] = T::a2fVertexOffset[i][j];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 2; ++j) edgeconn[i * 2 + j] = T::a2iEdgeConnection[i][j];
}

} // extern "C"
