/*
 * ParticlesToDensityB200.cpp
 *
 * Host side of the B200 density path.  Mirrors the call protocol of datatools::ParticlesToDensity
 * (plugins/datatools/src/ParticlesToDensity.cpp:163-383) and hands the particle lists of the incoming
 * MultiParticleDataCall to libmmsurf as raw (pointer, type, stride) triples -- no per-particle virtual accessors.
 */
#include "ParticlesToDensityB200.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <numeric>
#include <vector>

#include "datatools/table/TableDataCall.h"
#include "mmcore/param/BoolParam.h"
#include "mmcore/param/ColorParam.h"
#include "mmcore/param/FilePathParam.h"
#include "mmcore/param/EnumParam.h"
#include "mmcore/param/FloatParam.h"
#include "mmcore/param/IntParam.h"
#include "mmcore/param/StringParam.h"
#include "mmcore/utility/log/Log.h"

using namespace megamol;
using namespace megamol::b200surf;
using megamol::core::utility::log::Log;
using geocalls::MultiParticleDataCall;
using geocalls::VolumetricDataCall;

bool ParticlesToDensityB200::IsAvailable() {
    mms_ctx* probe = nullptr;
    mms_config cfg{0, 0};
    if (mms_create(&probe, &cfg) != MMS_OK)
        return false;
    mms_destroy(probe);
    return true;
}

ParticlesToDensityB200::ParticlesToDensityB200()
        : aggregatorSlot("aggregator", "algorithm for the aggregation")
        , xResSlot("sizex", "The size of the volume in numbers of voxels")
        , yResSlot("sizey", "The size of the volume in numbers of voxels")
        , zResSlot("sizez", "The size of the volume in numbers of voxels")
        , cyclXSlot("cyclX", "Considers cyclic boundary conditions in X direction")
        , cyclYSlot("cyclY", "Considers cyclic boundary conditions in Y direction")
        , cyclZSlot("cyclZ", "Considers cyclic boundary conditions in Z direction")
        , normalizeSlot("normalize", "Normalize the output volume")
        , sigmaSlot("sigma", "Sigma for Gauss in multiple of rad")
        , surfaceSlot("forSurfaceReconstruction", "Set true if this volume is used for surface reconstruction")
        , deviceSlot("device", "CUDA device ordinal the volume is computed on")
        , devicesSlot("devices", "Several CUDA devices, comma separated (e.g. 0,1,2,3): the volume is computed in z-slabs with halo, one per "
                                 "device (position aggregator of the bump mode); empty = the single device of 'device'")
        , memLocSlot("memoryLocation", "RAM: the volume is copied to host memory (the contract of VolumetricDataCall::GetData()); VRAM: it "
                                       "stays in importable device memory (MemLoc = VRAM, GetData() = nullptr, see ShareDensity())")
        , modeSlot("mode", "Density semantics: ParticlesToDensity bump kernel or QuickSurf Gaussian")
        , qsQualitySlot("quicksurf::quality", "Quality: 0 low .. 3 ultra (Gaussian cut-off 2.0/2.5/3.0/4.0 sigma)")
        , qsRadScaleSlot("quicksurf::radiusScale", "Radius scale")
        , qsColourSlot("quicksurf::colour", "Also build the density-weighted colour volume (coloured isosurface)")
        , colorTableFileSlot("color::colorTableFilename", "Path to the file containing a custom color table")
        , coloringMode0Slot("color::coloringMode0", "The first coloring mode")
        , coloringMode1Slot("color::coloringMode1", "The second coloring mode")
        , coloringModeWeightSlot("color::colorWeighting", "Weighting factor between the two coloring modes")
        , minGradColorSlot("color::minGradColor", "The color for the minimum value for gradient coloring")
        , midGradColorSlot("color::midGradColor", "The color for the middle value for gradient coloring")
        , maxGradColorSlot("color::maxGradColor", "The color for the maximum value for gradient coloring")
        , qsRefCellsSlot("quicksurf::referenceCandidates",
              "Sum over the candidate set of protein_cuda's QuickSurf (all atoms of the acceleration cells around a voxel's 8^3 block, no radial "
              "cut-off): the reference's density up to fp32 summation order, at ~15x the cost of the radial cut-off")
        , qsGridSpacingSlot("quicksurf::gridSpacing", "Grid spacing of QuickSurf's own grid set-up (0 = use sizex/sizey/sizez on the bounding box)")
        , outDataSlot("outData", "Provides a density volume for the particles")
        , outParticlesSlot("outParticles", "Provides the particles in grid form (vector aggregator only)")
        , outInfoSlot("outInfo", "Provides information about the grid (vector aggregator only)")
        , inDataSlot("inData", "takes the particle data") {

    auto* ep = new core::param::EnumParam(0);
    ep->SetTypePair(0, "PosToSingleCell_Volume");
    ep->SetTypePair(1, "IColToSingleCell_Volume");
    ep->SetTypePair(2, "IVecToSingleCell_Volume");
    this->aggregatorSlot << ep;
    this->MakeSlotAvailable(&this->aggregatorSlot);

    using VDC = VolumetricDataCall;
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_GET_DATA), &ParticlesToDensityB200::getDataCallback);
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_GET_EXTENTS), &ParticlesToDensityB200::getExtentCallback);
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_GET_METADATA), &ParticlesToDensityB200::getMetadataCallback);
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_START_ASYNC), &ParticlesToDensityB200::dummyCallback);
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_STOP_ASYNC), &ParticlesToDensityB200::dummyCallback);
    this->outDataSlot.SetCallback(VDC::ClassName(), VDC::FunctionName(VDC::IDX_TRY_GET_DATA), &ParticlesToDensityB200::dummyCallback);
    this->MakeSlotAvailable(&this->outDataSlot);

    this->outParticlesSlot.SetCallback(MultiParticleDataCall::ClassName(), MultiParticleDataCall::FunctionName(0), &ParticlesToDensityB200::getDataCallback);
    this->outParticlesSlot.SetCallback(MultiParticleDataCall::ClassName(), MultiParticleDataCall::FunctionName(1), &ParticlesToDensityB200::getExtentCallback);
    this->MakeSlotAvailable(&this->outParticlesSlot);

    using TDC = datatools::table::TableDataCall;
    this->outInfoSlot.SetCallback(TDC::ClassName(), TDC::FunctionName(0), &ParticlesToDensityB200::getDataCallback);
    this->outInfoSlot.SetCallback(TDC::ClassName(), TDC::FunctionName(1), &ParticlesToDensityB200::getExtentCallback);
    this->MakeSlotAvailable(&this->outInfoSlot);

    this->xResSlot << new core::param::IntParam(16);
    this->MakeSlotAvailable(&this->xResSlot);
    this->yResSlot << new core::param::IntParam(16);
    this->MakeSlotAvailable(&this->yResSlot);
    this->zResSlot << new core::param::IntParam(16);
    this->MakeSlotAvailable(&this->zResSlot);

    this->cyclXSlot << new core::param::BoolParam(true);
    this->MakeSlotAvailable(&this->cyclXSlot);
    this->cyclYSlot << new core::param::BoolParam(true);
    this->MakeSlotAvailable(&this->cyclYSlot);
    this->cyclZSlot << new core::param::BoolParam(true);
    this->MakeSlotAvailable(&this->cyclZSlot);

    this->normalizeSlot << new core::param::BoolParam(true);
    this->MakeSlotAvailable(&this->normalizeSlot);

    this->sigmaSlot << new core::param::FloatParam(1.0f, FLT_MIN);
    this->MakeSlotAvailable(&this->sigmaSlot);

    this->surfaceSlot << new core::param::BoolParam(false);
    this->MakeSlotAvailable(&this->surfaceSlot);

    this->deviceSlot << new core::param::IntParam(0, 0);
    this->MakeSlotAvailable(&this->deviceSlot);
    this->devicesSlot << new core::param::StringParam("");
    this->MakeSlotAvailable(&this->devicesSlot);
    auto* ml = new core::param::EnumParam(static_cast<int>(geocalls::MemoryLocation::RAM));
    ml->SetTypePair(static_cast<int>(geocalls::MemoryLocation::RAM), "RAM");
    ml->SetTypePair(static_cast<int>(geocalls::MemoryLocation::VRAM), "VRAM");
    this->memLocSlot << ml;
    this->MakeSlotAvailable(&this->memLocSlot);

    auto* mp = new core::param::EnumParam(0);
    mp->SetTypePair(0, "ParticlesToDensity_Bump");
    mp->SetTypePair(1, "QuickSurf_Gaussian");
    this->modeSlot << mp;
    this->MakeSlotAvailable(&this->modeSlot);
    this->qsQualitySlot << new core::param::IntParam(2, 0, 3);
    this->MakeSlotAvailable(&this->qsQualitySlot);
    this->qsRadScaleSlot << new core::param::FloatParam(1.0f, 0.0f);
    this->MakeSlotAvailable(&this->qsRadScaleSlot);
    this->qsColourSlot << new core::param::BoolParam(false);
    this->MakeSlotAvailable(&this->qsColourSlot);

    // colouring of a molecule input: names, types and defaults of the reference QuickSurf module (QuickSurf.cpp:62-109)
    using protein_calls::ProteinColor;
    std::string filename("colors.txt");
    ProteinColor::ReadColorTableFromFile(filename, this->fileColorTable); // (no such file: the built-in table, ProteinColor.cpp:64-75)
    this->colorTableFileSlot.SetParameter(
        new core::param::FilePathParam(filename, core::param::FilePathParam::FilePathFlags_::Flag_File_ToBeCreated));
    this->MakeSlotAvailable(&this->colorTableFileSlot);
    auto* cm0 = new core::param::EnumParam(static_cast<int>(ProteinColor::ColoringMode::CHAIN));
    auto* cm1 = new core::param::EnumParam(static_cast<int>(ProteinColor::ColoringMode::ELEMENT));
    for (int cCnt = 0; cCnt < static_cast<int>(ProteinColor::ColoringMode::MODE_COUNT); ++cCnt) {
        const auto name = ProteinColor::GetName(static_cast<ProteinColor::ColoringMode>(cCnt));
        cm0->SetTypePair(cCnt, name.c_str());
        cm1->SetTypePair(cCnt, name.c_str());
    }
    this->coloringMode0Slot << cm0;
    this->coloringMode1Slot << cm1;
    this->MakeSlotAvailable(&this->coloringMode0Slot);
    this->MakeSlotAvailable(&this->coloringMode1Slot);
    this->coloringModeWeightSlot.SetParameter(new core::param::FloatParam(0.5f, 0.0f, 1.0f));
    this->MakeSlotAvailable(&this->coloringModeWeightSlot);
    this->minGradColorSlot.SetParameter(new core::param::ColorParam("#146496"));
    this->MakeSlotAvailable(&this->minGradColorSlot);
    this->midGradColorSlot.SetParameter(new core::param::ColorParam("#f0f0f0"));
    this->MakeSlotAvailable(&this->midGradColorSlot);
    this->maxGradColorSlot.SetParameter(new core::param::ColorParam("#ae3b32"));
    this->MakeSlotAvailable(&this->maxGradColorSlot);
    ProteinColor::MakeRainbowColorTable(100, this->rainbowColorTable);
    this->qsRefCellsSlot << new core::param::BoolParam(false);
    this->MakeSlotAvailable(&this->qsRefCellsSlot);
    this->qsGridSpacingSlot << new core::param::FloatParam(0.0f, 0.0f);
    this->MakeSlotAvailable(&this->qsGridSpacingSlot);

    this->inDataSlot.SetCompatibleCall<geocalls::MultiParticleDataCallDescription>();
    this->inDataSlot.SetCompatibleCall<protein_calls::MolecularDataCallDescription>();
    this->MakeSlotAvailable(&this->inDataSlot);
}

ParticlesToDensityB200::~ParticlesToDensityB200() {
    this->Release();
}

bool ParticlesToDensityB200::create() {
    return true; // the context is created lazily on the device selected by the "device" parameter
}

void ParticlesToDensityB200::release() {
    if (this->ctx != nullptr) {
        mms_destroy(this->ctx);
        this->ctx = nullptr;
    }
    if (this->group != nullptr) {
        mms_slabs_destroy(this->group);
        this->group = nullptr;
        this->groupActive = false;
    }
    this->metadata.MinValues = nullptr;
    this->metadata.MaxValues = nullptr;
    for (auto& s : this->metadata.SliceDists)
        s = nullptr;
}

bool ParticlesToDensityB200::dummyCallback(core::Call&) {
    return true;
}

bool ParticlesToDensityB200::anythingDirty() const {
    return this->aggregatorSlot.IsDirty() || this->xResSlot.IsDirty() || this->yResSlot.IsDirty() || this->zResSlot.IsDirty() ||
           this->cyclXSlot.IsDirty() || this->cyclYSlot.IsDirty() || this->cyclZSlot.IsDirty() || this->normalizeSlot.IsDirty() ||
           this->sigmaSlot.IsDirty() || this->deviceSlot.IsDirty() || this->devicesSlot.IsDirty() || this->memLocSlot.IsDirty() || this->modeSlot.IsDirty() || this->qsQualitySlot.IsDirty() ||
           this->qsRadScaleSlot.IsDirty() || this->qsColourSlot.IsDirty() || this->qsGridSpacingSlot.IsDirty() || this->colorTableFileSlot.IsDirty() ||
           this->coloringMode0Slot.IsDirty() || this->coloringMode1Slot.IsDirty() || this->coloringModeWeightSlot.IsDirty() ||
           this->minGradColorSlot.IsDirty() || this->midGradColorSlot.IsDirty() || this->maxGradColorSlot.IsDirty() ||
           this->qsRefCellsSlot.IsDirty();
}

void ParticlesToDensityB200::resetDirty() {
    for (auto* s : {&aggregatorSlot, &xResSlot, &yResSlot, &zResSlot, &cyclXSlot, &cyclYSlot, &cyclZSlot, &normalizeSlot, &sigmaSlot, &deviceSlot, &devicesSlot, &memLocSlot, &modeSlot,
             &qsQualitySlot, &qsRadScaleSlot, &qsColourSlot, &qsGridSpacingSlot, &qsRefCellsSlot, &coloringMode0Slot, &coloringMode1Slot,
             &coloringModeWeightSlot, &minGradColorSlot, &midGradColorSlot, &maxGradColorSlot})
        s->ResetDirty();
}

bool ParticlesToDensityB200::getExtentCallback(core::Call& c) {
    auto* out = dynamic_cast<VolumetricDataCall*>(&c);
    auto* outGrid = dynamic_cast<MultiParticleDataCall*>(&c);
    auto* outInfo = dynamic_cast<datatools::table::TableDataCall*>(&c);
    auto* in = this->inDataSlot.CallAs<core::AbstractGetData3DCall>(); // MultiParticleDataCall or MolecularDataCall: 0 = data, 1 = extent
    if (in == nullptr)
        return false;
    const unsigned int frameID = out != nullptr ? out->FrameID() : (outGrid != nullptr ? outGrid->FrameID() : 0);
    in->SetFrameID(frameID, true);
    if (!(*in)(1)) {
        Log::DefaultLog.WriteError("ParticlesToDensityB200: could not get current frame extents (%u)", frameID);
        return false;
    }
    core::AbstractGetData3DCall* o3 = out != nullptr ? static_cast<core::AbstractGetData3DCall*>(out) : static_cast<core::AbstractGetData3DCall*>(outGrid);
    if (o3 != nullptr) {
        o3->AccessBoundingBoxes().SetObjectSpaceBBox(in->GetBoundingBoxes().ObjectSpaceBBox());
        o3->AccessBoundingBoxes().SetObjectSpaceClipBox(in->GetBoundingBoxes().ObjectSpaceClipBox());
        o3->AccessBoundingBoxes().MakeScaledWorld(1.0f);
        o3->SetFrameCount(in->FrameCount());
    }
    if (outInfo != nullptr) {
        outInfo->SetDataHash(this->datahash);
        outInfo->SetUnlocker(nullptr);
        outInfo->SetFrameCount(in->FrameCount());
    }
    return true;
}

bool ParticlesToDensityB200::getMetadataCallback(core::Call& c) {
    // the reference answers GET_METADATA with its extent callback and only fills the metadata in GET_DATA
    // (ParticlesToDensity.cpp:100-102,249-293); we additionally publish what is already known
    if (!this->getExtentCallback(c))
        return false;
    auto* out = dynamic_cast<VolumetricDataCall*>(&c);
    auto* in = this->inDataSlot.CallAs<core::AbstractGetData3DCall>();
    if (out != nullptr && in != nullptr) {
        this->fillMetadata(in);
        out->SetMetadata(&this->metadata);
    }
    return true;
}

void ParticlesToDensityB200::fillMetadata(core::AbstractGetData3DCall* in) {
    auto& md = this->metadata;
    md.Components = this->isVector ? 3 : 1;
    md.GridType = geocalls::GridType_t::CARTESIAN;
    md.Resolution[0] = static_cast<size_t>(this->xResSlot.Param<core::param::IntParam>()->Value());
    md.Resolution[1] = static_cast<size_t>(this->yResSlot.Param<core::param::IntParam>()->Value());
    md.Resolution[2] = static_cast<size_t>(this->zResSlot.Param<core::param::IntParam>()->Value());
    if (this->ownGrid && this->has_data)
        for (int a = 0; a < 3; ++a)
            md.Resolution[a] = static_cast<size_t>(this->gridUsed.res[a]);
    md.ScalarType = geocalls::ScalarType_t::FLOATING_POINT;
    md.ScalarLength = sizeof(float);
    for (int k = 0; k < 3; ++k) { // the vector volume reports the range of the magnitudes for all three components (:262-272)
        this->minValue[k] = this->minDens;
        this->maxValue[k] = this->maxDens;
    }
    md.MinValues = this->minValue; // owned by the module, allocated once (the reference leaks a new[] per call)
    md.MaxValues = this->maxValue;
    const auto bbox = in->AccessBoundingBoxes().ObjectSpaceBBox();
    md.Extents[0] = bbox.Width();
    md.Extents[1] = bbox.Height();
    md.Extents[2] = bbox.Depth();
    md.NumberOfFrames = 1;
    for (int a = 0; a < 3; ++a) {
        this->sliceDists[a] = md.Extents[a] / static_cast<float>(md.Resolution[a] - 1);
        md.SliceDists[a] = &this->sliceDists[a];
        md.IsUniform[a] = true;
    }
    md.Origin[0] = bbox.Left();
    md.Origin[1] = bbox.Bottom();
    md.Origin[2] = bbox.Back();
    if (this->ownGrid && this->has_data) { // QuickSurf's own grid: padded origin, spacing = quicksurf::gridSpacing
        for (int a = 0; a < 3; ++a) {
            md.Extents[a] = this->gridUsed.extent[a];
            md.Origin[a] = this->gridUsed.min[a];
            this->sliceDists[a] = md.Extents[a] / static_cast<float>(md.Resolution[a] - 1);
        }
    }
    md.MemLoc = this->volumeOnDevice ? geocalls::MemoryLocation::VRAM : geocalls::MemoryLocation::RAM;
}

bool ParticlesToDensityB200::ShareDensity(mms_share* volume, mms_share* rgb) {
    if (!this->has_data || !this->volumeOnDevice || this->ctx == nullptr || this->groupActive)
        return false;
    if (mms_share_density(this->ctx, volume, rgb) != MMS_OK) {
        Log::DefaultLog.WriteError("ParticlesToDensityB200: %s", mms_last_error(this->ctx));
        return false;
    }
    return true;
}

void ParticlesToDensityB200::surfaceBBox(core::AbstractGetData3DCall* in) {
    // "forSurfaceReconstruction" (ParticlesToDensity.cpp:749-799): grow the box by 10 % plus a two-voxel margin and make
    // the voxels cubic by CHANGING sizex / sizey; the modified box only lives in the incoming call for this request.
    const int sz = this->zResSlot.Param<core::param::IntParam>()->Value();
    auto bb = in->AccessBoundingBoxes().ObjectSpaceBBox();
    const float scale = 1.1f;
    float w = bb.Width(), hgt = bb.Height(), d = bb.Depth();
    float spacing = (d * scale) / sz;
    const float newDepth = (d * scale) + 2 * spacing;
    spacing = newDepth / sz;
    auto fit = [&](float range, core::param::ParamSlot& slot) {
        float n = (range * scale) + 2 * spacing;
        int res = static_cast<int>(n / spacing);
        const float rest = n / spacing - static_cast<float>(res);
        n += (1 - rest) * spacing;
        res += 1;
        slot.Param<core::param::IntParam>()->SetValue(res);
        return n;
    };
    const float newWidth = fit(w, this->xResSlot);
    const float newHeight = fit(hgt, this->yResSlot);
    const float l = bb.Left() - (newWidth - w) / 2, b = bb.Bottom() - (newHeight - hgt) / 2, k = bb.Back() - (newDepth - d) / 2;
    const float r = bb.Right() + (newWidth - w) / 2, t = bb.Top() + (newHeight - hgt) / 2, f = bb.Front() + (newDepth - d) / 2;
    in->AccessBoundingBoxes().SetObjectSpaceBBox(l, b, k, r, t, f);
    in->AccessBoundingBoxes().SetObjectSpaceClipBox(l, b, k, r, t, f);
}

bool ParticlesToDensityB200::computeVolume(core::AbstractGetData3DCall* in) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    const int device = this->deviceSlot.Param<core::param::IntParam>()->Value();
    if (this->ctx == nullptr || this->ctxDevice != device) {
        if (this->ctx != nullptr)
            mms_destroy(this->ctx);
        this->ctx = nullptr;
        mms_config cfg{device, 0};
        if (mms_create(&this->ctx, &cfg) != MMS_OK) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: %s", mms_last_error(nullptr));
            return false;
        }
        this->ctxDevice = device;
        this->volumeOnDevice = false;
    }
    const bool wantDevice =
        this->memLocSlot.Param<core::param::EnumParam>()->Value() == static_cast<int>(geocalls::MemoryLocation::VRAM);
    if (wantDevice != this->volumeOnDevice) {
        if (mms_share_enable(this->ctx, wantDevice ? 1 : 0) != MMS_OK) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: %s", mms_last_error(this->ctx));
            return false;
        }
        this->volumeOnDevice = wantDevice;
    }
    const auto bbox = in->AccessBoundingBoxes().ObjectSpaceBBox();
    mms_grid grid{};
    grid.min[0] = bbox.Left(), grid.min[1] = bbox.Bottom(), grid.min[2] = bbox.Back();
    grid.extent[0] = bbox.Width(), grid.extent[1] = bbox.Height(), grid.extent[2] = bbox.Depth();
    grid.res[0] = this->xResSlot.Param<core::param::IntParam>()->Value();
    grid.res[1] = this->yResSlot.Param<core::param::IntParam>()->Value();
    grid.res[2] = this->zResSlot.Param<core::param::IntParam>()->Value();
    grid.cyclic[0] = this->cyclXSlot.Param<core::param::BoolParam>()->Value();
    grid.cyclic[1] = this->cyclYSlot.Param<core::param::BoolParam>()->Value();
    grid.cyclic[2] = this->cyclZSlot.Param<core::param::BoolParam>()->Value();
    mms_params p{};
    p.mode = this->modeSlot.Param<core::param::EnumParam>()->Value() == 1 ? MMS_MODE_QS_GAUSS : MMS_MODE_P2D_BUMP;
    p.aggregator = this->aggregatorSlot.Param<core::param::EnumParam>()->Value();
    p.normalize = p.mode == MMS_MODE_P2D_BUMP && this->normalizeSlot.Param<core::param::BoolParam>()->Value();
    p.sigma = this->sigmaSlot.Param<core::param::FloatParam>()->Value();
    static const float kGaussLim[4] = {2.0f, 2.5f, 3.0f, 4.0f}; // QuickSurf.cpp:580-587
    p.radscale = this->qsRadScaleSlot.Param<core::param::FloatParam>()->Value();
    p.gausslim = kGaussLim[this->qsQualitySlot.Param<core::param::IntParam>()->Value() & 3];
    p.colour = p.mode == MMS_MODE_QS_GAUSS && this->qsColourSlot.Param<core::param::BoolParam>()->Value();
    const bool gaussian = p.mode == MMS_MODE_QS_GAUSS;
    if (gaussian && this->qsRefCellsSlot.Param<core::param::BoolParam>()->Value())
        p.mode = MMS_MODE_QS_GAUSS_REFCELLS;
    if (gaussian)
        grid.cyclic[0] = grid.cyclic[1] = grid.cyclic[2] = 0; // QuickSurf has no periodic images

    std::vector<mms_list> lists;
    std::vector<const void*> dirs;
    std::vector<uint32_t> dirStrides;
    size_t total = 0;
    auto* mpdc = dynamic_cast<MultiParticleDataCall*>(in);
    if (auto* mol = dynamic_cast<protein_calls::MolecularDataCall*>(in)) {
        // one FLOAT_XYZR list with interleaved FLOAT_RGBA colours, what QuickSurf::calculateSurface(MolecularDataCall&) assembles
        // (QuickSurf.cpp:369-386)
        const size_t n = mol->AtomCount();
        if (n > 0 && (mol->AtomTypeCount() == 0 || mol->AtomPositions() == nullptr || mol->AtomTypeIndices() == nullptr)) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: MolecularDataCall without atom types or positions");
            return false;
        }
        // the atoms' colours: the reference module's colour table (QuickSurf.cpp:596-616 -> ProteinColor::MakeWeightedColorTable, the
        // UNMODIFIED protein_calls code): two colouring modes blended by color::colorWeighting.  A mode that walks the molecule's
        // structure (chains / molecules / residues) on a call that carries none would index empty arrays in the reference; here it falls
        // back to the element colours.
        using protein_calls::ProteinColor;
        if (this->colorTableFileSlot.IsDirty()) {
            ProteinColor::ReadColorTableFromFile(this->colorTableFileSlot.Param<core::param::FilePathParam>()->Value(), this->fileColorTable);
            this->colorTableFileSlot.ResetDirty();
        }
        const bool hasStructure = mol->ResidueCount() > 0 && mol->MoleculeCount() > 0 && mol->ChainCount() > 0 && mol->Residues() != nullptr &&
                                  mol->ResidueTypeNameCount() > 0;
        auto usable = [&](int m) {
            const auto mode = static_cast<ProteinColor::ColoringMode>(m);
            switch (mode) {
            case ProteinColor::ColoringMode::ELEMENT:
            case ProteinColor::ColoringMode::RAINBOW:
            case ProteinColor::ColoringMode::HEIGHTMAP_COLOR:
            case ProteinColor::ColoringMode::HEIGHTMAP_VALUE: return mode;
            case ProteinColor::ColoringMode::BFACTOR: return mol->AtomBFactors() ? mode : ProteinColor::ColoringMode::ELEMENT;
            case ProteinColor::ColoringMode::CHARGE: return mol->AtomCharges() ? mode : ProteinColor::ColoringMode::ELEMENT;
            case ProteinColor::ColoringMode::OCCUPANCY: return mol->AtomOccupancies() ? mode : ProteinColor::ColoringMode::ELEMENT;
            case ProteinColor::ColoringMode::BINDINGSITE:
            case ProteinColor::ColoringMode::PER_ATOM_FLOAT: return ProteinColor::ColoringMode::ELEMENT; // need calls this module does not have
            default: return hasStructure ? mode : ProteinColor::ColoringMode::ELEMENT;
            }
        };
        const auto mode0 = usable(this->coloringMode0Slot.Param<core::param::EnumParam>()->Value());
        const auto mode1 = usable(this->coloringMode1Slot.Param<core::param::EnumParam>()->Value());
        const std::vector<glm::vec3> smallColorTable = {glm::make_vec3(this->minGradColorSlot.Param<core::param::ColorParam>()->Value().data()),
            glm::make_vec3(this->midGradColorSlot.Param<core::param::ColorParam>()->Value().data()),
            glm::make_vec3(this->maxGradColorSlot.Param<core::param::ColorParam>()->Value().data())};
        const float weight = this->coloringModeWeightSlot.Param<core::param::FloatParam>()->Value();
        if (n > 0)
            ProteinColor::MakeWeightedColorTable(*mol, mode0, mode1, weight, 1.0 - weight, this->atomColorTable, smallColorTable, this->fileColorTable,
                this->rainbowColorTable, nullptr, nullptr, true);
        this->atoms.resize(8 * n);
        for (size_t i = 0; i < n; ++i) {
            const auto& type = mol->AtomTypes()[mol->AtomTypeIndices()[i]];
            float* a = &this->atoms[8 * i];
            a[0] = mol->AtomPositions()[3 * i + 0], a[1] = mol->AtomPositions()[3 * i + 1], a[2] = mol->AtomPositions()[3 * i + 2];
            a[3] = type.Radius();
            const glm::vec3 col = this->atomColorTable[i];
            a[4] = col.r, a[5] = col.g, a[6] = col.b, a[7] = 1.0f;
        }
        if (n > 0) {
            mms_list l{};
            l.vtx = this->atoms.data();
            l.vtx_type = MMS_VERT_FLOAT_XYZR;
            l.vtx_stride = 32;
            l.col = this->atoms.data() + 4;
            l.col_type = MMS_COL_FLOAT_RGBA;
            l.col_stride = 32;
            l.count = n;
            l.irange[1] = 1.0f;
            total = n;
            lists.push_back(l);
            dirs.push_back(nullptr);
            dirStrides.push_back(0u);
        }
    }
    for (unsigned int i = 0; mpdc != nullptr && i < mpdc->GetParticleListCount(); ++i) {
        const auto& parts = mpdc->AccessParticles(i);
        if (parts.GetVertexDataType() == MultiParticleDataCall::Particles::VERTDATA_NONE)
            continue;
        mms_list l{};
        l.vtx = parts.GetVertexData();
        l.vtx_type = static_cast<int32_t>(parts.GetVertexDataType());
        l.vtx_stride = parts.GetVertexDataStride();
        l.col = parts.GetColourData();
        l.col_type = static_cast<int32_t>(parts.GetColourDataType());
        l.col_stride = parts.GetColourDataStride();
        l.count = parts.GetCount();
        l.global_radius = parts.GetGlobalRadius();
        for (int k = 0; k < 4; ++k)
            l.global_rgba[k] = parts.GetGlobalColour()[k];
        l.irange[0] = parts.GetMinColourIndexValue();
        l.irange[1] = parts.GetMaxColourIndexValue();
        total += l.count;
        lists.push_back(l);
        const bool hasDir = parts.GetDirDataType() == geocalls::SimpleSphericalParticles::DIRDATA_FLOAT_XYZ;
        dirs.push_back(hasDir ? parts.GetDirData() : nullptr);
        dirStrides.push_back(hasDir ? parts.GetDirDataStride() : 0u);
    }
    auto fail = [&](const char* what) {
        Log::DefaultLog.WriteError("ParticlesToDensityB200: %s: %s", what, mms_last_error(this->ctx));
        return false;
    };
    // ---- several devices: z-slabs behind one handle (mms_slabs_*), same host contract ---------------------------------------------------
    std::vector<int32_t> devs;
    {
        const std::string text = this->devicesSlot.Param<core::param::StringParam>()->Value();
        size_t pos = 0;
        while (pos < text.size()) {
            size_t end = text.find(',', pos);
            if (end == std::string::npos) end = text.size();
            const std::string tok = text.substr(pos, end - pos);
            if (tok.find_first_of("0123456789") != std::string::npos) devs.push_back(std::atoi(tok.c_str()));
            pos = end + 1;
        }
    }
    // A volume of 2^32 voxels or more does not fit one context's 32-bit voxel indices: it is computed in z-chunks on the one device -- a
    // slab group that names the device several times (the reference's QuickSurf chunks its volume in z as well,
    // CUDAQuickSurf.cu:1050-1126, 1406-1447)
    if (devs.size() <= 1) {
        const unsigned long long nvox = static_cast<unsigned long long>(grid.res[0]) * grid.res[1] * grid.res[2];
        if (nvox >= (1ull << 32) - 1) {
            const int dev = devs.empty() ? this->deviceSlot.Param<core::param::IntParam>()->Value() : devs[0];
            const unsigned long long chunks = std::min<unsigned long long>((nvox >> 31) + 1, 16ull);
            devs.assign(static_cast<size_t>(chunks), dev);
        }
    }
    this->groupActive = false;
    if (devs.size() > 1) {
        if (this->volumeOnDevice) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: 'memoryLocation' = VRAM hands out ONE device allocation; it cannot be combined with 'devices'");
            return false;
        }
        if (!((p.mode == MMS_MODE_P2D_BUMP && p.aggregator == 0) || p.mode == MMS_MODE_QS_GAUSS)) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: 'devices' computes the position aggregator of the bump mode or the QuickSurf Gaussian; use 'device' for the other modes");
            return false;
        }
        if (this->group == nullptr || devs != this->groupDevices) {
            if (this->group != nullptr) mms_slabs_destroy(this->group);
            this->group = nullptr;
            if (mms_slabs_create(&this->group, devs.data(), static_cast<int32_t>(devs.size())) != MMS_OK) {
                Log::DefaultLog.WriteError("ParticlesToDensityB200: %s", mms_slabs_last_error(nullptr));
                return false;
            }
            this->groupDevices = devs;
        }
        auto gfail = [&](const char* what) {
            Log::DefaultLog.WriteError("ParticlesToDensityB200: %s: %s", what, mms_slabs_last_error(this->group));
            return false;
        };
        this->gridUsed = grid;
        this->ownGrid = false;
        float mm2[2] = {0, 0};
        if (mms_slabs_clear_particles(this->group) != MMS_OK || mms_slabs_set_grid(this->group, &grid) != MMS_OK ||
            mms_slabs_set_params(this->group, &p) != MMS_OK ||
            mms_slabs_push_particles(this->group, static_cast<int32_t>(lists.size()), lists.data()) != MMS_OK ||
            mms_slabs_compute_density(this->group) != MMS_OK || mms_slabs_get_density_range(this->group, mm2) != MMS_OK ||
            mms_slabs_get_density(this->group, &this->hostVolume) != MMS_OK)
            return gfail("slab group");
        this->minDens = mm2[0], this->maxDens = mm2[1];
        this->isVector = false;
        this->hasColour = false;
        this->groupActive = true;
        Log::DefaultLog.WriteInfo("ParticlesToDensityB200: Captured density %f -> %f", this->minDens, this->maxDens);
        if (p.normalize) {
            this->minDens = 0.0f;
            this->maxDens = 1.0f;
        }
        const std::chrono::duration<float, std::milli> msg = std::chrono::high_resolution_clock::now() - t0;
        Log::DefaultLog.WriteInfo("ParticlesToDensityB200: creation of %u x %u x %u volume from %llu particles on %zu devices took %f ms.", grid.res[0],
            grid.res[1], grid.res[2], static_cast<unsigned long long>(total), devs.size(), msg.count());
        return true;
    }
    if (mms_clear_particles(this->ctx) != MMS_OK)
        return fail("clear_particles");
    if (mms_push_particles_dir(this->ctx, static_cast<int32_t>(lists.size()), lists.data(), dirs.data(), dirStrides.data()) != MMS_OK)
        return fail("push_particles");
    const float gridSpacing = this->qsGridSpacingSlot.Param<core::param::FloatParam>()->Value();
    this->ownGrid = gaussian && gridSpacing > 0.0f;
    if (this->ownGrid) {
        // QuickSurf's grid set-up (QuickSurf.cpp:456-480): the bounding box grown by a padding derived from the largest radius, then
        // ceil(extent / gridspacing) voxels of exactly that spacing, origin = the padded minimum.  The molecule path skips the
        // padding ("we ignore the padding and the radscale", :345-360).
        float pad = 0.0f;
        if (mpdc != nullptr) {
            float rmax = 0.0f;
            if (mms_get_max_radius(this->ctx, &rmax) != MMS_OK)
                return fail("get_max_radius");
            float gridpadding = p.radscale * rmax * 1.5f;
            const float pi = std::acos(-1.0f);
            const float padrad = static_cast<float>(0.4 * std::sqrt(4.0 / 3.0 * pi * gridpadding * gridpadding * gridpadding));
            pad = std::max(gridpadding, padrad);
        }
        const float lo[3] = {bbox.Left() - pad, bbox.Bottom() - pad, bbox.Back() - pad};
        const float hi[3] = {bbox.Right() + pad, bbox.Top() + pad, bbox.Front() + pad};
        for (int a = 0; a < 3; ++a) {
            grid.min[a] = lo[a];
            grid.res[a] = std::max(2, static_cast<int>(std::ceil((hi[a] - lo[a]) / gridSpacing)));
            grid.extent[a] = static_cast<float>(grid.res[a] - 1) * gridSpacing; // node i sits at origin + i * gridspacing
        }
    }
    this->gridUsed = grid;
    if (mms_set_grid(this->ctx, &grid) != MMS_OK)
        return fail("set_grid");
    if (mms_set_params(this->ctx, &p) != MMS_OK)
        return fail("set_params");
    if (mms_compute_density(this->ctx) != MMS_OK)
        return fail("compute_density");
    float mm[2] = {0, 0};
    if (mms_get_density_range(this->ctx, mm) != MMS_OK)
        return fail("get_density_range");
    this->minDens = mm[0], this->maxDens = mm[1];
    this->isVector = p.mode == MMS_MODE_P2D_BUMP && p.aggregator == 2;
    if (this->isVector) {
        if (!this->buildVectorOutputs(grid, p.normalize != 0))
            return fail("get_vector_field");
    } else if (this->volumeOnDevice) {
        this->hostVolume = nullptr; // VRAM: consumers import the device memory (ShareDensity) or adopt it (IsoSurfaceB200)
    } else if (mms_get_density(this->ctx, &this->hostVolume, nullptr) != MMS_OK) // RAM contract of VolumetricDataCall::GetData()
        return fail("get_density");
    this->hasColour = p.colour != 0;
    Log::DefaultLog.WriteInfo("ParticlesToDensityB200: Captured density %f -> %f", this->minDens, this->maxDens);
    if (p.normalize) {
        this->minDens = 0.0f;
        this->maxDens = 1.0f;
    }
    const std::chrono::duration<float, std::milli> ms = std::chrono::high_resolution_clock::now() - t0;
    Log::DefaultLog.WriteInfo("ParticlesToDensityB200: creation of %u x %u x %u volume from %llu particles took %f ms.", grid.res[0],
        grid.res[1], grid.res[2], static_cast<unsigned long long>(total), ms.count());
    return true;
}

/**
 * Aggregator 2: the host-side tail of createVolumeCPU (ParticlesToDensity.cpp:634-667 bookkeeping, :684-727 compaction).  The
 * per-voxel arithmetic (v = sum(w d)/sum(w), |v|, v/|v|, range, normalisation) already happened on the device; what is left is
 * the table / grid-particle bookkeeping: colours (|v| - min)/(max - min), the seven table columns, and the removal of the
 * zero-length vectors by a sort on |v|, descending.  The reference sorts with simultaneous_sort (order of equal magnitudes
 * unspecified); this module keeps equal magnitudes in voxel order.
 */
bool ParticlesToDensityB200::buildVectorOutputs(const mms_grid& grid, bool normalize) {
    const float *vec = nullptr, *mag = nullptr, *dir = nullptr;
    if (mms_get_vector_field(this->ctx, &vec, &mag, &dir) != MMS_OK)
        return false;
    this->hostVolume = vec;
    const size_t sx = grid.res[0], sy = grid.res[1], sz = grid.res[2], n = sx * sy * sz;
    const float mn = this->minDens, mx = this->maxDens;
    float sd[3];
    for (int a = 0; a < 3; ++a)
        sd[a] = grid.extent[a] / static_cast<float>(grid.res[a] - 1);
    std::vector<size_t> order(n);
    std::iota(order.begin(), order.end(), size_t(0));
    std::stable_sort(order.begin(), order.end(), [mag](size_t a, size_t b) { return mag[a] > mag[b]; });
    size_t kept = 0;
    while (kept < n && mag[order[kept]] != 0.0f) // std::find(densities, 0.0f) on the sorted magnitudes (:696-697)
        ++kept;
    this->gridPos.resize(3 * kept);
    this->directions.resize(3 * kept);
    this->colors.resize(kept);
    this->infoData.resize(7 * kept);
    for (size_t r = 0; r < kept; ++r) {
        const size_t i = order[r];
        const size_t x = i % sx, y = (i / sx) % sy, z = i / (sx * sy);
        const float pos[3] = {grid.min[0] + sd[0] * x, grid.min[1] + sd[1] * y, grid.min[2] + sd[2] * z}; // :440-442
        const float colour = (mag[i] - mn) / (mx - mn);
        for (int k = 0; k < 3; ++k) {
            this->gridPos[3 * r + k] = pos[k];
            this->directions[3 * r + k] = dir[3 * i + k];
            this->infoData[7 * r + k] = pos[k];
            this->infoData[7 * r + 3 + k] = dir[3 * i + k];
        }
        this->colors[r] = colour;
        this->infoData[7 * r + 6] = normalize ? colour : mag[i];
    }
    return true;
}

bool ParticlesToDensityB200::getDataCallback(core::Call& c) {
    auto* in = this->inDataSlot.CallAs<core::AbstractGetData3DCall>();
    if (in == nullptr)
        return false;
    auto* outVol = dynamic_cast<VolumetricDataCall*>(&c);
    auto* outGrid = dynamic_cast<MultiParticleDataCall*>(&c);

    if (outVol != nullptr || outGrid != nullptr) {
        const unsigned int frameID = outVol != nullptr ? outVol->FrameID() : outGrid->FrameID();
        do {
            in->SetFrameID(frameID, true);
            if (!(*in)(1)) {
                Log::DefaultLog.WriteError("ParticlesToDensityB200: Unable to get extents.");
                return false;
            }
            if (!(*in)(0)) {
                Log::DefaultLog.WriteError("ParticlesToDensityB200: Unable to get data.");
                return false;
            }
        } while (in->FrameID() != frameID);
        if (this->time != in->FrameID() || this->in_datahash != in->DataHash() || this->anythingDirty() || !this->has_data) {
            if (this->surfaceSlot.Param<core::param::BoolParam>()->Value())
                this->surfaceBBox(in);
            if (!this->computeVolume(in))
                return false;
            this->time = in->FrameID();
            this->in_datahash = in->DataHash();
            ++this->datahash;
            this->resetDirty();
            this->has_data = true;
        }
    }
    if (outVol != nullptr) {
        outVol->SetFrameID(this->time);
        outVol->SetData(const_cast<float*>(this->hostVolume));
        this->fillMetadata(in);
        outVol->SetMetadata(&this->metadata);
        outVol->SetDataHash(this->datahash);
    }
    const bool vectorParam = this->aggregatorSlot.Param<core::param::EnumParam>()->Value() == 2;
    if (auto* outInfo = dynamic_cast<datatools::table::TableDataCall*>(&c)) { // table rows exist only for the vector aggregator
        if (vectorParam) { // :325-377
            using CT = datatools::table::TableDataCall::ColumnType;
            static const char* const kNames[7] = {"PositionX", "PositionY", "PositionZ", "VelocityX", "VelocityY", "VelocityZ", "VelocityMag"};
            for (int k = 0; k < 7; ++k)
                this->info[k].SetName(kNames[k]).SetType(CT::QUANTITATIVE);
            const bool have = this->has_data && this->isVector;
            if (!have) {
                this->infoData.clear();
                outInfo->SetDataHash(0);
            } else {
                const auto bb = in->AccessBoundingBoxes().ObjectSpaceBBox();
                this->info[0].SetMinimumValue(bb.Left()).SetMaximumValue(bb.Right());
                this->info[1].SetMinimumValue(bb.Bottom()).SetMaximumValue(bb.Top());
                this->info[2].SetMinimumValue(bb.Back()).SetMaximumValue(bb.Front());
                for (int k = 3; k < 6; ++k)
                    this->info[k].SetMinimumValue(-1.0f).SetMaximumValue(1.0f);
                this->info[6].SetMinimumValue(this->minDens).SetMaximumValue(this->maxDens);
                outInfo->SetDataHash(this->datahash);
            }
            outInfo->Set(this->info.size(), this->infoData.size() / this->info.size(), this->info.data(), this->infoData.data());
        } else {
            outInfo->SetDataHash(this->datahash);
            outInfo->Set(0, 0, nullptr, nullptr);
        }
    }
    if (outGrid != nullptr) { // grid particles exist only for the vector aggregator (:297-315)
        outGrid->SetFrameID(this->time);
        outGrid->SetDataHash(this->datahash);
        if (vectorParam && this->isVector) {
            outGrid->SetParticleListCount(1);
            auto& gp = outGrid->AccessParticles(0);
            gp.SetCount(this->colors.size());
            if (gp.GetCount() > 0) {
                gp.SetVertexData(MultiParticleDataCall::Particles::VERTDATA_FLOAT_XYZ, this->gridPos.data());
                gp.SetDirData(geocalls::SimpleSphericalParticles::DIRDATA_FLOAT_XYZ, this->directions.data());
                gp.SetColourData(geocalls::SimpleSphericalParticles::COLDATA_FLOAT_I, this->colors.data());
                gp.SetGlobalRadius(in->AccessBoundingBoxes().ObjectSpaceBBox().Width() /
                                   static_cast<float>(this->xResSlot.Param<core::param::IntParam>()->Value()) / 5.0f);
            }
        } else {
            outGrid->SetParticleListCount(0);
        }
    }
    in->Unlock();
    return true;
}
