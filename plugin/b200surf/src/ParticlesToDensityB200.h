/*
 * ParticlesToDensityB200.h -- drop-in for datatools::ParticlesToDensity (plugins/datatools/src/ParticlesToDensity.h):
 * same slots (inData / outData / outParticles / outInfo), same parameters with the same defaults, same call protocol;
 * the volume is computed by libmmsurf on a B200 instead of createVolumeCPU.
 * "inData" additionally accepts a protein_calls::MolecularDataCall, the other input of the reference's QuickSurf module
 * (QuickSurf::calculateSurface(MolecularDataCall&), plugins/protein_cuda/src/QuickSurf.cpp:326-404): one sphere per atom with
 * the radius of its atom type, coloured with the atom type's colour.
 */
#pragma once

#include <array>
#include <cstddef>
#include <limits>
#include <vector>

#include "datatools/table/TableDataCall.h"

#include "geometry_calls/MultiParticleDataCall.h"
#include "geometry_calls/VolumetricDataCall.h"
#include "mmcore/CalleeSlot.h"
#include "mmcore/CallerSlot.h"
#include "mmcore/Module.h"
#include "mmcore/param/ParamSlot.h"
#include "protein_calls/ProteinColor.h"
#include "mmstd/data/AbstractGetData3DCall.h"
#include "protein_calls/MolecularDataCall.h"

#include "mmsurf.h"

namespace megamol::b200surf {

class ParticlesToDensityB200 : public core::Module {
public:
    static const char* ClassName() {
        return "ParticlesToDensityB200";
    }
    static const char* Description() {
        return "Computes a density volume from particles on a B200 (drop-in for ParticlesToDensity)";
    }
    /** No CPU fallback: the module is only available where libmmsurf finds an sm_100 device. */
    static bool IsAvailable();

    ParticlesToDensityB200();
    ~ParticlesToDensityB200() override;

    /** Device-resident hand-off for IsoSurfaceB200: the context holding the volume of the last GetData. */
    mms_ctx* Context() const {
        return this->ctx;
    }
    std::size_t VolumeHash() const {
        return this->datahash;
    }
    /** CUDA device ordinal of Context(). */
    int Device() const {
        return this->ctxDevice;
    }
    /** Several devices ("devices" parameter): the slab group holding the volume of the last GetData, else NULL. */
    mms_slabs* Group() const {
        return this->groupActive ? this->group : nullptr;
    }
    const std::vector<int32_t>& GroupDevices() const {
        return this->groupDevices;
    }
    /**
     * Device-resident hand-off ('memoryLocation' = VRAM): the volume (and, with colours, the RGB volume) as importable device memory --
     * what VolumetricDataCall::SetData(uint32_t texture) + MemLoc VRAM (VolumetricDataCall.h:290-292) is for GL.  The caller closes the
     * descriptors.  false when the parameter is off or nothing has been computed.
     */
    bool ShareDensity(mms_share* volume, mms_share* rgb);
    /** true if the context also holds a density-weighted RGB volume (QuickSurf mode with colour). */
    bool HasColour() const {
        return this->hasColour;
    }

protected:
    bool create() override;
    void release() override;

private:
    bool getExtentCallback(core::Call& c);
    bool getMetadataCallback(core::Call& c);
    bool getDataCallback(core::Call& c);
    bool dummyCallback(core::Call& c);

    bool anythingDirty() const;
    void resetDirty();
    bool computeVolume(core::AbstractGetData3DCall* in);
    void fillMetadata(core::AbstractGetData3DCall* in);
    void surfaceBBox(core::AbstractGetData3DCall* in);
    bool buildVectorOutputs(const mms_grid& grid, bool normalize);

    core::param::ParamSlot aggregatorSlot, xResSlot, yResSlot, zResSlot, cyclXSlot, cyclYSlot, cyclZSlot, normalizeSlot,
        sigmaSlot, surfaceSlot;
    core::param::ParamSlot deviceSlot; // extra: CUDA device ordinal
    // extra: QuickSurf semantics as a kernel mode (names and defaults of protein_cuda::QuickSurf, QuickSurf.cpp:18-31,62-72)
    core::param::ParamSlot modeSlot, qsQualitySlot, qsRadScaleSlot, qsColourSlot, qsGridSpacingSlot, qsRefCellsSlot;
    // the reference QuickSurf module's colouring of a MolecularDataCall (plugins/protein_cuda/src/QuickSurf.cpp:25-28, 62-91, 281-320,
    // 596-616): two protein_calls::ProteinColor colouring modes blended by a weight, gradient colours, colour table file
    core::param::ParamSlot colorTableFileSlot, coloringMode0Slot, coloringMode1Slot, coloringModeWeightSlot, minGradColorSlot, midGradColorSlot,
        maxGradColorSlot;
    std::vector<glm::vec3> atomColorTable, fileColorTable, rainbowColorTable;
    core::CalleeSlot outDataSlot, outParticlesSlot, outInfoSlot;
    core::CallerSlot inDataSlot;

    mms_ctx* ctx = nullptr;
    int ctxDevice = -1;
    core::param::ParamSlot memLocSlot; // extra: RAM (the reference's contract) or VRAM (no host copy; consumers import the device memory)
    bool volumeOnDevice = false;
    core::param::ParamSlot devicesSlot; // extra: several CUDA devices, e.g. "0,1,2,3": the volume is computed in z-slabs (scalar bump mode)
    mms_slabs* group = nullptr;
    std::vector<int32_t> groupDevices;
    bool groupActive = false;
    const float* hostVolume = nullptr;
    std::size_t in_datahash = std::numeric_limits<std::size_t>::max();
    std::size_t datahash = 0;
    unsigned int time = 0;
    float minDens = 0.0f, maxDens = 0.0f;
    bool has_data = false;
    bool hasColour = false;
    geocalls::VolumetricDataCall::Metadata metadata;
    double minValue[3] = {0.0, 0.0, 0.0}, maxValue[3] = {0.0, 0.0, 0.0};
    // aggregator 2 (IVecToSingleCell_Volume): what "outParticles" and "outInfo" hand out (ParticlesToDensity.h:117-125)
    bool isVector = false;
    std::vector<float> gridPos, directions, colors, infoData;
    std::vector<float> atoms; // MolecularDataCall input: x y z r R G B A per atom, rebuilt per frame
    std::array<datatools::table::TableDataCall::ColumnInfo, 7> info;
    float sliceDists[3] = {0, 0, 0};
    mms_grid gridUsed{}; // the grid of the last compute (differs from bbox + sizex/y/z when QuickSurf's own grid set-up is on)
    bool ownGrid = false;
};

} // namespace megamol::b200surf
