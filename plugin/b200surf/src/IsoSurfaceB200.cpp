/*
 * IsoSurfaceB200.cpp
 *
 * Call protocol of trisoup_gl::volumetrics::IsoSurface (plugins/trisoup_gl/src/volumetrics/IsoSurface.cpp:94-223).
 * If the volume comes from a ParticlesToDensityB200 the density never leaves the GPU: the isosurface is extracted
 * on that module's context.  Any other VolumetricDataCall source is uploaded through mms_set_density.
 * Vertex positions use the node-centred frame of the metadata (Origin + idx * SliceDist, docs/volumes.md:14-16),
 * not the reference IsoSurface's half-voxel-shifted frame (IsoSurface.cpp:238-252; SURVEY.md 8a traps).
 */
#include "IsoSurfaceB200.h"

#include <chrono>

#include "ParticlesToDensityB200.h"
#include "mmcore/param/BoolParam.h"
#include "mmcore/param/EnumParam.h"
#include "mmcore/param/FloatParam.h"
#include "mmcore/param/IntParam.h"
#include "mmcore/param/StringParam.h"
#include "mmcore/utility/log/Log.h"

using namespace megamol;
using namespace megamol::b200surf;
using megamol::core::utility::log::Log;
using geocalls::VolumetricDataCall;
using geocalls_gl::CallTriMeshDataGL;

bool IsoSurfaceB200::IsAvailable() {
    return ParticlesToDensityB200::IsAvailable();
}

IsoSurfaceB200::IsoSurfaceB200()
        : inDataSlot("inData", "The slot for requesting input data")
        , outDataSlot("outData", "Gets the data")
        , attributeSlot("attr", "The attribute to show")
        , isoValueSlot("isoval", "The iso value")
        , deviceSlot("device", "CUDA device ordinal used for volumes that are not already device resident")
        , algorithmSlot("algorithm", "Triangulation: marching cubes (smooth normals, node-centred frame) or the CPU module's marching "
                                     "tetrahedra reproduced triangle for triangle")
        , deviceMeshSlot("deviceMesh", "Leave the mesh in importable device memory (ShareMesh()); CallTriMeshData then carries no object")
        , indexedMeshSlot("indexedMesh", "Hand out an indexed mesh (one vertex per crossed grid edge + 32-bit triangle indices) instead of the "
                                         "reference's unindexed triangle soup; marching cubes on a single device only") {

    this->inDataSlot.SetCompatibleCall<geocalls::VolumetricDataCallDescription>();
    this->MakeSlotAvailable(&this->inDataSlot);

    this->outDataSlot.SetCallback("CallTriMeshData", "GetData", &IsoSurfaceB200::outDataCallback);
    this->outDataSlot.SetCallback("CallTriMeshData", "GetExtent", &IsoSurfaceB200::outExtentCallback);
    this->MakeSlotAvailable(&this->outDataSlot);

    this->attributeSlot << new core::param::StringParam("0");
    this->MakeSlotAvailable(&this->attributeSlot);

    this->isoValueSlot << new core::param::FloatParam(0.5f);
    this->MakeSlotAvailable(&this->isoValueSlot);

    this->deviceSlot << new core::param::IntParam(0, 0);
    this->MakeSlotAvailable(&this->deviceSlot);

    auto* alg = new core::param::EnumParam(MMS_ISO_MARCHING_CUBES);
    alg->SetTypePair(MMS_ISO_MARCHING_CUBES, "MarchingCubes");
    alg->SetTypePair(MMS_ISO_MARCHING_TETS, "MarchingTetrahedra (as trisoup_gl::IsoSurface)");
    this->algorithmSlot << alg;
    this->MakeSlotAvailable(&this->algorithmSlot);

    this->deviceMeshSlot << new core::param::BoolParam(false);
    this->MakeSlotAvailable(&this->deviceMeshSlot);

    this->indexedMeshSlot << new core::param::BoolParam(false);
    this->MakeSlotAvailable(&this->indexedMeshSlot);
}

IsoSurfaceB200::~IsoSurfaceB200() {
    this->Release();
}

bool IsoSurfaceB200::create() {
    return true;
}

void IsoSurfaceB200::release() {
    if (this->ctx != nullptr) {
        mms_destroy(this->ctx);
        this->ctx = nullptr;
    }
    if (this->group != nullptr) {
        mms_slabs_destroy(this->group);
        this->group = nullptr;
    }
}

bool IsoSurfaceB200::outExtentCallback(core::Call& caller) {
    auto* tmd = dynamic_cast<CallTriMeshDataGL*>(&caller);
    if (tmd == nullptr)
        return false;
    tmd->AccessBoundingBoxes().Clear();
    auto* cvd = this->inDataSlot.CallAs<VolumetricDataCall>();
    if (cvd != nullptr)
        cvd->SetFrameID(tmd->FrameID(), tmd->IsFrameForced());
    if (cvd == nullptr || !(*cvd)(VolumetricDataCall::IDX_GET_EXTENTS) || !(*cvd)(VolumetricDataCall::IDX_GET_METADATA)) {
        tmd->SetDataHash(0);
        tmd->SetFrameCount(1);
    } else {
        tmd->SetDataHash(cvd->DataHash());
        tmd->SetExtent(cvd->FrameCount(), cvd->AccessBoundingBoxes());
    }
    tmd->SetUnlocker(nullptr);
    return true;
}

bool IsoSurfaceB200::buildMesh(VolumetricDataCall* cvd, float iso) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    const int algorithm = this->algorithmSlot.Param<core::param::EnumParam>()->Value();
    // device-resident hand-off: is the callee a ParticlesToDensityB200 whose context holds exactly this volume?  Then this module's OWN
    // context adopts that volume by reference (mms_adopt_density): no host round trip, and the mesh lives in this module's buffers --
    // several IsoSurfaceB200 behind one producer, or a producer that recomputes, cannot touch a mesh this module has handed out.
    mms_ctx* producer = nullptr;
    int device = this->deviceSlot.Param<core::param::IntParam>()->Value();
    const core::CalleeSlot* callee = cvd->PeekCalleeSlot();
    if (callee != nullptr) {
        auto parent = callee->Parent();
        auto* p2d = dynamic_cast<const ParticlesToDensityB200*>(parent.get());
        if (p2d != nullptr && p2d->Group() != nullptr && this->deviceMeshSlot.Param<core::param::BoolParam>()->Value()) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: 'deviceMesh' hands out ONE device allocation; it cannot follow a multi-device producer");
            return false;
        }
        if (p2d != nullptr && p2d->Group() != nullptr && this->indexedMeshSlot.Param<core::param::BoolParam>()->Value()) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: 'indexedMesh' needs the whole volume on one device (vertex ids cross z-slab borders)");
            return false;
        }
        if (p2d != nullptr && p2d->Group() != nullptr && p2d->VolumeHash() == cvd->DataHash() && algorithm == MMS_ISO_MARCHING_CUBES) {
            // a multi-device producer: this module's own slab group on the same devices adopts every slab's volume, the slabs' meshes are
            // concatenated (= the single-GPU order) into this module's pinned arrays
            if (this->group == nullptr || this->groupDevices != p2d->GroupDevices()) {
                if (this->group != nullptr) mms_slabs_destroy(this->group);
                this->group = nullptr;
                this->groupDevices = p2d->GroupDevices();
                if (mms_slabs_create(&this->group, this->groupDevices.data(), static_cast<int32_t>(this->groupDevices.size())) != MMS_OK) {
                    Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_slabs_last_error(nullptr));
                    return false;
                }
            }
            uint64_t nv = 0;
            const float *gp = nullptr, *gn = nullptr, *gc = nullptr; // gc stays NULL unless the slabs carry a QuickSurf colour volume
            if (mms_slabs_adopt_density(this->group, p2d->Group()) != MMS_OK || mms_slabs_extract_isosurface(this->group, iso) != MMS_OK ||
                mms_slabs_get_mesh(this->group, &nv, &gp, &gn) != MMS_OK || mms_slabs_get_mesh_colours(this->group, &gc) != MMS_OK) {
                Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_slabs_last_error(this->group));
                return false;
            }
            this->mesh.SetMaterial(nullptr);
            this->mesh.SetVertexData(static_cast<unsigned int>(nv), const_cast<float*>(gp), const_cast<float*>(gn), const_cast<float*>(gc),
                static_cast<float*>(nullptr), false);
            this->mesh.SetTriangleData(0, static_cast<unsigned int*>(nullptr), false);
            const std::chrono::duration<float, std::milli> msg = std::chrono::high_resolution_clock::now() - t0;
            Log::DefaultLog.WriteInfo("IsoSurfaceB200: %llu triangles at iso %f took %f ms (device-resident volume on %zu devices).",
                static_cast<unsigned long long>(nv / 3), iso, msg.count(), this->groupDevices.size());
            return true;
        }
        if (p2d != nullptr && p2d->Context() != nullptr && p2d->Group() == nullptr && p2d->VolumeHash() == cvd->DataHash()) {
            producer = p2d->Context();
            device = p2d->Device();
        }
    }
    if (this->ctx == nullptr || this->ctxDevice != device) {
        if (this->ctx != nullptr)
            mms_destroy(this->ctx);
        this->ctx = nullptr;
        mms_config cfg{device, 0};
        if (mms_create(&this->ctx, &cfg) != MMS_OK) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(nullptr));
            return false;
        }
        this->ctxDevice = device;
        this->meshOnDevice = false;
    }
    const bool wantDevice = this->deviceMeshSlot.Param<core::param::BoolParam>()->Value();
    if (wantDevice != this->meshOnDevice) {
        if (mms_share_enable(this->ctx, wantDevice ? 1 : 0) != MMS_OK) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(this->ctx));
            return false;
        }
        this->meshOnDevice = wantDevice;
    }
    mms_ctx* use = this->ctx;
    if (producer != nullptr) {
        if (mms_adopt_density(this->ctx, producer) != MMS_OK) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(this->ctx));
            return false;
        }
    } else {
        const auto* md = cvd->GetMetadata();
        if (md == nullptr || cvd->GetData() == nullptr || md->Components != 1 || md->GridType != geocalls::GridType_t::CARTESIAN) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: need a host-resident single-component cartesian float volume");
            return false;
        }
        mms_grid grid{};
        for (int a = 0; a < 3; ++a) {
            grid.min[a] = md->Origin[a];
            grid.extent[a] = md->Extents[a];
            grid.res[a] = static_cast<int32_t>(md->Resolution[a]);
            grid.cyclic[a] = 0;
        }
        if (algorithm == MMS_ISO_MARCHING_TETS) { // the CPU module places its cells in the object-space bounding box (IsoSurface.cpp:125, 238-252)
            const auto& bb = cvd->AccessBoundingBoxes().ObjectSpaceBBox();
            grid.min[0] = bb.Left(), grid.min[1] = bb.Bottom(), grid.min[2] = bb.Back();
            grid.extent[0] = bb.Width(), grid.extent[1] = bb.Height(), grid.extent[2] = bb.Depth();
        }
        if (mms_set_grid(this->ctx, &grid) != MMS_OK || mms_set_density(this->ctx, static_cast<const float*>(cvd->GetData())) != MMS_OK) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(this->ctx));
            return false;
        }
    }
    uint64_t nverts = 0;
    const float *pos = nullptr, *nrm = nullptr, *col = nullptr; // col stays NULL unless the volume carries colours (QuickSurf mode)
    const bool indexed = this->indexedMeshSlot.Param<core::param::BoolParam>()->Value();
    if (mms_set_mesh_indexed(use, indexed ? 1 : 0) != MMS_OK) {
        Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(use));
        return false;
    }
    if (indexed) {
        // opt-in: CallTriMeshData's indexed form (Mesh::SetVertexData + SetTriangleData(cnt, unsigned int*), CallTriMeshDataGL.h:897-1000)
        uint64_t ntris = 0;
        const uint32_t* idx = nullptr;
        if (this->meshOnDevice) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: 'indexedMesh' and 'deviceMesh' cannot be combined (ShareMesh() shares the triangle soup)");
            return false;
        }
        if (mms_set_isosurface_mode(use, algorithm) != MMS_OK || mms_extract_isosurface(use, iso) != MMS_OK ||
            mms_get_mesh_indexed(use, &nverts, &ntris, &pos, &nrm, &idx) != MMS_OK) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(use));
            return false;
        }
        this->mesh.SetMaterial(nullptr);
        this->mesh.SetVertexData(static_cast<unsigned int>(nverts), const_cast<float*>(pos), const_cast<float*>(nrm), static_cast<float*>(nullptr),
            static_cast<float*>(nullptr), false);
        this->mesh.SetTriangleData(static_cast<unsigned int>(ntris), const_cast<unsigned int*>(idx), false);
        const std::chrono::duration<float, std::milli> msi = std::chrono::high_resolution_clock::now() - t0;
        Log::DefaultLog.WriteInfo("IsoSurfaceB200: %llu triangles on %llu vertices (indexed) at iso %f took %f ms.", static_cast<unsigned long long>(ntris),
            static_cast<unsigned long long>(nverts), iso, msi.count());
        return true;
    }
    if (mms_set_isosurface_mode(use, algorithm) != MMS_OK || mms_extract_isosurface(use, iso) != MMS_OK ||
        (!this->meshOnDevice && mms_get_mesh(use, &nverts, &pos, &nrm, &col) != MMS_OK)) {
        Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(use));
        return false;
    }
    this->mesh.SetMaterial(nullptr);
    // the reference's contract (IsoSurface.cpp:171-181): unindexed soup, float positions + normals, no colours, 0 "triangles"
    // (the Mesh setters are overloaded on NON-const pointers; a const float* would select the catch-all "no data" overload)
    this->mesh.SetVertexData(static_cast<unsigned int>(nverts), const_cast<float*>(pos), const_cast<float*>(nrm), const_cast<float*>(col),
        static_cast<float*>(nullptr), false);
    this->mesh.SetTriangleData(0, static_cast<unsigned int*>(nullptr), false);
    const std::chrono::duration<float, std::milli> ms = std::chrono::high_resolution_clock::now() - t0;
    Log::DefaultLog.WriteInfo("IsoSurfaceB200: %llu triangles at iso %f took %f ms (%s volume).", static_cast<unsigned long long>(nverts / 3), iso,
        ms.count(), producer == nullptr ? "uploaded" : "device-resident");
    return true;
}

bool IsoSurfaceB200::ShareMesh(uint64_t* nverts, mms_share* positions, mms_share* normals, mms_share* colours) {
    if (!this->has_mesh || !this->meshOnDevice || this->ctx == nullptr)
        return false;
    if (mms_share_mesh(this->ctx, nverts, positions, normals, colours) != MMS_OK) {
        Log::DefaultLog.WriteError("IsoSurfaceB200: %s", mms_last_error(this->ctx));
        return false;
    }
    return true;
}

bool IsoSurfaceB200::outDataCallback(core::Call& caller) {
    auto* tmd = dynamic_cast<CallTriMeshDataGL*>(&caller);
    if (tmd == nullptr)
        return false;
    auto* cvd = this->inDataSlot.CallAs<VolumetricDataCall>();
    if (cvd != nullptr) {
        bool recalc = false;
        if (this->isoValueSlot.IsDirty()) {
            this->isoValueSlot.ResetDirty();
            recalc = true;
        }
        if (this->attributeSlot.IsDirty()) {
            this->attributeSlot.ResetDirty();
            recalc = true;
        }
        if (this->algorithmSlot.IsDirty()) {
            this->algorithmSlot.ResetDirty();
            recalc = true;
        }
        if (this->deviceMeshSlot.IsDirty()) {
            this->deviceMeshSlot.ResetDirty();
            recalc = true;
        }
        if (this->indexedMeshSlot.IsDirty()) {
            this->indexedMeshSlot.ResetDirty();
            recalc = true;
        }
        cvd->SetFrameID(tmd->FrameID(), tmd->IsFrameForced());
        if (!(*cvd)(VolumetricDataCall::IDX_GET_EXTENTS) || !(*cvd)(VolumetricDataCall::IDX_GET_METADATA) ||
            !(*cvd)(VolumetricDataCall::IDX_GET_DATA)) {
            recalc = false;
        } else if (this->dataHash != cvd->DataHash() || this->frameIdx != cvd->FrameID() || !this->has_mesh) {
            recalc = true;
        }
        if (recalc && cvd->GetScalarType() != VolumetricDataCall::ScalarType::FLOATING_POINT) {
            Log::DefaultLog.WriteError("IsoSurfaceB200: Only float volumes are supported ATM");
            recalc = false;
        }
        if (recalc) {
            if (!this->buildMesh(cvd, this->isoValueSlot.Param<core::param::FloatParam>()->Value()))
                return false;
            this->dataHash = cvd->DataHash();
            this->frameIdx = cvd->FrameID();
            this->has_mesh = true;
        }
    }
    tmd->SetDataHash(this->dataHash);
    tmd->SetFrameID(this->frameIdx);
    tmd->SetObjects(this->meshOnDevice ? 0 : 1, &this->mesh); // device mesh: nothing for a host-side consumer, see ShareMesh()
    tmd->SetUnlocker(nullptr);
    return true;
}
