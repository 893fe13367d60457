/*
 * IsoSurfaceB200.h -- drop-in for trisoup_gl::volumetrics::IsoSurface (plugins/trisoup_gl/src/volumetrics/IsoSurface.h):
 * slots inData (VolumetricDataCall) / outData ("CallTriMeshData": GetData, GetExtent), parameters attr, isoval.
 * Marching cubes with the classic table instead of marching tetrahedra, smooth normals; the mesh contract is the
 * reference's: one Mesh, unindexed triangle soup, float positions + float normals.  'indexedMesh' (off by default) switches to the
 * indexed form the call carries as well (SetVertexData with one vertex per crossed grid edge + SetTriangleData with 32-bit indices):
 * the same triangles in the same order, 28 instead of 72 bytes per triangle.
 */
#pragma once
#include <vector>

#include <cstddef>

#include "geometry_calls/VolumetricDataCall.h"
#include "geometry_calls_gl/CallTriMeshDataGL.h"
#include "mmcore/CalleeSlot.h"
#include "mmcore/CallerSlot.h"
#include "mmcore/Module.h"
#include "mmcore/param/ParamSlot.h"

#include "mmsurf.h"

namespace megamol::b200surf {

class IsoSurfaceB200 : public core::Module {
public:
    static const char* ClassName() {
        return "IsoSurfaceB200";
    }
    static const char* Description() {
        return "Extracts an iso-surface mesh from a volume on a B200 (drop-in for IsoSurface)";
    }
    static bool IsAvailable();

    IsoSurfaceB200();
    ~IsoSurfaceB200() override;

    /**
     * Device-resident hand-off ('deviceMesh' on): the triangle soup of the last GetData as importable device memory -- positions,
     * normals (and colours) with 3 floats per vertex, the layout CallTriMeshData hands out on the host.  A renderer imports the
     * descriptors (GL_EXT_memory_object_fd / Vulkan / CUDA) instead of re-buffering host arrays every frame
     * (trisoup_gl/src/ModernTrisoupRenderer.cpp:308-440).  The caller closes the descriptors.
     */
    bool ShareMesh(uint64_t* nverts, mms_share* positions, mms_share* normals, mms_share* colours);

protected:
    bool create() override;
    void release() override;

private:
    bool outDataCallback(core::Call& caller);
    bool outExtentCallback(core::Call& caller);
    bool buildMesh(geocalls::VolumetricDataCall* cvd, float iso);

    core::CallerSlot inDataSlot;
    core::CalleeSlot outDataSlot;
    core::param::ParamSlot attributeSlot, isoValueSlot, deviceSlot, algorithmSlot, deviceMeshSlot, indexedMeshSlot;
    bool meshOnDevice = false;

    mms_ctx* ctx = nullptr; // own context, used when the volume comes from a foreign (host) source
    int ctxDevice = -1;
    mms_slabs* group = nullptr; // own slab group, used behind a multi-device ParticlesToDensityB200
    std::vector<int32_t> groupDevices;
    std::size_t dataHash = 0;
    unsigned int frameIdx = 0;
    bool has_mesh = false;
    geocalls_gl::CallTriMeshDataGL::Mesh mesh;
};

} // namespace megamol::b200surf
