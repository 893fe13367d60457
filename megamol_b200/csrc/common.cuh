// common.cuh -- shared device-side types and helpers of libmmsurf (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mms {

constexpr int kMaxLists = 64;

/** Grid geometry as ParticlesToDensity::createVolumeCPU derives it (ParticlesToDensity.cpp:416-434). */
struct Geo {
    float mn[3];  // bbox Left/Bottom/Back
    float sd[3];  // sliceDist = range / float(res-1), computed on the HOST in fp32 exactly like the reference
    float isd[3]; // 1 / sliceDist (only for the slop-protected support bounds, never for the binning itself)
    int s[3];     // resolution
    int cyc[3];   // cyclX/Y/Z
    int z0, nz;   // slab: voxel planes [z0, z0+nz)
    int cshift;   // log2(cell edge in voxels)
    int nc[3];    // cell grid = ceil(s / cell)
    // A z-slab only bins what reaches it: its cell arrays cover czCount cell layers from global layer czBase on (wrapping on a periodic
    // axis) instead of all nc[2] -- per-GPU work must not grow with the number of slabs.  Local layer = cellZLocal(g, global layer).
    int czBase, czCount;
    float sigma;
    int agg;      // 0 position, 1 intensity-weighted, 2 direction-weighted (vector field)
    int mode;     // 0 P2D bump, 1 QuickSurf Gaussian
    float radscale, gausslim;
    int colour;
    // MMS_MODE_QS_GAUSS_REFCELLS: the reference's acceleration grid (CUDAQuickSurf.cu:1259-1282); qsAc == 0: radial cut-off
    float qsAc, qsInvAc;
    int qsCells[3];
};

/** One particle list on the device (mirror of mms_list with device pointers). */
struct ListDev {
    const char* vtx;
    const char* col;
    unsigned long long count;
    unsigned long long base; // global index of particle 0 of this list
    int vtype;
    unsigned vstride;
    int ctype;
    unsigned cstride;
    float grad;
    float gcol[4]; // global colour / 255
    float irange[2];
    int valign; // 16: float4 loads ok, 4: scalar float loads ok, 1: bytewise
    int calign;
    int gf[3];  // lists with a global radius: support half-width per axis (filter size / Gaussian cut-off), set per compute
    const char* dir; // DIRDATA_FLOAT_XYZ (aggregator 2) or nullptr = DIRDATA_NONE: the accessors deliver 0 (SimpleSphericalParticles.h:179-193)
    unsigned dstride;
    int dalign;
    // halo receive list (mms_halo_receive): the length lives on the device (written by the peers' halo_push_kernel), `count` is its bound
    const unsigned* countPtr;
    int radiusBound; // 1: `grad` bounds the per-particle radii (no device scan for the largest radius)
};

/** Number of particles of a list: host-known, or (halo receive list) read from device memory and clamped to the buffer's capacity. */
__device__ __forceinline__ unsigned long long listCount(const ListDev& l) {
    if (!l.countPtr) return l.count;
    const unsigned long long n = *l.countPtr;
    return n < l.count ? n : l.count;
}

/** Values produced on the device and consumed by later kernels without a host round trip. */
struct DevState {
    unsigned rmaxBits;   // max radius over all kept particles (positive float bits; atomicMax)
    unsigned kept;       // number of particles binned (not culled)
    unsigned minKey;     // density range, order-preserving uint keys (atomicMin/Max)
    unsigned maxKey;
    unsigned long long totalTris;
    unsigned pad[2];     // [0] error flag of the density kernels, [1] some kept particle lies outside the grid on a periodic axis
    unsigned nBig;       // crowded cells (more than kBigCell records) found by cell_order_kernel ...
    unsigned bigNext;    // ... and the work counter of cell_sort_big_kernel
    unsigned long long totalVerts; // indexed mesh: crossed grid edges = vertices (mcx_vertex_kernel + scan)
};

/** Local index (into the context's cell arrays) of global cell layer cz in [0, nc[2]); >= czCount: a layer this slab does not hold. */
__host__ __device__ __forceinline__ int cellZLocal(const Geo& g, int cz) {
    int l = cz - g.czBase;
    return l < 0 ? l + g.nc[2] : l;
}

__device__ __forceinline__ unsigned floatKey(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float keyFloat(unsigned k) {
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}

__device__ __forceinline__ int floorMod(int a, int m) {
    int r = a % m;
    return r < 0 ? r + m : r;
}

__device__ __forceinline__ float loadF32(const char* p, int align) {
    if (align >= 4) return *reinterpret_cast<const float*>(p);
    unsigned b = static_cast<unsigned char>(p[0]) | (static_cast<unsigned char>(p[1]) << 8) |
                 (static_cast<unsigned char>(p[2]) << 16) | (static_cast<unsigned>(static_cast<unsigned char>(p[3])) << 24);
    return __uint_as_float(b);
}
__device__ __forceinline__ double loadF64(const char* p, int align) {
    if (align >= 8) return *reinterpret_cast<const double*>(p);
    unsigned long long b = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b |= static_cast<unsigned long long>(static_cast<unsigned char>(p[i])) << (8 * i);
    return __longlong_as_double(static_cast<long long>(b));
}
__device__ __forceinline__ unsigned loadU16(const char* p) {
    return static_cast<unsigned char>(p[0]) | (static_cast<unsigned char>(p[1]) << 8);
}

/** Position + radius of particle j as the reference's accessors deliver them (Get_f: everything -> float). */
__device__ __forceinline__ float4 fetchParticle(const ListDev& l, unsigned long long j) {
    const char* p = l.vtx + j * l.vstride;
    float4 q;
    q.w = l.grad;
    switch (l.vtype) {
    case 1: // FLOAT_XYZ
        q.x = loadF32(p, l.valign), q.y = loadF32(p + 4, l.valign), q.z = loadF32(p + 8, l.valign);
        break;
    case 2: // FLOAT_XYZR
        if (l.valign >= 16) {
            q = *reinterpret_cast<const float4*>(p);
        } else {
            q.x = loadF32(p, l.valign), q.y = loadF32(p + 4, l.valign), q.z = loadF32(p + 8, l.valign);
            q.w = loadF32(p + 12, l.valign);
        }
        break;
    case 3: // SHORT_XYZ: raw unsigned short -> float (SimpleSphericalParticles.h:107-112)
        q.x = static_cast<float>(loadU16(p)), q.y = static_cast<float>(loadU16(p + 2)), q.z = static_cast<float>(loadU16(p + 4));
        break;
    case 4: // DOUBLE_XYZ narrowed per component
        q.x = static_cast<float>(loadF64(p, l.valign)), q.y = static_cast<float>(loadF64(p + 8, l.valign));
        q.z = static_cast<float>(loadF64(p + 16, l.valign));
        break;
    default: q.x = q.y = q.z = 0.0f;
    }
    return q;
}

/** Colour accessors cr/cg/cb/ca as floats (SimpleSphericalParticles.h:123-178). */
__device__ __forceinline__ float4 fetchColourRaw(const ListDev& l, unsigned long long j) {
    const char* p = l.col + j * l.cstride;
    float4 c;
    switch (l.ctype) {
    case 1: c = make_float4((float)(unsigned char)p[0], (float)(unsigned char)p[1], (float)(unsigned char)p[2], 255.0f); break;
    case 2: c = make_float4((float)(unsigned char)p[0], (float)(unsigned char)p[1], (float)(unsigned char)p[2], (float)(unsigned char)p[3]); break;
    case 3: c = make_float4(loadF32(p, l.calign), loadF32(p + 4, l.calign), loadF32(p + 8, l.calign), 1.0f); break;
    case 4: c = make_float4(loadF32(p, l.calign), loadF32(p + 4, l.calign), loadF32(p + 8, l.calign), loadF32(p + 12, l.calign)); break;
    case 5: c = make_float4(loadF32(p, l.calign), 0.0f, 0.0f, 0.0f); break;
    case 6: c = make_float4((float)loadU16(p), (float)loadU16(p + 2), (float)loadU16(p + 4), (float)loadU16(p + 6)); break;
    case 7: c = make_float4(static_cast<float>(loadF64(p, l.calign)), 0.0f, 0.0f, 0.0f); break;
    default: c = make_float4(l.gcol[0], l.gcol[1], l.gcol[2], l.gcol[3]);
    }
    return c;
}

/** Direction accessors dx/dy/dz as floats (ParticlesToDensity.cpp:484-486,497-499). */
__device__ __forceinline__ float4 fetchDirection(const ListDev& l, unsigned long long j) {
    if (!l.dir) return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    const char* p = l.dir + j * l.dstride;
    return make_float4(loadF32(p, l.dalign), loadF32(p + 4, l.dalign), loadF32(p + 8, l.dalign), 0.0f);
}

/** QuickSurf's colour conversion (QuickSurf.cpp:511-577): everything to [0,1] RGB. */
__device__ __forceinline__ float4 quicksurfColour(const ListDev& l, float4 c) {
    if (l.ctype == 1 || l.ctype == 2 || l.ctype == 6) {
        c.x = __fdiv_rn(c.x, 255.0f), c.y = __fdiv_rn(c.y, 255.0f), c.z = __fdiv_rn(c.z, 255.0f);
    } else if (l.ctype == 5 || l.ctype == 7) {
        const float v = __fdiv_rn(__fsub_rn(c.x, l.irange[0]), __fsub_rn(l.irange[1], l.irange[0]));
        c.x = c.y = c.z = v;
    }
    return c;
}

/** Home voxel: static_cast<int>((p - min) / sliceDist), individually rounded fp32 ops, truncation
 *  (ParticlesToDensity.cpp:563-568).  cvt.rzi saturates where C++ is undefined. */
__device__ __forceinline__ int homeVoxel(float p, float mn, float sd) {
    return __float2int_rz(__fdiv_rn(__fsub_rn(p, mn), sd));
}
/** filterSize = static_cast<int>(std::ceil(rad / sliceDist)) (:573-575). */
__device__ __forceinline__ int filterSize(float rad, float sd) {
    return __float2int_rz(ceilf(__fdiv_rn(rad, sd)));
}

} // namespace mms
