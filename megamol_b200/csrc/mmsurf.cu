// mmsurf.cu -- context management and the C ABI of libmmsurf (see include/mmsurf.h).
// Built for sm_100a only; there is no CPU path in this library.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/mmsurf.h"
#include "common.cuh"
#include "scan.cuh"
#include "bin.cuh"
#include "density.cuh"
#include "mc.cuh"
#include "mc_emit_v4.cuh" // independent cross-check of mc_emit_kernel (MMS_EMIT_V4=1, tests/test_gpu_variants.py)
#include "mc_indexed.cuh"
#include "mt.cuh"
#include "route.cuh"

using namespace mms;

static_assert(sizeof(mms_list) == 56 && sizeof(mms_grid) == 48 && sizeof(mms_params) == 40 && sizeof(mms_timings) == 32,
    "C ABI struct layout changed: update include/mmsurf.h users (megamol_b200/api.py, plugin/b200surf)");

namespace {

std::string g_createError;

/** Driver entry points of the virtual-memory API (exportable allocations), looked up through the runtime: the library links cudart only. */
struct Vmm {
    CUresult (*create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*reserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*addrFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*granularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*exportHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*importHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    bool ok = false;
    static const Vmm& get() {
        static const Vmm v = [] {
            Vmm t;
            auto load = [](const char* name, auto& fn) {
                void* p = nullptr;
                cudaDriverEntryPointQueryResult q{};
                if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
                    cudaGetLastError();
                    p = nullptr;
                }
                fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
                return p != nullptr;
            };
            bool all = load("cuMemCreate", t.create);
            all &= load("cuMemRelease", t.release);
            all &= load("cuMemAddressReserve", t.reserve);
            all &= load("cuMemAddressFree", t.addrFree);
            all &= load("cuMemMap", t.map);
            all &= load("cuMemUnmap", t.unmap);
            all &= load("cuMemSetAccess", t.setAccess);
            all &= load("cuMemGetAllocationGranularity", t.granularity);
            all &= load("cuMemExportToShareableHandle", t.exportHandle);
            all &= load("cuMemImportFromShareableHandle", t.importHandle);
            t.ok = all;
            return t;
        }();
        return v;
    }
    static CUmemAllocationProp prop(int device) {
        CUmemAllocationProp pr{};
        pr.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        pr.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        pr.location.id = device;
        pr.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        return pr;
    }
    /** Maps an allocation handle read-write for `device`; returns nullptr on failure. */
    static void* mapHandle(CUmemGenericAllocationHandle h, size_t bytes, int device) {
        const Vmm& v = get();
        CUdeviceptr va = 0;
        if (v.reserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) return nullptr;
        if (v.map(va, bytes, 0, h, 0) != CUDA_SUCCESS) {
            v.addrFree(va, bytes);
            return nullptr;
        }
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (v.setAccess(va, bytes, &acc, 1) != CUDA_SUCCESS) {
            v.unmap(va, bytes);
            v.addrFree(va, bytes);
            return nullptr;
        }
        return reinterpret_cast<void*>(va);
    }
};

/** Grow-only device buffer.  `shareable` buffers (mms_share_enable) come from the virtual-memory API with a POSIX file-descriptor handle,
 *  so that a renderer can import them (GL_EXT_memory_object_fd / VkImportMemoryFdInfoKHR / cuMemImportFromShareableHandle). */
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool shareable = false;
    int device = 0;
    CUmemGenericAllocationHandle handle = 0;
    bool allocShared(size_t bytes) {
        const Vmm& v = Vmm::get();
        if (!v.ok) return false;
        const CUmemAllocationProp pr = Vmm::prop(device);
        size_t gran = 0;
        if (v.granularity(&gran, &pr, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || gran == 0) return false;
        const size_t want = (bytes + gran - 1) / gran * gran;
        if (v.create(&handle, want, &pr, 0) != CUDA_SUCCESS) return false;
        p = Vmm::mapHandle(handle, want, device);
        if (!p) {
            v.release(handle);
            handle = 0;
            return false;
        }
        cap = want;
        return true;
    }
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        size_t want = std::max(bytes, cap + cap / 2);
        release();
        if (shareable) {
            if (allocShared(want) || allocShared(bytes)) return true;
            cap = 0, p = nullptr;
            return false;
        }
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            if (cudaMalloc(&p, want) != cudaSuccess) {
                cudaGetLastError();
                cap = 0;
                p = nullptr;
                return false;
            }
        }
        cap = want;
        return true;
    }
    void release() {
        if (p && handle) {
            const Vmm& v = Vmm::get();
            v.unmap(reinterpret_cast<CUdeviceptr>(p), cap);
            v.addrFree(reinterpret_cast<CUdeviceptr>(p), cap);
            v.release(handle);
        } else if (p)
            cudaFree(p);
        p = nullptr;
        handle = 0;
        cap = 0;
    }
    template<class T> T* as() { return static_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t bytes) {
        if (bytes <= cap) return true;
        if (p) cudaFreeHost(p);
        p = nullptr;
        const size_t want = std::max(bytes, cap + cap / 2);
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            cap = 0;
            p = nullptr;
            return false;
        }
        cap = want;
        return true;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template<class T> T* as() { return static_cast<T*>(p); }
};

/** Upload arena: device chunks that persist across frames (no cudaMalloc / cudaFree in steady state). Two arenas
 *  alternate so that the H2D copy of frame k+1 (copy stream) overlaps the kernels of frame k. */
struct Arena {
    struct Chunk {
        char* p;
        size_t cap, used;
    };
    std::vector<Chunk> chunks;
    cudaEvent_t consumed = nullptr; // recorded on the compute stream once the raw input has been re-written into cell order
    bool consumedSet = false;
    bool waited = true;             // the copy stream already waits for `consumed` in this frame
    char* alloc(size_t bytes, size_t mis) {
        for (auto& ch : chunks) {
            const size_t off = (ch.used + 255) & ~size_t(255);
            if (off + mis + bytes + 16 <= ch.cap) {
                ch.used = off + mis + bytes;
                return ch.p + off + mis;
            }
        }
        size_t want = std::max<size_t>(bytes + mis + 512, size_t(32) << 20);
        if (!chunks.empty()) want = std::max(want, chunks.back().cap + chunks.back().cap / 2);
        void* p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            cudaGetLastError();
            want = bytes + mis + 512;
            if (cudaMalloc(&p, want) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        chunks.push_back({static_cast<char*>(p), want, mis + bytes});
        return static_cast<char*>(p) + mis;
    }
    void reset() {
        for (auto& ch : chunks) ch.used = 0;
    }
    void release() {
        for (auto& ch : chunks) cudaFree(ch.p);
        chunks.clear();
    }
};

enum Ev { EV_H2D0, EV_H2D1, EV_BIN0, EV_BIN1, EV_DEN1, EV_NRM0, EV_NRM1, EV_MC0, EV_MC1, EV_DV0, EV_DV1, EV_DM0, EV_DM1, EV_T0, EV_T1, EV_EMIT0, EV_COUNT };

} // namespace

/** Host -> device copies of PAGEABLE memory (what a MegaMol data source hands out: std::vector / RawStorage): the driver stages those
 *  through one bounce buffer on the calling thread (~10 GB/s).  Large ones are instead cut into chunks that a few host threads copy
 *  into pinned slots of their own and send on streams of their own: the memcpy of one chunk overlaps the DMA of another.  Pinned or
 *  registered sources and small copies take the plain cudaMemcpyAsync. */
struct HostStager {
    static constexpr int T = 4;                      // host threads
    static constexpr size_t CH = 4u << 20;           // chunk size
    static constexpr size_t kMinBytes = 16u << 20;   // smaller copies: not worth the threads
    cudaStream_t st[T] = {};
    void* pin[T][2] = {};
    cudaEvent_t slotFree[T][2] = {};
    bool slotUsed[T][2] = {};
    cudaEvent_t start = nullptr, done[T] = {};
    bool ready = false, broken = false;
    bool init() {
        if (ready || broken) return ready;
        bool ok = cudaEventCreateWithFlags(&start, cudaEventDisableTiming) == cudaSuccess;
        for (int t = 0; t < T && ok; ++t) {
            ok = cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&done[t], cudaEventDisableTiming) == cudaSuccess;
            for (int k = 0; k < 2 && ok; ++k)
                ok = cudaMallocHost(&pin[t][k], CH) == cudaSuccess && cudaEventCreateWithFlags(&slotFree[t][k], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) {
            cudaGetLastError();
            release();
            broken = true;
            return false;
        }
        return ready = true;
    }
    void release() {
        for (int t = 0; t < T; ++t) {
            if (st[t]) cudaStreamSynchronize(st[t]);
            for (int k = 0; k < 2; ++k) {
                if (pin[t][k]) cudaFreeHost(pin[t][k]);
                if (slotFree[t][k]) cudaEventDestroy(slotFree[t][k]);
                pin[t][k] = nullptr, slotFree[t][k] = nullptr, slotUsed[t][k] = false;
            }
            if (done[t]) cudaEventDestroy(done[t]);
            if (st[t]) cudaStreamDestroy(st[t]);
            done[t] = nullptr, st[t] = nullptr;
        }
        if (start) cudaEventDestroy(start);
        start = nullptr;
        ready = false;
    }
    /** dst[0, bytes) <- src[0, bytes), ordered like a copy enqueued on `order` (it starts after what is in `order` now; work enqueued on
     *  `order` afterwards waits for it).  The source has been read completely when this returns. */
    cudaError_t copy(int device, char* dst, const char* src, size_t bytes, cudaStream_t order) {
        cudaError_t e = cudaEventRecord(start, order);
        if (e != cudaSuccess) return e;
        std::atomic<int> err{static_cast<int>(cudaSuccess)};
        const size_t nchunks = (bytes + CH - 1) / CH;
        auto work = [&](int t) {
            if (cudaSetDevice(device) != cudaSuccess || cudaStreamWaitEvent(st[t], start, 0) != cudaSuccess) {
                err = static_cast<int>(cudaErrorUnknown);
                return;
            }
            size_t round = 0;
            for (size_t c = static_cast<size_t>(t); c < nchunks; c += T, ++round) {
                const int k = static_cast<int>(round & 1u);
                const size_t off = c * CH, n = std::min(CH, bytes - off);
                cudaError_t r = slotUsed[t][k] ? cudaEventSynchronize(slotFree[t][k]) : cudaSuccess; // the slot's previous chunk has left
                if (r == cudaSuccess) {
                    std::memcpy(pin[t][k], src + off, n);
                    r = cudaMemcpyAsync(dst + off, pin[t][k], n, cudaMemcpyHostToDevice, st[t]);
                }
                if (r == cudaSuccess) r = cudaEventRecord(slotFree[t][k], st[t]);
                if (r != cudaSuccess) {
                    err = static_cast<int>(r);
                    return;
                }
                slotUsed[t][k] = true;
            }
            const cudaError_t r = cudaEventRecord(done[t], st[t]);
            if (r != cudaSuccess) err = static_cast<int>(r);
        };
        std::thread th[T];
        for (int t = 1; t < T; ++t) th[t] = std::thread(work, t);
        work(0);
        for (int t = 1; t < T; ++t) th[t].join();
        if (err.load() != static_cast<int>(cudaSuccess)) return static_cast<cudaError_t>(err.load());
        for (int t = 0; t < T; ++t)
            if ((e = cudaStreamWaitEvent(order, done[t], 0)) != cudaSuccess) return e;
        return cudaSuccess;
    }
};

struct mms_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    mms_grid grid{};
    bool haveGrid = false;
    int z0 = 0, nz = 0, cellZ0 = 0, cellNz = 0;
    mms_params params{};
    std::vector<ListDev> lists;
    Arena arena[2];
    int arenaCur = 0;
    cudaStream_t copyStream = nullptr;
    HostStager stager;
    cudaEvent_t uploadDone = nullptr;
    cudaEvent_t countReady = nullptr; // recorded behind the count's publish kernel: the host waits for THIS, not for the whole stream
    cudaEvent_t volReady = nullptr, volCopied = nullptr; // mms_prefetch_density: volume final on the compute stream / host copy complete
    bool volPrefetched = false;
    bool uploadPending = false;
    unsigned long long nparticles = 0;
    bool haveDensity = false, haveMesh = false, normalized = false;
    unsigned long long ntris = 0;
    unsigned long long launches = 0;
    int cshift = 2, reach = 2;
    int czBase = 0, czCount = 0; // cell layers this slab's cell arrays cover (czCount == 0: all), see Geo
    bool useGather = false, haveColour = false, splatV2ok = false;
    bool haveVector = false; // aggregator 2: rgb = vector volume, vol = |v|, dirVol = unit directions
    McGeo mcGeo{};
    cudaStream_t ownStream = nullptr;
    DevBuf rangeBuf; // {-min, max} as floats for device-side all-reduce + normalise
    bool haveCount = false, meshExternal = false, countPending = false;
    float qsAc = 0.0f;    // MMS_MODE_QS_GAUSS_REFCELLS: the reference's acceleration grid of this frame
    int qsCells[3] = {1, 1, 1};
    const float* adoptedVol = nullptr; // mms_adopt_density: another context's volume (and colour volume), by reference
    const float* adoptedRgb = nullptr;
    cudaEvent_t adoptReady = nullptr;
    const float* isoVol() { return adoptedVol ? adoptedVol : vol.as<float>(); }
    const float* isoRgb() { return adoptedVol ? adoptedRgb : rgb.as<float>(); }
    int isoMode = MMS_ISO_MARCHING_CUBES, countMode = MMS_ISO_MARCHING_CUBES;
    MtGeo mtGeo{};

    DevBuf routeCounts, routeOffsets, routeTile;
    DevBuf haloBuf, haloCounters; // mms_halo_*: receive buffer (float4 records) and counter block of this slab
    uint64_t haloCap = 0;
    bool haloColour = false;      // the halves carry cap colours behind their cap records (set by mms_halo_buffers from the parameters)
    unsigned haloFrame = 0;       // parity selects the counter word of the current frame
    int haloWaitPeers = -1;       // mms_halo_wait: arrivals the stream has to see before the received list is read (-1: nothing pending)
    PinBuf hRoute;
    DevBuf cellCount, cellStart, cursor, tileSums, recsA, recsB, auxA, auxB, vol, rgb, segCount, segOffset, meshPos, meshNrm,
        meshCol, triCount, home, dstate, dirVol, rmaxBuf, bigCells, s3Tables, vertCount, vertOffset, vertRec, meshIdx, cellOf;
    PinBuf hIdx;
    bool meshIndexed = false, countIndexed = false; // mms_set_mesh_indexed: the mode in force / the mode the last count ran in
    unsigned long long nverts = 0;                  // indexed mesh: vertices of the last count
    Geo s3Geo{};          // density_splat3_kernel: the geometry its cell tables (s3Tables) were built for
    int s3Reach = -1;
    std::vector<unsigned char> s3Host;
    PinBuf hState, hVol, hRgb, hPos, hNrm, hCol, hHome, hTri, hDir;
    cudaEvent_t ev[EV_COUNT]{};
    bool evSet[EV_COUNT]{};
    int smCount = 148;
    int countSlots = 0; // resident mc_count_kernel blocks of the device (occupancy query, once)

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    void rec(Ev e) {
        cudaEventRecord(ev[e], stream);
        evSet[e] = true;
    }
};

namespace {

struct DeviceGuard { // the host application may have another device current (SURVEY 8b threading)
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != prev) cudaSetDevice(prev);
    }
};

#define MMS_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return (ctx)->fail(MMS_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

const unsigned kVertSize[5] = {0, 12, 16, 6, 24};
const unsigned kColSize[8] = {0, 3, 4, 12, 16, 4, 8, 8};

bool isDevicePointer(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

/** Host -> device upload on the context's copy stream; pageable sources of 16 MB and more go through the HostStager. */
cudaError_t uploadHost(mms_ctx* c, char* dst, const char* src, size_t bytes) {
    if (bytes >= HostStager::kMinBytes && !getenv("MMS_NO_STAGER")) {
        cudaPointerAttributes a{};
        bool pageable = false;
        if (cudaPointerGetAttributes(&a, src) != cudaSuccess) cudaGetLastError(), pageable = true; // (older drivers: an error for plain host memory)
        else pageable = a.type == cudaMemoryTypeUnregistered;
        if (pageable && c->stager.init()) return c->stager.copy(c->device, dst, src, bytes, c->copyStream);
    }
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copyStream);
}

int alignOf(const void* p, unsigned stride, int want) {
    const uintptr_t u = reinterpret_cast<uintptr_t>(p);
    for (int a = want; a > 1; a >>= 1)
        if (u % a == 0 && stride % a == 0) return a;
    return 1;
}

Geo makeGeo(const mms_ctx* c) {
    Geo g{};
    for (int a = 0; a < 3; ++a) {
        g.mn[a] = c->grid.min[a];
        g.s[a] = c->grid.res[a];
        // sliceDist = rangeOS / static_cast<float>(s - 1)  (ParticlesToDensity.cpp:430-432), fp32 division on the host
        volatile float range = c->grid.extent[a];
        volatile float denom = static_cast<float>(c->grid.res[a] - 1);
        volatile float sd = range / denom;
        g.sd[a] = sd;
        g.isd[a] = 1.0f / sd;
        g.cyc[a] = c->grid.cyclic[a] != 0;
    }
    g.z0 = c->z0, g.nz = c->nz;
    g.cshift = c->cshift;
    for (int a = 0; a < 3; ++a) g.nc[a] = (g.s[a] + (1 << g.cshift) - 1) >> g.cshift;
    g.czBase = c->czCount > 0 ? c->czBase : 0, g.czCount = c->czCount > 0 ? c->czCount : g.nc[2];
    g.sigma = c->params.sigma;
    g.agg = c->params.aggregator;
    g.mode = c->params.mode == MMS_MODE_P2D_BUMP ? 0 : 1; // both Gaussian modes are mode 1 on the device; qsAc > 0 selects the reference cells
    g.qsAc = c->qsAc, g.qsInvAc = c->qsAc > 0.0f ? 1.0f / c->qsAc : 0.0f;
    for (int a = 0; a < 3; ++a) g.qsCells[a] = c->qsCells[a];
    g.radscale = c->params.radscale;
    g.gausslim = c->params.gausslim;
    g.colour = c->params.colour;
    return g;
}

int gridFor(unsigned long long n, int threads, int cap) {
    unsigned long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return static_cast<int>(std::min<unsigned long long>(b, static_cast<unsigned long long>(cap)));
}

__global__ void init_state_kernel(DevState* st) {
    st->rmaxBits = 0u;
    st->kept = 0u;
    st->minKey = 0xffffffffu;
    st->maxKey = 0u;
    st->totalTris = 0ull;
    st->pad[0] = 0u; // error flag
    st->pad[1] = 0u;
    st->nBig = 0u;
    st->bigNext = 0u;
    st->totalVerts = 0ull;
}

/** The state block goes to the host by a store into mapped pinned memory, not by a copy-engine transfer: a cudaMemcpyAsync would queue
 *  behind a volume read-back that is in flight on the same DMA engine (mms_prefetch_density) and stall the one host round trip of the
 *  isosurface (the triangle count) until that copy is done. */
__global__ void publish_state_kernel(const DevState* __restrict__ st, DevState* __restrict__ hostMapped) {
    *hostMapped = *st;
    __threadfence_system();
}

__global__ void range_to_float_kernel(const DevState* __restrict__ st, float* __restrict__ out) {
    out[0] = -keyFloat(st->minKey);
    out[1] = keyFloat(st->maxKey);
}

__global__ void __launch_bounds__(256) normalize_ptr_kernel(float* __restrict__ vol, size_t n, const float* __restrict__ negminMax) {
    const float mn = -negminMax[0], mx = negminMax[1];
    const float rcp = __fdiv_rn(1.0f, __fsub_rn(mx, mn));
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t n4 = n / 4;
    float4* v4 = reinterpret_cast<float4*>(vol);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = v4[i];
        v.x = __fmul_rn(__fsub_rn(v.x, mn), rcp), v.y = __fmul_rn(__fsub_rn(v.y, mn), rcp);
        v.z = __fmul_rn(__fsub_rn(v.z, mn), rcp), v.w = __fmul_rn(__fsub_rn(v.w, mn), rcp);
        v4[i] = v;
    }
    for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        vol[i] = __fmul_rn(__fsub_rn(vol[i], mn), rcp);
}

__global__ void __launch_bounds__(256) normalize_state_kernel(float* __restrict__ vol, size_t n, const DevState* __restrict__ st) {
    const float mn = keyFloat(st->minKey), mx = keyFloat(st->maxKey);
    const float rcp = __fdiv_rn(1.0f, __fsub_rn(mx, mn)); // 1.0f / (maxDens - minDens) (:677)
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t n4 = n / 4; // cudaMalloc'ed volume: 16-byte aligned
    float4* v4 = reinterpret_cast<float4*>(vol);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = v4[i];
        v.x = __fmul_rn(__fsub_rn(v.x, mn), rcp), v.y = __fmul_rn(__fsub_rn(v.y, mn), rcp);
        v.z = __fmul_rn(__fsub_rn(v.z, mn), rcp), v.w = __fmul_rn(__fsub_rn(v.w, mn), rcp);
        v4[i] = v;
    }
    for (size_t i = n4 * 4 + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        vol[i] = __fmul_rn(__fsub_rn(vol[i], mn), rcp);
}


/** Global density range of a slab group: every device reads all the slabs' {-min, max} (peer memory) and keeps the maximum of each. */
struct RangePeers {
    const float* r[16];
    int n;
};
__global__ void range_combine_kernel(RangePeers p, float* __restrict__ out) {
    float a = -INFINITY, b = -INFINITY;
    for (int i = 0; i < p.n; ++i) a = fmaxf(a, p.r[i][0]), b = fmaxf(b, p.r[i][1]);
    out[0] = a, out[1] = b;
}

size_t emitSmemBytes(bool colour) { return sizeof(McEmitShared) + 128 + (colour ? E_RECS * sizeof(float4) : 0); }
size_t emitV4SmemBytes(bool colour) { return sizeof(v4::McEmitV4Shared) + 128 + (colour ? v4::E_EDGES * sizeof(float4) : 0); }

/** 3-D TMA descriptor over the slab volume (x fastest), box = one 40 x 11 plane tile of mc_emit_kernel.  Returns false where TMA's
 *  rules do not hold (row pitch not a multiple of 16 bytes) or the driver entry point is missing: the kernel's cp.async variant runs. */
bool makeVolumeTensorMap(CUtensorMap* map, const float* vol, int sx, int sy, int nz) {
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q{};
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            fn = nullptr;
        }
        return reinterpret_cast<Encode>(fn);
    }();
    if (!encode || sx % 4 != 0 || (reinterpret_cast<uintptr_t>(vol) & 15u) || getenv("MMS_NO_TMA")) return false;
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(sx), static_cast<cuuint64_t>(sy), static_cast<cuuint64_t>(nz)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(sx) * 4, static_cast<cuuint64_t>(sx) * sy * 4};
    const cuuint32_t box[3] = {EPITCH, EHY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(vol), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
} // namespace

extern "C" {

int mms_version(void) { return 1; }

const char* mms_last_error(const mms_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int mms_create(mms_ctx** out, const mms_config* cfg) {
    if (!out) return MMS_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_createError = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                        " (libmmsurf has no CPU fallback)";
        cudaGetLastError();
        return MMS_ERR_CUDA;
    }
    const int dev = cfg ? cfg->device : 0;
    if (dev < 0 || dev >= ndev) {
        g_createError = "device ordinal out of range";
        return MMS_ERR_INVALID;
    }
    DeviceGuard guard(dev);
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, dev);
    if (prop.major < 10) {
        g_createError = std::string("device '") + prop.name + "' is not sm_100+; libmmsurf ships sm_100a code only";
        return MMS_ERR_CUDA;
    }
    auto* c = new mms_ctx();
    c->device = dev;
    c->smCount = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_createError = "cudaStreamCreate failed";
        delete c;
        return MMS_ERR_CUDA;
    }
    for (auto& ev : c->ev) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->uploadDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->volReady, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->volCopied, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->adoptReady, cudaEventDisableTiming);
    for (auto& a : c->arena) cudaEventCreateWithFlags(&a.consumed, cudaEventDisableTiming);
    c->params.mode = MMS_MODE_P2D_BUMP;
    c->params.sigma = 1.0f;
    c->params.normalize = 1;
    c->params.radscale = 1.0f;
    c->params.gausslim = 3.0f;
    cudaFuncSetAttribute(density_splat_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplatShared));
    cudaFuncSetAttribute(density_splat_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplatShared));
    cudaFuncSetAttribute(density_splat3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Splat3Shared));
    cudaFuncSetAttribute(mc_emit_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitSmemBytes(false));
    cudaFuncSetAttribute(mc_emit_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitSmemBytes(false));
    cudaFuncSetAttribute(mc_emit_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitSmemBytes(true));
    cudaFuncSetAttribute(mc_emit_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitSmemBytes(true));
    cudaFuncSetAttribute(v4::mc_emit_v4_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitV4SmemBytes(false));
    cudaFuncSetAttribute(v4::mc_emit_v4_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitV4SmemBytes(false));
    cudaFuncSetAttribute(v4::mc_emit_v4_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitV4SmemBytes(true));
    cudaFuncSetAttribute(v4::mc_emit_v4_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)emitV4SmemBytes(true));
    if (!c->dstate.ensure(sizeof(DevState)) || !c->hState.ensure(sizeof(DevState))) {
        g_createError = "allocation of the state block failed";
        delete c;
        return MMS_ERR_NOMEM;
    }
    if (cudaGetLastError() != cudaSuccess) {
        g_createError = "kernel attribute set-up failed (is this an sm_100a build on an sm_100 device?)";
        delete c;
        return MMS_ERR_CUDA;
    }
    *out = c;
    return MMS_OK;
}

int mms_clear_particles(mms_ctx* c) {
    if (!c) return MMS_ERR_INVALID;
    // no synchronisation, no frees: kernels already launched hold their pointers by value; the next pushes go to the OTHER
    // arena, whose previous contents were consumed two frames ago (guarded by its `consumed` event)
    c->lists.clear();
    c->nparticles = 0;
    c->arenaCur ^= 1;
    c->arena[c->arenaCur].reset();
    c->arena[c->arenaCur].waited = false;
    return MMS_OK;
}

int mms_destroy(mms_ctx* c) {
    if (!c) return MMS_ERR_INVALID;
    {
        DeviceGuard guard(c->device);
        mms_clear_particles(c);
        for (DevBuf* b : {&c->cellCount, &c->cellStart, &c->cursor, &c->tileSums, &c->recsA, &c->recsB, &c->auxA, &c->auxB, &c->vol,
                 &c->rgb, &c->segCount, &c->segOffset, &c->meshPos, &c->meshNrm, &c->meshCol, &c->triCount, &c->home, &c->dstate, &c->routeCounts, &c->routeOffsets, &c->routeTile, &c->rangeBuf, &c->dirVol, &c->rmaxBuf, &c->bigCells, &c->haloBuf, &c->haloCounters, &c->s3Tables, &c->vertCount, &c->vertOffset, &c->vertRec, &c->meshIdx, &c->cellOf})
            b->release();
        for (PinBuf* b : {&c->hState, &c->hVol, &c->hRgb, &c->hPos, &c->hNrm, &c->hCol, &c->hHome, &c->hTri, &c->hRoute, &c->hDir, &c->hIdx}) b->release();
        cudaStreamSynchronize(c->stream);
        cudaStreamSynchronize(c->copyStream);
        c->stager.release();
        for (auto& a : c->arena) {
            a.release();
            cudaEventDestroy(a.consumed);
        }
        cudaEventDestroy(c->uploadDone);
        if (c->countReady) cudaEventDestroy(c->countReady);
        cudaEventDestroy(c->volReady);
        cudaEventDestroy(c->volCopied);
        cudaEventDestroy(c->adoptReady);
        cudaStreamDestroy(c->copyStream);
        for (auto& ev : c->ev) cudaEventDestroy(ev);
        cudaStreamDestroy(c->ownStream ? c->ownStream : c->stream);
    }
    delete c;
    return MMS_OK;
}

int mms_set_grid(mms_ctx* c, const mms_grid* g) {
    if (!c || !g) return MMS_ERR_INVALID;
    for (int a = 0; a < 3; ++a) {
        if (g->res[a] < 2) return c->fail(MMS_ERR_INVALID, "resolution must be >= 2 per axis (sliceDist = extent/(res-1))");
        if (!(g->extent[a] > 0.0f) || !std::isfinite(g->extent[a]) || !std::isfinite(g->min[a]))
            return c->fail(MMS_ERR_INVALID, "bounding box extent must be positive and finite");
    }
    if (static_cast<unsigned long long>(g->res[0]) * g->res[1] * g->res[2] >= (1ull << 32))
        return c->fail(MMS_ERR_UNSUPPORTED, "volumes of 2^32 voxels or more are not supported");
    c->grid = *g;
    c->adoptedVol = c->adoptedRgb = nullptr;
    c->haveGrid = true;
    c->z0 = 0, c->nz = g->res[2];
    c->cellZ0 = 0, c->cellNz = g->res[2] - 1;
    c->haveDensity = c->haveMesh = c->haveCount = false;
    return MMS_OK;
}

int mms_set_slab(mms_ctx* c, int32_t z0, int32_t nz, int32_t cell_z0, int32_t cell_nz) {
    if (!c || !c->haveGrid) return MMS_ERR_INVALID;
    if (z0 < 0 || nz < 1 || z0 + nz > c->grid.res[2]) return c->fail(MMS_ERR_INVALID, "slab planes out of range");
    if (cell_nz < 0 || cell_z0 < z0 || (cell_nz > 0 && cell_z0 + cell_nz + 1 > z0 + nz))
        return c->fail(MMS_ERR_INVALID, "cell layers must lie inside the slab's planes");
    c->z0 = z0, c->nz = nz, c->cellZ0 = cell_z0, c->cellNz = cell_nz;
    c->haveDensity = c->haveMesh = c->haveCount = false;
    return MMS_OK;
}

int mms_set_params(mms_ctx* c, const mms_params* p) {
    if (!c || !p) return MMS_ERR_INVALID;
    if (p->mode != MMS_MODE_P2D_BUMP && p->mode != MMS_MODE_QS_GAUSS && p->mode != MMS_MODE_QS_GAUSS_REFCELLS)
        return c->fail(MMS_ERR_INVALID, "unknown mode %d", p->mode);
    if (p->mode == MMS_MODE_P2D_BUMP) {
        if (p->aggregator < 0 || p->aggregator > 2) return c->fail(MMS_ERR_INVALID, "unknown aggregator %d", p->aggregator);
        if (!(p->sigma > 0.0f)) return c->fail(MMS_ERR_INVALID, "sigma must be > 0");
    } else {
        if (!(p->radscale > 0.0f) || !(p->gausslim > 0.0f)) return c->fail(MMS_ERR_INVALID, "radscale and gausslim must be > 0");
    }
    c->params = *p;
    return MMS_OK;
}

static int pushLists(mms_ctx* c, int32_t nlists, const mms_list* lists, const void* const* dirs, const uint32_t* dirStrides) {
    if (!c || nlists < 0 || (nlists > 0 && !lists)) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    if (c->lists.size() + nlists > static_cast<size_t>(kMaxLists)) return c->fail(MMS_ERR_UNSUPPORTED, "more than %d particle lists", kMaxLists);
    Arena& arena = c->arena[c->arenaCur];
    cudaEventRecord(c->ev[EV_H2D0], c->copyStream);
    c->evSet[EV_H2D0] = true;
    for (int i = 0; i < nlists; ++i) {
        const mms_list& l = lists[i];
        if (l.vtx_type == MMS_VERT_NONE || l.count == 0) continue; // lists with VERTDATA_NONE are skipped (:464-466)
        if (l.vtx_type < 0 || l.vtx_type > 4 || l.col_type < 0 || l.col_type > 7) return c->fail(MMS_ERR_INVALID, "bad data type in list %d", i);
        if (!l.vtx) return c->fail(MMS_ERR_INVALID, "list %d has no vertex pointer", i);
        ListDev d{};
        d.count = l.count;
        d.base = c->nparticles;
        d.vtype = l.vtx_type;
        d.vstride = l.vtx_stride ? l.vtx_stride : kVertSize[l.vtx_type];
        d.ctype = l.col ? l.col_type : MMS_COL_NONE;
        d.cstride = l.col_stride ? l.col_stride : kColSize[d.ctype];
        d.grad = l.global_radius;
        for (int k = 0; k < 4; ++k) d.gcol[k] = static_cast<float>(l.global_rgba[k]) / 255.0f;
        d.irange[0] = l.irange[0], d.irange[1] = l.irange[1];
        const size_t vbytes = (l.count - 1) * static_cast<size_t>(d.vstride) + kVertSize[l.vtx_type];
        const size_t cbytes = d.ctype ? (l.count - 1) * static_cast<size_t>(d.cstride) + kColSize[d.ctype] : 0;
        const char* hv = static_cast<const char*>(l.vtx);
        const char* hc = static_cast<const char*>(l.col);
        if (isDevicePointer(l.vtx)) {
            d.vtx = hv;
            d.col = hc; // a device-resident list carries device colour pointers too
        } else {
            // one copy for interleaved vertex+colour (the MMPLD layout, MMPLDDataSource.cpp:160,197-198), else two
            const char* lo = hv;
            const char* hi = hv + vbytes;
            const uintptr_t uv = reinterpret_cast<uintptr_t>(hv), uc = reinterpret_cast<uintptr_t>(hc);
            const bool interleaved = d.ctype && uc + d.vstride >= uv && uc < uv + vbytes + d.vstride;
            if (interleaved) lo = std::min(lo, hc), hi = std::max(hi, hc + cbytes);
            // keep the source's alignment modulo 16 so that aligned host data stays aligned on the device
            const size_t mis = reinterpret_cast<uintptr_t>(lo) & 15u;
            if (!arena.waited) { // first upload into this arena for the new frame: its old contents must have been consumed
                if (arena.consumedSet) MMS_CUDA(c, cudaStreamWaitEvent(c->copyStream, arena.consumed, 0));
                arena.waited = true;
            }
            char* dbase = arena.alloc(static_cast<size_t>(hi - lo), mis);
            if (!dbase) return c->fail(MMS_ERR_NOMEM, "device allocation of %zu bytes for list %d failed", static_cast<size_t>(hi - lo), i);
            MMS_CUDA(c, uploadHost(c, dbase, lo, static_cast<size_t>(hi - lo)));
            d.vtx = dbase + (hv - lo);
            if (d.ctype) {
                if (interleaved) {
                    d.col = dbase + (hc - lo);
                } else {
                    char* dc = arena.alloc(cbytes, reinterpret_cast<uintptr_t>(hc) & 15u);
                    if (!dc) return c->fail(MMS_ERR_NOMEM, "device allocation of %zu bytes for colours of list %d failed", cbytes, i);
                    MMS_CUDA(c, uploadHost(c, dc, hc, cbytes));
                    d.col = dc;
                }
            }
            c->uploadPending = true;
        }
        if (dirs && dirs[i]) { // DIRDATA_FLOAT_XYZ; stride 0 = tightly packed (SimpleSphericalParticles.h:504-510)
            d.dstride = dirStrides && dirStrides[i] ? dirStrides[i] : 12u;
            const size_t dbytes = (l.count - 1) * static_cast<size_t>(d.dstride) + 12;
            const char* hd = static_cast<const char*>(dirs[i]);
            if (isDevicePointer(hd)) {
                d.dir = hd;
            } else {
                if (!arena.waited) {
                    if (arena.consumedSet) MMS_CUDA(c, cudaStreamWaitEvent(c->copyStream, arena.consumed, 0));
                    arena.waited = true;
                }
                char* dd = arena.alloc(dbytes, reinterpret_cast<uintptr_t>(hd) & 15u);
                if (!dd) return c->fail(MMS_ERR_NOMEM, "device allocation of %zu bytes for directions of list %d failed", dbytes, i);
                MMS_CUDA(c, uploadHost(c, dd, hd, dbytes));
                d.dir = dd;
                c->uploadPending = true;
            }
            d.dalign = alignOf(d.dir, d.dstride, 4) >= 4 ? 4 : 1;
        }
        d.valign = alignOf(d.vtx, d.vstride, l.vtx_type == MMS_VERT_DOUBLE_XYZ ? 8 : (l.vtx_type == MMS_VERT_FLOAT_XYZR ? 16 : 4));
        if (l.vtx_type == MMS_VERT_DOUBLE_XYZ && d.valign < 8) d.valign = 1;
        if ((l.vtx_type == MMS_VERT_FLOAT_XYZ || l.vtx_type == MMS_VERT_FLOAT_XYZR) && d.valign < 4) d.valign = 1;
        d.calign = d.ctype ? alignOf(d.col, d.cstride, d.ctype == MMS_COL_DOUBLE_I ? 8 : 4) : 4;
        if (d.ctype == MMS_COL_DOUBLE_I && d.calign < 8) d.calign = 1;
        if (d.calign < 4) d.calign = 1;
        c->lists.push_back(d);
        c->nparticles += l.count;
    }
    cudaEventRecord(c->ev[EV_H2D1], c->copyStream);
    c->evSet[EV_H2D1] = true;
    if (c->uploadPending) MMS_CUDA(c, cudaEventRecord(c->uploadDone, c->copyStream));
    // results of the previous compute stay readable: pushing frame k+1 while frame k is being read back is the streaming pattern
    return MMS_OK;
}

int mms_push_particles(mms_ctx* c, int32_t nlists, const mms_list* lists) { return pushLists(c, nlists, lists, nullptr, nullptr); }

int mms_push_particles_dir(mms_ctx* c, int32_t nlists, const mms_list* lists, const void* const* dirs, const uint32_t* dir_strides) {
    return pushLists(c, nlists, lists, dirs, dir_strides);
}

int mms_get_max_radius(mms_ctx* c, float* rmaxOut) {
    if (!c || !rmaxOut) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    float rmax = 0.0f;
    bool perParticle = false;
    for (const ListDev& l : c->lists) {
        if (l.vtype == MMS_VERT_FLOAT_XYZR && !l.radiusBound) perParticle = true;
        else if (l.grad > rmax && std::isfinite(l.grad)) rmax = l.grad;
    }
    if (perParticle) {
        cudaStream_t st = c->stream;
        if (c->uploadPending) MMS_CUDA(c, cudaStreamWaitEvent(st, c->uploadDone, 0)); // compute_density waits again; harmless
        if (!c->rmaxBuf.ensure(16)) return c->fail(MMS_ERR_NOMEM, "device allocation failed");
        MMS_CUDA(c, cudaMemsetAsync(c->rmaxBuf.p, 0, 4, st));
        for (const ListDev& l : c->lists)
            if (l.vtype == MMS_VERT_FLOAT_XYZR && !l.radiusBound) {
                radius_max_kernel<<<gridFor(l.count, 256, c->smCount * 16), 256, 0, st>>>(l, c->rmaxBuf.as<unsigned>());
                ++c->launches;
            }
        unsigned bits = 0;
        MMS_CUDA(c, cudaMemcpyAsync(&bits, c->rmaxBuf.p, 4, cudaMemcpyDeviceToHost, st));
        MMS_CUDA(c, cudaStreamSynchronize(st));
        float r = 0.0f;
        std::memcpy(&r, &bits, 4);
        rmax = std::max(rmax, r);
    }
    *rmaxOut = rmax;
    return MMS_OK;
}

int mms_compute_density(mms_ctx* c) {
    if (!c || !c->haveGrid) return c ? c->fail(MMS_ERR_INVALID, "mms_set_grid has not been called") : MMS_ERR_INVALID;
    if (c->nparticles >= (1ull << 32) - 1) return c->fail(MMS_ERR_UNSUPPORTED, "2^32 or more particles per context");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    c->adoptedVol = c->adoptedRgb = nullptr;
    if (c->uploadPending) {
        MMS_CUDA(c, cudaStreamWaitEvent(st, c->uploadDone, 0));
        c->uploadPending = false;
    }
    if (c->volPrefetched) { // a prefetch copy of the previous volume may still be reading it
        MMS_CUDA(c, cudaStreamWaitEvent(st, c->volCopied, 0));
        c->volPrefetched = false;
    }
    c->rec(EV_BIN0);
    init_state_kernel<<<1, 1, 0, st>>>(c->dstate.as<DevState>());
    ++c->launches;
    // ---- largest support radius -> reach in voxels -> cell size (the colouring needs reach <= cell/2) ----------
    {
        float rmax = 0.0f;
        bool perParticle = false;
        for (const ListDev& l : c->lists) {
            if (l.vtype == MMS_VERT_FLOAT_XYZR && !l.radiusBound) perParticle = true;
            else if (l.grad > rmax && std::isfinite(l.grad)) rmax = l.grad;
        }
        if (perParticle) {
            for (const ListDev& l : c->lists)
                if (l.vtype == MMS_VERT_FLOAT_XYZR && !l.radiusBound) {
                    radius_max_kernel<<<gridFor(l.count, 256, c->smCount * 16), 256, 0, st>>>(l, &c->dstate.as<DevState>()->rmaxBits);
                    ++c->launches;
                }
            MMS_CUDA(c, cudaMemcpyAsync(c->hState.p, c->dstate.p, sizeof(DevState), cudaMemcpyDeviceToHost, st));
            MMS_CUDA(c, cudaStreamSynchronize(st));
            float r = 0.0f;
            const unsigned bits = c->hState.as<DevState>()->rmaxBits;
            std::memcpy(&r, &bits, 4);
            rmax = std::max(rmax, r);
        }
        const Geo g0 = makeGeo(c);
        float epsMax = (c->params.mode == MMS_MODE_P2D_BUMP) ? c->params.sigma * rmax : c->params.gausslim * c->params.radscale * rmax;
        c->qsAc = 0.0f;
        if (c->params.mode == MMS_MODE_QS_GAUSS_REFCELLS) {
            // the reference's acceleration grid, its own fp32 expressions (CUDAQuickSurf.cu:1259-1264, 1279-1282)
            const float gridspacing = g0.sd[0];
            for (int a = 1; a < 3; ++a)
                if (std::fabs(g0.sd[a] - gridspacing) > 1e-5f * gridspacing)
                    return c->fail(MMS_ERR_UNSUPPORTED, "the reference candidate set needs one grid spacing for all axes (QuickSurf's gridspacing)");
            if (c->z0 != 0 || c->nz != c->grid.res[2]) return c->fail(MMS_ERR_UNSUPPORTED, "the reference candidate set is not available on z-slabs");
            float ac = c->params.gausslim * c->params.radscale * rmax;
            if (ac < gridspacing) ac = gridspacing;
            c->qsAc = ac;
            for (int a = 0; a < 3; ++a) c->qsCells[a] = std::max(static_cast<int>((c->grid.res[a] * gridspacing) / ac), 1);
            epsMax = 2.0f * ac + gridspacing; // an atom up to two cell sizes (+ a voxel) away from a tile can sit in one of its candidate cells
        }
        int need = 1;
        for (int a = 0; a < 3; ++a) need = std::max(need, static_cast<int>(std::ceil(epsMax / g0.sd[a] + 0.02f)));
        // aggregator 2 keeps four sums per voxel: the register-accumulating gather kernel has them (as the QuickSurf colour sums)
        c->useGather = c->params.mode != MMS_MODE_P2D_BUMP || need > 8 || (c->params.mode == MMS_MODE_P2D_BUMP && c->params.aggregator == 2);
        // a support box wider than a periodic axis: a voxel receives the same particle through several images (the reference's loop over the
        // un-wrapped box, ParticlesToDensity.cpp:583-603) -- the gather kernel enumerates them, the splat kernels keep one image per voxel
        for (int a = 0; a < 3; ++a)
            if (c->grid.cyclic[a] && 2 * need + 1 > c->grid.res[a]) c->useGather = true;
        if (c->useGather) {
            if (need > 96) return c->fail(MMS_ERR_UNSUPPORTED, "kernel support of %d voxels per side is not supported", need);
            c->cshift = need <= 16 ? 3 : 4; // gather: cells only organise the candidate stream
        } else if (need <= 2) c->cshift = 2;
        else if (need <= 4) c->cshift = 3;
        else c->cshift = 4;
        c->reach = need;
        // ---- the cell layers this slab has to hold: everything within the largest filter size (what the binning keeps) or reach (what the
        //      density kernels ask for) of its planes, + a voxel of margin; wrapped on a periodic axis -----------------------------------
        {
            float radEff = c->params.mode == MMS_MODE_P2D_BUMP ? rmax : c->params.gausslim * c->params.radscale * rmax;
            if (c->qsAc > 0.0f) radEff = 2.0f * c->qsAc;
            const int fz = static_cast<int>(std::ceil(radEff / g0.sd[2])) + 2;
            const int zr = std::max(fz, need + 1);
            const int s = c->grid.res[2], nc2 = (s + (1 << c->cshift) - 1) >> c->cshift, cs = c->cshift;
            int a = c->z0 - zr, b = c->z0 + c->nz - 1 + zr;
            if (!c->grid.cyclic[2]) {
                a = std::max(a, 0), b = std::min(b, s - 1);
                c->czBase = a >> cs, c->czCount = (b >> cs) - c->czBase + 1;
            } else if (b - a + 1 >= s) {
                c->czBase = 0, c->czCount = nc2;
            } else {
                // the planes a..b wrap around the axis; a range that comes back to the layer it started in has visited every layer
                const int first = (a < 0 ? a + s : a) >> cs;
                int layers = 1, prev = first;
                for (int v = a + 1; v <= b && layers < nc2; ++v) {
                    const int cz = (v < 0 ? v + s : v >= s ? v - s : v) >> cs;
                    if (cz != prev) ++layers, prev = cz;
                }
                c->czBase = first, c->czCount = layers;
            }
            if (c->czCount >= nc2 || c->czCount <= 0) c->czBase = 0, c->czCount = nc2;
        }
        // tight support box = the integers of an interval of length 2 eps / sliceDist (+ rounding slop): at most 3 per axis?
        c->splatV2ok = true;
        for (int a = 0; a < 3; ++a)
            if (!(2.0f * epsMax / g0.sd[a] < 2.95f)) c->splatV2ok = false;
    }
    const Geo g = makeGeo(c);
    for (ListDev& l : c->lists) { // global-radius lists: (int)ceil(rad / sliceDist) once on the host, same fp32 operations (:573-575)
        for (int a = 0; a < 3; ++a) {
            volatile float rad = (g.mode == 0) ? l.grad : g.gausslim * g.radscale * l.grad;
            volatile float q = rad / g.sd[a];
            l.gf[a] = static_cast<int>(std::ceil(q)) + (g.mode == 0 ? 0 : 1);
        }
    }
    const size_t ncells = static_cast<size_t>(g.nc[0]) * g.nc[1] * g.czCount; // (a slab: only the cell layers that can reach it)
    const size_t nvox = static_cast<size_t>(g.s[0]) * g.s[1] * g.nz;
    const bool colour = g.mode == 1 && c->params.colour != 0;
    if (g.qsAc > 0.0f)
        for (ListDev& l : c->lists)
            for (int a = 0; a < 3; ++a) l.gf[a] = static_cast<int>(std::ceil(2.0f * g.qsAc / g.sd[a])) + 1;
    const bool vector = g.mode == 0 && g.agg == 2;
    const int auxN = (g.mode == 0 && g.agg == 1) ? 1 : ((colour || vector) ? 4 : 0);
    if ((colour || vector) && !c->rgb.ensure(nvox * 12)) return c->fail(MMS_ERR_NOMEM, "device allocation failed (colour / vector volume)");
    if (vector && !c->dirVol.ensure(nvox * 12)) return c->fail(MMS_ERR_NOMEM, "device allocation failed (direction volume)");
    const size_t n = static_cast<size_t>(c->nparticles);
    const unsigned ntiles = static_cast<unsigned>((ncells + kScanTile - 1) / kScanTile);
    if (!c->cellCount.ensure(ncells * 4) || !c->cellStart.ensure((ncells + 1) * 4) || !c->cursor.ensure(ncells * 4) ||
        !c->tileSums.ensure(std::max<size_t>(ntiles, 1) * 4) || !c->recsA.ensure(std::max<size_t>(n, 1) * 16) ||
        !c->recsB.ensure(std::max<size_t>(n, 1) * 16) || !c->vol.ensure(nvox * 4) || !c->bigCells.ensure((n / kBigCell + 2) * 4) ||
        !c->cellOf.ensure(std::max<size_t>(n, 1) * 4))
        return c->fail(MMS_ERR_NOMEM, "device allocation failed (cells %zu, particles %zu, voxels %zu)", ncells, n, nvox);
    if (auxN && (!c->auxA.ensure(std::max<size_t>(n, 1) * 4 * auxN) || !c->auxB.ensure(std::max<size_t>(n, 1) * 4 * auxN)))
        return c->fail(MMS_ERR_NOMEM, "device allocation failed (aux)");
    int* homeOut = nullptr;
    if (c->params.want_home_voxels) {
        if (!c->home.ensure(std::max<size_t>(n, 1) * 12)) return c->fail(MMS_ERR_NOMEM, "device allocation failed (home voxels)");
        homeOut = c->home.as<int>();
    }
    MMS_CUDA(c, cudaMemsetAsync(c->cellCount.p, 0, ncells * 4, st));
    const int cap = c->smCount * 16;
    for (const ListDev& l : c->lists) {
        if (l.countPtr && c->haloWaitPeers >= 0) { // the received halo list: its records and its length must have arrived (mms_halo_wait)
            halo_wait_kernel<<<1, 1, 0, st>>>(c->haloCounters.as<unsigned>() + 2 + (c->haloFrame & 1u), static_cast<unsigned>(c->haloWaitPeers),
                c->haloCounters.as<unsigned>() + (c->haloFrame & 1u));
            ++c->launches;
            c->haloWaitPeers = -1;
        }
        bin_count_kernel<<<gridFor(l.count, 256, cap), 256, 0, st>>>(g, l, c->cellCount.as<unsigned>(), c->dstate.as<DevState>(), homeOut, c->cellOf.as<int>());
        ++c->launches;
    }
    exclusiveScan(c->cellCount.as<unsigned>(), c->cellStart.as<unsigned>(), c->cursor.as<unsigned>(), c->tileSums.as<unsigned>(),
        static_cast<unsigned>(ncells), nullptr, st, c->launches);
    for (const ListDev& l : c->lists) {
        bin_scatter_kernel<<<gridFor(l.count, 256, cap), 256, 0, st>>>(g, l, c->cursor.as<unsigned>(), c->recsA.as<float4>(),
            c->auxA.as<float>(), auxN, c->cellOf.as<int>());
        ++c->launches;
    }
    // the raw input is not needed any more: its arena may be overwritten by the upload of the frame after next
    MMS_CUDA(c, cudaEventRecord(c->arena[c->arenaCur].consumed, st));
    c->arena[c->arenaCur].consumedSet = true;
    if (n > 0) {
        // upper bound n threads; slots >= kept are never claimed, the kernel reads the segment table only
        DevState* ds = c->dstate.as<DevState>();
        cell_order_kernel<<<gridFor(n, 256, 1 << 30), 256, 0, st>>>(g, c->cellStart.as<unsigned>(), c->recsA.as<float4>(), c->auxA.as<float>(),
            c->recsB.as<float4>(), c->auxB.as<float>(), auxN, ds, c->bigCells.as<unsigned>(), &ds->nBig);
        // crowded cells (rare: coarse grids with wide kernels, clustered data) get a merge sort; no crowded cell -> the blocks leave at once
        cell_sort_big_kernel<<<c->smCount, kBigThreads, 0, st>>>(c->cellStart.as<unsigned>(), c->recsA.as<float4>(), c->auxA.as<float>(),
            c->recsB.as<float4>(), c->auxB.as<float>(), auxN, c->bigCells.as<unsigned>(), &ds->nBig, &ds->bigNext);
        c->launches += 2;
    }
    c->rec(EV_BIN1);
    if (c->useGather) {
        dim3 grid((g.s[0] + GT_X - 1) / GT_X, (g.s[1] + GT_Y - 1) / GT_Y, (g.nz + GT_Z - 1) / GT_Z);
        const float4* R = c->recsB.as<float4>();
        const float* A = c->auxB.as<float>();
        const unsigned* CS = c->cellStart.as<unsigned>();
        DevState* DS = c->dstate.as<DevState>();
        // GENERAL: several periodic images of a particle can reach one tile, or the reference's integer support box clips the kernel (sigma > 1)
        bool general = (g.mode == 0 && g.sigma > 1.0f) || g.qsAc > 0.0f; // (Gaussian mode is non-periodic: GENERAL there = the reference cells)
        const int gtile[3] = {GT_X, GT_Y, GT_Z};
        for (int a = 0; a < 3; ++a)
            if (g.cyc[a] && g.s[a] < gtile[a] + 2 * c->reach + 2) general = true;
        float* V = c->vol.as<float>();
        float* C3 = c->rgb.as<float>();
#define MMS_GATHER(M, COL, out) \
        (general ? density_gather_kernel<M, COL, true><<<grid, GT_THREADS, 0, st>>>(g, DS, R, A, auxN, CS, V, out, c->reach) \
                 : density_gather_kernel<M, COL, false><<<grid, GT_THREADS, 0, st>>>(g, DS, R, A, auxN, CS, V, out, c->reach))
        // Gaussian mode, radial cut-off, no periodic axis (QuickSurf's own case): the warp-patch kernel
        const bool gauss = g.mode == 1 && !general && !g.cyc[0] && !g.cyc[1] && !g.cyc[2] && !getenv("MMS_GATHER_GENERIC");
        const dim3 gridQ((g.s[0] + GT_X - 1) / GT_X, (g.s[1] + GP_Y - 1) / GP_Y, (g.nz + GT_Z - 1) / GT_Z);
        if (gauss && colour) density_gauss_kernel<true><<<gridQ, GQ_THREADS, 0, st>>>(g, DS, R, A, auxN, CS, V, C3, c->reach);
        else if (gauss) density_gauss_kernel<false><<<gridQ, GQ_THREADS, 0, st>>>(g, DS, R, A, auxN, CS, V, nullptr, c->reach);
        else if (vector) MMS_GATHER(0, true, C3);
        else if (g.mode == 0) MMS_GATHER(0, false, nullptr);
        else if (colour) MMS_GATHER(1, true, C3);
        else MMS_GATHER(1, false, nullptr);
#undef MMS_GATHER
    } else {
        dim3 grid((g.s[0] + CT_X - 1) / CT_X, (g.s[1] + CT_Y - 1) / CT_Y, (g.nz + CT_Z - 1) / CT_Z);
        // V2 (lanes walk a compacted hit list) where every support box is at most 3x3x3 voxels and one periodic image per particle is enough
        bool v2 = g.mode == 0 && auxN == 0 && c->splatV2ok && !getenv("MMS_SPLAT_V1");
        const int tileDim[3] = {CT_X, CT_Y, CT_Z};
        for (int a = 0; a < 3; ++a) // a periodic axis so short that one cell holds particles of two images of the tile: the general kernel
            if (g.cyc[a] && g.s[a] < tileDim[a] + 2 * c->reach + 2 + (1 << g.cshift)) v2 = false;
        if (v2) {
            Splat3Consts kc{};
            for (int a = 0; a < 3; ++a) {
                kc.isd[a] = 1.0f / g.sd[a];
                kc.per[a] = static_cast<float>(g.s[a]) * g.sd[a];
                kc.rper[a] = 1.0f / kc.per[a];
                kc.slack[a] = (0.05f + 1e-5f * static_cast<float>(g.s[a])) * g.sd[a];
            }
            // per-block-coordinate cell lists: a function of the geometry alone -> built on the host, uploaded when the geometry changes
            if (c->s3Reach != c->reach || std::memcmp(&c->s3Geo, &g, sizeof(Geo)) != 0 || !c->s3Tables.p) {
                const size_t bytes = splat3TableBytes(grid.x, grid.y, grid.z);
                if (!c->s3Tables.ensure(bytes)) return c->fail(MMS_ERR_NOMEM, "device allocation failed (splat tables)");
                c->s3Host.assign(bytes, 0);
                splat3BuildTables(g, c->reach, grid.x, grid.y, grid.z, c->s3Host.data());
                MMS_CUDA(c, cudaMemcpyAsync(c->s3Tables.p, c->s3Host.data(), bytes, cudaMemcpyHostToDevice, st)); // pageable: staged before it returns
                c->s3Geo = g, c->s3Reach = c->reach;
            }
            density_splat3_kernel<<<grid, CT_THREADS, sizeof(Splat3Shared), st>>>(g, kc, c->dstate.as<DevState>(), c->recsB.as<float4>(),
                c->cellStart.as<unsigned>(), c->vol.as<float>(), c->reach, c->s3Tables.as<unsigned char>());
        }
        else if (g.mode == 0)
            density_splat_kernel<0><<<grid, CT_THREADS, sizeof(SplatShared), st>>>(g, c->dstate.as<DevState>(), c->recsB.as<float4>(),
                c->auxB.as<float>(), auxN, c->cellStart.as<unsigned>(), c->vol.as<float>(), c->reach);
        else
            density_splat_kernel<1><<<grid, CT_THREADS, sizeof(SplatShared), st>>>(g, c->dstate.as<DevState>(), c->recsB.as<float4>(),
                c->auxB.as<float>(), auxN, c->cellStart.as<unsigned>(), c->vol.as<float>(), c->reach);
    }
    c->haveColour = colour;
    c->haveVector = vector;
    ++c->launches;
    if (vector) {
        vector_finalize_kernel<<<gridFor(nvox, 256, c->smCount * 16), 256, 0, st>>>(c->vol.as<float>(), c->rgb.as<float>(), c->dirVol.as<float>(), nvox,
            c->dstate.as<DevState>());
        ++c->launches;
    }
    c->rec(EV_DEN1);
    c->normalized = false;
    c->volPrefetched = false;
    if (c->params.mode == MMS_MODE_P2D_BUMP && c->params.normalize && !c->params.defer_normalize) {
        c->rec(EV_NRM0);
        // aggregator 2: the reference normalises the three COMPONENTS with the range of the magnitudes (:669-682); |v| itself stays as it is
        normalize_state_kernel<<<c->smCount * 16, 256, 0, st>>>(vector ? c->rgb.as<float>() : c->vol.as<float>(), vector ? nvox * 3 : nvox,
            c->dstate.as<DevState>());
        ++c->launches;
        c->rec(EV_NRM1);
        c->normalized = true;
    }
    MMS_CUDA(c, cudaMemcpyAsync(c->hState.p, c->dstate.p, sizeof(DevState), cudaMemcpyDeviceToHost, st));
    MMS_CUDA(c, cudaGetLastError());
    c->haveDensity = true;
    c->haveMesh = c->haveCount = false; // a count (segment offsets, cached geometry) of the old volume must not be emitted
    return MMS_OK;
}

/** Device-side failures are flagged in DevState::pad[0]; the host sees them wherever it synchronises anyway (volume / range / vector
 *  field read-back, the triangle count).  Purely device-resident sequences (mms_*_device) end in mms_count_isosurface, which checks. */
static int deviceErrorFromState(mms_ctx* c) {
    const DevState* hs = c->hState.as<DevState>();
    if (hs->pad[0] == 5)
        return c->fail(MMS_ERR_UNSUPPORTED, "halo exchange: the receive buffer overflowed (mms_halo_buffers capacity too small) or the other slabs' "
                                            "records did not arrive in time (mms_halo_wait)");
    if (hs->pad[0] == 2)
        return c->fail(MMS_ERR_UNSUPPORTED, "internal error: the single-image gather kernel ran on a short periodic axis");
    if (hs->pad[0] != 0)
        return c->fail(MMS_ERR_UNSUPPORTED, "internal error: the splat kernel's neighbourhood list overflowed (%d cells per axis)", CT_MAXAXIS);
    return MMS_OK;
}
static int checkDeviceError(mms_ctx* c) {
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    return deviceErrorFromState(c);
}

int mms_get_density_range(mms_ctx* c, float minmax[2]) {
    if (!c || !minmax) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    DeviceGuard guard(c->device);
    if (int rc = checkDeviceError(c)) return rc;
    const DevState* hs = c->hState.as<DevState>();
    minmax[0] = keyFloat(hs->minKey);
    minmax[1] = keyFloat(hs->maxKey);
    return MMS_OK;
}

int mms_normalize(mms_ctx* c, float mn, float mx) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    if (c->adoptedVol) return c->fail(MMS_ERR_INVALID, "an adopted volume belongs to its producer: normalise there");
    DeviceGuard guard(c->device);
    const size_t nvox = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz;
    volatile float range = mx - mn;
    volatile float rcp = 1.0f / range;
    c->rec(EV_NRM0);
    normalize_kernel<<<c->smCount * 8, 256, 0, c->stream>>>(c->haveVector ? c->rgb.as<float>() : c->vol.as<float>(), c->haveVector ? nvox * 3 : nvox, mn, rcp);
    ++c->launches;
    c->rec(EV_NRM1);
    MMS_CUDA(c, cudaGetLastError());
    c->normalized = true;
    c->volPrefetched = false;
    c->haveMesh = c->haveCount = false; // a count (segment offsets, cached geometry) of the old volume must not be emitted
    return MMS_OK;
}

int mms_set_stream(mms_ctx* c, void* stream) {
    if (!c) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    if (!c->ownStream) c->ownStream = c->stream;
    c->stream = stream ? static_cast<cudaStream_t>(stream) : c->ownStream;
    return MMS_OK;
}

int mms_density_range_device(mms_ctx* c, float** dev_negmin_max) {
    if (!c || !dev_negmin_max) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    DeviceGuard guard(c->device);
    if (!c->rangeBuf.ensure(16)) return c->fail(MMS_ERR_NOMEM, "allocation failed");
    range_to_float_kernel<<<1, 1, 0, c->stream>>>(c->dstate.as<DevState>(), c->rangeBuf.as<float>());
    ++c->launches;
    *dev_negmin_max = c->rangeBuf.as<float>();
    return MMS_OK;
}

int mms_normalize_device(mms_ctx* c, const float* dev_negmin_max) {
    if (!c || !dev_negmin_max) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    if (c->adoptedVol) return c->fail(MMS_ERR_INVALID, "an adopted volume belongs to its producer: normalise there");
    DeviceGuard guard(c->device);
    const size_t nvox = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz;
    c->rec(EV_NRM0);
    normalize_ptr_kernel<<<c->smCount * 16, 256, 0, c->stream>>>(c->haveVector ? c->rgb.as<float>() : c->vol.as<float>(), c->haveVector ? nvox * 3 : nvox,
        dev_negmin_max);
    ++c->launches;
    c->rec(EV_NRM1);
    MMS_CUDA(c, cudaGetLastError());
    c->normalized = true;
    c->volPrefetched = false;
    c->haveMesh = c->haveCount = false; // a count (segment offsets, cached geometry) of the old volume must not be emitted
    return MMS_OK;
}

int mms_get_density_device(mms_ctx* c, const float** dv, const float** drgb) {
    if (!c || !dv) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    *dv = c->isoVol();
    if (drgb) *drgb = c->haveColour ? c->isoRgb() : nullptr;
    return MMS_OK;
}

int mms_prefetch_density(mms_ctx* c) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    DeviceGuard guard(c->device);
    const size_t bytes = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz * 4;
    if (!c->hVol.ensure(bytes)) return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
    if (c->haveColour && !c->hRgb.ensure(bytes * 3)) return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes * 3);
    // the volume is final once everything enqueued so far has run; the copy stream takes it from there while the compute stream
    // goes on with the isosurface (which only reads the volume)
    MMS_CUDA(c, cudaEventRecord(c->volReady, c->stream));
    MMS_CUDA(c, cudaStreamWaitEvent(c->copyStream, c->volReady, 0));
    cudaEventRecord(c->ev[EV_DV0], c->copyStream);
    c->evSet[EV_DV0] = true;
    MMS_CUDA(c, cudaMemcpyAsync(c->hVol.p, c->isoVol(), bytes, cudaMemcpyDeviceToHost, c->copyStream));
    if (c->haveColour) MMS_CUDA(c, cudaMemcpyAsync(c->hRgb.p, c->isoRgb(), bytes * 3, cudaMemcpyDeviceToHost, c->copyStream));
    cudaEventRecord(c->ev[EV_DV1], c->copyStream);
    c->evSet[EV_DV1] = true;
    MMS_CUDA(c, cudaEventRecord(c->volCopied, c->copyStream));
    c->volPrefetched = true;
    return MMS_OK;
}

int mms_get_density(mms_ctx* c, const float** hv, const float** hrgb) {
    if (!c || !hv) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    DeviceGuard guard(c->device);
    if (c->volPrefetched) { // mms_prefetch_density already has the copy under way (or done)
        MMS_CUDA(c, cudaEventSynchronize(c->volCopied));
        if (int rc = checkDeviceError(c)) return rc;
        *hv = c->hVol.as<float>();
        if (hrgb) *hrgb = c->haveColour ? c->hRgb.as<float>() : nullptr;
        return MMS_OK;
    }
    const size_t bytes = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz * 4;
    if (!c->hVol.ensure(bytes)) return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
    const bool wantRgb = hrgb && c->haveColour;
    if (wantRgb && !c->hRgb.ensure(bytes * 3)) return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes * 3);
    c->rec(EV_DV0);
    MMS_CUDA(c, cudaMemcpyAsync(c->hVol.p, c->isoVol(), bytes, cudaMemcpyDeviceToHost, c->stream));
    if (wantRgb) MMS_CUDA(c, cudaMemcpyAsync(c->hRgb.p, c->isoRgb(), bytes * 3, cudaMemcpyDeviceToHost, c->stream));
    c->rec(EV_DV1);
    if (int rc = checkDeviceError(c)) return rc;
    *hv = c->hVol.as<float>();
    if (hrgb) *hrgb = wantRgb ? c->hRgb.as<float>() : nullptr;
    return MMS_OK;
}

int mms_set_density(mms_ctx* c, const float* volume) {
    if (!c || !volume || !c->haveGrid) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    const size_t bytes = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz * 4;
    if (!c->vol.ensure(bytes)) return c->fail(MMS_ERR_NOMEM, "device allocation of %zu bytes failed", bytes);
    c->adoptedVol = c->adoptedRgb = nullptr;
    MMS_CUDA(c, cudaMemcpyAsync(c->vol.p, volume, bytes, cudaMemcpyDefault, c->stream));
    init_state_kernel<<<1, 1, 0, c->stream>>>(c->dstate.as<DevState>());
    ++c->launches;
    MMS_CUDA(c, cudaMemcpyAsync(c->hState.p, c->dstate.p, sizeof(DevState), cudaMemcpyDeviceToHost, c->stream));
    c->haveDensity = true;
    c->haveColour = false;
    c->haveVector = false;
    c->haveMesh = c->haveCount = false; // a count (segment offsets, cached geometry) of the old volume must not be emitted
    c->volPrefetched = false;
    return MMS_OK;
}

int mms_adopt_density(mms_ctx* c, mms_ctx* p) {
    if (!c || !p || c == p) return MMS_ERR_INVALID;
    if (!p->haveDensity) return c->fail(MMS_ERR_INVALID, "the producer context has no density");
    if (p->device != c->device) return c->fail(MMS_ERR_INVALID, "mms_adopt_density: the two contexts live on different devices (%d, %d)", c->device, p->device);
    DeviceGuard guard(c->device);
    c->grid = p->grid;
    c->haveGrid = true;
    c->z0 = p->z0, c->nz = p->nz, c->cellZ0 = p->cellZ0, c->cellNz = p->cellNz;
    c->adoptedVol = p->adoptedVol ? p->adoptedVol : p->vol.as<float>();
    c->haveColour = p->haveColour;
    c->adoptedRgb = p->haveColour ? (p->adoptedVol ? p->adoptedRgb : p->rgb.as<float>()) : nullptr;
    c->haveVector = false;
    // everything the producer has enqueued so far (its density kernels, normalisation) comes first
    MMS_CUDA(c, cudaEventRecord(c->adoptReady, p->stream));
    MMS_CUDA(c, cudaStreamWaitEvent(c->stream, c->adoptReady, 0));
    init_state_kernel<<<1, 1, 0, c->stream>>>(c->dstate.as<DevState>());
    ++c->launches;
    MMS_CUDA(c, cudaMemcpyAsync(c->hState.p, c->dstate.p, sizeof(DevState), cudaMemcpyDeviceToHost, c->stream));
    c->haveDensity = true;
    c->haveMesh = c->haveCount = false;
    c->volPrefetched = false;
    return MMS_OK;
}

int mms_get_vector_field(mms_ctx* c, const float** hvec, const float** hmag, const float** hdir) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity || !c->haveVector) return c->fail(MMS_ERR_INVALID, "no vector field has been computed (aggregator 2)");
    DeviceGuard guard(c->device);
    const size_t bytes = static_cast<size_t>(c->grid.res[0]) * c->grid.res[1] * c->nz * 4;
    if ((hvec && !c->hRgb.ensure(bytes * 3)) || (hmag && !c->hVol.ensure(bytes)) || (hdir && !c->hDir.ensure(bytes * 3)))
        return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes * 3);
    if (c->volPrefetched) { // never two copies into hVol at once
        MMS_CUDA(c, cudaEventSynchronize(c->volCopied));
        c->volPrefetched = false;
    }
    c->rec(EV_DV0);
    if (hvec) MMS_CUDA(c, cudaMemcpyAsync(c->hRgb.p, c->rgb.p, bytes * 3, cudaMemcpyDeviceToHost, c->stream));
    if (hmag) MMS_CUDA(c, cudaMemcpyAsync(c->hVol.p, c->vol.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (hdir) MMS_CUDA(c, cudaMemcpyAsync(c->hDir.p, c->dirVol.p, bytes * 3, cudaMemcpyDeviceToHost, c->stream));
    c->rec(EV_DV1);
    if (int rc = checkDeviceError(c)) return rc;
    if (hvec) *hvec = c->hRgb.as<float>();
    if (hmag) *hmag = c->hVol.as<float>();
    if (hdir) *hdir = c->hDir.as<float>();
    return MMS_OK;
}

int mms_get_vector_field_device(mms_ctx* c, const float** dvec, const float** dmag, const float** ddir) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity || !c->haveVector) return c->fail(MMS_ERR_INVALID, "no vector field has been computed (aggregator 2)");
    if (dvec) *dvec = c->rgb.as<float>();
    if (dmag) *dmag = c->vol.as<float>();
    if (ddir) *ddir = c->dirVol.as<float>();
    return MMS_OK;
}

/** Count, first half: everything up to (not including) the host round trip.  countPending tells countFinish whether kernels are in flight. */
static int countLaunch(mms_ctx* c, float iso) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity) return c->fail(MMS_ERR_INVALID, "no density has been computed");
    DeviceGuard guard(c->device);
    c->countPending = false;
    McGeo m{};
    m.sx = c->grid.res[0], m.sy = c->grid.res[1];
    m.nzPlanes = c->nz, m.zPlane0 = c->z0, m.szGlobal = c->grid.res[2];
    m.cx = m.sx - 1, m.cy = m.sy - 1, m.cz0 = c->cellZ0, m.cnz = c->cellNz;
    m.nsegx = (m.cx + 31) / 32;
    const Geo g = makeGeo(c);
    for (int a = 0; a < 3; ++a) {
        m.org[a] = g.mn[a], m.sd[a] = g.sd[a];
        for (int n = 1; n <= 2; ++n) {
            volatile float den = static_cast<float>(n) * g.sd[a];
            volatile float r = 1.0f / den;
            m.rinv[a][n] = r;
        }
        m.rinv[a][0] = 0.0f;
    }
    m.iso = iso;
    m.maxTris = 0xffffffffu;
    m.layersPerBlock = EM_LAYERS;
    m.countLayers = CN_LAYERS;
    c->mcGeo = m;
    c->ntris = 0;
    c->haveCount = false; // set once the count has run and the device reported no error
    c->haveMesh = false;
    c->rec(EV_MC0);
    if (m.cnz <= 0) {
        c->haveCount = true; // an empty slab: nothing to count, nothing to emit
        return MMS_OK;
    }
    const size_t nseg = static_cast<size_t>(m.nsegx) * m.cy * m.cnz;
    if (nseg >= (1ull << 32) - 1) return c->fail(MMS_ERR_UNSUPPORTED, "too many cell segments");
    const unsigned ntiles = static_cast<unsigned>((nseg + kScanTile - 1) / kScanTile);
    if (!c->segCount.ensure(nseg * 4) || !c->segOffset.ensure((nseg + 1) * 4) || !c->tileSums.ensure(std::max<size_t>(ntiles, 1) * 4))
        return c->fail(MMS_ERR_NOMEM, "device allocation failed (marching-cubes segments)");
    unsigned char* tri = nullptr;
    if (c->params.want_cell_tricounts) {
        const size_t ncell = static_cast<size_t>(m.cx) * m.cy * m.cnz;
        if (!c->triCount.ensure(ncell)) return c->fail(MMS_ERR_NOMEM, "device allocation failed (cell counts)");
        tri = c->triCount.as<unsigned char>();
    }
    cudaStream_t st = c->stream;
    c->countMode = c->isoMode;
    if (c->isoMode == MMS_ISO_MARCHING_TETS) {
        if (c->haveColour) return c->fail(MMS_ERR_UNSUPPORTED, "the marching-tetrahedra mode has no colour output");
        MtGeo t{};
        t.sx = m.sx, t.sy = m.sy, t.zPlane0 = m.zPlane0, t.szGlobal = m.szGlobal, t.cx = m.cx, t.cy = m.cy, t.cz0 = m.cz0, t.cnz = m.cnz, t.nsegx = m.nsegx;
        for (int a = 0; a < 3; ++a) {
            t.mn[a] = c->grid.min[a], t.ext[a] = c->grid.extent[a];
            volatile float ext = c->grid.extent[a];
            volatile float s = static_cast<float>(c->grid.res[a]);
            volatile float cell = ext / s; // osbb.Width() / static_cast<float>(sx) (IsoSurface.cpp:238-240)
            t.cell[a] = cell;
        }
        t.iso = iso;
        c->mtGeo = t;
        dim3 grid(m.nsegx, (m.cy + MT_THREADS / 32 - 1) / (MT_THREADS / 32), m.cnz);
        mt_count_kernel<<<grid, MT_THREADS, 0, st>>>(t, c->isoVol(), c->segCount.as<unsigned>(), tri);
    } else {
        if (tri) MMS_CUDA(c, cudaMemsetAsync(tri, 0, static_cast<size_t>(m.cx) * m.cy * m.cnz, st)); // the kernel writes the non-empty cells only
        // Layers per block: the smallest count >= CN_LAYERS for which the grid is a whole number of waves of resident blocks or just under
        // it (C2: 1024 blocks of 16 layers on 888 slots ran a second wave that was 15 % full), large grids keep the default.
        {
            if (c->countSlots == 0) {
                int perSm = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, mc_count_kernel, MC_THREADS, 0) != cudaSuccess || perSm < 1) perSm = 4;
                c->countSlots = perSm * c->smCount;
            }
            const long long cols = static_cast<long long>((m.nsegx + CN_SEGS - 1) / CN_SEGS) * ((m.cy + CN_ROWS - 1) / CN_ROWS);
            int L = CN_LAYERS;
            const long long blocks0 = cols * ((m.cnz + L - 1) / L);
            if (blocks0 > c->countSlots && blocks0 < 4ll * c->countSlots) {
                const long long waves = (blocks0 + c->countSlots - 1) / c->countSlots;
                // fewest layers per block that bring the grid down to (waves - 1) full waves
                const long long zBlocks = std::max(1ll, (waves - 1) * c->countSlots / cols);
                const int L2 = static_cast<int>((m.cnz + zBlocks - 1) / zBlocks);
                if (L2 <= 2 * CN_LAYERS && static_cast<double>(blocks0) / (waves * c->countSlots) < 0.8) L = L2; // (a nearly full last wave stays)
            }
            c->mcGeo.countLayers = m.countLayers = L;
        }
        dim3 grid((m.nsegx + CN_SEGS - 1) / CN_SEGS, (m.cy + CN_ROWS - 1) / CN_ROWS, (m.cnz + m.countLayers - 1) / m.countLayers);
        // indexed mesh: the count kernel writes the node segments' mask records as a by-product (validated and allocated below)
        const bool ix = c->meshIndexed && !c->haveColour && c->z0 == 0 && c->nz == c->grid.res[2] && m.cz0 == 0 && m.cnz == c->grid.res[2] - 1 &&
                        static_cast<long long>(m.sx) * m.sy < (1ll << 31);
        const size_t nvs = static_cast<size_t>((m.sx + 31) / 32) * m.sy * m.szGlobal;
        if (ix && (nvs >= (1ull << 32) - 1 || !c->vertCount.ensure(nvs * 4) || !c->vertOffset.ensure((nvs + 1) * 4) || !c->vertRec.ensure(nvs * 16)))
            return c->fail(MMS_ERR_NOMEM, "device allocation failed (vertex segments)");
        mc_count_kernel<<<grid, MC_THREADS, 0, st>>>(m, c->isoVol(), c->segCount.as<unsigned>(), tri, ix ? c->vertRec.as<uint4>() : nullptr,
            ix ? c->vertCount.as<unsigned>() : nullptr);
    }
    ++c->launches;
    DevState* ds = c->dstate.as<DevState>();
    exclusiveScan(c->segCount.as<unsigned>(), c->segOffset.as<unsigned>(), nullptr, c->tileSums.as<unsigned>(), static_cast<unsigned>(nseg),
        &ds->totalTris, st, c->launches);
    c->countIndexed = false;
    if (c->meshIndexed) {
        // indexed mesh: one vertex per crossed grid edge, numbered by node segment (mc_indexed.cuh)
        if (c->isoMode != MMS_ISO_MARCHING_CUBES) return c->fail(MMS_ERR_UNSUPPORTED, "the indexed mesh is a marching-cubes output");
        if (c->haveColour) return c->fail(MMS_ERR_UNSUPPORTED, "the indexed mesh has no colour output yet");
        if (c->z0 != 0 || c->nz != c->grid.res[2] || m.cz0 != 0 || m.cnz != c->grid.res[2] - 1)
            return c->fail(MMS_ERR_UNSUPPORTED, "the indexed mesh needs the whole volume in one context (vertex ids cross z-slab borders)");
        if (static_cast<long long>(m.sx) * m.sy >= (1ll << 31)) return c->fail(MMS_ERR_UNSUPPORTED, "the indexed mesh needs planes below 2^31 voxels");
        const size_t nvs = static_cast<size_t>((m.sx + 31) / 32) * m.sy * m.szGlobal;
        if (nvs >= (1ull << 32) - 1) return c->fail(MMS_ERR_UNSUPPORTED, "too many node segments");
        const unsigned vtiles = static_cast<unsigned>((nvs + kScanTile - 1) / kScanTile);
        if (!c->vertCount.ensure(nvs * 4) || !c->vertOffset.ensure((nvs + 1) * 4) || !c->vertRec.ensure(nvs * 16) || !c->tileSums.ensure(std::max<size_t>(std::max(ntiles, vtiles), 1) * 4))
            return c->fail(MMS_ERR_NOMEM, "device allocation failed (vertex segments)");
        // the node rows without a cell row (mc_count_kernel has written all the others): last row, last plane, a lone last segment
        const int nsvI = (m.sx + 31) / 32;
        {
            const McxRange lastRow{0, m.sy - 1, m.sy, 0};
            mcx_mask_kernel<<<dim3(nsvI, 1, m.szGlobal), MCX_THREADS, 0, st>>>(m, lastRow, c->isoVol(), c->vertRec.as<uint4>(), c->vertCount.as<unsigned>());
            ++c->launches;
            if (m.sy > 1) {
                const McxRange lastPlane{0, 0, m.sy - 1, m.szGlobal - 1};
                mcx_mask_kernel<<<dim3(nsvI, (m.sy - 1 + MCX_WARPS - 1) / MCX_WARPS, 1), MCX_THREADS, 0, st>>>(m, lastPlane, c->isoVol(), c->vertRec.as<uint4>(),
                    c->vertCount.as<unsigned>());
                ++c->launches;
            }
            if (nsvI > m.nsegx && m.sy > 1 && m.szGlobal > 1) {
                const McxRange lastSeg{nsvI - 1, 0, m.sy - 1, 0};
                mcx_mask_kernel<<<dim3(1, (m.sy - 1 + MCX_WARPS - 1) / MCX_WARPS, m.szGlobal - 1), MCX_THREADS, 0, st>>>(m, lastSeg, c->isoVol(),
                    c->vertRec.as<uint4>(), c->vertCount.as<unsigned>());
                ++c->launches;
            }
        }
        exclusiveScan(c->vertCount.as<unsigned>(), c->vertOffset.as<unsigned>(), nullptr, c->tileSums.as<unsigned>(), static_cast<unsigned>(nvs),
            &ds->totalVerts, st, c->launches);
        c->countIndexed = true;
    }
    publish_state_kernel<<<1, 1, 0, st>>>(c->dstate.as<DevState>(), c->hState.as<DevState>());
    ++c->launches;
    MMS_CUDA(c, cudaGetLastError());
    if (!c->countReady) MMS_CUDA(c, cudaEventCreateWithFlags(&c->countReady, cudaEventDisableTiming));
    MMS_CUDA(c, cudaEventRecord(c->countReady, st));
    c->countPending = true;
    return MMS_OK;
}

/** Count, second half: the one host round trip (the mesh size decides the allocation). */
static int countFinish(mms_ctx* c, uint64_t* ntris) {
    if (ntris) *ntris = 0;
    if (!c->countPending) return MMS_OK; // empty slab
    DeviceGuard guard(c->device);
    c->countPending = false;
    MMS_CUDA(c, cudaEventSynchronize(c->countReady)); // (not the stream: a speculative emit kernel may already be running behind the count)
    // the state block also carries the density kernels' error flag: a truncated volume must not become a mesh
    if (int rc = deviceErrorFromState(c)) return rc;
    c->ntris = c->hState.as<DevState>()->totalTris;
    c->nverts = c->countIndexed ? c->hState.as<DevState>()->totalVerts : 0;
    if (c->countIndexed && c->nverts >= (1ull << 32)) return c->fail(MMS_ERR_UNSUPPORTED, "more than 2^32 vertices do not fit 32-bit indices");
    c->haveCount = true;
    if (ntris) *ntris = c->ntris;
    return MMS_OK;
}

int mms_count_isosurface(mms_ctx* c, float iso, uint64_t* ntris) {
    if (ntris) *ntris = 0;
    if (int rc = countLaunch(c, iso)) return rc;
    return countFinish(c, ntris);
}

/** mc_emit_kernel into P / N / C (the default marching-cubes emission); maxTris: see McGeo. */
static void launchMcEmit(mms_ctx* c, float* P, float* N, float* C, unsigned maxTris) {
    McGeo m = c->mcGeo;
    m.maxTris = maxTris;
    cudaStream_t st = c->stream;
    // a block marches up to EM_LAYERS cell layers; on small volumes fewer, until the grid holds two waves of blocks (4 per SM)
    int L = EM_LAYERS;
    const long long columns = static_cast<long long>(m.nsegx) * ((m.cy + EY - 1) / EY);
    while (L > 8 && columns * ((m.cnz + L - 1) / L) < 8ll * c->smCount) L /= 2;
    m.layersPerBlock = L;
    const dim3 gridE(m.nsegx, (m.cy + EY - 1) / EY, (m.cnz + L - 1) / L);
    CUtensorMap map{};
    const bool tma = makeVolumeTensorMap(&map, c->isoVol(), m.sx, m.sy, m.nzPlanes);
    const float* V = c->isoVol();
    const float* RGB = c->isoRgb();
    const unsigned* S = c->segOffset.as<unsigned>();
    if (c->haveColour) {
        if (tma) mc_emit_kernel<true, true><<<gridE, MC_THREADS, emitSmemBytes(true), st>>>(m, map, V, RGB, S, P, N, C);
        else mc_emit_kernel<true, false><<<gridE, MC_THREADS, emitSmemBytes(true), st>>>(m, map, V, RGB, S, P, N, C);
    } else {
        if (tma) mc_emit_kernel<false, true><<<gridE, MC_THREADS, emitSmemBytes(false), st>>>(m, map, V, nullptr, S, P, N, nullptr);
        else mc_emit_kernel<false, false><<<gridE, MC_THREADS, emitSmemBytes(false), st>>>(m, map, V, nullptr, S, P, N, nullptr);
    }
    ++c->launches;
}

int mms_emit_isosurface(mms_ctx* c, float* pos, float* nrm, float* col, uint64_t first_triangle) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveCount) return c->fail(MMS_ERR_INVALID, "mms_count_isosurface has not been called");
    DeviceGuard guard(c->device);
    const McGeo& m = c->mcGeo;
    cudaStream_t st = c->stream;
    const bool own = pos == nullptr;
    if (c->countIndexed) {
        if (!own) return c->fail(MMS_ERR_UNSUPPORTED, "the indexed mesh is emitted into the library's own buffers (mms_get_mesh_indexed*)");
        if (c->ntris > 0) {
            const size_t vbytes = static_cast<size_t>(c->nverts) * 12, ibytes = static_cast<size_t>(c->ntris) * 12;
            auto grow = [&](DevBuf& b, size_t bytes) { return b.cap >= bytes || b.ensure(bytes + bytes / 8) || b.ensure(bytes); };
            if (!grow(c->meshPos, vbytes) || !grow(c->meshNrm, vbytes) || !grow(c->meshIdx, ibytes))
                return c->fail(MMS_ERR_NOMEM, "device allocation of the indexed mesh (%llu vertices, %llu triangles) failed", c->nverts, c->ntris);
            c->rec(EV_EMIT0);
            dim3 gv((m.sx + 31) / 32, (m.sy + MCX_WARPS - 1) / MCX_WARPS, m.szGlobal);
            mcx_vertex_kernel<<<gv, MCX_THREADS, 0, st>>>(m, c->isoVol(), c->vertRec.as<uint4>(), c->vertOffset.as<unsigned>(), c->meshPos.as<float>(),
                c->meshNrm.as<float>());
            dim3 gi(m.nsegx, (m.cy + MCX_WARPS - 1) / MCX_WARPS, m.cnz);
            mcx_index_kernel<<<gi, MCX_THREADS, 0, st>>>(m, c->vertRec.as<uint4>(), c->segOffset.as<unsigned>(), c->vertOffset.as<unsigned>(),
                c->meshIdx.as<unsigned>());
            c->launches += 2;
        }
        c->rec(EV_MC1);
        MMS_CUDA(c, cudaGetLastError());
        c->haveMesh = true;
        c->meshExternal = false;
        return MMS_OK;
    }
    if (c->ntris > 0) {
        float *P = pos, *N = nrm, *C = col;
        if (own) {
            // 1/8 headroom: the triangle count of a time series wobbles from frame to frame; growing a multi-GB buffer is a stall
            const size_t mbytes = static_cast<size_t>(c->ntris) * 36, mwant = mbytes + mbytes / 8;
            auto grow = [&](DevBuf& b) { return b.cap >= mbytes || b.ensure(mwant) || b.ensure(mbytes); };
            if (!grow(c->meshPos) || !grow(c->meshNrm) || (c->haveColour && !grow(c->meshCol)))
                return c->fail(MMS_ERR_NOMEM, "device allocation of the mesh (%llu triangles, %llu bytes) failed", c->ntris, c->ntris * 72ull);
            P = c->meshPos.as<float>(), N = c->meshNrm.as<float>(), C = c->haveColour ? c->meshCol.as<float>() : nullptr;
        } else {
            if (!nrm) return c->fail(MMS_ERR_INVALID, "a normal buffer is required");
            if (c->haveColour && !col) return c->fail(MMS_ERR_INVALID, "colour mode needs a colour buffer");
            P += first_triangle * 9, N += first_triangle * 9;
            if (C) C += first_triangle * 9;
        }
        c->rec(EV_EMIT0);
        if (c->countMode == MMS_ISO_MARCHING_TETS) {
            dim3 grid(m.nsegx, (m.cy + MT_THREADS / 32 - 1) / (MT_THREADS / 32), m.cnz);
            mt_emit_kernel<<<grid, MT_THREADS, 0, st>>>(c->mtGeo, c->isoVol(), c->segOffset.as<unsigned>(), P, N);
            ++c->launches;
            c->rec(EV_MC1);
            MMS_CUDA(c, cudaGetLastError());
            c->haveMesh = true;
            c->meshExternal = !own;
            return MMS_OK;
        }
        CUtensorMap map{};
        const bool tma = makeVolumeTensorMap(&map, c->isoVol(), m.sx, m.sy, m.nzPlanes);
        const float* V = c->isoVol();
        const float* RGB = c->isoRgb();
        const unsigned* S = c->segOffset.as<unsigned>();
        if (getenv("MMS_EMIT_V4")) {
            dim3 g4(m.nsegx, (m.cy + v4::EY - 1) / v4::EY, (m.cnz + v4::EM_STEPS * v4::EZ - 1) / (v4::EM_STEPS * v4::EZ));
            if (c->haveColour) {
                if (tma) v4::mc_emit_v4_kernel<true, true><<<g4, MC_THREADS, emitV4SmemBytes(true), st>>>(m, map, V, RGB, S, P, N, C);
                else v4::mc_emit_v4_kernel<true, false><<<g4, MC_THREADS, emitV4SmemBytes(true), st>>>(m, map, V, RGB, S, P, N, C);
            } else {
                if (tma) v4::mc_emit_v4_kernel<false, true><<<g4, MC_THREADS, emitV4SmemBytes(false), st>>>(m, map, V, nullptr, S, P, N, nullptr);
                else v4::mc_emit_v4_kernel<false, false><<<g4, MC_THREADS, emitV4SmemBytes(false), st>>>(m, map, V, nullptr, S, P, N, nullptr);
            }
            ++c->launches;
        } else {
            launchMcEmit(c, P, N, C, 0xffffffffu);
        }
    }
    c->rec(EV_MC1);
    MMS_CUDA(c, cudaGetLastError());
    c->haveMesh = true;
    c->meshExternal = !own;
    return MMS_OK;
}

int mms_set_isosurface_mode(mms_ctx* c, int32_t mode) {
    if (!c) return MMS_ERR_INVALID;
    if (mode != MMS_ISO_MARCHING_CUBES && mode != MMS_ISO_MARCHING_TETS) return c->fail(MMS_ERR_INVALID, "unknown isosurface mode %d", mode);
    c->isoMode = mode;
    c->haveCount = false;
    c->haveMesh = c->haveCount = false; // a count (segment offsets, cached geometry) of the old volume must not be emitted
    return MMS_OK;
}

int mms_extract_isosurface(mms_ctx* c, float iso) {
    if (!c) return MMS_ERR_INVALID;
    if (int rc = countLaunch(c, iso)) return rc;
    // Speculative emission: the only thing the host needs the triangle count for is the SIZE of the mesh buffers.  Where buffers of an
    // earlier frame exist (steady state of a time series; they keep 1/8 headroom), the emit kernel is launched right behind the count
    // and the scan, with the buffers' capacity as its limit, and the host round trip leaves the GPU's critical path; only a frame that
    // outgrows the buffers is emitted again after growing them.
    uint64_t capTris = std::min(c->meshPos.cap, c->meshNrm.cap) / 36;
    if (c->haveColour) capTris = std::min<uint64_t>(capTris, c->meshCol.cap / 36);
    const bool speculate = c->countPending && !c->countIndexed && c->countMode == MMS_ISO_MARCHING_CUBES && capTris > 0 && !getenv("MMS_EMIT_V4") &&
                           !getenv("MMS_NO_SPECULATION");
    if (speculate) {
        DeviceGuard guard(c->device);
        c->rec(EV_EMIT0);
        launchMcEmit(c, c->meshPos.as<float>(), c->meshNrm.as<float>(), c->haveColour ? c->meshCol.as<float>() : nullptr,
            static_cast<unsigned>(std::min<uint64_t>(capTris, 0xffffffffull)));
        c->rec(EV_MC1);
    }
    if (int rc = countFinish(c, nullptr)) return rc;
    if (speculate && c->ntris <= capTris) {
        MMS_CUDA(c, cudaGetLastError());
        c->haveMesh = true;
        c->meshExternal = false;
        return MMS_OK;
    }
    if (speculate) { // the frame outgrew the buffers: the truncated emission must have left them before they are re-allocated
        DeviceGuard guard(c->device);
        MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return mms_emit_isosurface(c, nullptr, nullptr, nullptr, 0);
}

int mms_set_mesh_indexed(mms_ctx* c, int32_t on) {
    if (!c) return MMS_ERR_INVALID;
    c->meshIndexed = on != 0;
    c->haveMesh = c->haveCount = false; // a count made for the other format must not be emitted
    return MMS_OK;
}

int mms_get_mesh_indexed_device(mms_ctx* c, uint64_t* nverts, uint64_t* ntris, const float** pos, const float** nrm, const uint32_t** idx) {
    if (!c || !nverts || !ntris) return MMS_ERR_INVALID;
    if (!c->haveMesh || !c->countIndexed) return c->fail(MMS_ERR_INVALID, "no indexed isosurface has been extracted (mms_set_mesh_indexed)");
    *nverts = c->ntris ? c->nverts : 0, *ntris = c->ntris;
    if (pos) *pos = c->ntris ? c->meshPos.as<float>() : nullptr;
    if (nrm) *nrm = c->ntris ? c->meshNrm.as<float>() : nullptr;
    if (idx) *idx = c->ntris ? c->meshIdx.as<uint32_t>() : nullptr;
    return MMS_OK;
}

int mms_get_mesh_indexed(mms_ctx* c, uint64_t* nverts, uint64_t* ntris, const float** pos, const float** nrm, const uint32_t** idx) {
    if (!c || !nverts || !ntris) return MMS_ERR_INVALID;
    if (!c->haveMesh || !c->countIndexed) return c->fail(MMS_ERR_INVALID, "no indexed isosurface has been extracted (mms_set_mesh_indexed)");
    DeviceGuard guard(c->device);
    *nverts = c->ntris ? c->nverts : 0, *ntris = c->ntris;
    if (pos) *pos = nullptr;
    if (nrm) *nrm = nullptr;
    if (idx) *idx = nullptr;
    const size_t vbytes = static_cast<size_t>(*nverts) * 12, ibytes = static_cast<size_t>(c->ntris) * 12;
    if (ibytes) {
        auto growPin = [&](PinBuf& b, size_t bytes) { return b.cap >= bytes || b.ensure(bytes + bytes / 8) || b.ensure(bytes); };
        if ((pos && !growPin(c->hPos, vbytes)) || (nrm && !growPin(c->hNrm, vbytes)) || (idx && !growPin(c->hIdx, ibytes)))
            return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", vbytes + vbytes + ibytes);
        c->rec(EV_DM0);
        if (pos) MMS_CUDA(c, cudaMemcpyAsync(c->hPos.p, c->meshPos.p, vbytes, cudaMemcpyDeviceToHost, c->stream));
        if (nrm) MMS_CUDA(c, cudaMemcpyAsync(c->hNrm.p, c->meshNrm.p, vbytes, cudaMemcpyDeviceToHost, c->stream));
        if (idx) MMS_CUDA(c, cudaMemcpyAsync(c->hIdx.p, c->meshIdx.p, ibytes, cudaMemcpyDeviceToHost, c->stream));
        c->rec(EV_DM1);
        MMS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (pos) *pos = c->hPos.as<float>();
        if (nrm) *nrm = c->hNrm.as<float>();
        if (idx) *idx = c->hIdx.as<uint32_t>();
    }
    return MMS_OK;
}

int mms_get_mesh_device(mms_ctx* c, uint64_t* nverts, const float** pos, const float** nrm, const float** col) {
    if (!c || !nverts) return MMS_ERR_INVALID;
    if (!c->haveMesh) return c->fail(MMS_ERR_INVALID, "no isosurface has been extracted");
    if (c->countIndexed) return c->fail(MMS_ERR_INVALID, "the mesh is indexed: use mms_get_mesh_indexed_device");
    if (c->meshExternal) return c->fail(MMS_ERR_INVALID, "the mesh was emitted into caller-supplied memory");
    *nverts = c->ntris * 3;
    if (pos) *pos = c->ntris ? c->meshPos.as<float>() : nullptr;
    if (nrm) *nrm = c->ntris ? c->meshNrm.as<float>() : nullptr;
    if (col) *col = (c->ntris && c->haveColour) ? c->meshCol.as<float>() : nullptr;
    return MMS_OK;
}

int mms_get_mesh(mms_ctx* c, uint64_t* nverts, const float** pos, const float** nrm, const float** col) {
    if (!c || !nverts) return MMS_ERR_INVALID;
    if (!c->haveMesh) return c->fail(MMS_ERR_INVALID, "no isosurface has been extracted");
    if (c->countIndexed) return c->fail(MMS_ERR_INVALID, "the mesh is indexed: use mms_get_mesh_indexed");
    if (c->meshExternal) return c->fail(MMS_ERR_INVALID, "the mesh was emitted into caller-supplied memory");
    DeviceGuard guard(c->device);
    const size_t bytes = static_cast<size_t>(c->ntris) * 36;
    *nverts = c->ntris * 3;
    if (pos) *pos = nullptr;
    if (nrm) *nrm = nullptr;
    if (col) *col = nullptr;
    if (bytes) {
        const bool wantCol = col && c->haveColour;
        auto growPin = [&](PinBuf& b) { return b.cap >= bytes || b.ensure(bytes + bytes / 8) || b.ensure(bytes); };
        if ((pos && !growPin(c->hPos)) || (nrm && !growPin(c->hNrm)) || (wantCol && !growPin(c->hCol)))
            return c->fail(MMS_ERR_NOMEM, "pinned allocation of %zu bytes failed", bytes);
        c->rec(EV_DM0);
        if (pos) MMS_CUDA(c, cudaMemcpyAsync(c->hPos.p, c->meshPos.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (nrm) MMS_CUDA(c, cudaMemcpyAsync(c->hNrm.p, c->meshNrm.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        if (wantCol) MMS_CUDA(c, cudaMemcpyAsync(c->hCol.p, c->meshCol.p, bytes, cudaMemcpyDeviceToHost, c->stream));
        c->rec(EV_DM1);
        MMS_CUDA(c, cudaStreamSynchronize(c->stream));
        if (pos) *pos = c->hPos.as<float>();
        if (nrm) *nrm = c->hNrm.as<float>();
        if (wantCol) *col = c->hCol.as<float>();
    }
    return MMS_OK;
}

int mms_get_home_voxels(mms_ctx* c, const int32_t** home, uint64_t* nparticles) {
    if (!c || !home || !nparticles) return MMS_ERR_INVALID;
    if (!c->haveDensity || !c->params.want_home_voxels) return c->fail(MMS_ERR_INVALID, "home voxels were not requested (params.want_home_voxels)");
    DeviceGuard guard(c->device);
    const size_t bytes = static_cast<size_t>(c->nparticles) * 12;
    *nparticles = c->nparticles;
    *home = nullptr;
    if (!bytes) return MMS_OK;
    if (!c->hHome.ensure(bytes)) return c->fail(MMS_ERR_NOMEM, "pinned allocation failed");
    MMS_CUDA(c, cudaMemcpyAsync(c->hHome.p, c->home.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    *home = c->hHome.as<int32_t>();
    return MMS_OK;
}

int mms_get_cell_tricounts(mms_ctx* c, const uint8_t** counts, uint64_t* ncells) {
    if (!c || !counts || !ncells) return MMS_ERR_INVALID;
    if (!c->haveMesh || !c->params.want_cell_tricounts) return c->fail(MMS_ERR_INVALID, "cell counts were not requested (params.want_cell_tricounts)");
    DeviceGuard guard(c->device);
    const size_t n = static_cast<size_t>(c->grid.res[0] - 1) * (c->grid.res[1] - 1) * std::max(c->cellNz, 0);
    *ncells = n;
    *counts = nullptr;
    if (!n) return MMS_OK;
    if (!c->hTri.ensure(n)) return c->fail(MMS_ERR_NOMEM, "pinned allocation failed");
    MMS_CUDA(c, cudaMemcpyAsync(c->hTri.p, c->triCount.p, n, cudaMemcpyDeviceToHost, c->stream));
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    *counts = c->hTri.as<uint8_t>();
    return MMS_OK;
}

int mms_synchronize(mms_ctx* c) {
    if (!c) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    MMS_CUDA(c, cudaStreamSynchronize(c->copyStream));
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    return MMS_OK;
}

int mms_get_timings(mms_ctx* c, mms_timings* t) {
    if (!c || !t) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    MMS_CUDA(c, cudaStreamSynchronize(c->copyStream));
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    auto el = [&](Ev a, Ev b) -> float {
        float ms = 0.0f;
        if (c->evSet[a] && c->evSet[b] && cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) == cudaSuccess) return ms;
        cudaGetLastError();
        return 0.0f;
    };
    t->h2d = el(EV_H2D0, EV_H2D1);
    t->bin = el(EV_BIN0, EV_BIN1);
    t->density = el(EV_BIN1, EV_DEN1);
    t->normalize = el(EV_NRM0, EV_NRM1);
    t->mc = el(EV_MC0, EV_MC1);
    t->d2h_volume = el(EV_DV0, EV_DV1);
    t->d2h_mesh = el(EV_DM0, EV_DM1);
    t->mc_emit = el(EV_EMIT0, EV_MC1);
    return MMS_OK;
}

int mms_route_particles(mms_ctx* c, const mms_list* list, int32_t nslabs, const int32_t* plane_lo, const int32_t* plane_hi, void* send_buf,
    uint64_t capacity_records, uint64_t* counts) {
    if (!c || !list || !plane_lo || !plane_hi || !counts || nslabs < 1) return MMS_ERR_INVALID;
    if (!c->haveGrid) return c->fail(MMS_ERR_INVALID, "mms_set_grid has not been called");
    if (nslabs > kMaxSlabs) return c->fail(MMS_ERR_UNSUPPORTED, "more than %d slabs", kMaxSlabs);
    if (list->vtx_type != MMS_VERT_FLOAT_XYZ && list->vtx_type != MMS_VERT_FLOAT_XYZR)
        return c->fail(MMS_ERR_UNSUPPORTED, "routing handles FLOAT_XYZ / FLOAT_XYZR lists");
    DeviceGuard guard(c->device);
    if (!isDevicePointer(list->vtx)) return c->fail(MMS_ERR_INVALID, "mms_route_particles needs a device-resident list");
    ListDev d{};
    d.vtx = static_cast<const char*>(list->vtx);
    d.count = list->count;
    d.vtype = list->vtx_type;
    d.vstride = list->vtx_stride ? list->vtx_stride : kVertSize[list->vtx_type];
    d.grad = list->global_radius;
    d.valign = alignOf(d.vtx, d.vstride, list->vtx_type == MMS_VERT_FLOAT_XYZR ? 16 : 4);
    if (d.valign < 4 || d.vstride % 4) return c->fail(MMS_ERR_UNSUPPORTED, "routing needs 4-byte aligned records");
    const Geo g = makeGeo(c);
    RouteGeo r{};
    r.zmin = g.mn[2], r.sdz = g.sd[2], r.sz = g.s[2], r.cyc = g.cyc[2], r.nslabs = nslabs;
    for (int i = 0; i < nslabs; ++i) {
        r.lo[i] = plane_lo[i], r.hi[i] = plane_hi[i];
        if (plane_lo[i] <= plane_hi[i]) r.enabled |= 1u << i; // lo > hi switches a slab off (e.g. the caller's own)
    }
    r.sigma = g.sigma, r.radscale = g.radscale, r.gausslim = g.gausslim, r.mode = g.mode;
    for (int i = 0; i < nslabs; ++i) counts[i] = 0;
    if (d.count == 0) return MMS_OK;
    const unsigned nwarps = static_cast<unsigned>(std::min<unsigned long long>((d.count + 255) / 256, static_cast<unsigned long long>(c->smCount) * 64));
    const unsigned long long chunk = ((d.count + nwarps - 1) / nwarps + 31) / 32 * 32;
    const unsigned nent = nwarps * static_cast<unsigned>(nslabs);
    const unsigned ntiles = (nent + kScanTile - 1) / kScanTile;
    if (!c->routeCounts.ensure(nent * 4) || !c->routeOffsets.ensure((nent + 1 + kMaxSlabs + 1) * 4) || !c->routeTile.ensure(std::max(ntiles, 1u) * 4) ||
        !c->hRoute.ensure((nslabs + 1) * 4))
        return c->fail(MMS_ERR_NOMEM, "allocation failed (routing tables)");
    cudaStream_t st = c->stream;
    const int blocks = static_cast<int>((static_cast<size_t>(nwarps) * 32 + 255) / 256);
    route_count_kernel<<<blocks, 256, 0, st>>>(r, d, chunk, c->routeCounts.as<unsigned>(), nwarps);
    ++c->launches;
    exclusiveScan(c->routeCounts.as<unsigned>(), c->routeOffsets.as<unsigned>(), nullptr, c->routeTile.as<unsigned>(), nent, nullptr, st, c->launches);
    // slab d starts at offsets[d * nwarps]; the grand total sits at offsets[nent]
    unsigned* h = c->hRoute.as<unsigned>();
    unsigned* heads = c->routeOffsets.as<unsigned>() + nent + 1;
    route_heads_kernel<<<1, 32, 0, st>>>(c->routeOffsets.as<unsigned>(), nwarps, nslabs, heads);
    ++c->launches;
    MMS_CUDA(c, cudaMemcpyAsync(h, heads, (nslabs + 1) * 4, cudaMemcpyDeviceToHost, st));
    MMS_CUDA(c, cudaStreamSynchronize(st));
    for (int i = 0; i < nslabs; ++i) counts[i] = h[i + 1] - h[i];
    if (h[nslabs] > capacity_records)
        return c->fail(MMS_ERR_NOMEM, "send buffer too small: %u records needed, %llu available", h[nslabs], static_cast<unsigned long long>(capacity_records));
    route_scatter_kernel<<<blocks, 256, 0, st>>>(r, d, chunk, c->routeOffsets.as<unsigned>(), nwarps, static_cast<unsigned*>(send_buf), static_cast<int>(d.vstride / 4));
    ++c->launches;
    MMS_CUDA(c, cudaGetLastError());
    return MMS_OK;
}

int mms_halo_buffers(mms_ctx* c, uint64_t cap, void** buf, void** counters) {
    if (!c || !buf || !counters || cap == 0) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    const bool fresh = c->haloCounters.p == nullptr;
    // per frame half: cap x y z r records, followed by their cap RGBA colours where the QuickSurf colour volume is on
    c->haloColour = c->params.mode != MMS_MODE_P2D_BUMP && c->params.colour != 0;
    if (!c->haloBuf.ensure(cap * 16 * 2 * (c->haloColour ? 2 : 1)) || !c->haloCounters.ensure(16)) return c->fail(MMS_ERR_NOMEM, "allocation of the halo receive buffer (%llu records) failed",
        static_cast<unsigned long long>(cap));
    if (fresh) MMS_CUDA(c, cudaMemset(c->haloCounters.p, 0, 16));
    c->haloCap = cap;
    *buf = c->haloBuf.p;
    *counters = c->haloCounters.p;
    return MMS_OK;
}

int mms_halo_push(mms_ctx* c, int32_t nslabs, int32_t mine, const int32_t* plane_lo, const int32_t* plane_hi, void* const* peer_bufs,
    void* const* peer_counters, uint64_t cap) {
    if (!c || !plane_lo || !plane_hi || !peer_bufs || !peer_counters || nslabs < 1 || mine < 0 || mine >= nslabs) return MMS_ERR_INVALID;
    if (!c->haveGrid) return c->fail(MMS_ERR_INVALID, "mms_set_grid has not been called");
    if (nslabs > kMaxSlabs) return c->fail(MMS_ERR_UNSUPPORTED, "more than %d slabs", kMaxSlabs);
    if (!c->haloCounters.p) return c->fail(MMS_ERR_INVALID, "mms_halo_buffers has not been called");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->stream;
    if (c->uploadPending) MMS_CUDA(c, cudaStreamWaitEvent(st, c->uploadDone, 0)); // (compute_density waits again; harmless)
    const Geo g = makeGeo(c);
    RouteGeo r{};
    r.zmin = g.mn[2], r.sdz = g.sd[2], r.sz = g.s[2], r.cyc = g.cyc[2], r.nslabs = nslabs;
    HaloPeers hp{};
    ++c->haloFrame;
    const unsigned word = c->haloFrame & 1u;
    for (int i = 0; i < nslabs; ++i) {
        r.lo[i] = plane_lo[i], r.hi[i] = plane_hi[i];
        if (i != mine && plane_lo[i] <= plane_hi[i] && peer_bufs[i] && peer_counters[i]) {
            r.enabled |= 1u << i;
            hp.buf[i] = static_cast<float4*>(peer_bufs[i]) + static_cast<size_t>(word) * cap * (c->haloColour ? 2 : 1); // frames alternate between the buffer's halves
            hp.counter[i] = static_cast<unsigned*>(peer_counters[i]) + word;
            hp.arrive[i] = static_cast<unsigned*>(peer_counters[i]) + 2 + word;
        }
    }
    hp.cap = static_cast<unsigned>(std::min<uint64_t>(cap, 0xffffffffull));
    hp.colour = c->haloColour ? 1 : 0;
    r.sigma = g.sigma, r.radscale = g.radscale, r.gausslim = g.gausslim, r.mode = g.mode;
    if (cap != c->haloCap) return c->fail(MMS_ERR_INVALID, "capacity_records must be the capacity every slab passed to mms_halo_buffers");
    if (r.enabled)
        for (const ListDev& l : c->lists) {
            if (l.countPtr) continue; // a received list is never forwarded
            halo_push_kernel<<<gridFor(l.count, 256, c->smCount * 16), 256, 0, st>>>(r, l, hp);
            ++c->launches;
        }
    // my own counters of the NEXT frame are cleared (nobody touches them before this frame's signal), then the peers learn that my
    // records of this frame have landed
    halo_signal_kernel<<<1, 32, 0, st>>>(hp, r.enabled, c->haloCounters.as<unsigned>() + (word ^ 1u), c->haloCounters.as<unsigned>() + 2 + (word ^ 1u));
    ++c->launches;
    MMS_CUDA(c, cudaGetLastError());
    return MMS_OK;
}

int mms_halo_wait(mms_ctx* c, int32_t npeers) {
    if (!c || npeers < 0) return MMS_ERR_INVALID;
    if (!c->haloCounters.p || !c->haloCap) return c->fail(MMS_ERR_INVALID, "mms_halo_buffers has not been called");
    // deferred: the wait kernel goes into the stream right before the first kernel that READS the received list (mms_compute_density
    // bins the context's own lists first), so the wait for the slowest neighbour hides behind work that does not need its records
    c->haloWaitPeers = npeers;
    return MMS_OK;
}

int mms_halo_receive(mms_ctx* c, float radius_bound) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haloCounters.p || !c->haloCap) return c->fail(MMS_ERR_INVALID, "mms_halo_buffers has not been called");
    if (c->lists.size() + 1 > static_cast<size_t>(kMaxLists)) return c->fail(MMS_ERR_UNSUPPORTED, "more than %d particle lists", kMaxLists);
    ListDev d{};
    d.vtx = static_cast<const char*>(c->haloBuf.p) + static_cast<size_t>(c->haloFrame & 1u) * c->haloCap * 16 * (c->haloColour ? 2 : 1); // this frame's half
    if (c->haloColour) { // converted colours behind the records: FLOAT_RGBA, for which quicksurfColour is the identity
        d.col = d.vtx + c->haloCap * 16;
        d.ctype = MMS_COL_FLOAT_RGBA;
        d.cstride = 16;
    }
    d.count = c->haloCap; // the bound; the length is read on the device
    d.countPtr = c->haloCounters.as<unsigned>() + (c->haloFrame & 1u);
    d.base = c->nparticles;
    d.vtype = MMS_VERT_FLOAT_XYZR;
    d.vstride = 16;
    d.valign = 16;
    d.calign = c->haloColour ? 16 : 4;
    d.grad = radius_bound;
    d.radiusBound = 1;
    c->lists.push_back(d);
    c->nparticles += c->haloCap;
    return MMS_OK;
}

int mms_timer_start(mms_ctx* c) {
    if (!c) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    c->rec(EV_T0);
    return MMS_OK;
}

int mms_timer_stop(mms_ctx* c, float* ms) {
    if (!c || !ms) return MMS_ERR_INVALID;
    DeviceGuard guard(c->device);
    c->rec(EV_T1);
    MMS_CUDA(c, cudaEventSynchronize(c->ev[EV_T1]));
    MMS_CUDA(c, cudaEventElapsedTime(ms, c->ev[EV_T0], c->ev[EV_T1]));
    return MMS_OK;
}

uint64_t mms_launch_count(const mms_ctx* c) { return c ? c->launches : 0; }

int mms_device_alloc(int32_t device, size_t bytes, void** ptr) {
    if (!ptr) return MMS_ERR_INVALID;
    DeviceGuard guard(device);
    if (cudaMalloc(ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return MMS_ERR_NOMEM;
    }
    return MMS_OK;
}

int mms_device_free(int32_t device, void* ptr) {
    DeviceGuard guard(device);
    return cudaFree(ptr) == cudaSuccess ? MMS_OK : MMS_ERR_CUDA;
}

int mms_ipc_export(int32_t device, const void* devptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!devptr || !handle) return MMS_ERR_INVALID;
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, const_cast<void*>(devptr)) != cudaSuccess) {
        cudaGetLastError();
        return MMS_ERR_CUDA;
    }
    std::memcpy(handle, &h, 64);
    return MMS_OK;
}

int mms_ipc_open(int32_t device, const unsigned char handle[64], void** ptr) {
    if (!handle || !ptr) return MMS_ERR_INVALID;
    DeviceGuard guard(device);
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    if (cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        return MMS_ERR_CUDA;
    }
    return MMS_OK;
}

int mms_ipc_close(int32_t device, void* ptr) {
    DeviceGuard guard(device);
    return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? MMS_OK : MMS_ERR_CUDA;
}

// ---- device-resident hand-off ------------------------------------------------------------------------------------
int mms_share_enable(mms_ctx* c, int32_t on) {
    if (!c) return MMS_ERR_INVALID;
    if (on && !Vmm::get().ok) return c->fail(MMS_ERR_UNSUPPORTED, "the driver has no virtual-memory API (cuMemCreate ...)");
    DeviceGuard guard(c->device);
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    for (DevBuf* b : {&c->vol, &c->rgb, &c->meshPos, &c->meshNrm, &c->meshCol}) {
        if (b->shareable != (on != 0)) b->release(); // re-allocated from the other kind of memory on the next frame
        b->shareable = on != 0;
        b->device = c->device;
    }
    c->haveDensity = c->haveMesh = c->haveCount = false;
    return MMS_OK;
}

static int shareOf(mms_ctx* c, DevBuf& b, uint64_t bytes, mms_share* out) {
    if (!out) return MMS_OK;
    *out = mms_share{-1, 0, 0, 0, 0};
    if (!bytes) return MMS_OK;
    if (!b.shareable || !b.handle) return c->fail(MMS_ERR_INVALID, "mms_share_enable was not in force when this buffer was produced");
    int fd = -1;
    if (Vmm::get().exportHandle(&fd, b.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || fd < 0)
        return c->fail(MMS_ERR_CUDA, "cuMemExportToShareableHandle failed");
    out->fd = fd;
    out->alloc_bytes = b.cap;
    out->offset = 0;
    out->bytes = bytes;
    return MMS_OK;
}

int mms_share_density(mms_ctx* c, mms_share* vol, mms_share* rgb) {
    if (!c) return MMS_ERR_INVALID;
    if (!c->haveDensity || c->adoptedVol) return c->fail(MMS_ERR_INVALID, "no density of this context's own has been computed");
    DeviceGuard guard(c->device);
    if (int rc = checkDeviceError(c)) return rc; // (synchronises: the importer has no stream of ours to wait on)
    const uint64_t nvox = static_cast<uint64_t>(c->grid.res[0]) * c->grid.res[1] * c->nz;
    if (int rc = shareOf(c, c->vol, nvox * 4, vol)) return rc;
    return shareOf(c, c->rgb, (c->haveColour || c->haveVector) ? nvox * 12 : 0, rgb);
}

int mms_share_mesh(mms_ctx* c, uint64_t* nverts, mms_share* pos, mms_share* nrm, mms_share* col) {
    if (!c || !nverts) return MMS_ERR_INVALID;
    if (!c->haveMesh || c->meshExternal) return c->fail(MMS_ERR_INVALID, "no isosurface has been extracted into library memory");
    if (c->countIndexed) return c->fail(MMS_ERR_UNSUPPORTED, "mms_share_mesh shares the triangle soup; the indexed mesh is read with mms_get_mesh_indexed_device");
    DeviceGuard guard(c->device);
    MMS_CUDA(c, cudaStreamSynchronize(c->stream));
    *nverts = c->ntris * 3;
    const uint64_t bytes = c->ntris * 36;
    if (int rc = shareOf(c, c->meshPos, bytes, pos)) return rc;
    if (int rc = shareOf(c, c->meshNrm, bytes, nrm)) return rc;
    return shareOf(c, c->meshCol, c->haveColour ? bytes : 0, col);
}

int mms_share_open(int32_t device, const mms_share* s, void** devptr) {
    if (!s || !devptr || s->fd < 0 || !s->alloc_bytes || !Vmm::get().ok) return MMS_ERR_INVALID;
    DeviceGuard guard(device);
    cudaFree(nullptr); // a context on this device, should the importer be a fresh process
    CUmemGenericAllocationHandle h = 0;
    if (Vmm::get().importHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(s->fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS)
        return MMS_ERR_CUDA;
    void* p = Vmm::mapHandle(h, s->alloc_bytes, device);
    Vmm::get().release(h); // the mapping keeps the memory alive
    if (!p) return MMS_ERR_CUDA;
    *devptr = static_cast<char*>(p) + s->offset;
    return MMS_OK;
}

int mms_share_close(int32_t device, void* devptr, const mms_share* s) {
    if (!devptr || !s || !Vmm::get().ok) return MMS_ERR_INVALID;
    DeviceGuard guard(device);
    const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(static_cast<char*>(devptr) - s->offset);
    if (Vmm::get().unmap(va, s->alloc_bytes) != CUDA_SUCCESS) return MMS_ERR_CUDA;
    Vmm::get().addrFree(va, s->alloc_bytes);
    return MMS_OK;
}

void* mms_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void mms_free_pinned(void* p) {
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// several GPUs behind one handle (single process)
// ---------------------------------------------------------------------------------------------------------------------------------
} // extern "C"

struct mms_slabs {
    std::vector<mms_ctx*> ctx;
    std::vector<int> dev;
    std::string err;
    mms_grid grid{};
    bool haveGrid = false, haveDensity = false, haveMesh = false;
    mms_params params{};
    struct Plan {
        int cellZ0, cellNz, z0, nz;
    };
    std::vector<Plan> plan;
    std::vector<cudaEvent_t> evPushed, evRange;
    std::vector<void*> haloBuf, haloCtr;
    std::vector<DevBuf> combined; // per device: the combined {-min, max}
    uint64_t haloCap = 0, nparticles = 0;
    float radiusBound = 0.0f;
    bool perParticleRadii = false;
    PinBuf hVol, hPos, hNrm, hRgb, hCol;
    uint64_t ntris = 0;
    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int failFrom(int code, mms_ctx* c) {
        err = c->err;
        return code;
    }
};

extern "C" {

const char* mms_slabs_last_error(const mms_slabs* s) { return s ? s->err.c_str() : g_createError.c_str(); }
int32_t mms_slabs_count(const mms_slabs* s) { return s ? static_cast<int32_t>(s->ctx.size()) : 0; }
mms_ctx* mms_slabs_context(mms_slabs* s, int32_t i) { return (s && i >= 0 && i < static_cast<int32_t>(s->ctx.size())) ? s->ctx[i] : nullptr; }

int mms_slabs_destroy(mms_slabs* s) {
    if (!s) return MMS_ERR_INVALID;
    for (size_t g = 0; g < s->ctx.size(); ++g) {
        DeviceGuard guard(s->dev[g]);
        cudaStreamSynchronize(s->ctx[g]->stream);
        if (g < s->evPushed.size()) cudaEventDestroy(s->evPushed[g]), cudaEventDestroy(s->evRange[g]);
        if (g < s->combined.size()) s->combined[g].release();
    }
    for (mms_ctx* c : s->ctx) mms_destroy(c);
    s->hVol.release(), s->hPos.release(), s->hNrm.release(), s->hRgb.release(), s->hCol.release();
    delete s;
    return MMS_OK;
}

int mms_slabs_create(mms_slabs** out, const int32_t* devices, int32_t n) {
    if (!out || !devices || n < 1) return MMS_ERR_INVALID;
    *out = nullptr;
    if (n > kMaxSlabs) {
        g_createError = "too many devices for one slab group";
        return MMS_ERR_UNSUPPORTED;
    }
    auto* s = new mms_slabs();
    for (int g = 0; g < n; ++g) {
        mms_config cfg{devices[g], 0};
        mms_ctx* c = nullptr;
        if (int rc = mms_create(&c, &cfg)) {
            mms_slabs_destroy(s);
            return rc;
        }
        s->ctx.push_back(c);
        s->dev.push_back(devices[g]);
    }
    s->combined.resize(n);
    for (int g = 0; g < n; ++g) {
        DeviceGuard guard(devices[g]);
        for (int h = 0; h < n; ++h) {
            if (h == g || devices[h] == devices[g]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[g], devices[h]);
            if (!can) {
                g_createError = "the devices of a slab group need peer access to each other (NVLink / PCIe P2P)";
                mms_slabs_destroy(s);
                return MMS_ERR_UNSUPPORTED;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[h], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                g_createError = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                cudaGetLastError();
                mms_slabs_destroy(s);
                return MMS_ERR_CUDA;
            }
            cudaGetLastError();
        }
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&b, cudaEventDisableTiming);
        s->evPushed.push_back(a), s->evRange.push_back(b);
        if (!s->combined[g].ensure(16)) {
            g_createError = "allocation failed";
            mms_slabs_destroy(s);
            return MMS_ERR_NOMEM;
        }
    }
    s->params = s->ctx[0]->params;
    *out = s;
    return MMS_OK;
}

int mms_slabs_set_grid(mms_slabs* s, const mms_grid* grid) {
    if (!s || !grid) return MMS_ERR_INVALID;
    const int G = static_cast<int>(s->ctx.size()), sz = grid->res[2], ncell = sz - 1;
    s->plan.clear();
    for (int g = 0; g < G; ++g) { // megamol_b200/slabs.py plan_slabs
        const int c0 = static_cast<int>(static_cast<long long>(g) * ncell / G), c1 = static_cast<int>(static_cast<long long>(g + 1) * ncell / G);
        int p0 = std::max(c0 - 1, 0), p1 = std::min(c1 + 1, sz - 1);
        if (c1 <= c0) p0 = p1 = std::min(c0, sz - 1); // more devices than cell layers: this one idles on a one-plane slab
        s->plan.push_back({c0, c1 - c0, p0, p1 - p0 + 1});
    }
    for (int g = 0; g < G; ++g) {
        if (int rc = mms_set_grid(s->ctx[g], grid)) return s->failFrom(rc, s->ctx[g]);
        if (G > 1)
            if (int rc = mms_set_slab(s->ctx[g], s->plan[g].z0, s->plan[g].nz, s->plan[g].cellZ0, s->plan[g].cellNz)) return s->failFrom(rc, s->ctx[g]);
    }
    s->grid = *grid;
    s->haveGrid = true;
    s->haveDensity = s->haveMesh = false;
    return MMS_OK;
}

int mms_slabs_set_params(mms_slabs* s, const mms_params* p) {
    if (!s || !p) return MMS_ERR_INVALID;
    const int G = static_cast<int>(s->ctx.size());
    // halo records travel as x y z r (+ RGBA where the QuickSurf colour volume is on): the P2D bump with aggregator 0, or the QuickSurf
    // Gaussian with the radial cut-off
    if (G > 1 && !((p->mode == MMS_MODE_P2D_BUMP && p->aggregator == 0) || p->mode == MMS_MODE_QS_GAUSS))
        return s->fail(MMS_ERR_UNSUPPORTED, "a slab group of several devices computes ParticlesToDensity with aggregator 0 or the QuickSurf "
                                            "Gaussian with the radial cut-off");
    mms_params q = *p;
    if (G > 1) q.defer_normalize = 1; // the range is global: normalised after the slabs' ranges have been combined
    for (mms_ctx* c : s->ctx)
        if (int rc = mms_set_params(c, &q)) return s->failFrom(rc, c);
    s->params = *p;
    return MMS_OK;
}

int mms_slabs_clear_particles(mms_slabs* s) {
    if (!s) return MMS_ERR_INVALID;
    for (mms_ctx* c : s->ctx) mms_clear_particles(c);
    s->nparticles = 0;
    s->radiusBound = 0.0f;
    s->perParticleRadii = false;
    return MMS_OK;
}

int mms_slabs_push_particles(mms_slabs* s, int32_t nlists, const mms_list* lists) {
    if (!s || nlists < 0 || (nlists > 0 && !lists)) return MMS_ERR_INVALID;
    const int G = static_cast<int>(s->ctx.size());
    for (int i = 0; i < nlists; ++i) {
        const mms_list& l = lists[i];
        if (l.vtx_type == MMS_VERT_NONE || l.count == 0) continue;
        if (l.vtx_type < 0 || l.vtx_type > 4 || l.col_type < 0 || l.col_type > 7 || !l.vtx) return s->fail(MMS_ERR_INVALID, "bad list %d", i);
        if (G > 1 && isDevicePointer(l.vtx)) return s->fail(MMS_ERR_UNSUPPORTED, "a slab group takes host lists (they are split over the devices' PCIe links)");
        const unsigned vstride = l.vtx_stride ? l.vtx_stride : kVertSize[l.vtx_type];
        const int ctype = l.col ? l.col_type : MMS_COL_NONE;
        const unsigned cstride = l.col_stride ? l.col_stride : kColSize[ctype];
        if (l.vtx_type == MMS_VERT_FLOAT_XYZR) s->perParticleRadii = true;
        else if (std::isfinite(l.global_radius)) s->radiusBound = std::max(s->radiusBound, l.global_radius);
        for (int g = 0; g < G; ++g) {
            const uint64_t i0 = l.count * static_cast<uint64_t>(g) / G, i1 = l.count * static_cast<uint64_t>(g + 1) / G;
            if (i1 <= i0) continue;
            mms_list sub = l;
            sub.vtx = static_cast<const char*>(l.vtx) + i0 * vstride;
            sub.vtx_stride = vstride;
            if (ctype) sub.col = static_cast<const char*>(l.col) + i0 * cstride, sub.col_stride = cstride;
            sub.count = i1 - i0;
            if (int rc = mms_push_particles(s->ctx[g], 1, &sub)) return s->failFrom(rc, s->ctx[g]);
        }
        s->nparticles += l.count;
    }
    return MMS_OK;
}

int mms_slabs_compute_density(mms_slabs* s) {
    if (!s || !s->haveGrid) return s ? s->fail(MMS_ERR_INVALID, "mms_slabs_set_grid has not been called") : MMS_ERR_INVALID;
    const int G = static_cast<int>(s->ctx.size());
    s->haveDensity = s->haveMesh = false;
    if (G == 1) {
        if (int rc = mms_compute_density(s->ctx[0])) return s->failFrom(rc, s->ctx[0]);
        s->haveDensity = true;
        return MMS_OK;
    }
    // per-particle radii: the largest one sizes everybody's sort cells (one device scan per share; the only host wait of this call)
    float bound = s->radiusBound;
    if (s->perParticleRadii)
        for (mms_ctx* c : s->ctx) {
            float r = 0.0f;
            if (int rc = mms_get_max_radius(c, &r)) return s->failFrom(rc, c);
            bound = std::max(bound, r);
        }
    // ---- halo exchange: push into the peers' buffers, ordered by events --------------------------------------------------------------
    const uint64_t cap = std::max<uint64_t>(s->nparticles, 1);
    s->haloBuf.assign(G, nullptr), s->haloCtr.assign(G, nullptr);
    for (int g = 0; g < G; ++g)
        if (int rc = mms_halo_buffers(s->ctx[g], cap, &s->haloBuf[g], &s->haloCtr[g])) return s->failFrom(rc, s->ctx[g]);
    std::vector<int32_t> lo(G), hi(G);
    for (int g = 0; g < G; ++g) lo[g] = s->plan[g].z0, hi[g] = s->plan[g].z0 + s->plan[g].nz - 1;
    for (int g = 0; g < G; ++g) {
        mms_ctx* c = s->ctx[g];
        if (int rc = mms_halo_push(c, G, g, lo.data(), hi.data(), s->haloBuf.data(), s->haloCtr.data(), cap)) return s->failFrom(rc, c);
        DeviceGuard guard(s->dev[g]);
        MMS_CUDA(c, cudaEventRecord(s->evPushed[g], c->stream));
    }
    for (int g = 0; g < G; ++g) {
        mms_ctx* c = s->ctx[g];
        {
            DeviceGuard guard(s->dev[g]);
            for (int h = 0; h < G; ++h)
                if (h != g) MMS_CUDA(c, cudaStreamWaitEvent(c->stream, s->evPushed[h], 0)); // every peer's push into MY buffer is complete
        }
        if (int rc = mms_halo_receive(c, bound)) return s->failFrom(rc, c);
        if (int rc = mms_compute_density(c)) return s->failFrom(rc, c);
    }
    // ---- global range -> normalise (ParticlesToDensity.cpp:669-682) -----------------------------------------------------------------
    if (s->params.mode == MMS_MODE_P2D_BUMP && s->params.normalize && !s->params.defer_normalize) {
        RangePeers rp{};
        rp.n = G;
        for (int g = 0; g < G; ++g) {
            mms_ctx* c = s->ctx[g];
            float* r = nullptr;
            if (int rc = mms_density_range_device(c, &r)) return s->failFrom(rc, c);
            rp.r[g] = r;
            DeviceGuard guard(s->dev[g]);
            MMS_CUDA(c, cudaEventRecord(s->evRange[g], c->stream));
        }
        for (int g = 0; g < G; ++g) {
            mms_ctx* c = s->ctx[g];
            {
                DeviceGuard guard(s->dev[g]);
                for (int h = 0; h < G; ++h)
                    if (h != g) MMS_CUDA(c, cudaStreamWaitEvent(c->stream, s->evRange[h], 0));
                range_combine_kernel<<<1, 1, 0, c->stream>>>(rp, s->combined[g].as<float>());
                ++c->launches;
            }
            if (int rc = mms_normalize_device(c, s->combined[g].as<float>())) return s->failFrom(rc, c);
        }
    }
    s->haveDensity = true;
    return MMS_OK;
}

int mms_slabs_get_density_range(mms_slabs* s, float minmax[2]) {
    if (!s || !minmax) return MMS_ERR_INVALID;
    if (!s->haveDensity) return s->fail(MMS_ERR_INVALID, "no density has been computed");
    float mn = INFINITY, mx = -INFINITY;
    for (mms_ctx* c : s->ctx) {
        float r[2];
        if (int rc = mms_get_density_range(c, r)) return s->failFrom(rc, c);
        mn = std::min(mn, r[0]), mx = std::max(mx, r[1]);
    }
    minmax[0] = mn, minmax[1] = mx;
    return MMS_OK;
}

int mms_slabs_get_density(mms_slabs* s, const float** hv) {
    if (!s || !hv) return MMS_ERR_INVALID;
    if (!s->haveDensity) return s->fail(MMS_ERR_INVALID, "no density has been computed");
    const int G = static_cast<int>(s->ctx.size());
    if (G == 1) {
        if (int rc = mms_get_density(s->ctx[0], hv, nullptr)) return s->failFrom(rc, s->ctx[0]);
        return MMS_OK;
    }
    const size_t plane = static_cast<size_t>(s->grid.res[0]) * s->grid.res[1];
    if (!s->hVol.ensure(plane * s->grid.res[2] * 4)) return s->fail(MMS_ERR_NOMEM, "pinned allocation of the volume failed");
    for (int g = 0; g < G; ++g) { // every slab sends the planes of its own cell layers (the last one also the final plane)
        mms_ctx* c = s->ctx[g];
        const int p0 = s->plan[g].cellZ0, p1 = g == G - 1 ? s->grid.res[2] : s->plan[g].cellZ0 + s->plan[g].cellNz;
        if (p1 <= p0) continue;
        DeviceGuard guard(s->dev[g]);
        MMS_CUDA(c, cudaMemcpyAsync(s->hVol.as<float>() + plane * p0, c->vol.as<float>() + plane * (p0 - s->plan[g].z0), plane * (p1 - p0) * 4,
            cudaMemcpyDeviceToHost, c->stream));
    }
    for (mms_ctx* c : s->ctx)
        if (int rc = checkDeviceError(c)) return s->failFrom(rc, c);
    *hv = s->hVol.as<float>();
    return MMS_OK;
}

int mms_slabs_adopt_density(mms_slabs* s, mms_slabs* p) {
    if (!s || !p || s == p) return MMS_ERR_INVALID;
    if (!p->haveDensity) return s->fail(MMS_ERR_INVALID, "the producer group has no density");
    if (s->dev != p->dev) return s->fail(MMS_ERR_INVALID, "mms_slabs_adopt_density: the two groups must use the same devices in the same order");
    for (size_t g = 0; g < s->ctx.size(); ++g)
        if (int rc = mms_adopt_density(s->ctx[g], p->ctx[g])) return s->failFrom(rc, s->ctx[g]);
    s->grid = p->grid, s->plan = p->plan, s->params = p->params;
    s->haveGrid = s->haveDensity = true;
    s->haveMesh = false;
    return MMS_OK;
}

int mms_slabs_extract_isosurface(mms_slabs* s, float iso) {
    if (!s) return MMS_ERR_INVALID;
    if (!s->haveDensity) return s->fail(MMS_ERR_INVALID, "no density has been computed");
    for (mms_ctx* c : s->ctx) // all counts in flight before the first host wait
        if (int rc = countLaunch(c, iso)) return s->failFrom(rc, c);
    s->ntris = 0;
    for (mms_ctx* c : s->ctx) {
        uint64_t n = 0;
        if (int rc = countFinish(c, &n)) return s->failFrom(rc, c);
        s->ntris += n;
    }
    for (mms_ctx* c : s->ctx)
        if (int rc = mms_emit_isosurface(c, nullptr, nullptr, nullptr, 0)) return s->failFrom(rc, c);
    s->haveMesh = true;
    return MMS_OK;
}

int mms_slabs_get_mesh(mms_slabs* s, uint64_t* nverts, const float** pos, const float** nrm) {
    if (!s || !nverts) return MMS_ERR_INVALID;
    if (!s->haveMesh) return s->fail(MMS_ERR_INVALID, "no isosurface has been extracted");
    *nverts = s->ntris * 3;
    if (pos) *pos = nullptr;
    if (nrm) *nrm = nullptr;
    const size_t bytes = static_cast<size_t>(s->ntris) * 36;
    if (!bytes) return MMS_OK;
    if ((pos && !s->hPos.ensure(bytes)) || (nrm && !s->hNrm.ensure(bytes))) return s->fail(MMS_ERR_NOMEM, "pinned allocation of the mesh (%zu bytes) failed", bytes);
    size_t off = 0;
    for (size_t g = 0; g < s->ctx.size(); ++g) { // slab order = cell-linear order; every slab over its own link
        mms_ctx* c = s->ctx[g];
        const size_t b = static_cast<size_t>(c->ntris) * 36;
        if (b) {
            DeviceGuard guard(s->dev[g]);
            if (pos) MMS_CUDA(c, cudaMemcpyAsync(s->hPos.as<char>() + off, c->meshPos.p, b, cudaMemcpyDeviceToHost, c->stream));
            if (nrm) MMS_CUDA(c, cudaMemcpyAsync(s->hNrm.as<char>() + off, c->meshNrm.p, b, cudaMemcpyDeviceToHost, c->stream));
        }
        off += b;
    }
    for (size_t g = 0; g < s->ctx.size(); ++g) {
        DeviceGuard guard(s->dev[g]);
        MMS_CUDA(s->ctx[g], cudaStreamSynchronize(s->ctx[g]->stream));
    }
    if (pos) *pos = s->hPos.as<float>();
    if (nrm) *nrm = s->hNrm.as<float>();
    return MMS_OK;
}

int mms_slabs_get_colour_volume(mms_slabs* s, const float** hrgb) {
    if (!s || !hrgb) return MMS_ERR_INVALID;
    *hrgb = nullptr;
    if (!s->haveDensity) return s->fail(MMS_ERR_INVALID, "no density has been computed");
    const int G = static_cast<int>(s->ctx.size());
    if (G == 1) {
        const float* v = nullptr;
        if (int rc = mms_get_density(s->ctx[0], &v, hrgb)) return s->failFrom(rc, s->ctx[0]);
        return MMS_OK;
    }
    if (!s->ctx[0]->haveColour) return MMS_OK; // no colour volume in this mode: NULL, like mms_get_density
    const size_t plane = static_cast<size_t>(s->grid.res[0]) * s->grid.res[1] * 3;
    if (!s->hRgb.ensure(plane * s->grid.res[2] * 4)) return s->fail(MMS_ERR_NOMEM, "pinned allocation of the colour volume failed");
    for (int g = 0; g < G; ++g) { // the planes of the slab's own cell layers, as in mms_slabs_get_density
        mms_ctx* c = s->ctx[g];
        const int p0 = s->plan[g].cellZ0, p1 = g == G - 1 ? s->grid.res[2] : s->plan[g].cellZ0 + s->plan[g].cellNz;
        if (p1 <= p0) continue;
        DeviceGuard guard(s->dev[g]);
        MMS_CUDA(c, cudaMemcpyAsync(s->hRgb.as<float>() + plane * p0, c->isoRgb() + plane * (p0 - s->plan[g].z0), plane * (p1 - p0) * 4,
            cudaMemcpyDeviceToHost, c->stream));
    }
    for (mms_ctx* c : s->ctx)
        if (int rc = checkDeviceError(c)) return s->failFrom(rc, c);
    *hrgb = s->hRgb.as<float>();
    return MMS_OK;
}

int mms_slabs_get_mesh_colours(mms_slabs* s, const float** col) {
    if (!s || !col) return MMS_ERR_INVALID;
    *col = nullptr;
    if (!s->haveMesh) return s->fail(MMS_ERR_INVALID, "no isosurface has been extracted");
    if (s->ctx.size() == 1) {
        uint64_t n = 0;
        if (int rc = mms_get_mesh(s->ctx[0], &n, nullptr, nullptr, col)) return s->failFrom(rc, s->ctx[0]);
        return MMS_OK;
    }
    const size_t bytes = static_cast<size_t>(s->ntris) * 36;
    if (!bytes || !s->ctx[0]->haveColour) return MMS_OK;
    if (!s->hCol.ensure(bytes)) return s->fail(MMS_ERR_NOMEM, "pinned allocation of the mesh colours (%zu bytes) failed", bytes);
    size_t off = 0;
    for (size_t g = 0; g < s->ctx.size(); ++g) {
        mms_ctx* c = s->ctx[g];
        const size_t b = static_cast<size_t>(c->ntris) * 36;
        if (b) {
            DeviceGuard guard(s->dev[g]);
            MMS_CUDA(c, cudaMemcpyAsync(s->hCol.as<char>() + off, c->meshCol.p, b, cudaMemcpyDeviceToHost, c->stream));
        }
        off += b;
    }
    for (size_t g = 0; g < s->ctx.size(); ++g) {
        DeviceGuard guard(s->dev[g]);
        MMS_CUDA(s->ctx[g], cudaStreamSynchronize(s->ctx[g]->stream));
    }
    *col = s->hCol.as<float>();
    return MMS_OK;
}

} // extern "C"

