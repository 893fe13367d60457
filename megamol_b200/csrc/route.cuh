// route.cuh -- stable partition of a device-resident particle list by destination z-slab (multi-GPU exchange).
//
// A particle goes to every slab its support box [Z-f, Z+f] touches (periodic images in cyclic mode); Z and f are the
// reference's home voxel / filter size (ParticlesToDensity.cpp:567,575), computed with the same individually rounded fp32
// operations as the binning kernels.  Each warp owns one contiguous chunk of the list:
//   route_count_kernel    per (slab, warp) record counts
//   (exclusive scan over the slab-major count matrix: scan.cuh)
//   route_scatter_kernel  re-walks the chunk 32 particles at a time; a ballot per destination gives every record its
//                         position -> the send buffer holds, per destination, the records in their original order
// so that after the all-to-all-v (receivers concatenate in source-rank order) the global particle order is preserved.
#pragma once
#include "common.cuh"

namespace mms {

constexpr int kMaxSlabs = 16;

struct RouteGeo {
    float zmin, sdz;
    int sz, cyc;
    int nslabs;
    int lo[kMaxSlabs], hi[kMaxSlabs]; // plane range [lo, hi] each slab computes (incl. its halo planes); lo > hi: slab switched off
    unsigned enabled;                 // bit d: slab d takes part
    float sigma, radscale, gausslim;
    int mode;
};

__device__ __forceinline__ unsigned routeMask(const RouteGeo& r, const ListDev& l, unsigned long long j) {
    const float4 p = fetchParticle(l, j);
    const int Z = homeVoxel(p.z, r.zmin, r.sdz);
    int f;
    if (r.mode == 0) f = filterSize(p.w, r.sdz);
    else f = filterSize(r.gausslim * r.radscale * p.w, r.sdz) + 1;
    unsigned m = 0;
    if (!(p.w > 0.0f) || !isfinite(p.w) || !isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) return 0u; // never contributes (bin.cuh)
    if (!r.cyc) {
        for (int d = 0; d < r.nslabs; ++d)
            if (Z + f >= r.lo[d] && Z - f <= r.hi[d]) m |= 1u << d;
    } else if (2 * f + 1 >= r.sz) {
        m = r.enabled;
    } else {
        const int zw = (static_cast<unsigned>(Z) < static_cast<unsigned>(r.sz)) ? Z : floorMod(Z, r.sz);
        const int a = zw - f, b = zw + f;
        for (int d = 0; d < r.nslabs; ++d) {
            const int lo = r.lo[d], hi = r.hi[d];
            if ((b >= lo && a <= hi) || (b - r.sz >= lo && a - r.sz <= hi) || (b + r.sz >= lo && a + r.sz <= hi)) m |= 1u << d;
        }
    }
    return m & r.enabled;
}

/** offsets[d * nwarps] for d = 0..nslabs (slab starts + grand total) -> one compact array: a single small D2H copy */
__global__ void route_heads_kernel(const unsigned* __restrict__ offsets, unsigned nwarps, int nslabs, unsigned* __restrict__ heads) {
    const int d = threadIdx.x;
    if (d <= nslabs) heads[d] = offsets[static_cast<size_t>(d) * nwarps];
}

__global__ void __launch_bounds__(256) route_count_kernel(RouteGeo r, ListDev l, unsigned long long chunk, unsigned* __restrict__ counts,
    unsigned nwarps) {
    const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const unsigned long long beg = static_cast<unsigned long long>(w) * chunk, end = min(beg + chunk, l.count);
    unsigned cnt[kMaxSlabs];
#pragma unroll
    for (int d = 0; d < kMaxSlabs; ++d) cnt[d] = 0;
    for (unsigned long long j0 = beg; j0 < end; j0 += 32) {
        const unsigned long long j = j0 + lane;
        const unsigned m = j < end ? routeMask(r, l, j) : 0u;
#pragma unroll
        for (int d = 0; d < kMaxSlabs; ++d)
            if (d < r.nslabs) cnt[d] += __popc(__ballot_sync(0xffffffffu, (m >> d) & 1u));
    }
    if (lane == 0)
        for (int d = 0; d < r.nslabs; ++d) counts[static_cast<size_t>(d) * nwarps + w] = cnt[d];
}

/** recWords = record size in 4-byte words (the list's stride): records travel as they are (vertex + interleaved colour). */
__global__ void __launch_bounds__(256) route_scatter_kernel(RouteGeo r, ListDev l, unsigned long long chunk, const unsigned* __restrict__ offsets,
    unsigned nwarps, unsigned* __restrict__ out, int recWords) {
    const unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nwarps) return;
    const unsigned long long beg = static_cast<unsigned long long>(w) * chunk, end = min(beg + chunk, l.count);
    unsigned base[kMaxSlabs];
#pragma unroll
    for (int d = 0; d < kMaxSlabs; ++d) base[d] = d < r.nslabs ? offsets[static_cast<size_t>(d) * nwarps + w] : 0u;
    const unsigned lt = (1u << lane) - 1u;
    for (unsigned long long j0 = beg; j0 < end; j0 += 32) {
        const unsigned long long j = j0 + lane;
        const unsigned m = j < end ? routeMask(r, l, j) : 0u;
        const unsigned* src = reinterpret_cast<const unsigned*>(l.vtx + j * l.vstride);
#pragma unroll
        for (int d = 0; d < kMaxSlabs; ++d) {
            if (d >= r.nslabs) break;
            const unsigned b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
            if ((m >> d) & 1u) {
                unsigned* dst = out + static_cast<size_t>(base[d] + __popc(b & lt)) * recWords;
                for (int k = 0; k < recWords; ++k) dst[k] = src[k];
            }
            base[d] += __popc(b);
        }
    }
}

/** Receive buffers of the other slabs (peer memory: CUDA IPC mappings or peer access) and their record counters. */
struct HaloPeers {
    float4* buf[kMaxSlabs];       // this frame's half of the peer's receive buffer
    unsigned* counter[kMaxSlabs]; // this frame's record counter of the peer
    unsigned* arrive[kMaxSlabs];  // this frame's arrival counter of the peer: +1 once ALL my pushes of the frame have landed
    unsigned cap;
    int colour;                   // QuickSurf colour volume: the record's converted RGBA goes to slot cap + i of the same half
};

/** After the push kernels of a frame (same stream): clear MY counters of the next frame, then tell every peer that my records of this
 *  frame are complete.  The peers' halo_wait_kernel spins on that count -- the "every push has completed" point needs no collective
 *  and no host.  (The clear comes first: a peer can only push or signal the next frame after it has seen this signal.) */
__global__ void halo_signal_kernel(HaloPeers hp, unsigned enabled, unsigned* myNextCounter, unsigned* myNextArrive) {
    if (threadIdx.x == 0) *myNextCounter = 0u, *myNextArrive = 0u;
    __threadfence_system(); // (the push kernels before this one in the stream have completed: their stores are performed)
    __syncthreads();
    if (threadIdx.x < kMaxSlabs && ((enabled >> threadIdx.x) & 1u)) atomicAdd_system(hp.arrive[threadIdx.x], 1u);
}

/** Stream-ordered wait for `npeers` arrivals (see halo_signal_kernel).  Gives up after ~4 s: the record counter is then poisoned, which
 *  bin_count_kernel reports like an overflow of the receive buffer (the frame fails at the next host-synchronising call). */
__global__ void halo_wait_kernel(const unsigned* arrive, unsigned npeers, unsigned* myCounter) {
    const long long t0 = clock64();
    while (*reinterpret_cast<const volatile unsigned*>(arrive) < npeers) {
        if (clock64() - t0 > (8ll << 30)) {
            atomicExch(myCounter, 0xffffffffu);
            break;
        }
        __nanosleep(200);
    }
    __threadfence_system();
}

/**
 * Halo exchange in ONE kernel, no host round trip, no collective: every record of the list that another slab needs (routeMask; the
 * caller switches its own slab off) is appended, as an x y z r record (plus its RGBA colour where the QuickSurf colour volume is on), straight to that slab's receive buffer in peer memory.  A warp
 * claims its slots with one system-scope atomicAdd per destination (ballot-aggregated), so the NVLink atomics stay few.  The order of
 * arrival is arbitrary -- the canonical in-cell order of the binning makes the density independent of it.
 */
__global__ void __launch_bounds__(256) halo_push_kernel(RouteGeo r, ListDev l, HaloPeers hp) {
    const unsigned lane = threadIdx.x & 31, lt = (1u << lane) - 1u;
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long count = listCount(l), rounded = (count + 31) & ~31ull;
    for (unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < rounded; j += stride) {
        const unsigned m = j < count ? routeMask(r, l, j) : 0u;
        unsigned all = __reduce_or_sync(0xffffffffu, m);
        if (!all) continue;
        const float4 p = m ? fetchParticle(l, j) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        // the colour the binning of a single context would attach to this record (bin_scatter_kernel); the receiver's list is FLOAT_RGBA,
        // for which the conversion is the identity
        const float4 col = (m && hp.colour) ? quicksurfColour(l, fetchColourRaw(l, j)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        for (; all; all &= all - 1) {
            const int d = __ffs(all) - 1;
            const unsigned b = __ballot_sync(0xffffffffu, (m >> d) & 1u);
            unsigned base = 0;
            if (lane == static_cast<unsigned>(__ffs(b) - 1)) base = atomicAdd_system(hp.counter[d], static_cast<unsigned>(__popc(b)));
            base = __shfl_sync(0xffffffffu, base, __ffs(b) - 1);
            if ((m >> d) & 1u) {
                const unsigned slot = base + __popc(b & lt);
                if (slot < hp.cap) { // (an overflow shows in the counter: the receiver reports it)
                    hp.buf[d][slot] = p;
                    if (hp.colour) hp.buf[d][hp.cap + slot] = col;
                }
            }
        }
    }
}

} // namespace mms
