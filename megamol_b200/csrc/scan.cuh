// scan.cuh -- exclusive prefix sum over uint32 counters (cell histogram, marching-cubes segment counts).
// Three launches: per-tile reduce, scan of the tile sums (one block), per-tile scan + offset.
// out has n+1 entries: out[n] = total.  Optionally mirrors out[0..n) into `copy` (the scatter cursors).
#pragma once
#include "common.cuh"

namespace mms {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;  // per thread: one 16-byte vector, so that a warp reads and writes 512 contiguous bytes
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned warpInclusiveScan(unsigned v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

/** Block-wide exclusive scan of one value per thread; returns the exclusive prefix, *total = block sum. */
__device__ __forceinline__ unsigned blockExclusiveScan(unsigned v, unsigned* total, unsigned* warpSums /* [32] smem */) {
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    const unsigned inc = warpInclusiveScan(v);
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned w = lane < nwarp ? warpSums[lane] : 0u;
        const unsigned winc = warpInclusiveScan(w);
        warpSums[lane] = winc - w; // exclusive
        if (lane == 31) warpSums[32] = winc;
    }
    __syncthreads();
    const unsigned res = warpSums[warp] + inc - v;
    *total = warpSums[32];
    __syncthreads();
    return res;
}

/** Four consecutive counters of a thread: one aligned 16-byte load where the tile is complete (cudaMalloc'ed arrays), else scalar. */
__device__ __forceinline__ uint4 scanLoad4(const unsigned* __restrict__ in, size_t i, unsigned n) {
    if (i + 3 < n) return *reinterpret_cast<const uint4*>(in + i);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (i < n) v.x = in[i];
    if (i + 1 < n) v.y = in[i + 1];
    if (i + 2 < n) v.z = in[i + 2];
    return v;
}

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ tileSums, unsigned n) {
    __shared__ unsigned ws[33];
    const size_t i = (static_cast<size_t>(blockIdx.x) * kScanThreads + threadIdx.x) * kScanItems;
    const uint4 v = scanLoad4(in, i, n);
    unsigned total;
    blockExclusiveScan(v.x + v.y + v.z + v.w, &total, ws);
    if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tilesums_kernel(unsigned* __restrict__ tileSums, unsigned ntiles, unsigned* __restrict__ grandTotal,
    unsigned long long* __restrict__ total64) {
    __shared__ unsigned ws[33];
    unsigned carry = 0;
    for (unsigned base = 0; base < ntiles; base += blockDim.x) {
        const unsigned i = base + threadIdx.x;
        const unsigned v = i < ntiles ? tileSums[i] : 0u;
        unsigned total;
        const unsigned ex = blockExclusiveScan(v, &total, ws);
        if (i < ntiles) tileSums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        *grandTotal = carry;
        if (total64) *total64 = carry;
    }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const unsigned* __restrict__ in, const unsigned* __restrict__ tileSums,
    unsigned* __restrict__ out, unsigned* __restrict__ copy, unsigned n) {
    __shared__ unsigned ws[33];
    const size_t i = (static_cast<size_t>(blockIdx.x) * kScanThreads + threadIdx.x) * kScanItems;
    const uint4 v = scanLoad4(in, i, n);
    unsigned total;
    const unsigned ex = blockExclusiveScan(v.x + v.y + v.z + v.w, &total, ws) + tileSums[blockIdx.x];
    const uint4 o = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
    if (i + 3 < n) { // (out[n], the grand total, belongs to the tile-sums kernel: a full vector never reaches it)
        *reinterpret_cast<uint4*>(out + i) = o;
        if (copy) *reinterpret_cast<uint4*>(copy + i) = o;
    } else {
        const unsigned e[3] = {o.x, o.y, o.z};
        for (int k = 0; k < 3; ++k)
            if (i + k < n) {
                out[i + k] = e[k];
                if (copy) copy[i + k] = e[k];
            }
    }
}

/** out[0..n] (n+1 entries). tileSums needs ceil(n/kScanTile) entries. out[n] is written by the tilesums kernel. */
inline void exclusiveScan(const unsigned* in, unsigned* out, unsigned* copy, unsigned* tileSums, unsigned n,
    unsigned long long* total64, cudaStream_t st, unsigned long long& launches) {
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(unsigned), st);
        if (total64) cudaMemsetAsync(total64, 0, sizeof(unsigned long long), st);
        return;
    }
    const unsigned ntiles = (n + kScanTile - 1) / kScanTile;
    scan_reduce_kernel<<<ntiles, kScanThreads, 0, st>>>(in, tileSums, n);
    scan_tilesums_kernel<<<1, 1024, 0, st>>>(tileSums, ntiles, out + n, total64);
    scan_apply_kernel<<<ntiles, kScanThreads, 0, st>>>(in, tileSums, out, copy, n);
    launches += 3;
}

} // namespace mms
