// bin.cuh -- counting sort of particles into the uniform cell grid.
//
//   bin_count_kernel    home voxel (bit-exact vs ParticlesToDensity.cpp:563-568) -> cell id -> histogram
//   (exclusive scan of the histogram: scan.cuh)
//   bin_scatter_kernel  recompute the cell id, claim a slot in the cell's segment, write the 16-byte record
//   cell_order_kernel   canonical order INSIDE each cell (lexicographic on the record's bit pattern), so that
//                       the density kernel's summation order does not depend on which thread claimed which
//                       slot: results are bit-reproducible run to run and across slab decompositions.
// Integer atomics only (histogram / slot claim); no floating-point atomics anywhere.
#pragma once
#include "common.cuh"

namespace mms {

struct Binned {
    float4 p;   // x y z r
    int X, Y, Z; // home voxel (un-wrapped, as the reference computes it)
    int cell;   // linear cell id, -1 = does not contribute to this slab
};

__device__ __forceinline__ int wrapIndex(int X, int s) { // floor-mod, without the integer division in the common case
    return (static_cast<unsigned>(X) < static_cast<unsigned>(s)) ? X : floorMod(X, s);
}

__device__ __forceinline__ bool axisReaches(int X, int f, int lo, int hi, int s, bool cyc) {
    // does the support box [X-f, X+f] (periodic images if cyc) touch voxel range [lo, hi]?
    if (!cyc) return X + f >= lo && X - f <= hi;
    if (lo == 0 && hi == s - 1) return true; // the whole periodic axis
    if (2 * f + 1 >= s) return true;
    const int xw = wrapIndex(X, s);
    const int a = xw - f, b = xw + f;
    return (b >= lo && a <= hi) || (b - s >= lo && a - s <= hi) || (b + s >= lo && a + s <= hi);
}

/** Bump mode: does the interval [p - eps, p + eps], eps = sigma * rad, contain a grid node of this axis?  If it contains none, every
 *  voxel the reference visits for this particle has dis >= eps and receives nothing (ParticlesToDensity.cpp:472-476): the particle is
 *  dropped before the sort.  With voxels much larger than the kernel (the modules' default 16^3 grid) that is almost every particle.
 *  Same bounds and slop as the density kernels' tight support box, so nothing that contributes is ever dropped. */
__device__ __forceinline__ bool supportHasNode(float p, float rad, const Geo& g, int a) {
    const float eps = g.sigma * rad, isd = g.isd[a];
    const float aa = p - g.mn[a];
    const float vlo = (aa - eps) * isd, vhi = (aa + eps) * isd;
    const float slop = fmaxf(fmaxf(fabsf(vlo), fabsf(vhi)), 1.0f) * 4e-6f;
    return __float2int_ru(vlo - slop) <= __float2int_rd(vhi + slop);
}

__device__ __forceinline__ Binned binParticle(const Geo& g, const ListDev& l, unsigned long long j) {
    Binned q;
    q.p = fetchParticle(l, j);
    q.X = homeVoxel(q.p.x, g.mn[0], g.sd[0]);
    q.Y = homeVoxel(q.p.y, g.mn[1], g.sd[1]);
    q.Z = homeVoxel(q.p.z, g.mn[2], g.sd[2]);
    q.cell = -1;
    const float r = q.p.w;
    if (!(r > 0.0f) || !isfinite(r)) return q; // rad == 0 early-out (:523); r < 0 or NaN never contributes
    if (!isfinite(q.p.x) || !isfinite(q.p.y) || !isfinite(q.p.z)) return q;
    int fx, fy, fz;
    if (l.vtype != 2) { // global radius: the host computed the filter sizes once with the same IEEE operations
        fx = l.gf[0], fy = l.gf[1], fz = l.gf[2];
    } else if (g.mode == 0) {
        fx = filterSize(r, g.sd[0]), fy = filterSize(r, g.sd[1]), fz = filterSize(r, g.sd[2]);
    } else {
        const float cut = g.qsAc > 0.0f ? 2.0f * g.qsAc : g.gausslim * g.radscale * r; // reference cells: an atom two cell sizes from a tile can be its candidate
        fx = filterSize(cut, g.sd[0]) + 1, fy = filterSize(cut, g.sd[1]) + 1, fz = filterSize(cut, g.sd[2]) + 1;
    }
    if (g.mode == 0 && !(supportHasNode(q.p.x, q.p.w, g, 0) && supportHasNode(q.p.y, q.p.w, g, 1) && supportHasNode(q.p.z, q.p.w, g, 2))) return q;
    if (!axisReaches(q.X, fx, 0, g.s[0] - 1, g.s[0], g.cyc[0])) return q;
    if (!axisReaches(q.Y, fy, 0, g.s[1] - 1, g.s[1], g.cyc[1])) return q;
    if (!axisReaches(q.Z, fz, g.z0, g.z0 + g.nz - 1, g.s[2], g.cyc[2])) return q;
    const int xw = g.cyc[0] ? wrapIndex(q.X, g.s[0]) : min(max(q.X, 0), g.s[0] - 1);
    const int yw = g.cyc[1] ? wrapIndex(q.Y, g.s[1]) : min(max(q.Y, 0), g.s[1] - 1);
    const int zw = g.cyc[2] ? wrapIndex(q.Z, g.s[2]) : min(max(q.Z, 0), g.s[2] - 1);
    const int lz = cellZLocal(g, zw >> g.cshift);
    if (lz >= g.czCount) return q; // (cannot happen: the host sizes the slab's cell range from the largest filter size)
    q.cell = (xw >> g.cshift) + g.nc[0] * ((yw >> g.cshift) + g.nc[1] * lz);
    return q;
}

__global__ void __launch_bounds__(256) bin_count_kernel(Geo g, ListDev l, unsigned* __restrict__ cellCount,
    DevState* __restrict__ st, int* __restrict__ homeOut, int* __restrict__ cellOut) {
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    float rmax = 0.0f;
    unsigned kept = 0;
    bool outside = false; // a kept particle whose home voxel lies outside the grid on a periodic axis (binned through the wrap)
    const unsigned long long count = listCount(l);
    if (l.countPtr && threadIdx.x == 0 && blockIdx.x == 0 && *l.countPtr > l.count) st->pad[0] = 5u; // halo receive buffer overflowed
    for (unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < count; j += stride) {
        const Binned q = binParticle(g, l, j);
        if (homeOut) {
            int* h = homeOut + 3 * (l.base + j);
            h[0] = q.X, h[1] = q.Y, h[2] = q.Z;
        }
        cellOut[l.base + j] = q.cell; // the scatter pass does not repeat the divisions and the culling tests
        if (q.cell >= 0) {
            atomicAdd(&cellCount[q.cell], 1u);
            rmax = fmaxf(rmax, q.p.w);
            ++kept;
            outside |= (g.cyc[0] && static_cast<unsigned>(q.X) >= static_cast<unsigned>(g.s[0])) ||
                       (g.cyc[1] && static_cast<unsigned>(q.Y) >= static_cast<unsigned>(g.s[1])) ||
                       (g.cyc[2] && static_cast<unsigned>(q.Z) >= static_cast<unsigned>(g.s[2]));
        }
    }
    if (outside) st->pad[1] = 1u; // density_splat3_kernel: its pre-test has to normalise the coordinates first
    rmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(rmax))); // positive floats order like uints
    kept = __reduce_add_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0 && kept) {
        atomicMax(&st->rmaxBits, __float_as_uint(rmax));
        atomicAdd(&st->kept, kept);
    }
}

/** aux: 0 floats (plain), 1 float (aggregator 1: intensity) or 4 floats (QuickSurf colour; aggregator 2: direction) per record.
 *  cellIn = the cell ids bin_count_kernel left per particle (-1: culled).  (A per-slot cell id for cell_order_kernel was measured and
 *  rejected: the extra scattered 4-byte store costs this kernel more -- 95 -> 135 us on C2 -- than the divisions it saves there.) */
__global__ void __launch_bounds__(256) bin_scatter_kernel(Geo g, ListDev l, unsigned* __restrict__ cursor,
    float4* __restrict__ recs, float* __restrict__ aux, int auxN, const int* __restrict__ cellIn) {
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    const unsigned long long count = listCount(l);
    for (unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < count; j += stride) {
        const int cell = cellIn[l.base + j];
        if (cell < 0) continue;
        const unsigned slot = atomicAdd(&cursor[cell], 1u);
        recs[slot] = fetchParticle(l, j);
        if (auxN == 1) {
            aux[slot] = fetchColourRaw(l, j).x; // iAcc->Get_f (ParticlesToDensity.cpp:483,515)
        } else if (auxN == 4) {
            reinterpret_cast<float4*>(aux)[slot] = g.mode == 0 ? fetchDirection(l, j) : quicksurfColour(l, fetchColourRaw(l, j));
        }
    }
}

/** Lexicographic order on the records' bit patterns (x, y, z, r as unsigned words, then the aux words) = the order of the 128-bit
 *  concatenation: two 64-bit compares.  -1: a before b, 0: bit-identical, +1: b before a. */
__device__ __forceinline__ int recCompare(const float4& a, const float* aa, const float4& b, const float* ab, int auxN) {
    const unsigned long long a0 = static_cast<unsigned long long>(__float_as_uint(a.x)) << 32 | __float_as_uint(a.y);
    const unsigned long long a1 = static_cast<unsigned long long>(__float_as_uint(a.z)) << 32 | __float_as_uint(a.w);
    const unsigned long long b0 = static_cast<unsigned long long>(__float_as_uint(b.x)) << 32 | __float_as_uint(b.y);
    const unsigned long long b1 = static_cast<unsigned long long>(__float_as_uint(b.z)) << 32 | __float_as_uint(b.w);
    if (a0 != b0) return a0 < b0 ? -1 : 1;
    if (a1 != b1) return a1 < b1 ? -1 : 1;
    for (int i = 0; i < auxN; ++i) {
        const unsigned x = __float_as_uint(aa[i]), y = __float_as_uint(ab[i]);
        if (x != y) return x < y ? -1 : 1;
    }
    return 0;
}

__device__ __forceinline__ bool recLess(const float4& a, const float* aa, const float4& b, const float* ab, int auxN) {
    const unsigned ka[4] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w)};
    const unsigned kb[4] = {__float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (ka[i] != kb[i]) return ka[i] < kb[i];
    for (int i = 0; i < auxN; ++i) {
        const unsigned x = __float_as_uint(aa[i]), y = __float_as_uint(ab[i]);
        if (x != y) return x < y;
    }
    return false;
}

constexpr unsigned kBigCell = 1024;  // cells with more records than this are sorted by cell_sort_big_kernel
constexpr int kBigThreads = 512;

/**
 * One thread per sorted slot: rank of my record among the records of my cell (ties: slot order, the tied
 * records are bit-identical so their order cannot change any sum), written to the second buffer.
 */
__global__ void __launch_bounds__(256) cell_order_kernel(Geo g, const unsigned* __restrict__ cellStart,
    const float4* __restrict__ in, const float* __restrict__ auxIn, float4* __restrict__ out, float* __restrict__ auxOut,
    int auxN, const DevState* __restrict__ st, unsigned* __restrict__ bigCells, unsigned* __restrict__ nBig) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->kept) return;
    const float4 me = in[i];
    // my cell (records only exist for contributing particles, so the clamped/wrapped home is in range)
    const int X = homeVoxel(me.x, g.mn[0], g.sd[0]), Y = homeVoxel(me.y, g.mn[1], g.sd[1]), Z = homeVoxel(me.z, g.mn[2], g.sd[2]);
    const int xw = g.cyc[0] ? wrapIndex(X, g.s[0]) : min(max(X, 0), g.s[0] - 1);
    const int yw = g.cyc[1] ? wrapIndex(Y, g.s[1]) : min(max(Y, 0), g.s[1] - 1);
    const int zw = g.cyc[2] ? wrapIndex(Z, g.s[2]) : min(max(Z, 0), g.s[2] - 1);
    const unsigned cell = static_cast<unsigned>((xw >> g.cshift) + g.nc[0] * ((yw >> g.cshift) + g.nc[1] * cellZLocal(g, zw >> g.cshift)));
    const unsigned b = cellStart[cell], e = cellStart[cell + 1];
    if (e - b > kBigCell) { // crowded cell: ranking by all pairs would be quadratic; cell_sort_big_kernel sorts it
        if (i == b) bigCells[atomicAdd(nBig, 1u)] = cell;
        return;
    }
    float myAux[4] = {0, 0, 0, 0};
    for (int k = 0; k < auxN; ++k) myAux[k] = auxIn[static_cast<size_t>(i) * auxN + k];
    unsigned rank = 0;
    for (unsigned k = b; k < e; ++k) {
        const float4 o = in[k];
        float oa[4] = {0, 0, 0, 0};
        for (int q = 0; q < auxN; ++q) oa[q] = auxIn[static_cast<size_t>(k) * auxN + q];
        const int c = recCompare(o, oa, me, myAux, auxN);
        rank += (c < 0 || (c == 0 && k < i)) ? 1u : 0u; // (k == i compares equal and is not before itself)
    }
    out[b + rank] = me;
    for (int k = 0; k < auxN; ++k) auxOut[static_cast<size_t>(b + rank) * auxN + k] = myAux[k];
}

/**
 * Canonical order inside CROWDED cells (coarse grids with wide kernels, clustered data): a block takes a cell from the list
 * cell_order_kernel left behind, sorts chunks of kBigThreads records by all-pairs ranking in shared memory and merges them bottom-up,
 * ping-ponging between the two record buffers; every element finds its place in the merged run by a binary search in the other run
 * (ties: the left run first; tied records are bit-identical).  O(k log^2 k) per cell instead of O(k^2).  The launch is unconditional and
 * tiny; with no crowded cell every block leaves at once.
 */
__global__ void __launch_bounds__(kBigThreads) cell_sort_big_kernel(const unsigned* __restrict__ cellStart, float4* recA, float* auxA,
    float4* recB, float* auxB, int auxN, const unsigned* __restrict__ bigCells, const unsigned* __restrict__ nBig, unsigned* __restrict__ next) {
    __shared__ float4 sRec[kBigThreads];
    __shared__ float sAux[kBigThreads][4];
    __shared__ unsigned sCell;
    const int tid = threadIdx.x;
    const unsigned n = *nBig;
    for (;;) {
        __syncthreads();
        if (tid == 0) sCell = atomicAdd(next, 1u);
        __syncthreads();
        if (sCell >= n) return;
        const unsigned cell = bigCells[sCell];
        const unsigned b = cellStart[cell], k = cellStart[cell + 1] - b;
        // ---- sorted chunks: A -> B ----------------------------------------------------------------------------------
        for (unsigned c0 = 0; c0 < k; c0 += kBigThreads) {
            const unsigned m = min(static_cast<unsigned>(kBigThreads), k - c0);
            __syncthreads();
            if (static_cast<unsigned>(tid) < m) {
                sRec[tid] = recA[b + c0 + tid];
                for (int q = 0; q < auxN; ++q) sAux[tid][q] = auxA[static_cast<size_t>(b + c0 + tid) * auxN + q];
            }
            __syncthreads();
            if (static_cast<unsigned>(tid) < m) {
                const float4 me = sRec[tid];
                float ma[4] = {0, 0, 0, 0};
                for (int q = 0; q < auxN; ++q) ma[q] = sAux[tid][q];
                unsigned rank = 0;
                for (unsigned o = 0; o < m; ++o) {
                    if (o == static_cast<unsigned>(tid)) continue;
                    const float4 ot = sRec[o];
                    if (recLess(ot, sAux[o], me, ma, auxN) || (!recLess(me, ma, ot, sAux[o], auxN) && o < static_cast<unsigned>(tid))) ++rank;
                }
                recB[b + c0 + rank] = me;
                for (int q = 0; q < auxN; ++q) auxB[static_cast<size_t>(b + c0 + rank) * auxN + q] = ma[q];
            }
        }
        // ---- bottom-up merge, B -> A -> B ... ----------------------------------------------------------------------------
        float4* src = recB;
        float* srcAux = auxB;
        float4* dst = recA;
        float* dstAux = auxA;
        for (unsigned L = kBigThreads; L < k; L <<= 1) {
            __syncthreads(); // the previous pass (global writes of this block) is complete and visible
            for (unsigned i = tid; i < k; i += kBigThreads) {
                const unsigned run = i / L, pairBase = (run >> 1) * 2 * L;
                const bool left = (run & 1u) == 0;
                // the other run of my pair
                const unsigned oBeg = left ? min(pairBase + L, k) : pairBase, oEnd = left ? min(pairBase + 2 * L, k) : pairBase + L;
                const float4 me = src[b + i];
                float ma[4] = {0, 0, 0, 0};
                for (int q = 0; q < auxN; ++q) ma[q] = srcAux[static_cast<size_t>(b + i) * auxN + q];
                // left element: number of right elements strictly less than me; right element: number of left elements <= me
                unsigned lo = oBeg, hi = oEnd;
                while (lo < hi) {
                    const unsigned mid = (lo + hi) >> 1;
                    const float4 ot = src[b + mid];
                    float oa[4] = {0, 0, 0, 0};
                    for (int q = 0; q < auxN; ++q) oa[q] = srcAux[static_cast<size_t>(b + mid) * auxN + q];
                    const bool before = left ? recLess(ot, oa, me, ma, auxN) : !recLess(me, ma, ot, oa, auxN);
                    if (before) lo = mid + 1; else hi = mid;
                }
                const unsigned pos = pairBase + (i - (left ? pairBase : pairBase + L)) + (lo - oBeg);
                dst[b + pos] = me;
                for (int q = 0; q < auxN; ++q) dstAux[static_cast<size_t>(b + pos) * auxN + q] = ma[q];
            }
            float4* t = src; src = dst; dst = t;
            float* ta = srcAux; srcAux = dstAux; dstAux = ta;
        }
        __syncthreads();
        if (src != recB) { // an odd number of merge passes left the result in A
            for (unsigned i = tid; i < k; i += kBigThreads) {
                recB[b + i] = recA[b + i];
                for (int q = 0; q < auxN; ++q) auxB[static_cast<size_t>(b + i) * auxN + q] = auxA[static_cast<size_t>(b + i) * auxN + q];
            }
        }
    }
}

} // namespace mms
