// density.cuh -- density volume from cell-sorted particles, no floating-point atomics, fixed summation order.
//
// density_tile_kernel ("owner-warp splat"): a block owns a 32x16x16-voxel tile kept in shared memory, each
// of its 8 warps owns a private 16x8x8 sub-tile.  The block streams the particles of all cells whose support
// can reach the tile (cell rows are contiguous in the sorted array, so this is a handful of coalesced
// segments) through shared memory in chunks; every warp culls the chunk against its sub-tile (one candidate
// per lane + ballot) and then walks the survivors IN ORDER, the 32 lanes covering the voxels of the
// particle's tight support box.  A voxel is only ever touched by its owner warp and candidates are visited
// in (cell z, cell y, cell x, canonical in-cell) order, so every voxel's sum has one fixed order -- the
// same for every launch, tile decomposition and z-slab decomposition.
//
// Arithmetic of the P2D-bump mode follows ParticlesToDensity.cpp:577-620 and :472-476 operation by
// operation with round-to-nearest intrinsics (no FMA contraction):
//   pos = float(h)*sliceDist + minOS;  d = |pos - p|;  dis = sqrt(dx*dx + dy*dy + dz*dz)
//   dis >= sigma*rad ? 0 : exp(-1 / (1 - ((1/eps)*dis)^2))
// h is the UN-wrapped voxel index, so periodic images get the true distance (:605-613).
#pragma once
#include "common.cuh"

namespace mms {

constexpr int WTX = 16, WTY = 8, WTZ = 8;       // warp sub-tile
constexpr int BWX = 2, BWY = 2, BWZ = 2;        // warps per block tile
constexpr int BTX = WTX * BWX, BTY = WTY * BWY, BTZ = WTZ * BWZ;
constexpr int DT_WARPS = BWX * BWY * BWZ;
constexpr int DT_THREADS = DT_WARPS * 32;
constexpr int WT_SY = 18, WT_SZ = 171;          // padded strides: 3x3x3 lane pattern is bank-conflict free
constexpr int WT_FLOATS = WT_SZ * WTZ;
constexpr int DT_CHUNK = DT_THREADS;            // candidates staged per round (one per thread)
constexpr int DT_MAXSEG = 512;                  // cell-row segments per batch
constexpr int DT_MAXAXIS = 64;                  // cells per axis in a tile's neighbourhood list

/** A staged candidate, pre-digested once per block. 16 words. */
struct Cand {
    float x, y, z, eps;       // position, kernel radius (P2D: sigma*rad, QS: cut-off)
    float k0, weight;         // P2D: 1/eps; QS: w_p = -log2(e)/(2 (r*radscale)^2) | aggregator-1 intensity
    int lox, loy;             // tight support box, lower corner, un-wrapped but shifted into [-s, 2s)
    int loz;
    unsigned dims;            // bx | by<<10 | bz<<20   (0 = empty)
    int offx, offy, offz;     // true un-wrapped voxel index = shifted index + off  (non-zero only for homes outside [0,s))
    float cr, cg, cb;         // QS colour
};

struct TileShared {
    float tile[DT_WARPS][WT_FLOATS];
    Cand cand[DT_CHUNK];
    unsigned segBegin[DT_MAXSEG];
    unsigned segPrefix[DT_MAXSEG + 1];
    int axisCells[3][DT_MAXAXIS];
    int axisCount[3];
    unsigned scanTmp[33];
};

/** Ordered, duplicate-free list of the cells along one axis whose particles can reach voxels [t0, t1]. */
__device__ inline int buildAxisCells(int t0, int t1, int reach, int s, bool cyc, int sh, int nc, int* out) {
    int a = t0 - reach, b = t1 + reach;
    if (!cyc) {
        a = max(a, 0), b = min(b, s - 1);
        int n = 0;
        for (int c = a >> sh; c <= (b >> sh) && n < DT_MAXAXIS; ++c) out[n++] = c;
        return (b >> sh) - (a >> sh) + 1 > DT_MAXAXIS ? -1 : n;
    }
    if (b - a + 1 >= s) { // whole axis, each cell once
        if (nc > DT_MAXAXIS) return -1;
        for (int c = 0; c < nc; ++c) out[c] = c;
        return nc;
    }
    // un-wrapped order: high-end image first, then the main piece, then the low-end image
    int n = 0;
    int lastAdded = -1; // cells are added in pieces; keep every cell once
    auto addRange = [&](int v0, int v1) {
        for (int c = v0 >> sh; c <= (v1 >> sh); ++c) {
            bool dup = false;
            for (int k = 0; k < n; ++k) dup |= (out[k] == c);
            if (!dup) {
                if (n >= DT_MAXAXIS) { n = DT_MAXAXIS + 1; return; }
                out[n++] = c;
            }
        }
        (void)lastAdded;
    };
    if (a < 0) addRange(a + s, s - 1);
    if (n <= DT_MAXAXIS) addRange(max(a, 0), min(b, s - 1));
    if (n <= DT_MAXAXIS && b >= s) addRange(0, b - s);
    return n > DT_MAXAXIS ? -1 : n;
}

template<int MODE>
__device__ __forceinline__ void digest(const Geo& g, const float4 p, float auxI, const float4 auxC, Cand& c) {
    c.x = p.x, c.y = p.y, c.z = p.z;
    c.weight = auxI;
    c.cr = auxC.x, c.cg = auxC.y, c.cb = auxC.z;
    const int H[3] = {homeVoxel(p.x, g.mn[0], g.sd[0]), homeVoxel(p.y, g.mn[1], g.sd[1]), homeVoxel(p.z, g.mn[2], g.sd[2])};
    const float pos[3] = {p.x, p.y, p.z};
    float eps;
    if (MODE == 0) {
        eps = __fmul_rn(g.sigma, p.w);       // sigma * rad (:526)
        c.k0 = __fdiv_rn(1.0f, eps);         // (1.0f / epsilon) (:475)
    } else {
        const float sr = __fmul_rn(p.w, g.radscale);
        eps = __fmul_rn(g.gausslim, sr);
        c.k0 = __fdiv_rn(-1.4426950408889634f, __fmul_rn(__fmul_rn(2.0f, sr), sr));
    }
    c.eps = eps;
    int lo[3], hi[3], off[3];
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // tight bounds: a voxel further than eps along one axis cannot be inside the kernel support.
        // The slop only has to beat fp32 rounding of this bound; contributions within 0.48% of the
        // support radius are exactly 0.0f anyway (exp(-x) underflows for x > 104).
        const float vlo = __fdiv_rn(__fsub_rn(__fsub_rn(pos[a], eps), g.mn[a]), g.sd[a]);
        const float vhi = __fdiv_rn(__fsub_rn(__fadd_rn(pos[a], eps), g.mn[a]), g.sd[a]);
        const float slop = fmaxf(fmaxf(fabsf(vlo), fabsf(vhi)), 1.0f) * 2e-6f;
        int l = __float2int_ru(vlo - slop), h = __float2int_rd(vhi + slop);
        if (MODE == 0) { // the reference's support box around the home voxel (:573-579)
            const int f = filterSize(p.w, g.sd[a]);
            l = max(l, H[a] - f), h = min(h, H[a] + f);
        }
        off[a] = 0;
        if (g.cyc[a]) {
            if (h - l + 1 > g.s[a]) h = l + g.s[a] - 1; // cannot happen for f <= (s-1)/2; keeps images unique
            const int k = (H[a] >= 0 && H[a] < g.s[a]) ? 0 : (H[a] - floorMod(H[a], g.s[a]));
            off[a] = k, l -= k, h -= k;
        } else {
            l = max(l, 0), h = min(h, g.s[a] - 1);
        }
        if (h < l) empty = true;
        lo[a] = l, hi[a] = h;
    }
    c.lox = lo[0], c.loy = lo[1], c.loz = lo[2];
    c.offx = off[0], c.offy = off[1], c.offz = off[2];
    const int bx = hi[0] - lo[0] + 1, by = hi[1] - lo[1] + 1, bz = hi[2] - lo[2] + 1;
    c.dims = (empty || bx > 1023 || by > 1023 || bz > 1023) ? 0u : (unsigned)bx | ((unsigned)by << 10) | ((unsigned)bz << 20);
    // boxes wider than 1023 voxels per axis are rejected on the host (MMS_ERR_UNSUPPORTED)
}

__device__ __forceinline__ bool axisHits(int lo, int n, int w0, int wn, int s, bool cyc) {
    // does [lo, lo+n) (periodic if cyc; lo in [-s, 2s)) touch [w0, w0+wn)?
    const int hi = lo + n - 1, w1 = w0 + wn - 1;
    bool r = hi >= w0 && lo <= w1;
    if (cyc) r = r || (hi - s >= w0 && lo - s <= w1) || (hi + s >= w0 && lo + s <= w1);
    return r;
}

template<int MODE, bool COLOUR>
__global__ void __launch_bounds__(DT_THREADS) density_tile_kernel(Geo g, DevState* st,
    const float4* __restrict__ recs, const float* __restrict__ aux, int auxN, const unsigned* __restrict__ cellStart,
    float* __restrict__ vol, float* __restrict__ rgb) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    TileShared& sh = *reinterpret_cast<TileShared*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bt0[3] = {(int)blockIdx.x * BTX, (int)blockIdx.y * BTY, g.z0 + (int)blockIdx.z * BTZ};
    const int w0x = bt0[0] + (warp % BWX) * WTX, w0y = bt0[1] + ((warp / BWX) % BWY) * WTY, w0z = bt0[2] + (warp / (BWX * BWY)) * WTZ;
    float* mine = sh.tile[warp];
    for (int i = lane; i < WT_FLOATS; i += 32) mine[i] = 0.0f;

    // neighbourhood of the block tile in cell space
    if (tid < 3) {
        const float rmax = __uint_as_float(st->rmaxBits);
        int reach;
        if (MODE == 0) {
            const int f = filterSize(rmax, g.sd[tid]);
            const int t = __float2int_ru(__fdiv_rn(__fmul_rn(g.sigma, rmax), g.sd[tid])) + 1;
            reach = min(f, t);
        } else {
            reach = __float2int_ru(__fdiv_rn(g.gausslim * g.radscale * rmax, g.sd[tid])) + 2;
        }
        reach = max(reach, 0);
        const int hiV = (tid == 2) ? min(bt0[2] + BTZ, g.z0 + g.nz) - 1 : min(bt0[tid] + (tid == 0 ? BTX : BTY), g.s[tid]) - 1;
        sh.axisCount[tid] = buildAxisCells(bt0[tid], hiV, reach, g.s[tid], g.cyc[tid] != 0, g.cshift, g.nc[tid], sh.axisCells[tid]);
    }
    __syncthreads();
    const int ncx = sh.axisCount[0], ncy = sh.axisCount[1], ncz = sh.axisCount[2];
    if ((ncx < 0 || ncy < 0 || ncz < 0) && tid == 0) st->pad[0] = 1u; // neighbourhood list overflow -> host reports it
    // x cells form at most a few runs of consecutive cell ids; a (y,z) row contributes one segment per run.
    // Rows are enumerated z-major, then y, then x-run: the canonical candidate order.
    // (axisCount < 0 -> neighbourhood too large for the list: host guards against it.)
    int nruns = 0;
    int runStart[4], runEnd[4];
    for (int k = 0; k < ncx && nruns < 4; ++k) {
        const int c = sh.axisCells[0][k];
        if (nruns > 0 && c == runEnd[nruns - 1] + 1) runEnd[nruns - 1] = c;
        else { runStart[nruns] = c; runEnd[nruns] = c; ++nruns; }
    }
    const int nrows = (ncx > 0 && ncy > 0 && ncz > 0) ? ncy * ncz : 0;
    const int nsegTotal = nrows * nruns;

    for (int segBase = 0; segBase < nsegTotal; segBase += DT_MAXSEG) {
        const int nseg = min(DT_MAXSEG, nsegTotal - segBase);
        __syncthreads();
        // segment table + prefix sums
        unsigned carry = 0;
        for (int b0 = 0; b0 < nseg; b0 += DT_THREADS) {
            const int sI = b0 + tid;
            unsigned len = 0, beg = 0;
            if (sI < nseg) {
                const int gs = segBase + sI;
                const int row = gs / nruns, run = gs - row * nruns;
                const int cz = sh.axisCells[2][row / ncy], cy = sh.axisCells[1][row % ncy];
                const size_t rowBase = (static_cast<size_t>(cz) * g.nc[1] + cy) * g.nc[0];
                beg = cellStart[rowBase + runStart[run]];
                len = cellStart[rowBase + runEnd[run] + 1] - beg;
                sh.segBegin[sI] = beg;
            }
            unsigned total;
            const unsigned ex = blockExclusiveScan(len, &total, sh.scanTmp);
            if (sI < nseg) sh.segPrefix[sI] = carry + ex;
            carry += total;
        }
        if (tid == 0) sh.segPrefix[nseg] = carry;
        __syncthreads();
        const unsigned ncand = sh.segPrefix[nseg];

        for (unsigned chunk = 0; chunk < ncand; chunk += DT_CHUNK) {
            const unsigned nin = min((unsigned)DT_CHUNK, ncand - chunk);
            __syncthreads(); // previous chunk fully consumed
            if ((unsigned)tid < nin) {
                const unsigned pos = chunk + tid;
                int lo = 0, hi = nseg; // last segment with prefix <= pos
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sh.segPrefix[mid] <= pos) lo = mid; else hi = mid;
                }
                const unsigned idx = sh.segBegin[lo] + (pos - sh.segPrefix[lo]);
                const float4 p = recs[idx];
                float aI = 1.0f;
                float4 aC = make_float4(1.f, 1.f, 1.f, 1.f);
                if (auxN == 1) aI = aux[idx];
                else if (auxN == 4) aC = reinterpret_cast<const float4*>(aux)[idx];
                digest<MODE>(g, p, aI, aC, sh.cand[tid]);
            }
            __syncthreads();

            for (unsigned base = 0; base < nin; base += 32) {
                const unsigned ci = base + lane;
                bool hit = false;
                if (ci < nin) {
                    const Cand& c = sh.cand[ci];
                    const unsigned d = c.dims;
                    hit = d != 0 && axisHits(c.lox, d & 1023, w0x, WTX, g.s[0], g.cyc[0] != 0) &&
                          axisHits(c.loy, (d >> 10) & 1023, w0y, WTY, g.s[1], g.cyc[1] != 0) &&
                          axisHits(c.loz, (d >> 20) & 1023, w0z, WTZ, g.s[2], g.cyc[2] != 0);
                }
                unsigned mask = __ballot_sync(0xffffffffu, hit);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const Cand& c = sh.cand[base + j];
                    const int bx = c.dims & 1023, by = (c.dims >> 10) & 1023, bz = (c.dims >> 20) & 1023;
                    const int bxy = bx * by, nvox = bxy * bz;
                    const float rbx = __frcp_rn((float)bx), rbxy = __frcp_rn((float)bxy);
                    const bool smallBox = nvox <= 4096;
                    for (int i = lane; i < nvox; i += 32) {
                        int iz, iy, ix;
                        if (smallBox) { // exact for these ranges: (i+0.5)/n is never within rounding of an integer
                            iz = __float2int_rz(((float)i + 0.5f) * rbxy);
                            const int rem = i - iz * bxy;
                            iy = __float2int_rz(((float)rem + 0.5f) * rbx);
                            ix = rem - iy * bx;
                        } else {
                            iz = i / bxy;
                            const int rem = i - iz * bxy;
                            iy = rem / bx;
                            ix = rem - iy * bx;
                        }
                        int hx = c.lox + ix, hy = c.loy + iy, hz = c.loz + iz; // shifted un-wrapped index
                        int tx = hx, ty = hy, tz = hz;
                        if (g.cyc[0]) tx = hx < 0 ? hx + g.s[0] : (hx >= g.s[0] ? hx - g.s[0] : hx);
                        if (g.cyc[1]) ty = hy < 0 ? hy + g.s[1] : (hy >= g.s[1] ? hy - g.s[1] : hy);
                        if (g.cyc[2]) tz = hz < 0 ? hz + g.s[2] : (hz >= g.s[2] ? hz - g.s[2] : hz);
                        const unsigned lx = tx - w0x, ly = ty - w0y, lz = tz - w0z;
                        if (lx >= (unsigned)WTX || ly >= (unsigned)WTY || lz >= (unsigned)WTZ) continue;
                        hx += c.offx, hy += c.offy, hz += c.offz; // the reference's hx/hy/hz
                        const float px = __fadd_rn(__fmul_rn((float)hx, g.sd[0]), g.mn[0]);
                        const float py = __fadd_rn(__fmul_rn((float)hy, g.sd[1]), g.mn[1]);
                        const float pz = __fadd_rn(__fmul_rn((float)hz, g.sd[2]), g.mn[2]);
                        const float dx = fabsf(__fsub_rn(px, c.x)), dy = fabsf(__fsub_rn(py, c.y)), dz = fabsf(__fsub_rn(pz, c.z));
                        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                        float* cell = mine + lx + ly * WT_SY + lz * WT_SZ;
                        if (MODE == 0) {
                            const float dis = __fsqrt_rn(d2);
                            if (dis >= c.eps) continue;
                            const float q = __fmul_rn(c.k0, dis);
                            const float w = expf(__fdiv_rn(-1.0f, __fsub_rn(1.0f, __fmul_rn(q, q))));
                            *cell = __fadd_rn(*cell, g.agg == 1 ? __fmul_rn(w, c.weight) : w);
                        } else {
                            if (!(d2 < __fmul_rn(c.eps, c.eps))) continue;
                            const float w = exp2f(__fmul_rn(d2, c.k0));
                            *cell = __fadd_rn(*cell, w);
                            // colour volume handled by the gather kernel variant (QS mode): see density_gather_kernel
                        }
                    }
                    __syncwarp(); // the next candidate may touch voxels this one wrote from other lanes
                }
            }
        }
    }
    __syncthreads();

    // write-out: rows of 32 consecutive voxels (128 B), plus the block's min/max
    float vmin = INFINITY, vmax = -INFINITY;
    const int zEnd = g.z0 + g.nz;
    for (int r = warp; r < BTY * BTZ; r += DT_WARPS) {
        const int ly = r % BTY, lz = r / BTY;
        const int x = bt0[0] + lane, y = bt0[1] + ly, z = bt0[2] + lz;
        if (x < g.s[0] && y < g.s[1] && z < zEnd) {
            const int w = (lane / WTX) + BWX * ((ly / WTY) + BWY * (lz / WTZ));
            const float v = sh.tile[w][(lane % WTX) + (ly % WTY) * WT_SY + (lz % WTZ) * WT_SZ];
            vol[x + static_cast<size_t>(g.s[0]) * (y + static_cast<size_t>(g.s[1]) * (z - g.z0))] = v;
            vmin = fminf(vmin, v), vmax = fmaxf(vmax, v);
        }
    }
    unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if (lane == 0 && kmin <= kmax) {
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

__global__ void __launch_bounds__(256) normalize_kernel(float* __restrict__ vol, size_t n, float mn, float rcp) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        vol[i] = __fmul_rn(__fsub_rn(vol[i], mn), rcp);
}

} // namespace mms
