// density.cuh -- density volume from cell-sorted particles: no floating-point atomics, one fixed summation order.
//
// density_splat_kernel ("coloured cell splat").  A block owns a 32x16x16-voxel tile in shared memory.  The
// particles that can reach the tile live in the tile's cells plus one ring of neighbour cells (the cell edge C
// is chosen on the host so that every support box reaches at most C/2 voxels beyond its home cell).  Cells are
// coloured 2x2x2 by the parity of their GLOBAL cell coordinates: two different cells of one colour have
// disjoint footprints, so during one colour phase every warp can splat a different cell with plain
// load-add-store on shared memory and no two warps ever touch the same voxel.  Inside a cell the particles
// are walked in their canonical order by ONE warp, the 32 lanes covering the voxels of the particle's tight
// support box.  Every voxel therefore receives its contributions in the order
//        (colour phase, canonical in-cell order)
// which depends on nothing but the data: bit-identical run to run and for every tile / z-slab decomposition.
// Cells of one phase are handed to warps dynamically (integer counter in shared memory) -- the assignment
// does not influence any sum.
//
// Arithmetic of the P2D-bump mode follows ParticlesToDensity.cpp:577-620 and :472-476 with individually
// rounded operations:   pos = float(h)*sliceDist + minOS;  d = |pos - p|;  dis = sqrt(dx*dx + dy*dy + dz*dz);
//                       q = (1/eps)*dis;  den = 1 - q*q        -- bit-identical to the reference up to here --
//                       w = exp(-1/den)  evaluated as ex2(-log2(e) * rcp(den)) on the SFU (|rel err| < 3e-6).
// h is the UN-wrapped voxel index, so periodic images get the true distance (:605-613).
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include <type_traits>

namespace mms {

constexpr int CT_X = 32, CT_Y = 16, CT_Z = 16;       // block tile (voxels)
constexpr int CT_SY = 35, CT_SZ = 585;               // padded strides (= 3 and 9 mod 32): a 3x3x3 lane pattern hits 27 banks
constexpr int CT_FLOATS = CT_SZ * CT_Z;
constexpr int CT_WARPS = 8;
constexpr int CT_THREADS = CT_WARPS * 32;
constexpr int CT_MAXAXIS = 16;                       // cells per axis in a tile's neighbourhood (<= 32/4 + 3)
constexpr int CT_MAXCELLS = 576;                     // neighbourhood cells (C = 4: at most 11 x 7 x 7 = 539)

/** A digested particle (per-warp staging, read back with broadcast LDS.128). 12 words; the shared-memory budget
 *  (tile 36.6 KB + 12 KB of these + tables) is trimmed to 56 KB so that FOUR blocks fit one SM. */
struct Dig {
    float x, y, z, eps;
    float k0, weight, f0x, f0y; // k0: P2D 1/eps, QS w_p;  f0*: float(true un-wrapped index of the box's first voxel)
    float f0z;
    int base;                   // shared-memory offset of the box's first voxel
    unsigned dims;              // bx | by<<8 | bz<<16  (0 = nothing to do)
    unsigned mask27;            // fast path (all dims <= 3): valid lanes of the fixed 3x3x3 lane pattern
};

struct SplatShared {
    float tile[CT_FLOATS];
    Dig dig[CT_WARPS][32];
    unsigned cellB[CT_MAXCELLS];
    unsigned cellN[CT_MAXCELLS];       // particles in the cell
    int axisCells[3][CT_MAXAXIS];
    int axisCount[3];
    int colList[3][4][CT_MAXAXIS]; // per axis, per colour: positions in axisCells
    int colCount[3][4];
    int ncol[3];
    int phaseCount[64];            // non-empty cells per colour phase
    int phaseStart[65];
    unsigned short phaseCells[CT_MAXCELLS];
    int tl0[3], tl1[3];            // tile voxel range (inclusive), clipped to the grid / slab
    int perVoxelWrap[3];           // degenerate cyclic axis: wrap every voxel instead of choosing one image per particle
};
static_assert(sizeof(SplatShared) <= 57088, "density_splat_kernel must fit four blocks per SM");

/** Ordered, duplicate-free list of the cells along one axis whose particles can reach voxels [t0, t1]. */
__host__ __device__ inline int buildAxisCells(int t0, int t1, int reach, int s, bool cyc, int sh, int nc, int* out, int cap) {
    int a = t0 - reach, b = t1 + reach;
    int n = 0;
    if (!cyc) {
        a = max(a, 0), b = min(b, s - 1);
        for (int c = a >> sh; c <= (b >> sh); ++c) {
            if (n >= cap) return -1;
            out[n++] = c;
        }
        return n;
    }
    if (b - a + 1 >= s) { // whole axis, each cell once
        if (nc > cap) return -1;
        for (int c = 0; c < nc; ++c) out[c] = c;
        return nc;
    }
    bool overflow = false;
    auto addRange = [&](int v0, int v1) {
        for (int c = v0 >> sh; c <= (v1 >> sh); ++c) {
            bool dup = false;
            for (int k = 0; k < n; ++k) dup |= (out[k] == c);
            if (dup) continue;
            if (n >= cap) { overflow = true; return; }
            out[n++] = c;
        }
    };
    if (a < 0) addRange(a + s, s - 1);
    addRange(max(a, 0), min(b, s - 1));
    if (b >= s) addRange(0, b - s);
    return overflow ? -1 : n;
}

/** Colour of cell c along one axis: parity, except that an irregular periodic axis (odd cell count or a partial
 *  last cell) gives its last two cells private colours so that same-coloured cells stay >= one full cell apart
 *  across the wrap. */
__device__ __forceinline__ int axisColour(int c, int nc, bool irregular) {
    if (nc < 3) return min(c, 3);
    if (!irregular) return c & 1;
    if (c == nc - 2) return 2;
    if (c == nc - 1) return 3;
    return c & 1;
}

__device__ __forceinline__ float ex2Approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcpApprox(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

/** Contribution of one voxel; returns false if the voxel is outside the kernel support. */
template<int MODE, bool PRECISE = false>
__device__ __forceinline__ bool kernelValue(float d2, float eps, float k0, float& w, float lim2 = -1.0f) {
    if (MODE == 0) {
        const float dis = __fsqrt_rn(d2);
        if (dis >= eps) return false;
        const float q = __fmul_rn(k0, dis);
        const float den = __fsub_rn(1.0f, __fmul_rn(q, q));
        // PRECISE (aggregator 2): the vector field is a QUOTIENT of weighted sums, so a voxel whose only contributors sit in the far
        // tail of the bump still gets v = d -- as long as the weight does not flush to zero and keeps its relative accuracy there:
        // IEEE division and expf (2 ulp, subnormal results kept) instead of the SFU pair (rcp.approx in the exponent, ftz)
        w = PRECISE ? expf(__fdiv_rn(-1.0f, den)) : ex2Approx(-1.4426950408889634f * rcpApprox(den));
        return true;
    } else {
        if (!(d2 < (lim2 >= 0.0f ? lim2 : __fmul_rn(eps, eps)))) return false;
        w = ex2Approx(__fmul_rn(d2, k0));
        return true;
    }
}

/** The general splat kernel: any support box the cell size allows, aggregators 0 and 1, periodic axes of any length (per-voxel wrap).
 *  Supports of at most 3x3x3 voxels with aggregator 0 (C1, C2, C4) take density_splat3_kernel instead. */
template<int MODE>
__global__ void __launch_bounds__(CT_THREADS, 4) density_splat_kernel(Geo g, DevState* st, const float4* __restrict__ recs,
    const float* __restrict__ aux, int auxN, const unsigned* __restrict__ cellStart, float* __restrict__ vol, int reach) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    SplatShared& sh = *reinterpret_cast<SplatShared*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C = 1 << g.cshift;
    for (int i = tid; i < CT_FLOATS; i += CT_THREADS) sh.tile[i] = 0.0f;
    if (tid < 64) sh.phaseCount[tid] = 0;
    if (tid < 3) {
        const int a = tid;
        const int t0 = (a == 0 ? (int)blockIdx.x * CT_X : a == 1 ? (int)blockIdx.y * CT_Y : g.z0 + (int)blockIdx.z * CT_Z);
        const int lim = (a == 2) ? g.z0 + g.nz : g.s[a];
        const int t1 = min(t0 + (a == 0 ? CT_X : a == 1 ? CT_Y : CT_Z), lim) - 1;
        sh.tl0[a] = t0, sh.tl1[a] = t1;
        const bool cyc = g.cyc[a] != 0;
        // one periodic image per particle is enough unless the axis is so short that two images can hit the tile
        sh.perVoxelWrap[a] = (cyc && g.s[a] < (t1 - t0 + 1) + 2 * reach + 2) ? 1 : 0;
        const int n = buildAxisCells(t0, t1, reach, g.s[a], cyc, g.cshift, g.nc[a], sh.axisCells[a], CT_MAXAXIS);
        sh.axisCount[a] = n;
        const bool irregular = cyc && ((g.nc[a] & 1) || (g.s[a] & (C - 1)));
        sh.ncol[a] = (g.nc[a] < 3) ? min(g.nc[a], 4) : (irregular ? 4 : 2);
        for (int c = 0; c < 4; ++c) sh.colCount[a][c] = 0;
        for (int k = 0; k < n; ++k) {
            const int col = axisColour(sh.axisCells[a][k], g.nc[a], irregular);
            sh.colList[a][col][sh.colCount[a][col]++] = k;
        }
    }
    __syncthreads();
    const int nax = sh.axisCount[0], nay = sh.axisCount[1], naz = sh.axisCount[2];
    if (nax < 0 || nay < 0 || naz < 0 || nax * nay * naz > CT_MAXCELLS) {
        if (tid == 0) st->pad[0] = 1u; // cannot happen when the host picked the cell size from the reach
        return;
    }
    // segment table of the neighbourhood cells + per-phase lists of the non-empty ones (order inside a phase is free:
    // same-coloured cells never touch the same voxel)
    const int ncx = sh.ncol[0], ncy = sh.ncol[1], ncz = sh.ncol[2];
    const int nphase = ncx * ncy * ncz;
    const int nneigh = nax * nay * naz;
    for (int i = tid; i < nneigh; i += CT_THREADS) {
        const int kx = i % nax, ky = (i / nax) % nay, kz = i / (nax * nay);
        const int cxg = sh.axisCells[0][kx], cyg = sh.axisCells[1][ky], czg = sh.axisCells[2][kz];
        const int czl = cellZLocal(g, czg);
        const size_t cell = cxg + static_cast<size_t>(g.nc[0]) * (cyg + static_cast<size_t>(g.nc[1]) * czl);
        const unsigned b = czl < g.czCount ? cellStart[cell] : 0u, e = czl < g.czCount ? cellStart[cell + 1] : 0u;
        sh.cellB[i] = b;
        sh.cellN[i] = e - b;
        if (e > b) {
            const bool irx = g.cyc[0] && ((g.nc[0] & 1) || (g.s[0] & (C - 1))), iry = g.cyc[1] && ((g.nc[1] & 1) || (g.s[1] & (C - 1))),
                       irz = g.cyc[2] && ((g.nc[2] & 1) || (g.s[2] & (C - 1)));
            const int ph = axisColour(cxg, g.nc[0], irx) + ncx * (axisColour(cyg, g.nc[1], iry) + ncy * axisColour(czg, g.nc[2], irz));
            atomicAdd(&sh.phaseCount[ph], 1);
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int p = 0; p < nphase; ++p) {
            sh.phaseStart[p] = acc;
            acc += sh.phaseCount[p];
            sh.phaseCount[p] = 0;
        }
        sh.phaseStart[nphase] = acc;
    }
    __syncthreads();
    for (int i = tid; i < nneigh; i += CT_THREADS) {
        if (sh.cellN[i] > 0) {
            const int kx = i % nax, ky = (i / nax) % nay, kz = i / (nax * nay);
            const int cxg = sh.axisCells[0][kx], cyg = sh.axisCells[1][ky], czg = sh.axisCells[2][kz];
            const bool irx = g.cyc[0] && ((g.nc[0] & 1) || (g.s[0] & (C - 1))), iry = g.cyc[1] && ((g.nc[1] & 1) || (g.s[1] & (C - 1))),
                       irz = g.cyc[2] && ((g.nc[2] & 1) || (g.s[2] & (C - 1)));
            const int ph = axisColour(cxg, g.nc[0], irx) + ncx * (axisColour(cyg, g.nc[1], iry) + ncy * axisColour(czg, g.nc[2], irz));
            const int slot = atomicAdd(&sh.phaseCount[ph], 1);
            sh.phaseCells[sh.phaseStart[ph] + slot] = static_cast<unsigned short>(i);
        }
    }
    __syncthreads();

    // per-lane constants of the fixed 3x3x3 pattern
    const int pix = lane % 3, piy = (lane / 3) % 3, piz = lane / 9;
    const int plin = pix + piy * CT_SY + piz * CT_SZ;
    const float pfx = (float)pix, pfy = (float)piy, pfz = (float)piz;
    const int t0x = sh.tl0[0], t0y = sh.tl0[1], t0z = sh.tl0[2];
    const int t1x = sh.tl1[0], t1y = sh.tl1[1], t1z = sh.tl1[2];
    const bool pvx = sh.perVoxelWrap[0] != 0, pvy = sh.perVoxelWrap[1] != 0, pvz = sh.perVoxelWrap[2] != 0;
    const bool anyPv = pvx | pvy | pvz;
    const float isdx = __frcp_rn(g.sd[0]), isdy = __frcp_rn(g.sd[1]), isdz = __frcp_rn(g.sd[2]);
    Dig* myDig = sh.dig[warp];

    for (int phase = 0; phase < nphase; ++phase) {
        const int pBeg = sh.phaseStart[phase], pEnd = sh.phaseStart[phase + 1];
        // this warp's cells of the phase: pBeg + warp, + CT_WARPS, ...  (32 of them at a time, one per lane).  The particles of that
        // cell sequence are numbered by a warp scan of the cell sizes; a chunk = 32 consecutive particles of the sequence (a cell may
        // continue in the next chunk), each lane finds its cell by a binary search over the lanes' running totals.
        const int nMyCells = pBeg + warp < pEnd ? (pEnd - pBeg - warp + CT_WARPS - 1) / CT_WARPS : 0;
        for (int k0 = 0; k0 < nMyCells; k0 += 32) {
            unsigned cellBeg = 0, cellCnt = 0;
            if (k0 + lane < nMyCells) {
                const int ci = sh.phaseCells[pBeg + warp + CT_WARPS * (k0 + lane)];
                cellBeg = sh.cellB[ci], cellCnt = sh.cellN[ci];
            }
            const unsigned cellIncl = warpInclusiveScan(cellCnt), cellExcl = cellIncl - cellCnt;
            const unsigned nSeq = __shfl_sync(0xffffffffu, cellIncl, 31);
            for (unsigned t0 = 0; t0 < nSeq; t0 += 32) {
                const unsigned t = t0 + lane;
                unsigned owner = 0; // first lane whose running total exceeds t
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1)
                    if (__shfl_sync(0xffffffffu, cellIncl, owner + step - 1) <= t) owner += step;
                const unsigned src = __shfl_sync(0xffffffffu, cellBeg, owner & 31u) + (t - __shfl_sync(0xffffffffu, cellExcl, owner & 31u));
                const int cnt = static_cast<int>(min(32u, nSeq - t0));
                int myL0x = 0, myL0y = 0, myL0z = 0; // per-voxel-wrap mode only: tile-local box origin before wrapping
                bool myLive = false;                 // my particle reaches the tile
                if (lane < cnt) {
                    // ---- digest my particle -------------------------------------------------------------------
                    const float4 p = recs[src];
                    Dig d;
                    d.x = p.x, d.y = p.y, d.z = p.z;
                    d.weight = (auxN == 1) ? aux[src] : 1.0f;
                    float eps;
                    if (MODE == 0) {
                        eps = __fmul_rn(g.sigma, p.w);   // sigma * rad (:526)
                        d.k0 = __fdiv_rn(1.0f, eps);     // (1.0f / epsilon) (:475)
                    } else {
                        const float sr = __fmul_rn(p.w, g.radscale);
                        eps = __fmul_rn(g.gausslim, sr);
                        d.k0 = __fdiv_rn(-1.4426950408889634f, __fmul_rn(__fmul_rn(2.0f, sr), sr));
                    }
                    d.eps = eps;
                    const float pos[3] = {p.x, p.y, p.z};
                    const float isd[3] = {isdx, isdy, isdz};
                    const int tl0[3] = {t0x, t0y, t0z}, tl1[3] = {t1x, t1y, t1z};
                    const bool pv[3] = {pvx, pvy, pvz};
                    int l0[3], bd[3], tru[3];
                    bool empty = false;
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        // tight bounds of the support along this axis.  The slop only has to beat fp32 rounding of the
                        // bound itself: contributions within 0.48% of the support radius are exactly 0 (exp(-x), x > 104).
                        const float aa = pos[a] - g.mn[a];
                        const float vlo = (aa - eps) * isd[a], vhi = (aa + eps) * isd[a];
                        const float slop = fmaxf(fmaxf(fabsf(vlo), fabsf(vhi)), 1.0f) * 4e-6f;
                        int lo = __float2int_ru(vlo - slop), hi = __float2int_rd(vhi + slop);
                        if (MODE == 0 && g.sigma > 1.0f) { // the reference's box around the home voxel clips the kernel (:573-579)
                            const int H = homeVoxel(pos[a], g.mn[a], g.sd[a]);
                            const int f = filterSize(p.w, g.sd[a]);
                            lo = max(lo, H - f), hi = min(hi, H + f);
                        }
                        hi = min(hi, lo + 2 * reach); // footprint bound the colouring relies on (never binds for sane input)
                        int off = 0;                  // true index = normalised index + off
                        if (g.cyc[a]) {
                            if (lo < 0 || lo >= g.s[a]) {
                                const int m = floorMod(lo, g.s[a]);
                                off = lo - m, hi -= off, lo = m;
                            }
                            if (pv[a]) { // every voxel wraps individually
                                l0[a] = lo - tl0[a], bd[a] = hi - lo + 1, tru[a] = lo + off;
                            } else {
                                int kk = 0;
                                if (!(hi >= tl0[a] && lo <= tl1[a])) kk = -g.s[a]; // the image one period below
                                const int l = max(lo + kk, tl0[a]), h = min(hi + kk, tl1[a]);
                                l0[a] = l - tl0[a], bd[a] = h - l + 1, tru[a] = l - kk + off;
                            }
                        } else {
                            const int l = max(lo, tl0[a]), h = min(hi, tl1[a]);
                            l0[a] = l - tl0[a], bd[a] = h - l + 1, tru[a] = l;
                        }
                        if (bd[a] <= 0) empty = true;
                    }
                    d.f0x = (float)tru[0], d.f0y = (float)tru[1], d.f0z = (float)tru[2];
                    d.base = l0[0] + l0[1] * CT_SY + l0[2] * CT_SZ;
                    myL0x = l0[0], myL0y = l0[1], myL0z = l0[2];
                    if (empty || bd[0] > 255 || bd[1] > 255 || bd[2] > 255) {
                        d.dims = 0u, d.mask27 = 0u;
                    } else {
                        d.dims = (unsigned)bd[0] | ((unsigned)bd[1] << 8) | ((unsigned)bd[2] << 16);
                        unsigned m = 0u;
                        if (bd[0] <= 3 && bd[1] <= 3 && bd[2] <= 3) {
                            const unsigned mx = bd[0] == 1 ? 0x1249249u : (bd[0] == 2 ? 0x36DB6DBu : 0x7FFFFFFu);
                            const unsigned my = bd[1] == 1 ? 0x01C0E07u : (bd[1] == 2 ? 0x0FC7E3Fu : 0x7FFFFFFu);
                            const unsigned mz = bd[2] == 1 ? 0x00001FFu : (bd[2] == 2 ? 0x003FFFFu : 0x7FFFFFFu);
                            m = mx & my & mz;
                        }
                        d.mask27 = m;
                    }
                    myDig[lane] = d;
                    myLive = d.dims != 0u;
                }
                __syncwarp();
                // particles of the ring cells mostly do not reach the tile: walk the non-empty digests only (ascending = canonical order)
                const unsigned live = __ballot_sync(0xffffffffu, myLive);
                for (unsigned rest = live; rest; rest &= rest - 1) {
                    const int j = __ffs(rest) - 1;
                    const float4 A = reinterpret_cast<const float4*>(&myDig[j])[0];
                    const float4 B = reinterpret_cast<const float4*>(&myDig[j])[1];
                    const float4 Cc = reinterpret_cast<const float4*>(&myDig[j])[2];
                    const unsigned dims = __float_as_uint(Cc.z), mask27 = __float_as_uint(Cc.w);
                    const int sbase = __float_as_int(Cc.y);
                    if (mask27 != 0u && !anyPv) {
                        // ---- fast path: box <= 3x3x3, one pass, fixed lane pattern ------------------------------------
                        if ((mask27 >> lane) & 1u) {
                            const float vx = __fadd_rn(__fmul_rn(B.z + pfx, g.sd[0]), g.mn[0]);
                            const float vy = __fadd_rn(__fmul_rn(B.w + pfy, g.sd[1]), g.mn[1]);
                            const float vz = __fadd_rn(__fmul_rn(Cc.x + pfz, g.sd[2]), g.mn[2]);
                            const float dx = __fsub_rn(vx, A.x), dy = __fsub_rn(vy, A.y), dz = __fsub_rn(vz, A.z);
                            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            float w;
                            if (kernelValue<MODE>(d2, A.w, B.x, w)) {
                                float* cell = sh.tile + sbase + plin;
                                *cell = __fadd_rn(*cell, (MODE == 0 && g.agg == 1) ? __fmul_rn(w, B.y) : w);
                            }
                        }
                    } else {
                        // ---- general path: any box, optional per-voxel wrap ------------------------------------------
                        int4 L0 = make_int4(0, 0, 0, 0);
                        if (anyPv) L0 = make_int4(__shfl_sync(0xffffffffu, myL0x, j), __shfl_sync(0xffffffffu, myL0y, j), __shfl_sync(0xffffffffu, myL0z, j), 0);
                        const int bx = dims & 255, by = (dims >> 8) & 255, bz = dims >> 16;
                        const int bxy = bx * by, nvox = bxy * bz;
                        const float rbx = __frcp_rn((float)bx), rbxy = __frcp_rn((float)bxy);
                        for (int i = lane; i < nvox; i += 32) {
                            // exact for these ranges: (i+0.5)/n is never within rounding of an integer
                            const int iz = __float2int_rz(((float)i + 0.5f) * rbxy);
                            const int rem = i - iz * bxy;
                            const int iy = __float2int_rz(((float)rem + 0.5f) * rbx);
                            const int ix = rem - iy * bx;
                            int sidx = sbase + ix + iy * CT_SY + iz * CT_SZ;
                            if (anyPv) {
                                int lx = L0.x + ix, ly = L0.y + iy, lz = L0.z + iz;
                                if (pvx && lx + t0x >= g.s[0]) lx -= g.s[0];
                                if (pvy && ly + t0y >= g.s[1]) ly -= g.s[1];
                                if (pvz && lz + t0z >= g.s[2]) lz -= g.s[2];
                                if ((unsigned)lx > (unsigned)(t1x - t0x) || (unsigned)ly > (unsigned)(t1y - t0y) || (unsigned)lz > (unsigned)(t1z - t0z)) continue;
                                sidx = lx + ly * CT_SY + lz * CT_SZ;
                            }
                            const float vx = __fadd_rn(__fmul_rn(B.z + (float)ix, g.sd[0]), g.mn[0]);
                            const float vy = __fadd_rn(__fmul_rn(B.w + (float)iy, g.sd[1]), g.mn[1]);
                            const float vz = __fadd_rn(__fmul_rn(Cc.x + (float)iz, g.sd[2]), g.mn[2]);
                            const float dx = __fsub_rn(vx, A.x), dy = __fsub_rn(vy, A.y), dz = __fsub_rn(vz, A.z);
                            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            float w;
                            if (kernelValue<MODE>(d2, A.w, B.x, w)) {
                                float* cell = sh.tile + sidx;
                                *cell = __fadd_rn(*cell, (MODE == 0 && g.agg == 1) ? __fmul_rn(w, B.y) : w);
                            }
                        }
                    }
                    __syncwarp(); // the next particle of this cell may touch the same voxels from other lanes
                }
                __syncwarp();
            }
        }
        __syncthreads(); // colour phases are ordered
    }

    // write-out: rows of 32 consecutive voxels (128 B), plus the block's min/max
    float vmin = INFINITY, vmax = -INFINITY;
    for (int r = warp; r < CT_Y * CT_Z; r += CT_WARPS) {
        const int ly = r % CT_Y, lz = r / CT_Y;
        const int x = t0x + lane, y = t0y + ly, z = t0z + lz;
        if (x <= t1x && y <= t1y && z <= t1z) {
            const float v = sh.tile[lane + ly * CT_SY + lz * CT_SZ];
            vol[x + static_cast<size_t>(g.s[0]) * (y + static_cast<size_t>(g.s[1]) * (z - g.z0))] = v;
            vmin = fminf(vmin, v), vmax = fmaxf(vmax, v);
        }
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if (lane == 0 && kmin <= kmax) {
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// density_splat3_kernel: warp-owned sub-tiles (supports of at most 3x3x3 voxels: C1, C2, C4) -- no colour phases, no block barriers
// ---------------------------------------------------------------------------------------------------------------
// A warp OWNS a 32x8x4-voxel sub-tile (eight of them make the 32x16x16 block tile): nobody else ever touches its voxels, so there is
// nothing to colour and nothing to wait for.  The warp streams, in ascending GLOBAL cell order (z, y, x) and canonical in-cell order,
// the particles of every cell row that can reach the sub-tile -- the cells of a row are contiguous in the sorted record array, so a
// row is one or two coalesced ranges --, rejects those whose support box misses the sub-tile with a few float compares (most of the
// ring), and queues the survivors.  Every 32 queued particles are digested with full lanes: tight support box, 27-bit hit mask, the
// hits of all 32 expanded into one (particle, voxel) list that the lanes walk (real hits only: sqrt, rcp, ex2).  Hits of one round
// on the same voxel are applied in list order (match.any + rank).  Every voxel therefore receives its contributions in the order
//        (cell z, cell y, cell x, canonical in-cell order)
// which depends on nothing but the data: bit-identical run to run and for every tile / z-slab decomposition.
constexpr int S3_X = 32, S3_Y = 8, S3_Z = 4;         // voxels owned by one warp
constexpr int S3_SY = 35, S3_SZ = 297;               // padded strides (= 3 and 9 mod 32): a 3x3x3 pattern hits 27 different banks
constexpr int S3_FLOATS = S3_SZ * S3_Z;              // 1188
constexpr int S3_QCAP = 64;                          // survivor queue (ring)
constexpr int S3_HITCAP = 512;                       // (particle, voxel) pairs expanded at a time (32 x 27 worst case: further passes)
constexpr int S3_PRECELLS = 6;                       // y / z cells that can reach a sub-tile (8 + 2*2 voxels: at most 4 cells of 4 voxels)
constexpr int S3_MAXCELLS = 12;                      // cells along one axis that can reach a sub-tile (32/4 + 2 ring cells + wrap slack)

struct Splat3Warp {
    float tile[S3_FLOATS];
    float4 queue[S3_QCAP];
    unsigned short hits[S3_HITCAP];                  // (lane of the particle << 5) | voxel number 0..26 inside its 3x3x3 box, in summation order
    float2 bx[4], by[S3_PRECELLS], bz[S3_PRECELLS];  // pre-test bounds of the sub-tile per x run / y cell / z cell, the cell's periodic shift folded in
};
/** Cells along one axis whose particles can reach a voxel range, ascending global id; shift = what to add to the coordinate
 *  of a particle of that cell (the image next to the cell's own voxels) to get the periodic image that lies next to the range
 *  (0, -period or +period); mid = centre of the cell's voxels.  Both in OBJECT-space units. */
struct Splat3Axis {
    int n;
    int cell[S3_MAXCELLS];
    float shift[S3_MAXCELLS];
    float mid[S3_MAXCELLS];
};
/** The x cells of a block tile in runs of consecutive ids with one shift: contiguous record ranges. */
struct Splat3Runs {
    int nruns;                                       // -1: more than four runs (cannot happen when the host picked the cell size from the reach)
    int runA[4], runB[4];
    float runShift[4], runMid[4];
};
struct Splat3Shared {
    Splat3Warp w[CT_WARPS];
    float4 lut[27];                                  // voxel number -> (ix, iy, iz as floats, sub-tile offset as int bits)
    Splat3Axis axis[6];                              // y of the two sub-tile rows | z of the four sub-tile layers
    Splat3Runs runs;                                 // x of the block tile
};
static_assert(sizeof(Splat3Shared) <= 57088, "density_splat3_kernel must fit four blocks per SM");
/** Derived grid constants, computed once on the host: as kernel parameters they are constant-bank operands (no registers, nothing to
 *  re-derive inside the loops). */
struct Splat3Consts {
    float isd[3];    // 1 / sliceDist
    float per[3];    // period = s * sliceDist
    float rper[3];   // 1 / period
    float slack[3];  // slack of the pre-test, object units
};

/** Cell list of one axis range [t0, t1] of a sub-tile (a = axis, lim = end of the grid / slab along it): ascending global cell id -- the
 *  summation order must not depend on where the tile sits.  Depends on the grid geometry only, so the host builds the tables of all
 *  block coordinates once per geometry (splat3BuildTables) and the kernel just loads its seven. */
__host__ __device__ inline void splat3AxisTable(const Geo& g, int reach, int a, int t0, int lim, int extent, Splat3Axis& A) {
    const int t1 = (t0 + extent < lim ? t0 + extent : lim) - 1;
    int tmp[CT_MAXAXIS];
    const int n = t1 >= t0 ? buildAxisCells(t0, t1, reach, g.s[a], g.cyc[a] != 0, g.cshift, g.nc[a], tmp, S3_MAXCELLS) : 0;
    for (int i = 1; i < n; ++i) {
        const int c = tmp[i];
        int k = i - 1;
        for (; k >= 0 && tmp[k] > c; --k) tmp[k + 1] = tmp[k];
        tmp[k + 1] = c;
    }
    A.n = n;
    const int C = 1 << g.cshift, lo = t0 - reach, hi = t1 + reach;
    for (int i = 0; i < S3_MAXCELLS; ++i) A.cell[i] = 0, A.shift[i] = 0.0f, A.mid[i] = 0.0f;
    for (int i = 0; i < n; ++i) {
        const int v0 = tmp[i] * C, v1 = (v0 + C < g.s[a] ? v0 + C : g.s[a]) - 1; // the cell's voxels
        float shift = 0.0f;
        if (g.cyc[a] && !(v1 >= lo && v0 <= hi)) shift = (v1 + g.s[a] >= lo && v0 + g.s[a] <= hi) ? (float)g.s[a] : -(float)g.s[a];
        A.cell[i] = tmp[i], A.shift[i] = shift * g.sd[a], A.mid[i] = 0.5f * (float)(v0 + v1) * g.sd[a] + g.mn[a];
    }
}
__host__ __device__ inline void splat3RunTable(const Geo& g, const Splat3Axis& A, Splat3Runs& R) {
    const int C = 1 << g.cshift, n = A.n;
    int nr = 0;
    for (int k = 0; k < 4; ++k) R.runA[k] = R.runB[k] = 0, R.runShift[k] = R.runMid[k] = 0.0f;
    for (int i = 0; i < n;) {
        const int ca = A.cell[i];
        int cb = ca;
        for (++i; i < n && A.cell[i] == cb + 1 && A.shift[i] == A.shift[i - 1]; ++i) ++cb;
        if (nr < 4) {
            const int vEnd = (cb + 1) * C < g.s[0] ? (cb + 1) * C : g.s[0];
            R.runA[nr] = ca, R.runB[nr] = cb, R.runShift[nr] = A.shift[i - 1];
            R.runMid[nr] = 0.5f * (float)(ca * C + vEnd - 1) * g.sd[0] + g.mn[0];
        }
        ++nr;
    }
    R.nruns = n < 0 ? -1 : (nr <= 4 ? nr : -1);
}
/** Table layout in device memory: runs[gx] | axisY[2 * gy] | axisZ[4 * gz] (gx, gy, gz = grid of block tiles). */
inline size_t splat3TableBytes(int gx, int gy, int gz) { return sizeof(Splat3Runs) * gx + sizeof(Splat3Axis) * (2 * gy + 4 * gz); }
inline void splat3BuildTables(const Geo& g, int reach, int gx, int gy, int gz, unsigned char* out) {
    Splat3Runs* runs = reinterpret_cast<Splat3Runs*>(out);
    Splat3Axis* ay = reinterpret_cast<Splat3Axis*>(out + sizeof(Splat3Runs) * gx);
    Splat3Axis* az = ay + 2 * gy;
    for (int b = 0; b < gx; ++b) {
        Splat3Axis A;
        splat3AxisTable(g, reach, 0, b * CT_X, g.s[0], S3_X, A);
        splat3RunTable(g, A, runs[b]);
    }
    for (int b = 0; b < 2 * gy; ++b) splat3AxisTable(g, reach, 1, (b >> 1) * CT_Y + (b & 1) * S3_Y, g.s[1], S3_Y, ay[b]);
    for (int b = 0; b < 4 * gz; ++b) splat3AxisTable(g, reach, 2, g.z0 + (b >> 2) * CT_Z + (b & 3) * S3_Z, g.z0 + g.nz, S3_Z, az[b]);
}

__global__ void __launch_bounds__(CT_THREADS, 4) density_splat3_kernel(Geo g, Splat3Consts kc, DevState* st, const float4* __restrict__ recs,
    const unsigned* __restrict__ cellStart, float* __restrict__ vol, int reach, const unsigned char* __restrict__ tables) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Splat3Shared& sh = *reinterpret_cast<Splat3Shared*>(smemRaw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 27) {
        const int ix = tid % 3, iy = (tid / 3) % 3, iz = tid / 9;
        sh.lut[tid] = make_float4((float)ix, (float)iy, (float)iz, __int_as_float(ix + iy * S3_SY + iz * S3_SZ));
    }
    {   // the cell lists of the block's x range, its two y ranges and its four z ranges: warps 0..5 copy one axis table each, warp 6 the runs
        const unsigned char* ty = tables + sizeof(Splat3Runs) * gridDim.x;
        const int* src = nullptr;
        int* dst = nullptr;
        int nw = 0;
        if (warp < 2) src = reinterpret_cast<const int*>(ty + sizeof(Splat3Axis) * (2 * blockIdx.y + warp)), dst = reinterpret_cast<int*>(&sh.axis[warp]), nw = sizeof(Splat3Axis) / 4;
        else if (warp < 6) src = reinterpret_cast<const int*>(ty + sizeof(Splat3Axis) * (2 * gridDim.y + 4 * blockIdx.z + warp - 2)), dst = reinterpret_cast<int*>(&sh.axis[warp]), nw = sizeof(Splat3Axis) / 4;
        else if (warp == 6) src = reinterpret_cast<const int*>(tables + sizeof(Splat3Runs) * blockIdx.x), dst = reinterpret_cast<int*>(&sh.runs), nw = sizeof(Splat3Runs) / 4;
        for (int i = lane; i < nw; i += 32) dst[i] = __ldg(src + i);
    }
    __syncthreads(); // the only block barrier: from here on the warps are independent
    Splat3Warp& W = sh.w[warp];
    const int t0x = (int)blockIdx.x * CT_X, t0y = (int)blockIdx.y * CT_Y + (warp & 1) * S3_Y, t0z = g.z0 + (int)blockIdx.z * CT_Z + (warp >> 1) * S3_Z;
    const int t1x = min(t0x + S3_X, g.s[0]) - 1, t1y = min(t0y + S3_Y, g.s[1]) - 1, t1z = min(t0z + S3_Z, g.z0 + g.nz) - 1;
    if (t1y < t0y || t1z < t0z) return;
    const Splat3Axis& AY = sh.axis[warp & 1];
    const Splat3Axis& AZ = sh.axis[2 + (warp >> 1)];
    const int nruns = sh.runs.nruns;
    const bool anyWrapped = st->pad[1] != 0u; // set by bin_count_kernel
    if (nruns < 0 || AY.n < 0 || AZ.n < 0) {
        if (lane == 0) st->pad[0] = 1u; // cannot happen when the host picked the cell size from the reach
        return;
    }
    for (int i = lane; i < S3_FLOATS / 4; i += 32) reinterpret_cast<float4*>(W.tile)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    __syncwarp();
    const unsigned ltMask = (1u << lane) - 1u;
    unsigned qHead = 0, qTail = 0; // ring positions (uniform)

    // tight support box of a particle along one axis, clipped to the sub-tile: first voxel (tile-local), extent, true (un-wrapped) index
    auto axisBox = [&](float pos, float rad, float eps, float mn, float sd, float isd, int s, bool cyc, int tl0, int tl1, int& l0, int& bd, int& tru) {
        // The slop only has to beat fp32 rounding of the bound itself: contributions within 0.48% of the support radius are exactly 0
        // (exp(-x), x > 104).
        const float aa = pos - mn;
        const float vlo = (aa - eps) * isd, vhi = (aa + eps) * isd;
        const float slop = fmaxf(fmaxf(fabsf(vlo), fabsf(vhi)), 1.0f) * 4e-6f;
        int lo = __float2int_ru(vlo - slop), hi = __float2int_rd(vhi + slop);
        if (g.sigma > 1.0f) { // the reference's box around the home voxel clips the kernel (:573-579)
            const int H = homeVoxel(pos, mn, sd);
            const int f = filterSize(rad, sd);
            lo = max(lo, H - f), hi = min(hi, H + f);
        }
        hi = min(hi, lo + 2 * reach); // footprint bound the cell ring relies on (never binds for sane input)
        int kk = 0, off = 0;          // true index = normalised index + off; kk: the image one period below
        if (cyc) {
            if (lo < 0 || lo >= s) {
                const int mm = floorMod(lo, s);
                off = lo - mm, hi -= off, lo = mm;
            }
            if (!(hi >= tl0 && lo <= tl1)) kk = -s;
        }
        const int l = max(lo + kk, tl0), h = min(hi + kk, tl1);
        l0 = l - tl0, bd = h - l + 1, tru = l - kk + off;
    };

    // ---- a batch of queued particles: digest, hit masks, hit list, evaluation ----------------------------------------------------
    auto processBatch = [&](int cnt) {
        Dig dg{};
        bool myLive = false;
        if (lane < cnt) {
            const float4 p = W.queue[(qHead + lane) & (S3_QCAP - 1)];
            dg.x = p.x, dg.y = p.y, dg.z = p.z;
            const float eps = __fmul_rn(g.sigma, p.w); // sigma * rad (:526)
            dg.k0 = __fdiv_rn(1.0f, eps);               // (1.0f / epsilon) (:475)
            dg.eps = eps;
            int lx, ly, lz, bx, by, bz, tx, ty, tz;
            axisBox(p.x, p.w, eps, g.mn[0], g.sd[0], kc.isd[0], g.s[0], g.cyc[0] != 0, t0x, t1x, lx, bx, tx);
            axisBox(p.y, p.w, eps, g.mn[1], g.sd[1], kc.isd[1], g.s[1], g.cyc[1] != 0, t0y, t1y, ly, by, ty);
            axisBox(p.z, p.w, eps, g.mn[2], g.sd[2], kc.isd[2], g.s[2], g.cyc[2] != 0, t0z, t1z, lz, bz, tz);
            dg.f0x = (float)tx, dg.f0y = (float)ty, dg.f0z = (float)tz;
            dg.base = lx + ly * S3_SY + lz * S3_SZ;
            unsigned mk = 0u;
            if (bx > 0 && by > 0 && bz > 0) {
                if (bx <= 3 && by <= 3 && bz <= 3) {
                    const unsigned mx = bx == 1 ? 0x1249249u : (bx == 2 ? 0x36DB6DBu : 0x7FFFFFFu);
                    const unsigned my = by == 1 ? 0x01C0E07u : (by == 2 ? 0x0FC7E3Fu : 0x7FFFFFFu);
                    const unsigned mz = bz == 1 ? 0x00001FFu : (bz == 2 ? 0x003FFFFu : 0x7FFFFFFu);
                    mk = mx & my & mz;
                } else {
                    st->pad[0] = 4u; // the host promised boxes <= 3x3x3
                }
            }
            dg.mask27 = mk;
            myLive = mk != 0u;
        }
        // ---- A: hit mask of my particle ---------------------------------------------------------------------
        unsigned hm = 0u;
        if (myLive) {
            const float lim = __fmul_rn(__fmul_rn(dg.eps, dg.eps), 1.000001f);
            float qx[3], qy[3], qz[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float dx = __fsub_rn(__fadd_rn(__fmul_rn(dg.f0x + (float)i, g.sd[0]), g.mn[0]), dg.x);
                const float dy = __fsub_rn(__fadd_rn(__fmul_rn(dg.f0y + (float)i, g.sd[1]), g.mn[1]), dg.y);
                const float dz = __fsub_rn(__fadd_rn(__fmul_rn(dg.f0z + (float)i, g.sd[2]), g.mn[2]), dg.z);
                qx[i] = __fmul_rn(dx, dx), qy[i] = __fmul_rn(dy, dy), qz[i] = __fmul_rn(dz, dz);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float sxy = __fadd_rn(qx[i], qy[j]);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        if (__fadd_rn(sxy, qz[k]) < lim) hm |= 1u << (i + 3 * j + 9 * k);
                }
            hm &= dg.mask27;
        }
        const unsigned hcnt = __popc(hm);
        const unsigned incl = warpInclusiveScan(hcnt), excl = incl - hcnt;
        const unsigned total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned done = 0;
        int first = 0;
        while (done < total) {
            // the particles [first, last] whose hits fit the list together (a particle has <= 27 of them)
            const bool in = lane >= first && incl - done <= (unsigned)S3_HITCAP;
            const unsigned bal = __ballot_sync(0xffffffffu, in);
            const int last = 31 - __clz(bal);
            const unsigned nh = __shfl_sync(0xffffffffu, incl, last) - done;
            if (in) {
                unsigned pos = excl - done;
                for (unsigned mm = hm; mm; mm &= mm - 1) W.hits[pos++] = static_cast<unsigned short>(lane << 5 | (__ffs(mm) - 1));
            }
            __syncwarp();
            // ---- B: lanes = hits ------------------------------------------------------------------------------
            for (unsigned h0 = 0; h0 < nh; h0 += 32) {
                const unsigned h = h0 + lane;
                const bool act = h < nh;
                const unsigned code = act ? W.hits[h] : 0u;
                const int pl = code >> 5;
                const float px = __shfl_sync(0xffffffffu, dg.x, pl), py = __shfl_sync(0xffffffffu, dg.y, pl), pz = __shfl_sync(0xffffffffu, dg.z, pl);
                const float eps = __shfl_sync(0xffffffffu, dg.eps, pl), k0 = __shfl_sync(0xffffffffu, dg.k0, pl);
                const float f0x = __shfl_sync(0xffffffffu, dg.f0x, pl), f0y = __shfl_sync(0xffffffffu, dg.f0y, pl), f0z = __shfl_sync(0xffffffffu, dg.f0z, pl);
                const int sb = __shfl_sync(0xffffffffu, dg.base, pl);
                const float4 L = sh.lut[code & 31u];
                const float vx = __fadd_rn(__fmul_rn(f0x + L.x, g.sd[0]), g.mn[0]);
                const float vy = __fadd_rn(__fmul_rn(f0y + L.y, g.sd[1]), g.mn[1]);
                const float vz = __fadd_rn(__fmul_rn(f0z + L.z, g.sd[2]), g.mn[2]);
                const float dx = __fsub_rn(vx, px), dy = __fsub_rn(vy, py), dz = __fsub_rn(vz, pz);
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                float w = 0.0f;
                const bool hit = act && kernelValue<0>(d2, eps, k0, w);
                const int addr = sb + __float_as_int(L.w);
                // hits of this round on the same voxel: list order = canonical order
                const unsigned grp = __match_any_sync(0xffffffffu, hit ? addr : -1 - lane);
                const int rank = __popc(grp & ltMask);
                const int maxRank = __reduce_max_sync(0xffffffffu, hit ? rank : 0);
                for (int r = 0; r <= maxRank; ++r) {
                    if (hit && rank == r) W.tile[addr] = __fadd_rn(W.tile[addr], w);
                    __syncwarp();
                }
            }
            done += nh;
            first = last + 1;
        }
    };

    // ---- stream the cell rows ----------------------------------------------------------------------------------------------------
    // Conservative pre-test in object space: the support of a particle is [p - eps, p + eps]; the sub-tile's voxels span [lo, hi] (plus
    // a slack for the rounding of this test and the slop of the exact integer box, which follows in the digest).  On a periodic axis
    // the particle is first brought next to the voxels of its own cell (row / run centre +- half a period: robust against a home voxel
    // that the binning's exact division puts one cell further) and then moved to the image next to the sub-tile (the cell's shift).
    // With the cell's shift folded into the bounds the test of an in-box particle is p + eps >= lo - shift && p - eps <= hi - shift.
    const int nay = AY.n, naz = AZ.n;
    if (nay > S3_PRECELLS || naz > S3_PRECELLS) {
        if (lane == 0) st->pad[0] = 1u; // cannot happen: 8 + 2*2 voxels touch at most four 4-voxel cells
        return;
    }
    if (lane < 4 + 2 * S3_PRECELLS) {
        const int a = lane < 4 ? 0 : (lane < 4 + S3_PRECELLS ? 1 : 2), i = lane < 4 ? lane : (lane < 4 + S3_PRECELLS ? lane - 4 : lane - 4 - S3_PRECELLS);
        const int t0 = a == 0 ? t0x : (a == 1 ? t0y : t0z), t1 = a == 0 ? t1x : (a == 1 ? t1y : t1z);
        const float shift = a == 0 ? sh.runs.runShift[i] : (a == 1 ? AY.shift[i] : AZ.shift[i]);
        // (selects, not g.sd[a]: a dynamic index would make the compiler keep a local-memory copy of the parameter structs)
        const float sdA = a == 0 ? g.sd[0] : (a == 1 ? g.sd[1] : g.sd[2]), mnA = a == 0 ? g.mn[0] : (a == 1 ? g.mn[1] : g.mn[2]);
        const float slackA = a == 0 ? kc.slack[0] : (a == 1 ? kc.slack[1] : kc.slack[2]);
        const float2 b = make_float2(fmaf((float)t0, sdA, mnA) - slackA - shift, fmaf((float)t1, sdA, mnA) + slackA - shift);
        (a == 0 ? W.bx : (a == 1 ? W.by : W.bz))[i] = b;
    }
    __syncwarp();
    // Row descriptors: lane i owns (cell z, cell y, x run) number i of the sub-tile's neighbourhood, in summation order, and loads its
    // record range -- one round of table reads for all rows instead of one dependent read per row.
    const int nrows = naz * nay * nruns; // <= 4 * 4 * 4; in practice 12 .. 24
    for (int row0 = 0; row0 < nrows; row0 += 32) {
        unsigned myB = 0, myE = 0;
        int myCode = 0; // kz | ky << 4 | kr << 8
        if (row0 + lane < nrows) {
            const int i = row0 + lane;
            const int kr = i % nruns, t = i / nruns, ky = t % nay, kz = t / nay;
            const int czl = cellZLocal(g, AZ.cell[kz]);
            const size_t rowBase = (static_cast<size_t>(czl) * g.nc[1] + AY.cell[ky]) * g.nc[0];
            if (czl < g.czCount) myB = cellStart[rowBase + sh.runs.runA[kr]], myE = cellStart[rowBase + sh.runs.runB[kr] + 1];
            myCode = kz | ky << 4 | kr << 8;
        }
        const unsigned nonEmpty = __ballot_sync(0xffffffffu, myE > myB);
        // flattened walk over the chunks (32 records) of the non-empty rows; the records of the NEXT chunk are requested before the
        // current one is tested, so that the load latency overlaps the work
        unsigned rest = nonEmpty;
        unsigned base = 0, end = 0;
        int code = 0;
        auto nextChunk = [&]() { // -> false when the rows are exhausted
            base += 32;
            if (base >= end) {
                if (!rest) return false;
                const int r = __ffs(rest) - 1;
                rest &= rest - 1;
                base = __shfl_sync(0xffffffffu, myB, r), end = __shfl_sync(0xffffffffu, myE, r), code = __shfl_sync(0xffffffffu, myCode, r);
            }
            return true;
        };
        bool more = nextChunk();
        float4 pNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (more && base + lane < end) pNext = recs[base + lane];
        while (more) {
            const float4 p = pNext;
            bool keep = base + lane < end;
            const int kz = code & 15, ky = (code >> 4) & 15, kr = code >> 8;
            more = nextChunk();
            if (more && base + lane < end) pNext = recs[base + lane];
            if (keep) {
                const float eps = g.sigma * p.w;
                float ux = p.x, uy = p.y, uz = p.z;
                if (anyWrapped) { // some particle lies outside the box (binned through the periodic wrap): bring it next to its cell first
                    if (g.cyc[0]) ux = fmaf(-rintf((ux - sh.runs.runMid[kr]) * kc.rper[0]), kc.per[0], ux);
                    if (g.cyc[1]) uy = fmaf(-rintf((uy - AY.mid[ky]) * kc.rper[1]), kc.per[1], uy);
                    if (g.cyc[2]) uz = fmaf(-rintf((uz - AZ.mid[kz]) * kc.rper[2]), kc.per[2], uz);
                }
                const float2 bx = W.bx[kr], by = W.by[ky], bz = W.bz[kz];
                keep = ux + eps >= bx.x && ux - eps <= bx.y && uy + eps >= by.x && uy - eps <= by.y && uz + eps >= bz.x && uz - eps <= bz.y;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) W.queue[(qTail + __popc(bal & ltMask)) & (S3_QCAP - 1)] = p;
            qTail += __popc(bal);
            __syncwarp();
            if (qTail - qHead >= 32u) {
                processBatch(32);
                qHead += 32;
            }
        }
    }
    if (qTail != qHead) processBatch(static_cast<int>(qTail - qHead));
    __syncwarp();

    // write-out: 32 rows of 32 consecutive voxels (128 B), plus the sub-tile's min/max
    float vmin = INFINITY, vmax = -INFINITY;
    if (t0x + lane <= t1x) {
        float* out = vol + (t0x + lane) + static_cast<size_t>(g.s[0]) * (t0y + static_cast<size_t>(g.s[1]) * (t0z - g.z0));
        const size_t planeStride = static_cast<size_t>(g.s[0]) * g.s[1];
        const int ny = t1y - t0y + 1, nz = t1z - t0z + 1;
        for (int lz = 0; lz < nz; ++lz, out += planeStride) {
            const float* src = W.tile + lane + lz * S3_SZ;
            float* o = out;
#pragma unroll 4
            for (int ly = 0; ly < ny; ++ly, o += g.s[0], src += S3_SY) {
                const float v = *src;
                *o = v;
                vmin = fminf(vmin, v), vmax = fmaxf(vmax, v);
            }
        }
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if (lane == 0 && kmin <= kmax) {
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Wide supports: voxel gather
// ---------------------------------------------------------------------------------------------------------------
// density_gather_kernel: QuickSurf-Gaussian mode (supports of ~10 voxels, optional density-weighted RGB volume) and
// P2D supports wider than 8 voxels.  A block owns a 32x8x8-voxel tile; each thread keeps a column of 8 voxels (+ RGB)
// in registers.  The particles of every cell the tile's neighbourhood touches are streamed through shared memory in
// (cell z, cell y, cell x, canonical in-cell) order, 256 at a time; every thread walks the staged candidates in that
// order, rejects on the xy distance first, and accumulates in registers -- no atomics, one fixed order per voxel.
constexpr int GT_X = 32, GT_Y = 8, GT_Z = 8;
constexpr int GT_THREADS = 256;
constexpr int GT_CHUNK = 256;
constexpr int GT_MAXSEG = 512;

struct GCand {          // 48 bytes
    float x, y, z, k0;  // k0: QS w_p = -log2e/(2 (r radscale)^2); P2D 1/eps
    float lim, cr, cg, cb; // lim: QS cut-off^2, P2D eps;  colour (QS) / intensity in cr (P2D aggregator 1)
    int kx, ky, kz, pad;   // periodic image: the voxel index seen by this particle is t + k (k in {-s, 0, +s})
};

struct GatherShared {
    GCand cand[GT_CHUNK];
    unsigned segBegin[GT_MAXSEG];
    unsigned segPrefix[GT_MAXSEG + 1];
    int axisCells[3][CT_MAXAXIS];
    int axisCount[3];
    unsigned scanTmp[33];
};

/** GENERAL (host-selected): periodic axes so short that SEVERAL images of a particle reach one tile -- down to supports wider than
 *  the axis, where a voxel receives the same particle more than once, exactly as the reference's loop over the un-wrapped box does
 *  (ParticlesToDensity.cpp:577-613) -- and, in bump mode, the reference's integer support box home +- ceil(rad/sliceDist), which
 *  clips the kernel when sigma > 1 (:573-579).  Every (particle, image) pair is a candidate of its own, in (particle, image z, y, x)
 *  order.  !GENERAL: at most one image per particle and tile, no box test (sigma <= 1: the box never clips).
 *  MODE 1 (Gaussian, non-periodic) && GENERAL = MMS_MODE_QS_GAUSS_REFCELLS: the candidate set of the reference's CUDAQuickSurf instead
 *  of the radial cut-off -- every atom whose acceleration cell lies in the cell range of the voxel's 8x8x8 block
 *  (CUDAQuickSurf.cu:232-253), coordinates relative to the grid origin like the reference's (QuickSurf.cpp:553-556). */
template<int MODE, bool COLOUR, bool GENERAL>
__global__ void __launch_bounds__(GT_THREADS, 2) density_gather_kernel(Geo g, DevState* st, const float4* __restrict__ recs,
    const float* __restrict__ aux, int auxN, const unsigned* __restrict__ cellStart, float* __restrict__ vol, float* __restrict__ rgb,
    int reach) {
    __shared__ GatherShared sh;
    constexpr bool QSREF = MODE == 1 && GENERAL;
    const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
    const int t0x = (int)blockIdx.x * GT_X, t0y = (int)blockIdx.y * GT_Y, t0z = g.z0 + (int)blockIdx.z * GT_Z;
    const int t1x = min(t0x + GT_X, g.s[0]) - 1, t1y = min(t0y + GT_Y, g.s[1]) - 1, t1z = min(t0z + GT_Z, g.z0 + g.nz) - 1;
    if (tid < 3) {
        const int a = tid;
        const int t0 = a == 0 ? t0x : (a == 1 ? t0y : t0z), t1 = a == 0 ? t1x : (a == 1 ? t1y : t1z);
        sh.axisCount[a] = buildAxisCells(t0, t1, reach, g.s[a], g.cyc[a] != 0, g.cshift, g.nc[a], sh.axisCells[a], CT_MAXAXIS);
        // two periodic images of one particle reaching the same tile: the GENERAL variant's business (the host selects it)
        if (!GENERAL && g.cyc[a] && g.s[a] < (t1 - t0 + 1) + 2 * reach + 2) st->pad[0] = 2u;
    }
    __syncthreads();
    const int ncx = sh.axisCount[0], ncy = sh.axisCount[1], ncz = sh.axisCount[2];
    if (ncx < 0 || ncy < 0 || ncz < 0) {
        if (tid == 0) st->pad[0] = 1u;
        return;
    }
    int nruns = 0;
    int runStart[4], runEnd[4];
    for (int k = 0; k < ncx && nruns < 4; ++k) {
        const int c = sh.axisCells[0][k];
        if (nruns > 0 && c == runEnd[nruns - 1] + 1) runEnd[nruns - 1] = c;
        else { runStart[nruns] = c; runEnd[nruns] = c; ++nruns; }
    }
    const int nrows = ncy * ncz, nsegTotal = nrows * nruns;

    // warp = one row of the tile (an 8 x 4 patch per warp has better lane efficiency -- 63 % instead of 45 % -- but unbalances the warps
    // between the per-chunk barriers: 26.2 ms instead of 23.2 ms on C3)
    const int vxI = t0x + lane, vyI = t0y + ty;
    // QSREF: coorx = gridspacing * xindex, relative to the grid origin (CUDAQuickSurf.cu:255-257)
    const float vx = QSREF ? __fmul_rn(g.sd[0], (float)vxI) : __fadd_rn(__fmul_rn((float)vxI, g.sd[0]), g.mn[0]);
    const float vy = QSREF ? __fmul_rn(g.sd[1], (float)vyI) : __fadd_rn(__fmul_rn((float)vyI, g.sd[1]), g.mn[1]);
    // QSREF: acceleration-cell ranges of the reference's 8x8x8 thread blocks this tile covers (its own expressions, :232-253)
    auto abMin = [&](int b, int) { const int v = __float2int_rz(__fmul_rn(__fmaf_rn((float)(b * 8), g.sd[0], -g.qsAc), g.qsInvAc)); return v < 0 ? 0 : v; };
    auto abMax = [&](int b, int nc) { const int v = __float2int_rz(__fmul_rn(__fmaf_rn((float)((b + 1) * 8), g.sd[0], g.qsAc), g.qsInvAc)); return v >= nc - 1 ? nc - 1 : v; };
    int refLo[3] = {0, 0, 0}, refHi[3] = {0, 0, 0}, myXlo = 0, myXhi = 0;
    if (QSREF) {
        refLo[0] = abMin(t0x / 8, 0), refHi[0] = abMax(t1x / 8, g.qsCells[0]);
        refLo[1] = abMin(t0y / 8, 0), refHi[1] = abMax(t0y / 8, g.qsCells[1]);
        refLo[2] = abMin(t0z / 8, 0), refHi[2] = abMax(t0z / 8, g.qsCells[2]);
        myXlo = abMin(vxI / 8, 0), myXhi = abMax(vxI / 8, g.qsCells[0]);
    }
    float vz[GT_Z], acc[GT_Z], accR[GT_Z], accG[GT_Z], accB[GT_Z];
#pragma unroll
    for (int k = 0; k < GT_Z; ++k) {
        vz[k] = QSREF ? __fmul_rn(g.sd[2], (float)(t0z + k)) : __fadd_rn(__fmul_rn((float)(t0z + k), g.sd[2]), g.mn[2]);
        acc[k] = 0.0f, accR[k] = 0.0f, accG[k] = 0.0f, accB[k] = 0.0f;
    }

    for (int segBase = 0; segBase < nsegTotal; segBase += GT_MAXSEG) {
        const int nseg = min(GT_MAXSEG, nsegTotal - segBase);
        __syncthreads();
        unsigned carry = 0;
        for (int b0 = 0; b0 < nseg; b0 += GT_THREADS) {
            const int sI = b0 + tid;
            unsigned len = 0;
            if (sI < nseg) {
                const int gs = segBase + sI;
                const int row = gs / nruns, run = gs - row * nruns;
                const int cz = sh.axisCells[2][row / ncy], cy = sh.axisCells[1][row % ncy];
                const int czl = cellZLocal(g, cz);
                const size_t rowBase = (static_cast<size_t>(czl) * g.nc[1] + cy) * g.nc[0];
                const unsigned beg = czl < g.czCount ? cellStart[rowBase + runStart[run]] : 0u;
                len = czl < g.czCount ? cellStart[rowBase + runEnd[run] + 1] - beg : 0u;
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
        for (unsigned chunk = 0; chunk < ncand; chunk += GT_CHUNK) {
            const unsigned nin = min((unsigned)GT_CHUNK, ncand - chunk);
            __syncthreads();
            GCand c;
            bool valid = false;
            int qmin[3] = {0, 0, 0}, nq[3] = {1, 1, 1}, boxLo[3] = {0, 0, 0}, boxHi[3] = {0, 0, 0};
            float px3[3] = {0.0f, 0.0f, 0.0f}, eps2 = 0.0f;
            if ((unsigned)tid < nin) {
                const unsigned pos = chunk + tid;
                int lo = 0, hi = nseg;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sh.segPrefix[mid] <= pos) lo = mid; else hi = mid;
                }
                const unsigned idx = sh.segBegin[lo] + (pos - sh.segPrefix[lo]);
                const float4 p = recs[idx];
                c.x = p.x, c.y = p.y, c.z = p.z;
                c.cr = c.cg = c.cb = 1.0f;
                float eps;
                if (MODE == 0) {
                    eps = __fmul_rn(g.sigma, p.w);
                    c.k0 = __fdiv_rn(1.0f, eps);
                    c.lim = eps;
                    if (auxN == 1) c.cr = aux[idx];
                    if (COLOUR && auxN == 4) { // aggregator 2: the particle's direction (dxAcc/dyAcc/dzAcc, ParticlesToDensity.cpp:497-499)
                        const float4 d = reinterpret_cast<const float4*>(aux)[idx];
                        c.cr = d.x, c.cg = d.y, c.cb = d.z;
                    }
                } else {
                    const float sr = __fmul_rn(p.w, g.radscale);
                    eps = __fmul_rn(g.gausslim, sr);
                    c.k0 = __fdiv_rn(-1.4426950408889634f, __fmul_rn(__fmul_rn(2.0f, sr), sr));
                    c.lim = __fmul_rn(eps, eps);
                    if (QSREF) { // no cut-off; positions relative to the grid origin
                        c.lim = INFINITY;
                        c.x = __fsub_rn(p.x, g.mn[0]), c.y = __fsub_rn(p.y, g.mn[1]), c.z = __fsub_rn(p.z, g.mn[2]);
                    }
                    if (auxN == 4) {
                        const float4 col = reinterpret_cast<const float4*>(aux)[idx];
                        c.cr = col.x, c.cg = col.y, c.cb = col.z;
                    }
                }
                // periodic images that can reach this tile: voxel index seen by the particle = t + q*s must meet [H - R, H + R]
                const float pp[3] = {p.x, p.y, p.z};
                const int tl0[3] = {t0x, t0y, t0z}, tl1[3] = {t1x, t1y, t1z};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    qmin[a] = 0, nq[a] = 1, boxLo[a] = -(1 << 20), boxHi[a] = 1 << 20;
                    if (QSREF) { // the atom's acceleration cell: min(int(p * invcellsize), ncells - 1) (CUDASpatialSearch.cu:92-94)
                        const float rel = a == 0 ? c.x : (a == 1 ? c.y : c.z);
                        qmin[a] = min(__float2int_rz(__fmul_rn(rel, g.qsInvAc)), g.qsCells[a] - 1);
                        continue;
                    }
                    if (GENERAL && MODE == 0) { // the reference's integer support box, in un-wrapped voxel indices
                        const int H = homeVoxel(pp[a], g.mn[a], g.sd[a]), f = filterSize(p.w, g.sd[a]);
                        boxLo[a] = H - f, boxHi[a] = H + f;
                    }
                    if (g.cyc[a]) {
                        const int H = homeVoxel(pp[a], g.mn[a], g.sd[a]);
                        const int R = (GENERAL && MODE == 0) ? filterSize(p.w, g.sd[a]) : reach + 1;
                        const int a0 = H - R - tl1[a], a1 = H + R - tl0[a];
                        const int qlo = -((-a0 - floorMod(-a0, g.s[a])) / g.s[a]); // ceil(a0 / s)
                        const int qhi = (a1 - floorMod(a1, g.s[a])) / g.s[a];      // floor(a1 / s)
                        qmin[a] = qlo;
                        nq[a] = GENERAL ? max(qhi - qlo + 1, 0) : 1; // !GENERAL: at most one image works (checked per block above)
                        if (!GENERAL && tl0[a] + qlo * g.s[a] > H + R) qmin[a] = 0; // no image reaches this tile: gap2 rejects it
                    }
                }
                eps2 = eps * eps * 1.01f; // sphere / tile-box rejection (1 % slack for the rounding of this test; the exact test is per voxel)
                px3[0] = p.x, px3[1] = p.y, px3[2] = p.z;
                valid = true;
            }
            // squared distance from the particle to the tile's box in the frame of image (qx, qy, qz)
            auto imageKeep = [&](int qx, int qy, int qz) {
                if (QSREF) // qx, qy, qz = the atom's acceleration cell: inside the cell range of the tile's thread blocks?
                    return qx >= refLo[0] && qx <= refHi[0] && qy >= refLo[1] && qy <= refHi[1] && qz >= refLo[2] && qz <= refHi[2];
                const int q[3] = {qx, qy, qz};
                const int tl0[3] = {t0x, t0y, t0z}, tl1[3] = {t1x, t1y, t1z};
                float gap2 = 0.0f;
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const int k = q[a] * g.s[a];
                    const float lo_ = (float)(tl0[a] + k) * g.sd[a] + g.mn[a], hi_ = (float)(tl1[a] + k) * g.sd[a] + g.mn[a];
                    const float d = fmaxf(fmaxf(lo_ - px3[a], px3[a] - hi_), 0.0f);
                    gap2 += d * d;
                }
                return gap2 <= eps2;
            };
            // allowed tile-local voxel ranges of the integer box for image (qx, qy, qz), packed: x 2 x 6 bits, y and z 2 x 4 bits
            auto packBox = [&](int qx, int qy, int qz) {
                const int lx = min(max(boxLo[0] - qx * g.s[0] - t0x, 0), GT_X + 1), ux = min(max(boxHi[0] - qx * g.s[0] - t0x + 1, 0), GT_X + 1);
                const int ly = min(max(boxLo[1] - qy * g.s[1] - t0y, 0), GT_Y + 1), uy = min(max(boxHi[1] - qy * g.s[1] - t0y + 1, 0), GT_Y + 1);
                const int lz = min(max(boxLo[2] - qz * g.s[2] - t0z, 0), GT_Z + 1), uz = min(max(boxHi[2] - qz * g.s[2] - t0z + 1, 0), GT_Z + 1);
                return lx | ux << 6 | ly << 12 | uy << 16 | lz << 20 | uz << 24;
            };
            unsigned mine = 0; // my particle's images that pass the rejection test
            if (valid) {
                if (!GENERAL) mine = imageKeep(qmin[0], qmin[1], qmin[2]) ? 1u : 0u;
                else
                    for (int iz = 0; iz < nq[2]; ++iz)
                        for (int iy = 0; iy < nq[1]; ++iy)
                            for (int ix = 0; ix < nq[0]; ++ix) mine += imageKeep(qmin[0] + ix, qmin[1] + iy, qmin[2] + iz) ? 1u : 0u;
            }
            // order-preserving expansion of the surviving (particle, image) pairs, GT_CHUNK of them at a time
            unsigned ntotal;
            const unsigned slot0 = blockExclusiveScan(mine, &ntotal, sh.scanTmp);
            for (unsigned w0 = 0; w0 < ntotal; w0 += GT_CHUNK) {
                if (w0) __syncthreads(); // the previous window has been consumed
                if (mine && slot0 + mine > w0 && slot0 < w0 + GT_CHUNK) {
                    unsigned sl = slot0;
                    for (int iz = 0; iz < nq[2]; ++iz)
                        for (int iy = 0; iy < nq[1]; ++iy)
                            for (int ix = 0; ix < nq[0]; ++ix) {
                                const int qx = qmin[0] + ix, qy = qmin[1] + iy, qz = qmin[2] + iz;
                                if (GENERAL && !imageKeep(qx, qy, qz)) continue;
                                if (sl >= w0 && sl < w0 + GT_CHUNK) {
                                    c.kx = QSREF ? qx : qx * g.s[0], c.ky = QSREF ? qy : qy * g.s[1], c.kz = QSREF ? qz : qz * g.s[2];
                                    c.pad = (GENERAL && MODE == 0) ? packBox(qx, qy, qz) : 0;
                                    sh.cand[sl - w0] = c;
                                }
                                ++sl;
                            }
                }
                __syncthreads();
                const unsigned nstaged = min((unsigned)GT_CHUNK, ntotal - w0);
            for (unsigned j = 0; j < nstaged; ++j) {
                const float4 A = reinterpret_cast<const float4*>(&sh.cand[j])[0];
                const float4 B = reinterpret_cast<const float4*>(&sh.cand[j])[1];
                const int4 K = reinterpret_cast<const int4*>(&sh.cand[j])[2];
                int zLo = 0, zHi = GT_Z;
                if (GENERAL && MODE == 0) { // the reference's integer support box (clips the kernel when sigma > 1)
                    const int bx = K.w;
                    if (lane < (bx & 63) || lane >= ((bx >> 6) & 63) || ty < ((bx >> 12) & 15) || ty >= ((bx >> 16) & 15)) continue;
                    zLo = (bx >> 20) & 15, zHi = (bx >> 24) & 15;
                }
                if (QSREF && (K.x < myXlo || K.x > myXhi)) continue; // not in the cell range of MY 8-voxel block along x
                float px = vx, py = vy;
                if (!QSREF && (K.x | K.y | K.z)) { // periodic image: the reference's un-wrapped voxel index (ParticlesToDensity.cpp:605-613)
                    px = __fadd_rn(__fmul_rn((float)(vxI + K.x), g.sd[0]), g.mn[0]);
                    py = __fadd_rn(__fmul_rn((float)(vyI + K.y), g.sd[1]), g.mn[1]);
                }
                const float dx = __fsub_rn(px, A.x), dy = __fsub_rn(py, A.y);
                const float dxy2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                float lim2 = MODE == 0 ? B.x * B.x * 1.0001f : B.x;
                if (!QSREF && K.z == 0) { // the tile's planes are at least gz away from the particle: a column further out than sqrt(lim2 - gz^2) has no hit
                    const float gz = fmaxf(fmaxf(vz[0] - A.z, A.z - vz[GT_Z - 1]), 0.0f);
                    lim2 -= gz * gz * 0.999f;
                }
                if (!(dxy2 < lim2)) continue;
#pragma unroll
                for (int k = 0; k < GT_Z; ++k) {
                    float pz = vz[k];
                    if (!QSREF && K.z) pz = __fadd_rn(__fmul_rn((float)(t0z + k + K.z), g.sd[2]), g.mn[2]);
                    const float dz = __fsub_rn(pz, A.z);
                    const float d2 = __fadd_rn(dxy2, __fmul_rn(dz, dz)); // (no per-plane skip: branches here cost the eight independent
                                                                         //  ex2 chains their overlap -- measured 28.5 vs 23.2 ms)
                    float w;
                    if (MODE == 1) {
                        // branch-free: outside the cut-off the weight is an exact 0, and x + 0 == x -- the same bits as skipping, without
                        // eight divergent regions per candidate
                        const float e = ex2Approx(__fmul_rn(d2, A.w));
                        w = d2 < B.x ? e : 0.0f;
                        acc[k] = __fadd_rn(acc[k], w);
                        if (COLOUR) {
                            accR[k] = __fadd_rn(accR[k], __fmul_rn(w, B.y));
                            accG[k] = __fadd_rn(accG[k], __fmul_rn(w, B.z));
                            accB[k] = __fadd_rn(accB[k], __fmul_rn(w, B.w));
                        }
                    } else if ((!GENERAL || (k >= zLo && k < zHi)) && kernelValue<MODE, (MODE == 0 && COLOUR)>(d2, MODE == 0 ? B.x : 0.0f, A.w, w, B.x)) {
                        if (MODE == 0 && COLOUR) { // aggregator 2: weights += w, vol += w * dir (:501-505), product and sum rounded separately
                            acc[k] = __fadd_rn(acc[k], w);
                            accR[k] = __fadd_rn(accR[k], __fmul_rn(w, B.y));
                            accG[k] = __fadd_rn(accG[k], __fmul_rn(w, B.z));
                            accB[k] = __fadd_rn(accB[k], __fmul_rn(w, B.w));
                        } else if (MODE == 0) {
                            acc[k] = __fadd_rn(acc[k], g.agg == 1 ? __fmul_rn(w, B.y) : w);
                        } else {
                            acc[k] = __fadd_rn(acc[k], w);
                            if (COLOUR) {
                                accR[k] = __fadd_rn(accR[k], __fmul_rn(w, B.y));
                                accG[k] = __fadd_rn(accG[k], __fmul_rn(w, B.z));
                                accB[k] = __fadd_rn(accB[k], __fmul_rn(w, B.w));
                            }
                        }
                    }
                }
            }
            } // window
        }
    }
    float vmin = INFINITY, vmax = -INFINITY;
    if (vxI <= t1x && vyI <= t1y) {
#pragma unroll
        for (int k = 0; k < GT_Z; ++k) {
            const int z = t0z + k;
            if (z > t1z) break;
            const size_t o = vxI + static_cast<size_t>(g.s[0]) * (vyI + static_cast<size_t>(g.s[1]) * (z - g.z0));
            vol[o] = acc[k];
            if (COLOUR) rgb[3 * o + 0] = accR[k], rgb[3 * o + 1] = accG[k], rgb[3 * o + 2] = accB[k];
            vmin = fminf(vmin, acc[k]), vmax = fmaxf(vmax, acc[k]);
        }
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if (lane == 0 && kmin <= kmax && !(MODE == 0 && COLOUR)) { // aggregator 2: the range is that of |v|, taken by vector_finalize_kernel
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// density_gauss_kernel: QuickSurf-Gaussian mode with the radial cut-off on a non-periodic grid (C3)
// ---------------------------------------------------------------------------------------------------------------
// A WARP owns an 8x4x8-voxel patch (a thread: a column of eight voxels in registers; four patches make the 32x4x8 strip of a block, but
// the warps never talk to each other: no block barrier, no shared candidate stage).  The warp streams, in ascending (cell z, cell y,
// cell x) and canonical in-cell order, the records of every cell row its patch's neighbourhood touches -- the x cells of a row are one
// contiguous record range --, 32 records per round: each lane turns its record into a candidate (cut-off^2, exponent scale, colour),
// tests its sphere against the patch's box, and the warp then walks the surviving candidates of the round in record order (ballot
// mask; candidate fields broadcast from a 1 KB per-warp stage).  With a cut-off of ~10 voxels nearly every lane of the warp lies inside
// the circle of a candidate that reaches the patch at all.  A candidate that is skipped would have added an exact 0 to every voxel of
// the patch, so every voxel receives its non-zero terms in the order (cell z, cell y, cell x, canonical in-cell order): the same
// bits as density_gather_kernel's, for every tile and z-slab decomposition.  Distances and weights are rounded operation by operation
// (the cut-off test d2 < lim decides whether a term exists: it must not depend on contraction); the colour sums use fused multiply-adds.
struct GaussCand {          // 64 bytes
    float x, y, z, k0;      // k0 = -log2e / (2 (r radscale)^2)
    float lim, cr, cg, cb;  // lim = cut-off^2
    float dz2[GT_Z];        // (z_k - z)^2 of the patch's eight planes: the same for every lane, computed once by the lane that stages the atom
};
constexpr int GP_X = 8, GP_Y = 4; // voxel columns of a warp's patch
constexpr int GQ_WARPS = 4;       // independent warps per block: a 32x4x8 strip of four patches (20 warps per SM at <= 102 registers)
constexpr int GQ_THREADS = GQ_WARPS * 32;
struct GaussShared {
    GaussCand cand[GQ_WARPS][32];
};

template<bool COLOUR>
__global__ void __launch_bounds__(GQ_THREADS, 5) density_gauss_kernel(Geo g, DevState* st, const float4* __restrict__ recs,
    const float* __restrict__ aux, int auxN, const unsigned* __restrict__ cellStart, float* __restrict__ vol, float* __restrict__ rgb,
    int reach) {
    __shared__ GaussShared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0x = (int)blockIdx.x * GT_X, t0y = (int)blockIdx.y * GP_Y, t0z = g.z0 + (int)blockIdx.z * GT_Z;
    const int t1x = min(t0x + GT_X, g.s[0]) - 1, t1y = min(t0y + GP_Y, g.s[1]) - 1, t1z = min(t0z + GT_Z, g.z0 + g.nz) - 1;
    const int p0x = t0x + warp * GP_X, p0y = t0y;
    const int p1x = min(p0x + GP_X - 1, t1x), p1y = min(p0y + GP_Y - 1, t1y);
    if (p0x > t1x || p0y > t1y) return; // patch beyond the grid
    GaussCand* stage = sh.cand[warp];

    const int vxI = p0x + (lane & 7), vyI = p0y + (lane >> 3);
    const float vx = __fadd_rn(__fmul_rn((float)vxI, g.sd[0]), g.mn[0]), vy = __fadd_rn(__fmul_rn((float)vyI, g.sd[1]), g.mn[1]);
    float vz[GT_Z], acc[GT_Z], accR[GT_Z], accG[GT_Z], accB[GT_Z];
#pragma unroll
    for (int k = 0; k < GT_Z; ++k) {
        vz[k] = __fadd_rn(__fmul_rn((float)(t0z + k), g.sd[2]), g.mn[2]);
        acc[k] = 0.0f, accR[k] = 0.0f, accG[k] = 0.0f, accB[k] = 0.0f;
    }
    // the patch's box (node positions) for the sphere / box rejection
    const float bx0 = __fadd_rn(__fmul_rn((float)p0x, g.sd[0]), g.mn[0]), bx1 = __fadd_rn(__fmul_rn((float)p1x, g.sd[0]), g.mn[0]);
    const float by0 = __fadd_rn(__fmul_rn((float)p0y, g.sd[1]), g.mn[1]), by1 = __fadd_rn(__fmul_rn((float)p1y, g.sd[1]), g.mn[1]);
    const float bz0 = vz[0], bz1 = __fadd_rn(__fmul_rn((float)t1z, g.sd[2]), g.mn[2]);

    // cells whose particles can reach the patch (non-periodic: one contiguous range per axis)
    const int cx0 = max(p0x - reach, 0) >> g.cshift, cx1 = min(p1x + reach, g.s[0] - 1) >> g.cshift;
    const int cy0 = max(p0y - reach, 0) >> g.cshift, cy1 = min(p1y + reach, g.s[1] - 1) >> g.cshift;
    const int cz0 = max(t0z - reach, 0) >> g.cshift, cz1 = min(t1z + reach, g.s[2] - 1) >> g.cshift;
    const int nyc = cy1 - cy0 + 1, nrows = nyc * (cz1 - cz0 + 1);

    for (int row0 = 0; row0 < nrows; row0 += 32) {
        // lane i owns cell row number row0 + i (ascending cell z, then cell y) and loads its record range
        unsigned myB = 0, myE = 0;
        if (row0 + lane < nrows) {
            const int i = row0 + lane, cz = cz0 + i / nyc, cy = cy0 + i % nyc;
            const int czl = cellZLocal(g, cz);
            const size_t rowBase = (static_cast<size_t>(czl) * g.nc[1] + cy) * g.nc[0];
            if (czl < g.czCount) myB = cellStart[rowBase + cx0], myE = cellStart[rowBase + cx1 + 1];
        }
        unsigned rest = __ballot_sync(0xffffffffu, myE > myB);
        unsigned base = 0, end = 0;
        auto nextChunk = [&]() { // -> false when the rows are exhausted
            base += 32;
            if (base >= end) {
                if (!rest) return false;
                const int r = __ffs(rest) - 1;
                rest &= rest - 1;
                base = __shfl_sync(0xffffffffu, myB, r), end = __shfl_sync(0xffffffffu, myE, r);
            }
            return true;
        };
        bool more = nextChunk();
        float4 pNext = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (more && base + lane < end) pNext = recs[base + lane];
        while (more) {
            const float4 p = pNext;
            const unsigned idx = base + lane;
            const bool have = idx < end;
            more = nextChunk(); // the records of the next round are requested before this one is worked on
            if (more && base + lane < end) pNext = recs[base + lane];
            bool rel = false;
            if (have) {
                const float sr = __fmul_rn(p.w, g.radscale);
                const float eps = __fmul_rn(g.gausslim, sr);
                const float lim = __fmul_rn(eps, eps);
                // squared distance from the atom to the patch's box, 0.1 % slack for the rounding of this test (the exact test is per voxel)
                const float gx = fmaxf(fmaxf(bx0 - p.x, p.x - bx1), 0.0f), gy = fmaxf(fmaxf(by0 - p.y, p.y - by1), 0.0f);
                const float gz = fmaxf(fmaxf(bz0 - p.z, p.z - bz1), 0.0f);
                rel = (gx * gx + gy * gy + gz * gz) * 0.999f < lim;
                if (rel) { // a few per cent of the stream: only these pay for the division, the colour fetch and the stage
                    float4 col = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
                    if (COLOUR && auxN == 4) col = reinterpret_cast<const float4*>(aux)[idx];
                    float4* sg = reinterpret_cast<float4*>(&stage[lane]);
                    sg[0] = make_float4(p.x, p.y, p.z, __fdiv_rn(-1.4426950408889634f, __fmul_rn(__fmul_rn(2.0f, sr), sr)));
                    sg[1] = make_float4(lim, col.x, col.y, col.z);
                    float q[GT_Z];
#pragma unroll
                    for (int k = 0; k < GT_Z; ++k) {
                        const float dz = __fsub_rn(vz[k], p.z);
                        q[k] = __fmul_rn(dz, dz);
                    }
                    sg[2] = make_float4(q[0], q[1], q[2], q[3]);
                    sg[3] = make_float4(q[4], q[5], q[6], q[7]);
                }
            }
            unsigned bits = __ballot_sync(0xffffffffu, rel);
            __syncwarp();
            while (bits) {
                const int j = __ffs(bits) - 1;
                bits &= bits - 1;
                const float4* sg = reinterpret_cast<const float4*>(&stage[j]);
                const float4 A = sg[0], B = sg[1], Q0 = sg[2], Q1 = sg[3];
                const float dz2[GT_Z] = {Q0.x, Q0.y, Q0.z, Q0.w, Q1.x, Q1.y, Q1.z, Q1.w};
                const float dx = __fsub_rn(vx, A.x), dy = __fsub_rn(vy, A.y);
                const float dxy2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
#pragma unroll
                for (int k = 0; k < GT_Z; ++k) {
                    const float d2 = __fadd_rn(dxy2, dz2[k]);
                    // outside the cut-off nothing is added (predicated accumulation: one compare, no select; the exponential is evaluated
                    // regardless so that the eight chains stay independent)
                    const float w = ex2Approx(__fmul_rn(d2, A.w));
                    if (d2 < B.x) {
                        acc[k] = __fadd_rn(acc[k], w);
                        if (COLOUR) accR[k] = fmaf(w, B.y, accR[k]), accG[k] = fmaf(w, B.z, accG[k]), accB[k] = fmaf(w, B.w, accB[k]);
                    }
                }
            }
            __syncwarp(); // the stage is rewritten in the next round
        }
    }
    float vmin = INFINITY, vmax = -INFINITY;
    if (vxI <= t1x && vyI <= t1y) {
#pragma unroll
        for (int k = 0; k < GT_Z; ++k) {
            const int z = t0z + k;
            if (z > t1z) break;
            const size_t o = vxI + static_cast<size_t>(g.s[0]) * (vyI + static_cast<size_t>(g.s[1]) * (z - g.z0));
            vol[o] = acc[k];
            if (COLOUR) rgb[3 * o + 0] = accR[k], rgb[3 * o + 1] = accG[k], rgb[3 * o + 2] = accB[k];
            vmin = fminf(vmin, acc[k]), vmax = fmaxf(vmax, acc[k]);
        }
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if (lane == 0 && kmin <= kmax) {
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

/**
 * Aggregator 2 (IVecToSingleCell_Volume), the per-voxel pass after the accumulation (ParticlesToDensity.cpp:634-657):
 *   v = sum(w d) / (sum(w) == 0 ? 1 : sum(w));  density = sqrt(vx vx + vy vy + vz vz);  direction = density == 0 ? 0 : v / density
 * and the range of the densities (minDens starts at FLT_MAX, maxDens at 0).  In: weights = sum(w), vec = sum(w d).
 * Out: vec = v, weights -> density, dir = direction.  Individually rounded operations, the reference's evaluation order.
 */
__global__ void __launch_bounds__(256) vector_finalize_kernel(float* __restrict__ weights, float* __restrict__ vec, float* __restrict__ dir,
    size_t nvox, DevState* __restrict__ st) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    float vmin = INFINITY, vmax = 0.0f;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvox; i += stride) {
        const float w = weights[i];
        const float div = w == 0.0f ? 1.0f : w;
        const float x = __fdiv_rn(vec[3 * i + 0], div), y = __fdiv_rn(vec[3 * i + 1], div), z = __fdiv_rn(vec[3 * i + 2], div);
        const float den = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        vec[3 * i + 0] = x, vec[3 * i + 1] = y, vec[3 * i + 2] = z;
        weights[i] = den;
        const bool zero = den == 0.0f;
        dir[3 * i + 0] = zero ? 0.0f : __fdiv_rn(x, den);
        dir[3 * i + 1] = zero ? 0.0f : __fdiv_rn(y, den);
        dir[3 * i + 2] = zero ? 0.0f : __fdiv_rn(z, den);
        vmin = fminf(vmin, den), vmax = fmaxf(vmax, den); // std::max / std::min drop a NaN density the same way
    }
    const unsigned kmin = __reduce_min_sync(0xffffffffu, floatKey(vmin)), kmax = __reduce_max_sync(0xffffffffu, floatKey(vmax));
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&st->minKey, kmin);
        atomicMax(&st->maxKey, kmax);
    }
}

__global__ void __launch_bounds__(256) normalize_kernel(float* __restrict__ vol, size_t n, float mn, float rcp) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        vol[i] = __fmul_rn(__fsub_rn(vol[i], mn), rcp);
}

/** Largest radius over a list with per-particle radii (decides the cell size on the host). */
__global__ void __launch_bounds__(256) radius_max_kernel(ListDev l, unsigned* rmaxBits) {
    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
    float rmax = 0.0f;
    const unsigned long long count = listCount(l);
    for (unsigned long long j = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; j < count; j += stride) {
        const float r = fetchParticle(l, j).w;
        if (r > 0.0f && isfinite(r)) rmax = fmaxf(rmax, r);
    }
    rmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(rmax)));
    if ((threadIdx.x & 31) == 0 && rmax > 0.0f) atomicMax(rmaxBits, __float_as_uint(rmax));
}

} // namespace mms
