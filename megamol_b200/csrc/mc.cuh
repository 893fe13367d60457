// mc.cuh -- marching cubes on the device-resident density slab.
//
//   mc_count_kernel   classify every cell (cube index bit i <=> corner i < iso, corner order and case table of
//                     trisoup::volumetrics::MarchingCubeTables, MarchingCubeTables.cpp:11-16,58,278) and sum the
//                     triangle counts of each 32-cell x-segment with a warp reduction
//   (exclusive scan of the segment counts: scan.cuh -> triangle offsets in CELL-LINEAR order)
//   mc_emit_kernel    active tiles only (32x8x2 cells):
//                       A  stage density + one-node halo in shared memory (coalesced rows)
//                       B  find the crossed grid edges of the tile, compact them (ballot), and compute each crossed
//                          edge's vertex ONCE with full lanes: interpolation parameter, position coordinate,
//                          gradient normal -> one float4 record per edge in shared memory
//                       C  per 32-cell row: classify, warp-level prefix scan of the per-cell triangle counts, then the
//                          row's triangle corners are FLATTENED over the lanes (corner j of the row -> lane j mod 32):
//                          each lane looks up the edge record and writes 3+3 floats; consecutive lanes write
//                          consecutive 12-byte pieces, so every warp store covers one contiguous span of the output.
// Output order = cell-linear (x fastest, then y, then z), inside a cell the table's order: independent of the
// tile shape and of the z-slab decomposition.
#pragma once
#include "common.cuh"

namespace mms {

__constant__ unsigned long long kCaseWords[256] = {
#include "mc_case_words.inc"
};

struct McGeo {
    int sx, sy;        // volume resolution in x, y
    int nzPlanes;      // planes in the slab volume
    int zPlane0;       // global z index of plane 0 of the volume
    int szGlobal;      // global z resolution (gradient clamps at the GLOBAL border only)
    int cx, cy;        // cells in x, y (= s-1)
    int cz0, cnz;      // cell layers [cz0, cz0+cnz) in GLOBAL z handled by this context
    int nsegx;         // ceil(cx / 32)
    float org[3], sd[3];
    float rinv[3][3];  // rinv[a][n] = 1/(n*sd[a]), n = 1, 2 (gradient: one-sided / central)
    float iso;
};

// ---------------------------------------------------------------------------------------------------------------
// count
// ---------------------------------------------------------------------------------------------------------------
constexpr int MC_THREADS = 256;
constexpr int MCC_ROWS = 8;  // cell rows (y) per warp of the count kernel (more rows = fewer redundant row loads but more registers)

__device__ __forceinline__ int cubeIndexSmem(const float* f, int strideY, int strideZ, float iso) {
    // f points at corner 0; corners: (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
    int ci = 0;
    ci |= (f[0] < iso) ? 1 : 0;
    ci |= (f[1] < iso) ? 2 : 0;
    ci |= (f[1 + strideY] < iso) ? 4 : 0;
    ci |= (f[strideY] < iso) ? 8 : 0;
    ci |= (f[strideZ] < iso) ? 16 : 0;
    ci |= (f[1 + strideZ] < iso) ? 32 : 0;
    ci |= (f[1 + strideY + strideZ] < iso) ? 64 : 0;
    ci |= (f[strideY + strideZ] < iso) ? 128 : 0;
    return ci;
}

constexpr int MCC_LAYERS = 16; // cell layers (z) a warp marches through

/**
 * Count kernel, register-rolling in z: a warp owns a strip of 32 cells (x) by MCC_ROWS rows (y) and marches through
 * MCC_LAYERS cell layers.  Per layer it loads ONE new node plane of the strip (MCC_ROWS+1 rows of 128 contiguous bytes, node
 * 32 of the segment by lane 0) and turns it into "below iso" bits; the bits of the previous plane stay in registers and the
 * x+1 neighbour comes from a shuffle -- so every density value is loaded ~1.1 times and a cell costs ~20 instructions.
 * Per-cell lookups go to a shared copy of the count table (the constant cache would serialise per-lane indices).
 */
__global__ void __launch_bounds__(MC_THREADS) mc_count_kernel(McGeo m, const float* __restrict__ vol, unsigned* __restrict__ segCount,
    unsigned char* __restrict__ triCount) {
    __shared__ unsigned char sCount[256];
    sCount[threadIdx.x] = static_cast<unsigned char>(kCaseWords[threadIdx.x] & 15ull);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xseg = blockIdx.x;
    const int yBeg = (blockIdx.y * (MC_THREADS / 32) + warp) * MCC_ROWS;
    const int czBeg = m.cz0 + blockIdx.z * MCC_LAYERS;
    if (yBeg >= m.cy || czBeg >= m.cz0 + m.cnz) return;
    const int yEnd = min(yBeg + MCC_ROWS, m.cy), czEnd = min(czBeg + MCC_LAYERS, m.cz0 + m.cnz);
    const int x = xseg * 32 + lane;
    const int xa = min(x, m.sx - 1), xb = min(xseg * 32 + 32, m.sx - 1);
    const size_t plane = static_cast<size_t>(m.sx) * m.sy;
    // bits of one node plane of the strip: bit0 = my node below iso, bit1 = node x+1 below iso
    auto planeBits = [&](int zNode, unsigned (&out)[MCC_ROWS + 1]) {
        const float* p = vol + plane * (zNode - m.zPlane0);
        float v[MCC_ROWS + 1], e[MCC_ROWS + 1];
#pragma unroll
        for (int r = 0; r <= MCC_ROWS; ++r) {
            const size_t o = static_cast<size_t>(m.sx) * min(yBeg + r, m.sy - 1);
            v[r] = p[o + xa];
            e[r] = lane == 0 ? p[o + xb] : 0.0f;
        }
#pragma unroll
        for (int r = 0; r <= MCC_ROWS; ++r) {
            const unsigned me = v[r] < m.iso ? 1u : 0u;
            unsigned nb = __shfl_down_sync(0xffffffffu, me, 1);
            const unsigned last = __shfl_sync(0xffffffffu, e[r] < m.iso ? 1u : 0u, 0);
            if (lane == 31) nb = last;
            out[r] = me | (nb << 1);
        }
    };
    unsigned lo[MCC_ROWS + 1], hi[MCC_ROWS + 1];
    planeBits(czBeg, lo);
    for (int cz = czBeg; cz < czEnd; ++cz) {
        planeBits(cz + 1, hi);
#pragma unroll
        for (int r = 0; r < MCC_ROWS; ++r) {
            const int y = yBeg + r;
            if (y < yEnd) {
                // corners: 0 (x,y,z) 1 (x+1,y,z) 2 (x+1,y+1,z) 3 (x,y+1,z) 4 (x,y,z+1) 5 (x+1,y,z+1) 6 (x+1,y+1,z+1) 7 (x,y+1,z+1)
                const unsigned ci = (lo[r] & 1u) | (lo[r] & 2u) | ((lo[r + 1] >> 1) & 1u) << 2 | (lo[r + 1] & 1u) << 3 | (hi[r] & 1u) << 4 |
                                    ((hi[r] >> 1) & 1u) << 5 | ((hi[r + 1] >> 1) & 1u) << 6 | (hi[r + 1] & 1u) << 7;
                unsigned n = 0;
                if (x < m.cx) {
                    n = sCount[ci];
                    if (triCount) triCount[x + static_cast<size_t>(m.cx) * (y + static_cast<size_t>(m.cy) * (cz - m.cz0))] = static_cast<unsigned char>(n);
                }
                const unsigned tot = __reduce_add_sync(0xffffffffu, n);
                if (lane == 0) segCount[xseg + static_cast<size_t>(m.nsegx) * (y + static_cast<size_t>(m.cy) * (cz - m.cz0))] = tot;
            }
        }
#pragma unroll
        for (int r = 0; r <= MCC_ROWS; ++r) lo[r] = hi[r];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// emit: z-marching, software-pipelined
// ---------------------------------------------------------------------------------------------------------------
// A block owns a 32x8-cell column and marches through EM_STEPS steps of EZ = 2 cell layers.  The density planes live
// in a 7-slot ring in shared memory (5 planes of the current step incl. the gradient halo + the 2 planes the next
// step adds); the next step's planes are fetched with cp.async while the current step computes, so the global-load
// latency is paid once per block instead of once per tile, and no plane is loaded twice inside a column chunk.
constexpr int EX = 32, EY = 8, EZ = 2;                      // cells per step
constexpr int ENX = EX + 1, ENY = EY + 1, ENZ = EZ + 1;     // nodes per step
constexpr int EHX = EX + 3, EHY = EY + 3, EHZ = EZ + 3;     // nodes + gradient halo
constexpr int EHXP = EHX + 1;                               // padded row
constexpr int ERING = 8;                                    // plane slots (>= EHZ + EZ = 7; power of two: slot = (z+1) & 7)
constexpr int EM_STEPS = 16;                                // steps per block (32 cell layers)
constexpr int E_XEDGES = ENZ * ENY * EX;                    // 864  x-edges: ix < 32
constexpr int E_YEDGES = ENZ * EY * ENX;                    // 792  y-edges: iy < 8
constexpr int E_ZEDGES = EZ * ENY * ENX;                    // 594  z-edges: iz < 2
constexpr int E_YBASE = E_XEDGES, E_ZBASE = E_XEDGES + E_YEDGES, E_EDGES = E_XEDGES + E_YEDGES + E_ZEDGES;
constexpr int E_MAXROWTRIS = 160;

struct McEmitShared {
    float4 edge[E_EDGES];               // {interpolated coordinate along the edge's axis, nx, ny, nz}
    float ring[ERING][EHY][EHXP];       // plane with global node index z sits in slot (z + 1) mod ERING
    unsigned short crossList[E_EDGES];
    unsigned char triOwner[MC_THREADS / 32][E_MAXROWTRIS];
    float tabX[ENX], tabY[ENY];         // node positions float(idx)*sd + origin (ParticlesToDensity.cpp:605)
    unsigned segOff[EM_STEPS * EZ * EY + 1][2]; // per row of the chunk: first triangle, triangle count
    int stepActive[EM_STEPS];
    int anyActive;
    int ncross;
};

// per cube edge: low corner (dx,dy,dz) and axis: dx | dy<<1 | dz<<2 | axis<<3   (MarchingCubeTables.cpp:15-16, low node first)
//  e0 (0,0,0)x  e1 (1,0,0)y  e2 (0,1,0)x  e3 (0,0,0)y  e4 (0,0,1)x  e5 (1,0,1)y  e6 (0,1,1)x  e7 (0,0,1)y
//  e8 (0,0,0)z  e9 (1,0,0)z  e10 (1,1,0)z e11 (0,1,0)z
__device__ __forceinline__ unsigned edgeCode(int e) {
    const unsigned long long codes = (0ull) | (9ull << 5) | (2ull << 10) | (8ull << 15) | (4ull << 20) | (13ull << 25) |
                                     (6ull << 30) | (12ull << 35) | (16ull << 40) | (17ull << 45) | (19ull << 50) | (18ull << 55);
    return static_cast<unsigned>(codes >> (5 * e)) & 31u;
}

__device__ __forceinline__ int edgeIndex(int axis, int ix, int iy, int iz) {
    if (axis == 0) return (iz * ENY + iy) * EX + ix;
    if (axis == 1) return E_YBASE + (iz * EY + iy) * ENX + ix;
    return E_ZBASE + (iz * ENY + iy) * ENX + ix;
}

__device__ __forceinline__ int ringSlot(int zNode) { // zNode >= -1
    return (zNode + 1) & (ERING - 1);
}

__device__ __forceinline__ void cpAsync4(float* smemDst, const float* gmemSrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smemDst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmemSrc));
}

template<bool COLOUR>
__global__ void __launch_bounds__(MC_THREADS) mc_emit_kernel(McGeo m, const float* __restrict__ vol, const float* __restrict__ rgb,
    const unsigned* __restrict__ segOffset, float* __restrict__ outPos, float* __restrict__ outNrm, float* __restrict__ outCol) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    McEmitShared& sh = *reinterpret_cast<McEmitShared*>(smemRaw);
    float4* edgeCol = reinterpret_cast<float4*>(smemRaw + ((sizeof(McEmitShared) + 15) & ~size_t(15))); // COLOUR only
    const int x0 = blockIdx.x * EX, y0 = blockIdx.y * EY;
    const int zcBeg = m.cz0 + blockIdx.z * (EM_STEPS * EZ);           // first global cell layer of this block
    const int zcEnd = min(zcBeg + EM_STEPS * EZ, m.cz0 + m.cnz);      // exclusive
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- which rows / steps have triangles (one global read per row of the chunk) -----------------------------------
    if (threadIdx.x < EM_STEPS) sh.stepActive[threadIdx.x] = 0;
    if (threadIdx.x == 0) sh.anyActive = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < EM_STEPS * EZ * EY; r += MC_THREADS) {
        const int ly = r % EY, lzc = r / EY;
        const int cyi = y0 + ly, czi = zcBeg + lzc;
        unsigned off = 0, cnt = 0;
        if (cyi < m.cy && czi < zcEnd) {
            const size_t seg = blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0));
            off = segOffset[seg];
            cnt = segOffset[seg + 1] - off;
        }
        sh.segOff[r][0] = off, sh.segOff[r][1] = cnt;
        if (cnt) sh.stepActive[lzc / EZ] = 1, sh.anyActive = 1;
    }
    __syncthreads();
    if (!sh.anyActive) return;

    // ---- plane loader: rows of one plane, clamped to the GLOBAL grid and to the slab -------------------------------------
    auto loadPlane = [&](int zNode) { // all threads; asynchronous
        const int zg = min(max(zNode, 0), m.szGlobal - 1);
        const int zl = min(max(zg - m.zPlane0, 0), m.nzPlanes - 1);
        float* dst = &sh.ring[ringSlot(zNode)][0][0];
        for (int i = threadIdx.x; i < EHY * EHXP; i += MC_THREADS) {
            const int ix = i % EHXP, iy = i / EHXP;
            if (ix >= EHX) continue;
            const int x = min(max(x0 + ix - 1, 0), m.sx - 1), y = min(max(y0 + iy - 1, 0), m.sy - 1);
            cpAsync4(dst + iy * EHXP + ix, vol + x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
        }
    };
    // first active step: its five planes; later steps add two planes each
    int step = 0;
    while (step < EM_STEPS && !sh.stepActive[step]) ++step;
    int loadedUpTo = zcBeg + step * EZ - 2; // highest node plane present in the ring (none yet)
    if (threadIdx.x < ENX) sh.tabX[threadIdx.x] = __fadd_rn(__fmul_rn((float)(x0 + threadIdx.x), m.sd[0]), m.org[0]);
    else if (threadIdx.x < ENX + ENY) sh.tabY[threadIdx.x - ENX] = __fadd_rn(__fmul_rn((float)(y0 + threadIdx.x - ENX), m.sd[1]), m.org[1]);

    const float r1x = m.rinv[0][1], r2x = m.rinv[0][2], r1y = m.rinv[1][1], r2y = m.rinv[1][2], r1z = m.rinv[2][1], r2z = m.rinv[2][2];
    unsigned char* owner = sh.triOwner[warp];

    for (; step < EM_STEPS; ++step) {
        const int zc0 = zcBeg + step * EZ; // global cell layer = global node plane of the step's lowest cells
        if (zc0 >= zcEnd) break;
        if (!sh.stepActive[step]) continue;
        // planes zc0-1 .. zc0+3 must be in the ring.  In the dense case the previous step prefetched them and nothing is issued
        // here; otherwise the ring slots about to be overwritten may still be read by warps finishing the previous step.
        if (loadedUpTo < zc0 + EZ + 1) {
            __syncthreads();
            for (int z = max(loadedUpTo + 1, zc0 - 1); z <= zc0 + EZ + 1; ++z) loadPlane(z);
            loadedUpTo = zc0 + EZ + 1;
            asm volatile("cp.async.commit_group;");
        }
        asm volatile("cp.async.wait_group 0;");
        if (threadIdx.x == 0) sh.ncross = 0;
        __syncthreads(); // (a) the planes have landed for everybody, (b) every warp has left the previous step: edge[], crossList and
                         //     the ring slots the prefetch below overwrites are free
        // prefetch the two planes the next step adds (if that step is active) while this one computes
        const bool nextActive = step + 1 < EM_STEPS && sh.stepActive[step + 1] && zc0 + EZ < zcEnd;
        if (nextActive) {
            loadPlane(zc0 + EZ + 2);
            loadPlane(zc0 + EZ + 3);
            loadedUpTo = zc0 + EZ + 3;
            asm volatile("cp.async.commit_group;");
        }
        // halo coordinates (node + 1): plane iz of the step = node plane zc0 - 1 + iz = ring slot (slot0 + iz) mod ERING
        const int slot0 = ringSlot(zc0 - 1);
        auto H = [&](int iz, int iy, int ix) -> float {
            return sh.ring[(slot0 + iz) & (ERING - 1)][iy][ix];
        };

        // ---- B1: crossed edges -> crossList (order is irrelevant) ------------------------------------------------
        // rounds 0..26: one warp per node row, lanes = nodes 0..31; round 27: the 27 nodes of column 32, one per lane
        for (int r = warp; r < ENZ * ENY + 1; r += MC_THREADS / 32) {
            const bool lastCol = r == ENZ * ENY;
            const int rr = lastCol ? lane : r;
            const bool active = !lastCol || lane < ENZ * ENY;
            const int iy = rr % ENY, iz = (rr / ENY) % ENZ;
            const int ix = lastCol ? EX : lane;
            bool cx = false, cy = false, cz = false;
            if (active) {
                const bool b0 = H(iz + 1, iy + 1, ix + 1) < m.iso;
                if (ix < EX) cx = b0 != (H(iz + 1, iy + 1, ix + 2) < m.iso);
                if (iy < EY) cy = b0 != (H(iz + 1, iy + 2, ix + 1) < m.iso);
                if (iz < EZ) cz = b0 != (H(iz + 2, iy + 1, ix + 1) < m.iso);
            }
            const unsigned bx = __ballot_sync(0xffffffffu, cx), by = __ballot_sync(0xffffffffu, cy), bz = __ballot_sync(0xffffffffu, cz);
            const int nx = __popc(bx), ny = __popc(by), nz = __popc(bz);
            if (nx + ny + nz == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(&sh.ncross, nx + ny + nz);
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned lt = (1u << lane) - 1u;
            if (cx) sh.crossList[base + __popc(bx & lt)] = static_cast<unsigned short>(edgeIndex(0, ix, iy, iz));
            if (cy) sh.crossList[base + nx + __popc(by & lt)] = static_cast<unsigned short>(edgeIndex(1, ix, iy, iz));
            if (cz) sh.crossList[base + nx + ny + __popc(bz & lt)] = static_cast<unsigned short>(edgeIndex(2, ix, iy, iz));
        }
        __syncthreads();

        // ---- B2: one vertex per crossed edge ------------------------------------------------------------------------
        const int ncross = sh.ncross;
        // no node of this step touches the global border -> plain central differences
        const bool interior = x0 > 0 && x0 + EX < m.sx - 1 && y0 > 0 && y0 + EY < m.sy - 1 && zc0 > 0 && zc0 + EZ < m.szGlobal - 1;
        for (int c = threadIdx.x; c < ncross; c += MC_THREADS) {
            const int id = sh.crossList[c];
            int axis, ix, iy, iz;
            if (id < E_YBASE) { axis = 0; ix = id % EX; const int t = id / EX; iy = t % ENY; iz = t / ENY; }
            else if (id < E_ZBASE) { axis = 1; const int q = id - E_YBASE; ix = q % ENX; const int t = q / ENX; iy = t % EY; iz = t / EY; }
            else { axis = 2; const int q = id - E_ZBASE; ix = q % ENX; const int t = q / ENX; iy = t % ENY; iz = t / ENY; }
            const int jx = ix + (axis == 0), jy = iy + (axis == 1), jz = iz + (axis == 2);
            // gradient at a node: (f(+) - f(-)) * 1/(n*sd), samples clamped at the GLOBAL grid border
            auto grad = [&](int nx_, int ny_, int nz_, float& gx, float& gy, float& gz) {
                const int hx = nx_ + 1, hy = ny_ + 1, hz = nz_ + 1;
                if (interior) {
                    gx = __fmul_rn(__fsub_rn(H(hz, hy, hx + 1), H(hz, hy, hx - 1)), r2x);
                    gy = __fmul_rn(__fsub_rn(H(hz, hy + 1, hx), H(hz, hy - 1, hx)), r2y);
                    gz = __fmul_rn(__fsub_rn(H(hz + 1, hy, hx), H(hz - 1, hy, hx)), r2z);
                } else {
                    const int gxi = x0 + nx_, gyi = y0 + ny_, gzi = zc0 + nz_;
                    const int xm = gxi > 0 ? -1 : 0, xp = gxi < m.sx - 1 ? 1 : 0;
                    const int ym = gyi > 0 ? -1 : 0, yp = gyi < m.sy - 1 ? 1 : 0;
                    const int zm = gzi > 0 ? -1 : 0, zp = gzi < m.szGlobal - 1 ? 1 : 0;
                    gx = xp > xm ? __fmul_rn(__fsub_rn(H(hz, hy, hx + xp), H(hz, hy, hx + xm)), xp - xm == 2 ? r2x : r1x) : 0.0f;
                    gy = yp > ym ? __fmul_rn(__fsub_rn(H(hz, hy + yp, hx), H(hz, hy + ym, hx)), yp - ym == 2 ? r2y : r1y) : 0.0f;
                    gz = zp > zm ? __fmul_rn(__fsub_rn(H(hz + zp, hy, hx), H(hz + zm, hy, hx)), zp - zm == 2 ? r2z : r1z) : 0.0f;
                }
            };
            const float fa = H(iz + 1, iy + 1, ix + 1), fb = H(jz + 1, jy + 1, jx + 1);
            const float t01 = __fdiv_rn(__fsub_rn(m.iso, fa), __fsub_rn(fb, fa));
            const float pza = __fadd_rn(__fmul_rn((float)(zc0 + iz), m.sd[2]), m.org[2]);
            const float pzb = __fadd_rn(__fmul_rn((float)(zc0 + jz), m.sd[2]), m.org[2]);
            const float pa = axis == 0 ? sh.tabX[ix] : (axis == 1 ? sh.tabY[iy] : pza);
            const float pb = axis == 0 ? sh.tabX[jx] : (axis == 1 ? sh.tabY[jy] : pzb);
            float gax, gay, gaz, gbx, gby, gbz;
            grad(ix, iy, iz, gax, gay, gaz);
            grad(jx, jy, jz, gbx, gby, gbz);
            const float gx = __fadd_rn(gax, __fmul_rn(t01, __fsub_rn(gbx, gax)));
            const float gy = __fadd_rn(gay, __fmul_rn(t01, __fsub_rn(gby, gay)));
            const float gz = __fadd_rn(gaz, __fmul_rn(t01, __fsub_rn(gbz, gaz)));
            const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
            const float inv = len2 > 0.0f ? -rsqrtf(len2) : 0.0f; // SFU rsqrt: 2 ulp, normals are compared at 1e-4
            sh.edge[id] = make_float4(__fadd_rn(pa, __fmul_rn(t01, __fsub_rn(pb, pa))), __fmul_rn(gx, inv), __fmul_rn(gy, inv), __fmul_rn(gz, inv));
            if (COLOUR) { // node colour = rgb / rho (0 where rho == 0), interpolated with the same t
                auto nodeColour = [&](int nx_, int ny_, int nz_, float f, float& r, float& gg, float& b) {
                    const int x = min(x0 + nx_, m.sx - 1), y = min(y0 + ny_, m.sy - 1);
                    const int zl = min(max(zc0 + nz_ - m.zPlane0, 0), m.nzPlanes - 1);
                    const float* c = rgb + 3 * (x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
                    if (f > 0.0f) r = __fdiv_rn(c[0], f), gg = __fdiv_rn(c[1], f), b = __fdiv_rn(c[2], f);
                    else r = gg = b = 0.0f;
                };
                float ar, ag, ab, br, bg, bb;
                nodeColour(ix, iy, iz, fa, ar, ag, ab);
                nodeColour(jx, jy, jz, fb, br, bg, bb);
                edgeCol[id] = make_float4(__fadd_rn(ar, __fmul_rn(t01, __fsub_rn(br, ar))), __fadd_rn(ag, __fmul_rn(t01, __fsub_rn(bg, ag))),
                    __fadd_rn(ab, __fmul_rn(t01, __fsub_rn(bb, ab))), 0.0f);
            }
        }
        __syncthreads();

        // ---- C: triangles, one warp per 32-cell row ---------------------------------------------------------------------
        for (int r = warp; r < EY * EZ; r += MC_THREADS / 32) {
            const int ly = r % EY, lz = r / EY;
            const int cxi = x0 + lane;
            const int row = (step * EZ + lz) * EY + ly;
            const unsigned segOff = sh.segOff[row][0], segTris = sh.segOff[row][1];
            if (segTris == 0) continue;
            unsigned long long word = 0;
            if (cxi < m.cx) {
                int ci = 0;
                ci |= (H(lz + 1, ly + 1, lane + 1) < m.iso) ? 1 : 0;
                ci |= (H(lz + 1, ly + 1, lane + 2) < m.iso) ? 2 : 0;
                ci |= (H(lz + 1, ly + 2, lane + 2) < m.iso) ? 4 : 0;
                ci |= (H(lz + 1, ly + 2, lane + 1) < m.iso) ? 8 : 0;
                ci |= (H(lz + 2, ly + 1, lane + 1) < m.iso) ? 16 : 0;
                ci |= (H(lz + 2, ly + 1, lane + 2) < m.iso) ? 32 : 0;
                ci |= (H(lz + 2, ly + 2, lane + 2) < m.iso) ? 64 : 0;
                ci |= (H(lz + 2, ly + 2, lane + 1) < m.iso) ? 128 : 0;
                word = kCaseWords[ci];
            }
            const unsigned n = static_cast<unsigned>(word & 15ull);
            unsigned inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned first = inc - n; // my first triangle within the row
            for (unsigned k = 0; k < n; ++k) owner[first + k] = static_cast<unsigned char>(lane);
            __syncwarp();
            const unsigned wlo = static_cast<unsigned>(word >> 4), whi = static_cast<unsigned>(word >> 36); // 15 nibbles of edge ids
            const unsigned ncorn = segTris * 3;
            const size_t gbase = static_cast<size_t>(segOff) * 9;
            const float tz0 = __fadd_rn(__fmul_rn((float)(zc0 + lz), m.sd[2]), m.org[2]);
            const float tz1 = __fadd_rn(__fmul_rn((float)(zc0 + lz + 1), m.sd[2]), m.org[2]);
            for (unsigned j0 = 0; j0 < ncorn; j0 += 32) {
                const unsigned j = j0 + lane;
                const bool act = j < ncorn;
                const unsigned t = act ? j / 3 : 0;
                const unsigned L = owner[t];
                const unsigned oFirst = __shfl_sync(0xffffffffu, first, L);
                const unsigned oLo = __shfl_sync(0xffffffffu, wlo, L), oHi = __shfl_sync(0xffffffffu, whi, L);
                if (!act) continue;
                const unsigned slotc = 3 * (t - oFirst) + (j - 3 * t); // corner number inside the owner cell (0..14)
                const int e = static_cast<int>(slotc < 8 ? (oLo >> (4 * slotc)) & 15u : (oHi >> (4 * (slotc - 8))) & 15u);
                const unsigned code = edgeCode(e);
                const int dz = (code >> 2) & 1;
                const int ix = (int)L + (code & 1), iy = ly + ((code >> 1) & 1), iz = lz + dz;
                const int axis = code >> 3;
                const int eidx = edgeIndex(axis, ix, iy, iz);
                const float4 v = sh.edge[eidx];
                const float px = axis == 0 ? v.x : sh.tabX[ix];
                const float py = axis == 1 ? v.x : sh.tabY[iy];
                const float pz = axis == 2 ? v.x : (dz ? tz1 : tz0);
                float* op = outPos + gbase + static_cast<size_t>(j) * 3;
                float* on = outNrm + gbase + static_cast<size_t>(j) * 3;
                op[0] = px, op[1] = py, op[2] = pz;
                on[0] = v.y, on[1] = v.z, on[2] = v.w;
                if (COLOUR) {
                    const float4 cc = edgeCol[eidx];
                    float* oc = outCol + gbase + static_cast<size_t>(j) * 3;
                    oc[0] = cc.x, oc[1] = cc.y, oc[2] = cc.z;
                }
            }
            __syncwarp();
        }
    }
    asm volatile("cp.async.wait_group 0;");
}

} // namespace mms
