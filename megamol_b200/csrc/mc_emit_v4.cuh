// The round-1 emit kernel (two cell layers per step, four block barriers per step), kept as an independent implementation that
// tests/test_gpu_variants.py runs against mc_emit_kernel bit for bit (MMS_EMIT_V4=1; a debug knob, not part of the C ABI).
#pragma once
#include "mc.cuh"
namespace mms {
namespace v4 {
// ---------------------------------------------------------------------------------------------------------------
// emit: z-marching, software-pipelined
// ---------------------------------------------------------------------------------------------------------------
constexpr int EX = 32, EY = 8, EZ = 2;                      // cells per step
constexpr int ENX = EX + 1, ENY = EY + 1, ENZ = EZ + 1;     // nodes per step
constexpr int EHY = EY + 3;                                 // node rows + gradient halo (y0-1 .. y0+9)
constexpr int EPITCH = 40;                                  // x0-4 .. x0+35 (x0-1 .. x0+33 are needed): TMA wants the box start and width in
                                                            // multiples of 16 bytes (an unaligned start coordinate is an illegal instruction)
constexpr int EHX0 = 4;                                     // ring column of node 0
constexpr int EPLANE = 448;                                 // floats per ring slot: 11*40 = 440, padded to 14*128 bytes
constexpr int ERING = 8;                                    // plane slots (5 of the current step + 2 prefetched; power of two)
constexpr int EM_STEPS = 16;                                // steps per block (32 cell layers)
constexpr int E_XEDGES = ENZ * ENY * EX;                    // 864  x-edges: (plane*9 + row)*32 + ix,      ix < 32
constexpr int E_YEDGES = ENZ * EY * ENX;                    // 792  y-edges: (plane*8 + row)*33 + ix,      row < 8
constexpr int E_ZEDGES = EZ * ENY * ENX;                    // 594  z-edges: (plane*9 + row)*33 + ix,      plane < 2
constexpr int E_YBASE = E_XEDGES, E_ZBASE = E_XEDGES + E_YEDGES, E_EDGES = E_XEDGES + E_YEDGES + E_ZEDGES;
constexpr int E_MAXROWTRIS = 160;
constexpr int E_ROWS = EM_STEPS * EZ * EY;                  // 256 cell rows per block chunk
constexpr int E_TAB_Y = ENX, E_TAB_Z = ENX + ENY;           // node position table: 33 x, 9 y, 3 z (z per step)
constexpr unsigned E_PLANE_BYTES = EHY * EPITCH * 4;        // 1584

struct __align__(128) McEmitV4Shared {
    float ring[ERING * EPLANE];                 // plane with node index z sits in slot (z - zcBeg + 1) mod ERING      14336 B
    float4 edge[E_EDGES];                       // {interpolated coordinate along the edge's axis, nx, ny, nz}          36000 B
    union {
        unsigned short crossList[E_EDGES];      // phases X, V
        unsigned char triOwner[MC_THREADS / 32][E_MAXROWTRIS]; // phase C: (owner lane << 3) | triangle number inside the owner cell
    };
    unsigned char segCnt[E_ROWS];               // its triangle count (<= 160)
    unsigned edgeTab[EZ * EY][12];              // per (layer, row) of a step and cube edge: edge slot of cell 0 | flags
    float tab[E_TAB_Z + ENZ + 3];               // node positions float(idx)*sd + origin (ParticlesToDensity.cpp:605)
    unsigned below[ENZ * ENY];                  // "below iso" bits of the step's node rows, nodes 0..31
    unsigned col32;                             // ... of node 32 of every row (bit plane*9 + row)
    unsigned stepAct[MC_THREADS / 32];
    int ncross;
    unsigned long long mbar;
};
static_assert(sizeof(McEmitV4Shared) + 128 <= 57088, "mc_emit_v4_kernel must fit four blocks per SM");

/**
 * TMA: the planes arrive as 40x11x1 boxes of a 3-D tensor map over the slab volume (needs sx % 4 == 0; out-of-range elements are
 * zero-filled, which is harmless: gradients at the global border are one-sided and cells beyond the grid are masked).
 * !TMA: the same ring filled with 4-byte cp.async copies, coordinates clamped.
 */
template<bool COLOUR, bool TMA>
__global__ void __launch_bounds__(MC_THREADS) mc_emit_v4_kernel(McGeo m, const __grid_constant__ CUtensorMap volMap, const float* __restrict__ vol,
    const float* __restrict__ rgb, const unsigned* __restrict__ segOffset, float* __restrict__ outPos, float* __restrict__ outNrm,
    float* __restrict__ outCol) {
    extern __shared__ unsigned char smemRaw[];
    unsigned char* smemAligned = smemRaw + ((128u - (smemAddr(smemRaw) & 127u)) & 127u);
    McEmitV4Shared& sh = *reinterpret_cast<McEmitV4Shared*>(smemAligned);
    float4* edgeCol = reinterpret_cast<float4*>(smemAligned + sizeof(McEmitV4Shared)); // COLOUR only
    const int x0 = blockIdx.x * EX, y0 = blockIdx.y * EY;
    const int zcBeg = m.cz0 + blockIdx.z * (EM_STEPS * EZ);           // first global cell layer of this block
    const int zcEnd = min(zcBeg + EM_STEPS * EZ, m.cz0 + m.cnz);      // exclusive
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- which rows / steps have triangles (one global read per row of the chunk) -----------------------------------
    {
        const int ly = tid % EY, lzc = tid / EY; // E_ROWS == MC_THREADS
        const int cyi = y0 + ly, czi = zcBeg + lzc;
        unsigned off = 0, cnt = 0;
        if (cyi < m.cy && czi < zcEnd) {
            const size_t seg = blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0));
            off = segOffset[seg];
            cnt = segOffset[seg + 1] - off;
        }
        sh.segCnt[tid] = static_cast<unsigned char>(cnt);
        const unsigned bal = __ballot_sync(0xffffffffu, cnt != 0); // rows 16s .. 16s+15 belong to step s
        if (lane == 0) sh.stepAct[warp] = ((bal & 0xffffu) ? 1u : 0u) | ((bal >> 16) ? 2u : 0u);
    }
    if (tid < EZ * EY * 12) {
        const int row = tid / 12, e = tid % 12, ly = row % EY, lz = row / EY;
        const unsigned code = edgeCode(e);
        const int dx = code & 1, dy = (code >> 1) & 1, dz = (code >> 2) & 1, axis = code >> 3;
        unsigned slot;
        if (axis == 0) slot = ((lz + dz) * ENY + ly + dy) * EX + dx;
        else if (axis == 1) slot = E_YBASE + ((lz + dz) * EY + ly + dy) * ENX + dx;
        else slot = E_ZBASE + ((lz + dz) * ENY + ly + dy) * ENX + dx;
        sh.edgeTab[row][e] = slot | (ET_AX0 << axis) | (dx ? ET_DX : 0u) | (dy ? ET_DY : 0u) | (dz ? ET_DZ : 0u);
    }
    if (tid < ENX) sh.tab[tid] = __fadd_rn(__fmul_rn((float)(x0 + tid), m.sd[0]), m.org[0]);
    else if (tid < ENX + ENY) sh.tab[tid] = __fadd_rn(__fmul_rn((float)(y0 + tid - ENX), m.sd[1]), m.org[1]);
    if (TMA && tid == 0) mbarInit(&sh.mbar, 1);
    __syncthreads();
    unsigned stepMask = 0;
#pragma unroll
    for (int w = 0; w < MC_THREADS / 32; ++w) stepMask |= sh.stepAct[w] << (2 * w);
    if (!stepMask) return;

    // ---- plane loader -------------------------------------------------------------------------------------------------
    auto slotOf = [&](int zNode) { return (zNode - zcBeg + 1) & (ERING - 1); }; // zNode >= zcBeg - 1
    unsigned tmaParity = 0;
    auto loadPlanes = [&](int zFirst, int zLast) { // uniform; planes zFirst..zLast (node indices), asynchronous
        if (TMA) {
            if (tid == 0) {
                mbarExpectTx(&sh.mbar, static_cast<unsigned>(zLast - zFirst + 1) * E_PLANE_BYTES);
                for (int z = zFirst; z <= zLast; ++z)
                    tmaLoadPlane(&sh.ring[slotOf(z) * EPLANE], &volMap, x0 - EHX0, y0 - 1, z - m.zPlane0, &sh.mbar);
            }
        } else {
            for (int z = zFirst; z <= zLast; ++z) {
                const int zg = min(max(z, 0), m.szGlobal - 1);
                const int zl = min(max(zg - m.zPlane0, 0), m.nzPlanes - 1);
                float* dst = &sh.ring[slotOf(z) * EPLANE];
                for (int i = tid; i < EHY * (ENX + 2); i += MC_THREADS) { // nodes -1 .. 33 of every row
                    const int ix = i % (ENX + 2) - 1, iy = i / (ENX + 2);
                    const int x = min(max(x0 + ix, 0), m.sx - 1), y = min(max(y0 + iy - 1, 0), m.sy - 1);
                    cpAsync4(dst + iy * EPITCH + ix + EHX0, vol + x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
                }
            }
            asm volatile("cp.async.commit_group;");
        }
    };
    auto waitPlanes = [&]() {
        if (TMA) {
            mbarWait(&sh.mbar, tmaParity);
            tmaParity ^= 1u;
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
    };

    const float r1x = m.rinv[0][1], r2x = m.rinv[0][2], r1y = m.rinv[1][1], r2y = m.rinv[1][2], r1z = m.rinv[2][1], r2z = m.rinv[2][2];
    const float iso = m.iso;
    // validity of the block's nodes (TMA zero-fills beyond the grid; clamped copies would be harmless, zeros are not)
    const int nvx = min(ENX, m.sx - x0), nvy = min(ENY, m.sy - y0);
    const unsigned nodeValidX = nvx >= 32 ? 0xffffffffu : (1u << nvx) - 1u;            // nodes 0..31
    const unsigned xEdgeValid = nvx >= 33 ? 0xffffffffu : (1u << (nvx - 1)) - 1u;      // x-edge ix needs node ix+1
    const unsigned cellValidX = m.cx - x0 >= 32 ? 0xffffffffu : (1u << (m.cx - x0)) - 1u;
    unsigned char* owner = sh.triOwner[warp];

    int loadedUpTo = -0x40000000; // highest node plane present in (or on its way into) the ring
    bool pending = false;         // a load batch has been issued and not yet waited for
    while (stepMask) {
        const int step = __ffs(stepMask) - 1;
        stepMask &= stepMask - 1;
        const int zc0 = zcBeg + step * EZ; // global cell layer = global node plane of the step's lowest cells
        // planes zc0-1 .. zc0+3 must be in the ring.  In the dense case the previous step prefetched the two new ones.
        if (loadedUpTo < zc0 + EZ + 1) {
            if (pending) waitPlanes(), pending = false;
            __syncthreads(); // every warp has left phase V of the previous step: no ring slot is being read any more
            loadPlanes(max(loadedUpTo + 1, zc0 - 1), zc0 + EZ + 1);
            loadedUpTo = zc0 + EZ + 1;
            pending = true;
        }
        if (pending) waitPlanes(), pending = false;
        __syncthreads(); // S1: the planes have landed for everybody; every warp has left phase C of the previous step
        if (stepMask && (__ffs(stepMask) - 1) == step + 1) { // prefetch the two planes the next step adds while this one computes
            loadPlanes(zc0 + EZ + 2, zc0 + EZ + 3);
            loadedUpTo = zc0 + EZ + 3;
            pending = true;
        }
        const int slot0 = (step * EZ) & (ERING - 1);            // slot of halo plane 0 = node plane zc0 - 1
        auto planeOff = [&](int hz) { return ((slot0 + hz) & (ERING - 1)) * EPLANE; }; // hz = node plane + 1 (halo coordinates)
        const int nvz = min(ENZ, zcEnd - zc0 + 1);              // valid node planes of the step (cell layers beyond zcEnd are not ours)

        // ---- M: bit masks ------------------------------------------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int q = warp + 8 * k; // node row: plane q / 9, row q % 9
            if (q < ENZ * ENY) {
                const int p = (q * 57) >> 9, r = q - 9 * p;
                const unsigned b = __ballot_sync(0xffffffffu, sh.ring[planeOff(p + 1) + (r + 1) * EPITCH + lane + EHX0] < iso);
                if (lane == 0) sh.below[q] = b;
            }
        }
        if (warp == 7) {
            const int q = min(lane, ENZ * ENY - 1);
            const int p = (q * 57) >> 9, r = q - 9 * p;
            const unsigned b = __ballot_sync(0xffffffffu, lane < ENZ * ENY && sh.ring[planeOff(p + 1) + (r + 1) * EPITCH + EX + EHX0] < iso);
            if (lane == 0) sh.col32 = nvx >= ENX ? b : 0u, sh.ncross = 0;
        }
        if (warp == 6 && lane < ENZ) sh.tab[E_TAB_Z + lane] = __fadd_rn(__fmul_rn((float)(zc0 + lane), m.sd[2]), m.org[2]);
        __syncthreads(); // S2

        // ---- X: crossed edges -> crossList ----------------------------------------------------------------------------------
        if (tid < 96) {
            const int g = tid;
            unsigned mask = 0, base = 0, stride = 1;
            const unsigned c32 = sh.col32;
            if (g < 27) {           // x-edges of node row g = p*9 + r
                const int p = (g * 57) >> 9, r = g - 9 * p;
                const unsigned b = sh.below[g];
                if (r < nvy && p < nvz) mask = (b ^ ((b >> 1) | (((c32 >> g) & 1u) << 31))) & xEdgeValid;
                base = g * EX;
            } else if (g < 51) {    // y-edges between node rows r and r+1 of plane p, j = p*8 + r
                const int j = g - 27, p = j >> 3, r = j & 7;
                if (r + 1 < nvy && p < nvz) mask = (sh.below[p * ENY + r] ^ sh.below[p * ENY + r + 1]) & nodeValidX;
                base = E_YBASE + j * ENX;
            } else if (g < 69) {    // z-edges between planes p and p+1, j = p*9 + r
                const int j = g - 51, p = j >= ENY ? 1 : 0, r = j - ENY * p;
                if (r < nvy && p + 1 < nvz) mask = (sh.below[j] ^ sh.below[j + ENY]) & nodeValidX;
                base = E_ZBASE + j * ENX;
            } else if (g == 69) {   // y-edges of node column 32: bit p*8 + r
                const unsigned rows = (1u << (nvy - 1)) - 1u; // r + 1 < nvy
#pragma unroll
                for (int p = 0; p < ENZ; ++p)
                    if (p < nvz) mask |= (((c32 >> (ENY * p)) ^ (c32 >> (ENY * p + 1))) & 0xffu & rows) << (8 * p);
                if (nvx < ENX) mask = 0;
                base = E_YBASE + EX, stride = ENX;
            } else if (g == 70) {   // z-edges of node column 32: bit p*9 + r
                const unsigned rows = (1u << nvy) - 1u;
                mask = (c32 ^ (c32 >> ENY)) & (rows | (nvz > 2 ? rows << ENY : 0u));
                if (nvx < ENX || nvz < 2) mask = 0;
                base = E_ZBASE + EX, stride = ENX;
            }
            const unsigned cnt = __popc(mask);
            unsigned inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            unsigned wbase = 0;
            if (lane == 31 && inc) wbase = atomicAdd(&sh.ncross, static_cast<int>(inc));
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            unsigned pos = wbase + inc - cnt;
            while (mask) {
                const int b = __ffs(mask) - 1;
                mask &= mask - 1;
                sh.crossList[pos++] = static_cast<unsigned short>(base + b * stride);
            }
        }
        __syncthreads(); // S3

        // ---- V: one vertex per crossed edge --------------------------------------------------------------------------------
        const int ncross = sh.ncross;
        // no node of this step touches the global border -> plain central differences
        const bool interior = x0 > 0 && x0 + EX < m.sx - 1 && y0 > 0 && y0 + EY < m.sy - 1 && zc0 > 0 && zc0 + EZ < m.szGlobal - 1;
        for (int c = tid; c < ncross; c += MC_THREADS) {
            const int id = sh.crossList[c];
            int axis, ix, r, p;
            if (id < E_YBASE) { axis = 0; ix = id & 31; const int q = id >> 5; p = (q * 57) >> 9; r = q - 9 * p; }
            else if (id < E_ZBASE) { axis = 1; const int t = id - E_YBASE; const int q = (t * 993) >> 15; ix = t - ENX * q; p = q >> 3; r = q & 7; }
            else { axis = 2; const int t = id - E_ZBASE; const int q = (t * 993) >> 15; ix = t - ENX * q; p = q >= ENY ? 1 : 0; r = q - ENY * p; }
            const int offA = (r + 1) * EPITCH + ix + EHX0;
            const int offB = offA + (axis == 0 ? 1 : (axis == 1 ? EPITCH : 0));
            const int hzA = p + 1, hzB = hzA + (axis == 2 ? 1 : 0);
            // gradient at a node: (f(+) - f(-)) * 1/(n*sd), samples clamped at the GLOBAL grid border
            auto grad = [&](int off, int hz, int gxi, int gyi, int gzi, float& gx, float& gy, float& gz) {
                const float* P0 = sh.ring + planeOff(hz) + off;
                const float* Pm = sh.ring + planeOff(hz - 1) + off;
                const float* Pp = sh.ring + planeOff(hz + 1) + off;
                if (interior) {
                    gx = __fmul_rn(__fsub_rn(P0[1], P0[-1]), r2x);
                    gy = __fmul_rn(__fsub_rn(P0[EPITCH], P0[-EPITCH]), r2y);
                    gz = __fmul_rn(__fsub_rn(Pp[0], Pm[0]), r2z);
                } else {
                    const int xm = gxi > 0 ? 1 : 0, xp = gxi < m.sx - 1 ? 1 : 0;
                    const int ym = gyi > 0 ? 1 : 0, yp = gyi < m.sy - 1 ? 1 : 0;
                    const int zm = gzi > 0 ? 1 : 0, zp = gzi < m.szGlobal - 1 ? 1 : 0;
                    gx = xp + xm ? __fmul_rn(__fsub_rn(P0[xp], P0[-xm]), xp + xm == 2 ? r2x : r1x) : 0.0f;
                    gy = yp + ym ? __fmul_rn(__fsub_rn(P0[yp * EPITCH], P0[-ym * EPITCH]), yp + ym == 2 ? r2y : r1y) : 0.0f;
                    gz = zp + zm ? __fmul_rn(__fsub_rn(zp ? Pp[0] : P0[0], zm ? Pm[0] : P0[0]), zp + zm == 2 ? r2z : r1z) : 0.0f;
                }
            };
            const float fa = sh.ring[planeOff(hzA) + offA], fb = sh.ring[planeOff(hzB) + offB];
            // t = (iso - fa) / (fb - fa): SFU reciprocal (2 ulp; vertices are compared at 1e-4 of a cell); the IEEE division only
            // where the difference is too small for rcp.approx
            const float tnum = __fsub_rn(iso, fa), tden = __fsub_rn(fb, fa);
            const float t01 = fabsf(tden) > 1e-30f ? __fmul_rn(tnum, rcpApproxF(tden)) : __fdiv_rn(tnum, tden);
            const int ti = axis == 0 ? ix : (axis == 1 ? E_TAB_Y + r : E_TAB_Z + p);
            const float pa = sh.tab[ti], pb = sh.tab[ti + 1];
            float gax, gay, gaz, gbx, gby, gbz;
            const int gxi = x0 + ix, gyi = y0 + r, gzi = zc0 + p;
            grad(offA, hzA, gxi, gyi, gzi, gax, gay, gaz);
            grad(offB, hzB, gxi + (axis == 0), gyi + (axis == 1), gzi + (axis == 2), gbx, gby, gbz);
            const float gx = __fadd_rn(gax, __fmul_rn(t01, __fsub_rn(gbx, gax)));
            const float gy = __fadd_rn(gay, __fmul_rn(t01, __fsub_rn(gby, gay)));
            const float gz = __fadd_rn(gaz, __fmul_rn(t01, __fsub_rn(gbz, gaz)));
            const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
            const float inv = len2 > 0.0f ? -rsqrtf(len2) : 0.0f; // SFU rsqrt: 2 ulp, normals are compared at 1e-4
            sh.edge[id] = make_float4(__fadd_rn(pa, __fmul_rn(t01, __fsub_rn(pb, pa))), __fmul_rn(gx, inv), __fmul_rn(gy, inv), __fmul_rn(gz, inv));
            if (COLOUR) { // node colour = rgb / rho (0 where rho == 0), interpolated with the same t
                auto nodeColour = [&](int nx_, int ny_, int nz_, float f, float& cr, float& cg, float& cb) {
                    const int x = min(x0 + nx_, m.sx - 1), y = min(y0 + ny_, m.sy - 1);
                    const int zl = min(max(zc0 + nz_ - m.zPlane0, 0), m.nzPlanes - 1);
                    const float* cc = rgb + 3 * (x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
                    if (f > 0.0f) cr = __fdiv_rn(cc[0], f), cg = __fdiv_rn(cc[1], f), cb = __fdiv_rn(cc[2], f);
                    else cr = cg = cb = 0.0f;
                };
                float ar, ag, ab, br, bg, bb;
                nodeColour(ix, r, p, fa, ar, ag, ab);
                nodeColour(ix + (axis == 0), r + (axis == 1), p + (axis == 2), fb, br, bg, bb);
                edgeCol[id] = make_float4(__fadd_rn(ar, __fmul_rn(t01, __fsub_rn(br, ar))), __fadd_rn(ag, __fmul_rn(t01, __fsub_rn(bg, ag))),
                    __fadd_rn(ab, __fmul_rn(t01, __fsub_rn(bb, ab))), 0.0f);
            }
        }
        __syncthreads(); // S4

        // ---- C: triangles, one warp per 32-cell row ---------------------------------------------------------------------
#pragma unroll 1
        for (int rr = warp; rr < EY * EZ; rr += MC_THREADS / 32) {
            const int ly = rr % EY, lz = rr / EY;
            const int row = (step * EZ + lz) * EY + ly;
            const unsigned segTris = sh.segCnt[row];
            if (segTris == 0) continue;
            const unsigned segOff = segOffset[blockIdx.x + static_cast<size_t>(m.nsegx) * (y0 + ly + static_cast<size_t>(m.cy) * (zc0 + lz - m.cz0))];
            // cube index in permuted order from the (x, x+1) bit pairs of the four node rows
            const int q = lz * ENY + ly;
            const unsigned c32 = sh.col32;
            const unsigned p00 = __funnelshift_r(sh.below[q], (c32 >> q) & 1u, lane) & 3u;
            const unsigned p10 = __funnelshift_r(sh.below[q + 1], (c32 >> (q + 1)) & 1u, lane) & 3u;
            const unsigned p01 = __funnelshift_r(sh.below[q + ENY], (c32 >> (q + ENY)) & 1u, lane) & 3u;
            const unsigned p11 = __funnelshift_r(sh.below[q + ENY + 1], (c32 >> (q + ENY + 1)) & 1u, lane) & 3u;
            unsigned long long word = 0;
            if ((cellValidX >> lane) & 1u) word = __ldg(&kCasePerm.w[p00 | p10 << 2 | p01 << 4 | p11 << 6]);
            const unsigned n = static_cast<unsigned>(word) & 15u;
            unsigned inc = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += t;
            }
            const unsigned first = inc - n; // my first triangle within the row
#pragma unroll
            for (unsigned k = 0; k < 5; ++k) // a cell has at most five triangles: five predicated stores instead of a counted loop
                if (k < n) owner[first + k] = static_cast<unsigned char>(lane << 3 | k);
            __syncwarp();
            const unsigned wlo = static_cast<unsigned>(word >> 4), whi = static_cast<unsigned>(word >> 36); // 15 nibbles of edge ids
            const unsigned ncorn = segTris * 3;
            const size_t gbase = static_cast<size_t>(segOff) * 9;
            float* op = outPos + gbase + lane * 3; // this lane's corner of the current round; 32 corners = 96 floats per round
            float* on = outNrm + gbase + lane * 3;
            float* oc = COLOUR ? outCol + gbase + lane * 3 : nullptr;
            const float ty0 = sh.tab[E_TAB_Y + ly], ty1 = sh.tab[E_TAB_Y + ly + 1];
            const float tz0 = sh.tab[E_TAB_Z + lz], tz1 = sh.tab[E_TAB_Z + lz + 1];
            const unsigned* etab = sh.edgeTab[rr];
#pragma unroll 2 // two independent gather chains in flight: 3.08 -> 3.04 ms (unroll 4 spills: 3.16 ms)
            for (unsigned j = lane; j < ncorn + lane; j += 32, op += 96, on += 96) { // trip count uniform over the warp (shuffles inside)
                const bool act = j < ncorn;
                const unsigned t = act ? j / 3 : 0;
                const unsigned ok = owner[t];
                const unsigned L = ok >> 3;
                const unsigned oLo = __shfl_sync(0xffffffffu, wlo, L), oHi = __shfl_sync(0xffffffffu, whi, L);
                if (!act) continue;
                const unsigned slotc = 3 * (ok & 7u) + (j - 3 * t); // corner number inside the owner cell (0..14)
                const unsigned e = ((slotc < 8 ? oLo : oHi) >> ((4 * slotc) & 31u)) & 15u;
                const unsigned ent = etab[e];
                const unsigned eidx = (ent & 0xfffu) + L;
                const float4 v = sh.edge[eidx];
                float px = sh.tab[L + ((ent & ET_DX) ? 1 : 0)];
                float py = (ent & ET_DY) ? ty1 : ty0;
                float pz = (ent & ET_DZ) ? tz1 : tz0;
                if (ent & ET_AX0) px = v.x;
                if (ent & ET_AX1) py = v.x;
                if (ent & ET_AX2) pz = v.x;
                op[0] = px, op[1] = py, op[2] = pz;
                on[0] = v.y, on[1] = v.z, on[2] = v.w;
                if (COLOUR) {
                    const float4 cc = edgeCol[eidx];
                    float* o = oc + (j - lane) * 3;
                    o[0] = cc.x, o[1] = cc.y, o[2] = cc.z;
                }
            }
            __syncwarp();
        }
    }
    if (pending) waitPlanes(); // never leave a bulk copy in flight into a dying block's shared memory
}

} // namespace v4
} // namespace mms
