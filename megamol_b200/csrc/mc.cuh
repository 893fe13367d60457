// mc.cuh -- marching cubes on the device-resident density slab.
//
//   mc_count_kernel   "below iso" BIT MASKS: a warp turns a row of 32 nodes into one 32-bit word (coalesced 128-byte load,
//                     compare, ballot).  A thread then owns a whole 32-cell x-segment: the cells that produce triangles are
//                     found with a dozen bitwise operations on the 8 words of the segment's four node rows, and only those
//                     cells look up their triangle count (case table of trisoup::volumetrics::MarchingCubeTables,
//                     MarchingCubeTables.cpp:11-16,58,278; cube index bit i <=> corner i < iso).
//   (exclusive scan of the segment counts: scan.cuh -> triangle offsets in CELL-LINEAR order)
//   mc_emit_kernel    a block owns a 32x8-cell column and marches in z, ONE cell layer per iteration; the density planes live
//                     in a ring in shared memory that TMA (cp.async.bulk.tensor, one 40x11 box per plane, mbarrier completion)
//                     fills one iteration ahead.  The vertex records (one float4 per crossed grid edge) are kept per NODE PLANE
//                     (x- and y-edges) and per CELL LAYER (z-edges) in small rings, so every edge vertex of the column is computed
//                     exactly once while the march passes it.  Iteration j runs three independent pieces of work between two
//                     block barriers -- they touch different buffers, so no barrier separates them:
//                       MX(j+1) a warp per node row of plane j+1: "below iso" masks by ballot, crossed edges = XOR with the
//                               neighbouring masks (x: shifted, y: next row, z: plane j), ballot-compacted into a crossing list
//                       V(j)    one vertex per crossed edge of plane j / layer j-1, full lanes: interpolation parameter, coordinate,
//                               gradient normal -> one float4 record
//                       C(j-2)  a warp per 32-cell row of layer j-2: cube indices from the masks, warp prefix scan of the per-cell
//                               triangle counts, then the row's triangle corners are FLATTENED over the lanes (corner j -> lane
//                               j mod 32): a lane looks up its edge record through a 12-entry table and writes 3+3 floats;
//                               consecutive lanes write consecutive 12-byte pieces, every warp store covers one contiguous span.
// Output order = cell-linear (x fastest, then y, then z), inside a cell the table's order: independent of the
// tile shape and of the z-slab decomposition.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mms {

// word = ntri | edge0<<4 | edge1<<8 | ... (mc_case_words.inc is generated from the compiled reference table)
struct McCaseTable {
    unsigned long long w[256];
};
constexpr McCaseTable kCaseHost = {{
#include "mc_case_words.inc"
}};
// The kernels assemble the cube index from (x, x+1) bit pairs of four node rows, i.e. in the order
// c0 c1 | c3 c2 | c4 c5 | c7 c6; the table is stored under that permuted index.
constexpr int mcPermutedToCubeIndex(int i) {
    return (i & 0x33) | ((i & 0x44) << 1) | ((i & 0x88) >> 1);
}
constexpr McCaseTable mcMakePermuted() {
    McCaseTable t{};
    for (int i = 0; i < 256; ++i) t.w[i] = kCaseHost.w[mcPermutedToCubeIndex(i)];
    return t;
}
__device__ const McCaseTable kCasePerm = mcMakePermuted();

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
    int countLayers;    // mc_count_kernel: cell layers a block marches (the host picks it so that the grid fills whole waves of resident blocks)
    int layersPerBlock; // mc_emit_kernel: cell layers a block marches (<= EM_LAYERS; fewer on small volumes, so that the grid fills the machine)
    unsigned maxTris;  // mc_emit_kernel: cell rows whose triangles end beyond this many are skipped (speculative launch before the count is
                       // known on the host: the destination's capacity; 0xffffffff otherwise)
};

constexpr int MC_THREADS = 256;

// ---------------------------------------------------------------------------------------------------------------
// count
// ---------------------------------------------------------------------------------------------------------------
constexpr int CN_SEGS = 16;                 // x-segments (32 cells each) per block
constexpr int CN_ROWS = 16;                 // cell rows (y) per block
constexpr int CN_LAYERS = 16;               // cell layers (z) a block marches through (default; McGeo::countLayers)
constexpr int CN_WORDS = CN_SEGS + 1;       // mask words per node row: 16 segments + the first node of the next segment
constexpr int CN_NROWS = CN_ROWS + 1;       // node rows per plane

__global__ void __launch_bounds__(MC_THREADS) mc_count_kernel(McGeo m, const float* __restrict__ vol, unsigned* __restrict__ segCount,
    unsigned char* __restrict__ triCount, uint4* __restrict__ vrec, unsigned* __restrict__ vcount) {
    __shared__ unsigned sMask[3][CN_NROWS][CN_WORDS];
    __shared__ unsigned char sCount[256];
    sCount[threadIdx.x] = static_cast<unsigned char>(kCasePerm.w[threadIdx.x] & 15ull);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg0 = blockIdx.x * CN_SEGS, yBeg = blockIdx.y * CN_ROWS;
    const int czBeg = m.cz0 + blockIdx.z * m.countLayers, czEnd = min(czBeg + m.countLayers, m.cz0 + m.cnz);
    if (czBeg >= czEnd) return;
    const size_t plane = static_cast<size_t>(m.sx) * m.sy;
    const int nsegHere = min(CN_SEGS, m.nsegx - seg0);

    // one node plane -> mask words (nodes beyond the grid are clamped copies: they never make a valid cell look different)
    auto planeMasks = [&](int zNode, int buf) {
        const float* p = vol + plane * (zNode - m.zPlane0);
        // warp w converts segments w and w+8 of every node row: 2 x 17 words, 6 x 6 loads in flight
        const int xa = min((seg0 + warp) * 32 + lane, m.sx - 1), xb = min((seg0 + warp + 8) * 32 + lane, m.sx - 1);
#pragma unroll
        for (int r0 = 0; r0 < CN_NROWS; r0 += 3) {
            float va[3], vb[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const float* row = p + static_cast<size_t>(m.sx) * min(yBeg + r0 + u, m.sy - 1);
                va[u] = r0 + u < CN_NROWS ? row[xa] : 0.0f;
                vb[u] = r0 + u < CN_NROWS ? row[xb] : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const unsigned ba = __ballot_sync(0xffffffffu, va[u] < m.iso), bb = __ballot_sync(0xffffffffu, vb[u] < m.iso);
                if (lane == 0 && r0 + u < CN_NROWS) sMask[buf][r0 + u][warp] = ba, sMask[buf][r0 + u][warp + 8] = bb;
            }
        }
        // the node column after the block's last segment (bit 0 of word CN_SEGS)
        if (warp == (zNode & 7) && lane < CN_NROWS) {
            const int y = min(yBeg + lane, m.sy - 1), x = min((seg0 + CN_SEGS) * 32, m.sx - 1);
            sMask[buf][lane][CN_SEGS] = p[static_cast<size_t>(m.sx) * y + x] < m.iso ? 1u : 0u;
        }
    };

    planeMasks(czBeg, czBeg % 3);
    const int s = threadIdx.x & (CN_SEGS - 1), r = threadIdx.x >> 4; // this thread's segment and cell row
    const int xseg = seg0 + s, y = yBeg + r;
    const bool mine = s < nsegHere && y < m.cy;
    const int ncellsX = m.cx - xseg * 32; // valid cells of my segment
    const unsigned validCells = ncellsX >= 32 ? 0xffffffffu : (ncellsX > 0 ? (1u << ncellsX) - 1u : 0u);
    for (int cz = czBeg; cz < czEnd; ++cz) {
        planeMasks(cz + 1, (cz + 1) % 3);
        __syncthreads();
        if (mine) {
            const unsigned(*lo)[CN_WORDS] = sMask[cz % 3];
            const unsigned(*hi)[CN_WORDS] = sMask[(cz + 1) % 3];
            // (x, x+1) pairs live in the 33-bit value hiWord:loWord
            const unsigned a0 = lo[r][s], a1 = lo[r][s + 1], b0 = lo[r + 1][s], b1 = lo[r + 1][s + 1];
            const unsigned c0 = hi[r][s], c1 = hi[r][s + 1], d0 = hi[r + 1][s], d1 = hi[r + 1][s + 1];
            const unsigned a0s = __funnelshift_r(a0, a1, 1), b0s = __funnelshift_r(b0, b1, 1);
            const unsigned c0s = __funnelshift_r(c0, c1, 1), d0s = __funnelshift_r(d0, d1, 1);
            const unsigned all = a0 & a0s & b0 & b0s & c0 & c0s & d0 & d0s;
            const unsigned any = a0 | a0s | b0 | b0s | c0 | c0s | d0 | d0s;
            unsigned active = any & ~all & validCells;
            unsigned n = 0;
            unsigned char* tc = triCount ? triCount + (static_cast<size_t>(m.cx) * (y + static_cast<size_t>(m.cy) * (cz - m.cz0)) + xseg * 32) : nullptr;
            while (active) {
                const int b = __ffs(active) - 1;
                active &= active - 1;
                const unsigned ci = (__funnelshift_r(a0, a1, b) & 3u) | (__funnelshift_r(b0, b1, b) & 3u) << 2 |
                                    (__funnelshift_r(c0, c1, b) & 3u) << 4 | (__funnelshift_r(d0, d1, b) & 3u) << 6;
                const unsigned k = sCount[ci];
                n += k;
                if (tc) tc[b] = static_cast<unsigned char>(k);
            }
            segCount[xseg + static_cast<size_t>(m.nsegx) * (y + static_cast<size_t>(m.cy) * (cz - m.cz0))] = n;
            if (vrec) {
                // indexed mesh (mc_indexed.cuh): the mask record of node row (y, cz), segment xseg -- the bits are all here already.
                // Nodes beyond the grid are clamped copies, so an x-edge that leaves the grid is never crossed; y- and z-edges of
                // such copies are masked.  (The last node row / plane / a lone last node segment have no cell row: mcx_mask_kernel.)
                const int nvx = m.sx - xseg * 32;
                const unsigned nodeValid = nvx >= 32 ? 0xffffffffu : (1u << nvx) - 1u;
                const unsigned mx = a0 ^ a0s, my = (a0 ^ b0) & nodeValid, mz = (a0 ^ c0) & nodeValid;
                const size_t sv = xseg + static_cast<size_t>((m.sx + 31) >> 5) * (y + static_cast<size_t>(m.sy) * cz);
                vrec[sv] = make_uint4(a0, mx, my, mz);
                vcount[sv] = __popc(mx) + __popc(my) + __popc(mz);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// emit: z-marching column blocks, one cell layer per iteration, vertex records kept per node plane / cell layer
// ---------------------------------------------------------------------------------------------------------------
constexpr int EX = 32, EY = 8;                              // cells per layer of a block column
constexpr int ENX = EX + 1, ENY = EY + 1;                   // nodes per plane of a block column
constexpr int EHY = EY + 3;                                 // node rows + gradient halo (y0-1 .. y0+9)
constexpr int EPITCH = 40;                                  // x0-4 .. x0+35 (x0-1 .. x0+33 are needed): TMA wants the box start and width in
                                                            // multiples of 16 bytes (an unaligned start coordinate is an illegal instruction)
constexpr int EHX0 = 4;                                     // ring column of node 0
constexpr int EPLANE = 448;                                 // floats per ring slot: 11*40 = 440, padded to 14*128 bytes
constexpr int ERING = 6;                                    // plane slots: V(j) reads planes j-2 .. j+1 while plane j+2 is in flight
constexpr int EM_LAYERS = 64;                               // cell layers per block
// a crossed grid edge is named (axis, node row r, node ix): list entry = axis << 10 | r << 6 | ix; its record sits at r*33 + ix of
//   x-edges  node-plane buffer            (r < 9, ix < 32)
//   y-edges  node-plane buffer + E_YOFF   (r < 8, ix < 33)
//   z-edges  cell-layer buffer            (r < 9, ix < 33)
constexpr int E_YOFF = ENY * ENX;                           // 297
constexpr int E_PLREC = E_YOFF + EY * ENX;                  // 561 records per node plane
constexpr int E_ZEDGES = ENY * ENX;                         // 297 records per cell layer
constexpr int E_ZREC0 = 3 * E_PLREC;                        // records: three node-plane buffers, then two cell-layer buffers
constexpr int E_RECS = 3 * E_PLREC + 2 * E_ZEDGES;          // 2277
constexpr int E_LIST = 864;                                 // >= 9*32 + 8*33 + 9*33 = 849 crossings per iteration
constexpr int E_MAXROWTRIS = 160;
constexpr int E_ROWS = EM_LAYERS * EY;                      // 512 cell rows per block
constexpr int E_TAB_Y = ENX;                                // node position table: 33 x, 9 y
constexpr unsigned E_PLANE_BYTES = EHY * EPITCH * 4;

struct __align__(128) McEmitShared {
    float ring[ERING * EPLANE];                 // node plane p (relative to the block's first cell layer) sits in slot (p + 1) mod ERING
    float4 rec[E_RECS];                         // {interpolated coordinate along the edge's axis, nx, ny, nz}
    unsigned short list[2][E_LIST];             // crossing list of plane j in list[j & 1]
    unsigned short triOwner[MC_THREADS / 32][E_MAXROWTRIS]; // per triangle of a row: (owner lane & 15) << 12 | its three cube edges (nibbles)
    unsigned segOff[E_ROWS];                    // first triangle of every cell row of the block
    unsigned char segCnt[E_ROWS];               // its triangle count (<= 160)
    unsigned etab[MC_THREADS / 32][12];         // per warp (= cell row) and cube edge: record slot of cell 0 | flags
    uint2 below[4][ENY];                        // "below iso" bits of node plane j in below[j & 3]: .x nodes 0..31, .y node 32
    float tab[ENX + ENY + 2];                   // node positions float(idx)*sd + origin (ParticlesToDensity.cpp:605)
    unsigned actWarp[2][MC_THREADS / 32];
    int ncross[4];                              // length of the crossing list of plane j in ncross[j mod 3]
    unsigned long long mbar[2];                 // plane loads alternate between two mbarriers (see the loop)
};
static_assert(sizeof(McEmitShared) + 128 <= 57088, "mc_emit_kernel must fit four blocks per SM");
static_assert(E_RECS < 4096, "record slots are 12-bit fields of the edge table");

// per cube edge: low corner (dx,dy,dz) and axis: dx | dy<<1 | dz<<2 | axis<<3   (MarchingCubeTables.cpp:15-16, low node first)
//  e0 (0,0,0)x  e1 (1,0,0)y  e2 (0,1,0)x  e3 (0,0,0)y  e4 (0,0,1)x  e5 (1,0,1)y  e6 (0,1,1)x  e7 (0,0,1)y
//  e8 (0,0,0)z  e9 (1,0,0)z  e10 (1,1,0)z e11 (0,1,0)z
__host__ __device__ __forceinline__ unsigned edgeCode(int e) {
    const unsigned long long codes = (0ull) | (9ull << 5) | (2ull << 10) | (8ull << 15) | (4ull << 20) | (13ull << 25) |
                                     (6ull << 30) | (12ull << 35) | (16ull << 40) | (17ull << 45) | (19ull << 50) | (18ull << 55);
    return static_cast<unsigned>(codes >> (5 * e)) & 31u;
}

// flags of an etab entry (bits 0..11 = record slot of the row's cell 0)
constexpr unsigned ET_AX0 = 1u << 12, ET_AX1 = 1u << 13, ET_AX2 = 1u << 14, ET_DX = 1u << 15, ET_DY = 1u << 16, ET_DZ = 1u << 17;

__device__ __forceinline__ float rcpApproxF(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrtApproxF(float x) { // = rsqrtf(x) for normal x (gradients below 1e-19 do not occur), without its denormal branch
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ unsigned smemAddr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cpAsync4(float* smemDst, const float* gmemSrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smemAddr(smemDst)), "l"(gmemSrc));
}
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
    const unsigned a = smemAddr(bar);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmaLoadPlane(float* smemDst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smemAddr(smemDst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(z), "r"(smemAddr(bar)) : "memory");
}

/**
 * TMA: the planes arrive as 40x11x1 boxes of a 3-D tensor map over the slab volume (needs sx % 4 == 0; out-of-range elements are
 * zero-filled, which is harmless: gradients at the global border are one-sided and cells beyond the grid are masked).
 * !TMA: the same ring filled with 4-byte cp.async copies, coordinates clamped.
 *
 * Iteration j (j = first active layer - 1 .. last active layer + 2) runs MX(j+1), V(j) and C(j-2); what they share:
 *   ring      MX(j+1) reads plane j+1, V(j) planes j-2 .. j+1; the load of plane j+2 is in flight into the slot of plane j-4
 *   below     MX(j+1) writes [(j+1)&3] and reads [j&3]; C(j-2) reads [(j-2)&3], [(j-1)&3]
 *   list      MX(j+1) writes [(j+1)&1], V(j) reads [j&1]
 *   rec       V(j) writes plane buffer j mod 3 and layer buffer (j-1)&1; C(j-2) reads planes (j-2) mod 3, (j-1) mod 3, layer (j-2)&1
 * so ONE block barrier per iteration orders everything.  Inside an iteration the warps split the work by role = (warp - j) mod 8
 * (rotating, so that nobody is the slow one every time): roles 0..2 run MX (three node rows each), roles 3..7 walk the crossing
 * list of V; then every warp emits the triangles of cell row `warp` of layer j-2.
 */
template<bool COLOUR, bool TMA>
__global__ void __launch_bounds__(MC_THREADS, COLOUR ? 2 : 4) mc_emit_kernel(McGeo m, const __grid_constant__ CUtensorMap volMap,
    const float* __restrict__ vol, const float* __restrict__ rgb, const unsigned* __restrict__ segOffset, float* __restrict__ outPos,
    float* __restrict__ outNrm, float* __restrict__ outCol) {
    extern __shared__ unsigned char smemRaw[];
    unsigned char* smemAligned = smemRaw + ((128u - (smemAddr(smemRaw) & 127u)) & 127u);
    McEmitShared& sh = *reinterpret_cast<McEmitShared*>(smemAligned);
    float4* recCol = reinterpret_cast<float4*>(smemAligned + sizeof(McEmitShared)); // COLOUR only
    const int x0 = blockIdx.x * EX, y0 = blockIdx.y * EY;
    const int zcBeg = m.cz0 + blockIdx.z * m.layersPerBlock;               // first global cell layer (= node plane) of this block
    const int nLayers = min(m.layersPerBlock, m.cz0 + m.cnz - zcBeg);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- the block's cell rows: first triangle, triangle count, which layers have any (one global read per row) ---------
    static_assert(E_ROWS == 2 * MC_THREADS, "two cell rows per thread");
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int ly = tid % EY, k = tid / EY + h * (EM_LAYERS / 2);
        const int cyi = y0 + ly;
        unsigned off = 0, cnt = 0;
        if (cyi < m.cy && k < nLayers) {
            const size_t seg = blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (zcBeg + k - m.cz0));
            off = segOffset[seg];
            cnt = segOffset[seg + 1] - off;
            if (off + cnt > m.maxTris) cnt = 0; // beyond the destination's capacity (the host re-emits after growing the buffers)
        }
        sh.segOff[tid + h * MC_THREADS] = off;
        sh.segCnt[tid + h * MC_THREADS] = static_cast<unsigned char>(cnt);
        const unsigned bal = __ballot_sync(0xffffffffu, cnt != 0); // rows 8i .. 8i+7 of this warp belong to layer 32*h + 4*warp + i
        if (lane == 0)
            sh.actWarp[h][warp] = ((bal & 0xffu) ? 1u : 0u) | ((bal & 0xff00u) ? 2u : 0u) | ((bal & 0xff0000u) ? 4u : 0u) | ((bal >> 24) ? 8u : 0u);
    }
    if (tid < ENX) sh.tab[tid] = __fadd_rn(__fmul_rn((float)(x0 + tid), m.sd[0]), m.org[0]);
    else if (tid < ENX + ENY) sh.tab[tid] = __fadd_rn(__fmul_rn((float)(y0 + tid - ENX), m.sd[1]), m.org[1]);
    if (tid < 4) sh.ncross[tid] = 0;
    if (TMA && tid == 0) mbarInit(&sh.mbar[0], 1), mbarInit(&sh.mbar[1], 1);
    __syncthreads();
    unsigned actLo = 0, actHi = 0;
#pragma unroll
    for (int w = 0; w < MC_THREADS / 32; ++w) actLo |= sh.actWarp[0][w] << (4 * w), actHi |= sh.actWarp[1][w] << (4 * w);
    const unsigned long long act = static_cast<unsigned long long>(actHi) << 32 | actLo; // bit k <=> cell layer k of the block has triangles
    if (!act) return;
    const int kFirst = __ffsll(static_cast<long long>(act)) - 1, kLast = 63 - __clzll(static_cast<long long>(act));

    // ---- plane loader -------------------------------------------------------------------------------------------------
    // Load batch q signals mbar[q & 1] (phase parity (q >> 1) & 1).  Two barriers, because thread 0 arms the next batch right after its
    // own wait: with a single barrier a fast load could complete the NEXT phase before a slower warp has polled this one, and that
    // warp would then wait for a phase nobody arms.  Batch q+2 re-uses batch q's barrier one block barrier later: everybody is past it.
    unsigned qIssue = 0, qWait = 0;
    auto issuePlanes = [&](int pFirst, int pLast, int slotOff) { // uniform; asynchronous; plane pFirst goes to ring offset slotOff (floats)
        if (TMA) {
            if (tid == 0) {
                unsigned long long* bar = &sh.mbar[qIssue & 1u];
                mbarExpectTx(bar, static_cast<unsigned>(pLast - pFirst + 1) * E_PLANE_BYTES);
                for (int p = pFirst; p <= pLast; ++p, slotOff = slotOff == (ERING - 1) * EPLANE ? 0 : slotOff + EPLANE)
                    tmaLoadPlane(&sh.ring[slotOff], &volMap, x0 - EHX0, y0 - 1, zcBeg + p - m.zPlane0, bar);
            }
            ++qIssue;
        } else {
            for (int p = pFirst; p <= pLast; ++p, slotOff = slotOff == (ERING - 1) * EPLANE ? 0 : slotOff + EPLANE) {
                const int zg = min(max(zcBeg + p, 0), m.szGlobal - 1);
                const int zl = min(max(zg - m.zPlane0, 0), m.nzPlanes - 1);
                float* dst = &sh.ring[slotOff];
                for (int i = tid; i < EHY * (ENX + 2); i += MC_THREADS) { // nodes -1 .. 33 of every row
                    const int ix = i % (ENX + 2) - 1, iy = i / (ENX + 2);
                    const int x = min(max(x0 + ix, 0), m.sx - 1), y = min(max(y0 + iy - 1, 0), m.sy - 1);
                    cpAsync4(dst + iy * EPITCH + ix + EHX0, vol + x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
                }
            }
            asm volatile("cp.async.commit_group;");
        }
    };

    const float r1x = m.rinv[0][1], r2x = m.rinv[0][2], r1y = m.rinv[1][1], r2y = m.rinv[1][2], r1z = m.rinv[2][1], r2z = m.rinv[2][2];
    const float iso = m.iso;
    // validity of the block's nodes (TMA zero-fills beyond the grid; clamped copies would be harmless, zeros are not)
    const int nvx = min(ENX, m.sx - x0), nvy = min(ENY, m.sy - y0);
    const unsigned nodeValidX = nvx >= 32 ? 0xffffffffu : (1u << nvx) - 1u;            // nodes 0..31
    const unsigned xEdgeValid = nvx >= 33 ? 0xffffffffu : (1u << (nvx - 1)) - 1u;      // x-edge ix needs node ix+1
    const unsigned cellValidX = m.cx - x0 >= 32 ? 0xffffffffu : (1u << (m.cx - x0)) - 1u;
    const unsigned ltMask = (1u << lane) - 1u;
    const bool xyInterior = x0 > 0 && x0 + EX < m.sx - 1 && y0 > 0 && y0 + EY < m.sy - 1; // no node of the column touches the x/y border
    unsigned short* owner = sh.triOwner[warp];
    // this warp's cell row is always row `warp` of the layer: static part of its edge table (lanes 0..11)
    unsigned etStatic = 0;
    bool etZ = false, etUp = false;
    if (lane < 12) {
        const unsigned code = edgeCode(lane);
        const int dx = code & 1, dy = (code >> 1) & 1, dz = (code >> 2) & 1, axis = code >> 3;
        const unsigned slot = (warp + dy) * ENX + dx + (axis == 1 ? E_YOFF : 0);
        etStatic = slot | (ET_AX0 << axis) | (dx ? ET_DX : 0u) | (dy ? ET_DY : 0u) | (dz ? ET_DZ : 0u);
        etZ = axis == 2, etUp = dz != 0;
    }
    const float ty0 = sh.tab[E_TAB_Y + warp], ty1 = sh.tab[E_TAB_Y + warp + 1];

    // per-iteration state, rotated at the end of every iteration (no divisions in the loop)
    int oM2 = ((kFirst + 4) % ERING) * EPLANE, oM1 = ((kFirst + 5) % ERING) * EPLANE, o0 = (kFirst % ERING) * EPLANE; // ring offsets of planes
    int oP1 = ((kFirst + 1) % ERING) * EPLANE, oP2 = ((kFirst + 2) % ERING) * EPLANE;                                  // j-2 .. j+2, j = kFirst-1
    int n3M = (kFirst + 1) % 3, n3J = (kFirst + 2) % 3, n3P = kFirst % 3;                       // (j-1) mod 3, j mod 3, (j+1) mod 3
    int role = (warp - kFirst + 1) & 7;                                                           // (warp - j) mod 8
    unsigned long long actRem = act >> kFirst;                                                    // bit 0 <=> layer j+1 is active
    unsigned actHist = 1u;                                                                        // bits 0..3 <=> layers j+1, j, j-1, j-2 are active
    auto planeZ = [&](int p) { return __fadd_rn(__fmul_rn((float)(zcBeg + p), m.sd[2]), m.org[2]); }; // node position (ParticlesToDensity.cpp:605)
    float pzM2 = planeZ(kFirst - 3), pzM1 = planeZ(kFirst - 2), pz0 = planeZ(kFirst - 1);      // z of node planes j-2, j-1, j
    issuePlanes(kFirst - 1, kFirst, o0); // planes j, j+1
    bool pending = TMA; // a bulk load has been issued and not yet waited for
    // ONE thread polls the mbarrier (acquire) and the block barrier that follows hands the planes to everybody else: eight polling
    // warps cost 4 % of the kernel's instructions
    auto waitPlanes = [&]() {
        if (TMA && pending) {
            if (tid == 0) mbarWait(&sh.mbar[qWait & 1u], (qWait >> 1) & 1u);
            ++qWait;
            pending = false;
        }
    };
    if (!TMA) asm volatile("cp.async.wait_group 0;");
    waitPlanes();
    __syncthreads();
    for (int j = kFirst - 1; j <= kLast + 2; ++j) {
        if (j <= kLast) { // plane j+2 for the next iteration; its slot held plane j-4, last read two barriers ago
            issuePlanes(j + 2, j + 2, oP2);
            pending = TMA;
        }
        const bool actC = (actHist >> 3) & 1u, actJ = (actHist >> 1) & 1u, actJ1 = actHist & 1u; // layers j-2, j, j+1

        if (role < 3) {
            // ---- MX(j+1): masks of node plane j+1, crossed edges -> crossing list; three warps, three node rows each -------------
            if (role == 0 && lane == 0) sh.ncross[n3M] = 0; // counter of plane j+2: last read by V(j-1), next used by MX(j+2)
            if (actJ | actJ1) {
                const int jj = j + 1, r0 = role * 3;
                const float* P = sh.ring + oP1 + (r0 + 1) * EPITCH + EHX0; // node (row r0, x 0) of plane j+1
                uint2* belowJ = sh.below[jj & 3] + r0;
                const uint2* belowP = sh.below[j & 3] + r0;
                unsigned short* list = sh.list[jj & 1];
                // node 32 of rows r0 .. r0+3: bit i of cw; nodes 0..31: b[i]
                const unsigned cw = __ballot_sync(0xffffffffu, lane < 4 && r0 + lane < ENY && nvx >= ENX && P[lane * EPITCH + EX] < iso);
                unsigned b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) b[i] = (i < 3 || role < 2) ? __ballot_sync(0xffffffffu, P[i * EPITCH + lane] < iso) : 0u;
                unsigned mx[3], my[3], mz[3], y32 = 0, z32 = 0, tot = 0;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int r = r0 + i;
                    const unsigned c = (cw >> i) & 1u;
                    mx[i] = my[i] = mz[i] = 0;
                    if (r < nvy) {
                        mx[i] = (b[i] ^ ((b[i] >> 1) | (c << 31))) & xEdgeValid;
                        if (r < EY && r + 1 < nvy) my[i] = (b[i] ^ b[i + 1]) & nodeValidX, y32 |= (c ^ ((cw >> (i + 1)) & 1u)) << i;
                        if (actJ) { // z-edges of layer j: plane j below
                            const uint2 bp = belowP[i];
                            mz[i] = (b[i] ^ bp.x) & nodeValidX, z32 |= (c ^ bp.y) << i;
                        }
                    }
                    if (lane == 0) belowJ[i] = make_uint2(b[i], c);
                    tot += __popc(mx[i]) + __popc(my[i]) + __popc(mz[i]);
                }
                tot += __popc(y32) + __popc(z32);
                if (tot) {
                    unsigned pos = 0;
                    if (lane == 0) pos = static_cast<unsigned>(atomicAdd(&sh.ncross[n3P], static_cast<int>(tot)));
                    pos = __shfl_sync(0xffffffffu, pos, 0);
                    const unsigned me = (r0 << 6) | lane;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        if ((mx[i] >> lane) & 1u) list[pos + __popc(mx[i] & ltMask)] = static_cast<unsigned short>(me + (i << 6));
                        pos += __popc(mx[i]);
                        if ((my[i] >> lane) & 1u) list[pos + __popc(my[i] & ltMask)] = static_cast<unsigned short>(me + ((i << 6) | (1 << 10)));
                        pos += __popc(my[i]);
                        if ((mz[i] >> lane) & 1u) list[pos + __popc(mz[i] & ltMask)] = static_cast<unsigned short>(me + ((i << 6) | (2 << 10)));
                        pos += __popc(mz[i]);
                    }
                    // node column 32: lane i < 3 takes the y-edge, lane 4+i the z-edge of row r0+i
                    const unsigned m32 = y32 | (z32 << 4);
                    if ((m32 >> lane) & 1u)
                        list[pos + __popc(m32 & ltMask)] = static_cast<unsigned short>(((r0 + (lane & 3)) << 6) | EX | (lane < 4 ? 1 << 10 : 2 << 10));
                }
            }
        } else {
            // ---- V(j): one vertex per crossed edge of node plane j (x, y) and of cell layer j-1 (z); five warps, 160 entries a round ---
            const int ncross = sh.ncross[n3J];
            const int cBeg = ((role - 3) << 5) + lane;
            if (cBeg - lane < ncross) {
                const unsigned short* list = sh.list[j & 1];
                const int zg = zcBeg + j; // global index of node plane j
                // no node of this iteration touches the global border -> plain central differences, no per-round check
                const bool allInterior = xyInterior && zg - 1 > 0 && zg < m.szGlobal - 1;
                const int recPlane = n3J * E_PLREC, recLayer = E_ZREC0 + ((j + 1) & 1) * E_ZEDGES;
#pragma unroll 1
                for (int c = cBeg; c < ncross; c += 5 * 32) {
                    const unsigned ent = list[c];
                    const int axis = ent >> 10, r = (ent >> 6) & 15, ix = ent & 63;
                    const int offA = r * EPITCH + ix + (EPITCH + EHX0);
                    const int offB = offA + (axis == 0 ? 1 : (axis == 1 ? EPITCH : 0));
                    // node A is the edge's low node: in plane j-1 for a z-edge, else in plane j; node B is always in plane j
                    const float* ring = sh.ring;
                    const float* A0 = ring + ((axis == 2 ? oM1 : o0) + offA);
                    const float* AM = ring + ((axis == 2 ? oM2 : oM1) + offA);
                    const float* AP = ring + ((axis == 2 ? o0 : oP1) + offA);
                    const float* B0 = ring + (o0 + offB);
                    const float* BM = ring + (oM1 + offB);
                    const float* BP = ring + (oP1 + offB);
                    const int gxi = x0 + ix, gyi = y0 + r, gzi = zg - (axis == 2); // node A; node B = A + unit vector of the axis
                    bool slow = false;
                    if (!allInterior) {
                        const bool border = min(min(gxi, gyi), gzi) <= 0 || gxi + (axis == 0) >= m.sx - 1 || gyi + (axis == 1) >= m.sy - 1 || zg >= m.szGlobal - 1;
                        slow = __any_sync(__activemask(), border);
                    }
                    float gax, gay, gaz, gbx, gby, gbz;
                    if (!slow) { // gradient at a node: (f(+) - f(-)) * 1/(2 sd)
                        gax = __fmul_rn(__fsub_rn(A0[1], A0[-1]), r2x), gay = __fmul_rn(__fsub_rn(A0[EPITCH], A0[-EPITCH]), r2y);
                        gaz = __fmul_rn(__fsub_rn(AP[0], AM[0]), r2z);
                        gbx = __fmul_rn(__fsub_rn(B0[1], B0[-1]), r2x), gby = __fmul_rn(__fsub_rn(B0[EPITCH], B0[-EPITCH]), r2y);
                        gbz = __fmul_rn(__fsub_rn(BP[0], BM[0]), r2z);
                    } else { // samples clamped at the GLOBAL grid border: one-sided differences there
                        auto grad = [&](const float* P0, const float* Pm, const float* Pp, int gx_, int gy_, int gz_, float& gx, float& gy, float& gz) {
                            const int xm = gx_ > 0 ? 1 : 0, xp = gx_ < m.sx - 1 ? 1 : 0;
                            const int ym = gy_ > 0 ? 1 : 0, yp = gy_ < m.sy - 1 ? 1 : 0;
                            const int zm = gz_ > 0 ? 1 : 0, zp = gz_ < m.szGlobal - 1 ? 1 : 0;
                            gx = xp + xm ? __fmul_rn(__fsub_rn(P0[xp], P0[-xm]), xp + xm == 2 ? r2x : r1x) : 0.0f;
                            gy = yp + ym ? __fmul_rn(__fsub_rn(P0[yp * EPITCH], P0[-ym * EPITCH]), yp + ym == 2 ? r2y : r1y) : 0.0f;
                            gz = zp + zm ? __fmul_rn(__fsub_rn(zp ? Pp[0] : P0[0], zm ? Pm[0] : P0[0]), zp + zm == 2 ? r2z : r1z) : 0.0f;
                        };
                        grad(A0, AM, AP, gxi, gyi, gzi, gax, gay, gaz);
                        grad(B0, BM, BP, gxi + (axis == 0), gyi + (axis == 1), zg, gbx, gby, gbz);
                    }
                    const float fa = A0[0], fb = B0[0];
                    // t = (iso - fa) / (fb - fa): SFU reciprocal (2 ulp; vertices are compared at 1e-4 of a cell); the IEEE division only
                    // where the difference is too small for rcp.approx
                    const float tnum = __fsub_rn(iso, fa), tden = __fsub_rn(fb, fa);
                    const float t01 = fabsf(tden) > 1e-30f ? __fmul_rn(tnum, rcpApproxF(tden)) : __fdiv_rn(tnum, tden);
                    const int ti = axis == 0 ? ix : E_TAB_Y + r;
                    float pa = sh.tab[ti], pb = sh.tab[ti + 1];
                    if (axis == 2) pa = pzM1, pb = pz0;
                    const float gx = __fadd_rn(gax, __fmul_rn(t01, __fsub_rn(gbx, gax)));
                    const float gy = __fadd_rn(gay, __fmul_rn(t01, __fsub_rn(gby, gay)));
                    const float gz = __fadd_rn(gaz, __fmul_rn(t01, __fsub_rn(gbz, gaz)));
                    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
                    const float inv = len2 > 0.0f ? -rsqrtApproxF(len2) : 0.0f; // SFU rsqrt: 2 ulp, normals are compared at 1e-4
                    const int slot = r * ENX + ix + (axis == 0 ? recPlane : (axis == 1 ? recPlane + E_YOFF : recLayer));
                    sh.rec[slot] = make_float4(__fadd_rn(pa, __fmul_rn(t01, __fsub_rn(pb, pa))), __fmul_rn(gx, inv), __fmul_rn(gy, inv), __fmul_rn(gz, inv));
                    if (COLOUR) { // node colour = rgb / rho (0 where rho == 0), interpolated with the same t
                        auto nodeColour = [&](int nx_, int ny_, int gz_, float f, float& cr, float& cg, float& cb) {
                            const int x = min(nx_, m.sx - 1), y = min(ny_, m.sy - 1);
                            const int zl = min(max(gz_ - m.zPlane0, 0), m.nzPlanes - 1);
                            const float* cc = rgb + 3 * (x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl));
                            if (f > 0.0f) cr = __fdiv_rn(cc[0], f), cg = __fdiv_rn(cc[1], f), cb = __fdiv_rn(cc[2], f);
                            else cr = cg = cb = 0.0f;
                        };
                        float ar, ag, ab, br, bg, bb;
                        nodeColour(gxi, gyi, gzi, fa, ar, ag, ab);
                        nodeColour(gxi + (axis == 0), gyi + (axis == 1), zg, fb, br, bg, bb);
                        recCol[slot] = make_float4(__fadd_rn(ar, __fmul_rn(t01, __fsub_rn(br, ar))), __fadd_rn(ag, __fmul_rn(t01, __fsub_rn(bg, ag))),
                            __fadd_rn(ab, __fmul_rn(t01, __fsub_rn(bb, ab))), 0.0f);
                    }
                }
            }
        }

        // ---- C(j-2): triangles of cell layer k = j-2, one warp per 32-cell row --------------------------------------------------------
        if (actC) {
            const int k = j - 2;
            const int row = k * EY + warp;
            const unsigned segTris = sh.segCnt[row];
            if (segTris != 0) {
                const unsigned segOff = sh.segOff[row];
                // cube index in permuted order from the (x, x+1) bit pairs of the four node rows
                const uint2 m00 = sh.below[k & 3][warp], m10 = sh.below[k & 3][warp + 1];
                const uint2 m01 = sh.below[(k + 1) & 3][warp], m11 = sh.below[(k + 1) & 3][warp + 1];
                const unsigned p00 = __funnelshift_r(m00.x, m00.y, lane) & 3u, p10 = __funnelshift_r(m10.x, m10.y, lane) & 3u;
                const unsigned p01 = __funnelshift_r(m01.x, m01.y, lane) & 3u, p11 = __funnelshift_r(m11.x, m11.y, lane) & 3u;
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
                for (unsigned q = 0; q < 5; ++q) // a cell has at most five triangles: five predicated stores instead of a counted loop
                    if (q < n) owner[first + q] = static_cast<unsigned short>(lane << 12 | (static_cast<unsigned>(word >> (4 + 12 * q)) & 0xfffu));
                const unsigned tHalf = __shfl_sync(0xffffffffu, first, 16); // triangles from here on belong to lanes 16..31 (the table keeps 4 lane bits)
                // this layer's edge table: records of node planes k = j-2 (buffer (j+1) mod 3), k+1 (buffer (j-1) mod 3) and of cell layer k
                if (lane < 12) sh.etab[warp][lane] = etStatic + (etZ ? E_ZREC0 + (j & 1) * E_ZEDGES : (etUp ? n3M : n3P) * E_PLREC);
                __syncwarp();
                const unsigned ncorn = segTris * 3;
                const size_t gbase = static_cast<size_t>(segOff) * 9;
                float* op = outPos + gbase + lane * 3; // this lane's corner of the current round; 32 corners = 96 floats per round
                float* on = outNrm + gbase + lane * 3;
                float* oc = COLOUR ? outCol + gbase + lane * 3 : nullptr;
                const float tz0 = pzM2, tz1 = pzM1;
                const unsigned* etab = sh.etab[warp];
                // not unrolled: a row has 2.6 rounds on average, and unrolling by two costs more in the remainder logic than the second
                // gather chain gains (C2: 2.635 -> 2.522 ms; by three: 2.68 ms; profiles/rejected/r2_emit_micro_variants.md)
#pragma unroll 1
                for (unsigned jc = lane; jc < ncorn; jc += 32, op += 96, on += 96) {
                    const unsigned t = jc / 3;
                    const unsigned ok = owner[t];
                    const unsigned L = (ok >> 12) + (t >= tHalf ? 16u : 0u);
                    const unsigned e = (ok >> (4 * (jc - 3 * t))) & 15u; // this corner's cube edge
                    const unsigned ent = etab[e];
                    const unsigned eidx = (ent & 0xfffu) + L;
                    const float4 v = sh.rec[eidx];
                    float px = sh.tab[L + ((ent >> 15) & 1u)]; // ET_DX
                    float py = (ent & ET_DY) ? ty1 : ty0;
                    float pz = (ent & ET_DZ) ? tz1 : tz0;
                    if (ent & ET_AX0) px = v.x;
                    if (ent & ET_AX1) py = v.x;
                    if (ent & ET_AX2) pz = v.x;
                    op[0] = px, op[1] = py, op[2] = pz;
                    on[0] = v.y, on[1] = v.z, on[2] = v.w;
                    if (COLOUR) {
                        const float4 cc = recCol[eidx];
                        float* o = oc + (jc - lane) * 3;
                        o[0] = cc.x, o[1] = cc.y, o[2] = cc.z;
                    }
                }
            }
        }
        if (!TMA) asm volatile("cp.async.wait_group 0;");
        waitPlanes(); // planes .. j+2 have landed
        __syncthreads();
        // rotate: j -> j+1
        oM2 = oM1, oM1 = o0, o0 = oP1, oP1 = oP2, oP2 = oP2 == (ERING - 1) * EPLANE ? 0 : oP2 + EPLANE;
        const int t3 = n3M;
        n3M = n3J, n3J = n3P, n3P = t3;
        role = (role + 7) & 7;
        actRem >>= 1;
        actHist = ((actHist << 1) | (static_cast<unsigned>(actRem) & 1u)) & 15u;
        pzM2 = pzM1, pzM1 = pz0, pz0 = planeZ(j + 1);
    }
    // (the loop's last iteration issues no load and every issued load has been waited for: nothing is in flight into a dying block)
}

} // namespace mms
