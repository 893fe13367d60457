// mc_indexed.cuh -- marching cubes with an INDEXED triangle mesh (opt-in; the reference's contract is the unindexed soup of mc.cuh).
//
// trisoup's CallTriMeshData carries indexed meshes as well (SetVertexData + SetTriangleData with 32-bit indices,
// mesh_gl/CallTriMeshDataGL.h:922-1000 and geometry_calls/CallTriMeshData.h): every crossed grid edge becomes ONE vertex
// (24 bytes: position + normal) and a triangle is three indices (12 bytes) -- ~28 bytes per triangle instead of the soup's 72,
// for the HBM writes, the PCIe read-back and the renderer's vertex fetch alike.
//
//   vertex numbering   a crossed edge belongs to its LOW node; the nodes are walked in rows of 32 ("node segments", x fastest, then y,
//                      then z), inside a segment node by node, per node x-edge, y-edge, z-edge.  So
//                          id(node i of segment s, axis a) = vertOffset[s] + popc(mx & lt_i) + popc(my & lt_i) + popc(mz & lt_i)
//                                                            + (a >= 1 ? mx_i : 0) + (a == 2 ? my_i : 0)
//                      with mx/my/mz the segment's crossing masks -- any kernel that knows the "below iso" bits around a cell can name
//                      the vertex of each of its edges without a search or a hash.
//   mcx_vertex_kernel  <false>: crossing count per node segment (-> exclusive scan = vertOffset); <true>: the vertices, with exactly the
//                      arithmetic of mc_emit_kernel's V stage (gradient normal by central differences clamped at the global border,
//                      SFU reciprocal for the interpolation parameter): position and normal are BIT-IDENTICAL to the soup's corners.
//   mcx_index_kernel   a warp per 32-cell row: cube indices from the bit masks, the row's triangle corners flattened over the lanes
//                      (as in mc_emit_kernel), each lane writes one 32-bit index: fully coalesced 128-byte stores.
// Triangle order = the soup's (cell-linear, the table's order inside a cell): expanding the indices reproduces the soup exactly.
#pragma once
#include "mc.cuh"

namespace mms {

constexpr int MCX_THREADS = 256;
constexpr int MCX_WARPS = MCX_THREADS / 32;

__device__ __forceinline__ float mcxSample(const McGeo& m, const float* __restrict__ vol, int x, int y, int z) {
    return __ldg(vol + x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * (z - m.zPlane0)));
}

/** Gradient at node (x, y, z): (f(+) - f(-)) * 1/(n*sd), samples clamped at the GLOBAL grid border (one-sided there). */
__device__ __forceinline__ void mcxGradient(const McGeo& m, const float* __restrict__ vol, int x, int y, int z, float& gx, float& gy, float& gz) {
    const int xm = x > 0 ? 1 : 0, xp = x < m.sx - 1 ? 1 : 0;
    const int ym = y > 0 ? 1 : 0, yp = y < m.sy - 1 ? 1 : 0;
    const int zm = z > 0 ? 1 : 0, zp = z < m.szGlobal - 1 ? 1 : 0;
    gx = xp + xm ? __fmul_rn(__fsub_rn(mcxSample(m, vol, x + xp, y, z), mcxSample(m, vol, x - xm, y, z)), m.rinv[0][xp + xm]) : 0.0f;
    gy = yp + ym ? __fmul_rn(__fsub_rn(mcxSample(m, vol, x, y + yp, z), mcxSample(m, vol, x, y - ym, z)), m.rinv[1][yp + ym]) : 0.0f;
    gz = zp + zm ? __fmul_rn(__fsub_rn(mcxSample(m, vol, x, y, z + zp), mcxSample(m, vol, x, y, z - zm)), m.rinv[2][zp + zm]) : 0.0f;
}

/** The vertex on the crossed edge from node A = (x, y, z) to A + unit(axis); fa, fb = the two node values.  Operation for operation the
 *  V stage of mc_emit_kernel (mc.cuh). */
__device__ __forceinline__ void mcxEdgeVertex(const McGeo& m, const float* __restrict__ vol, int x, int y, int z, int axis, float fa, float fb,
    float* __restrict__ pos, float* __restrict__ nrm) {
    float gax, gay, gaz, gbx, gby, gbz;
    mcxGradient(m, vol, x, y, z, gax, gay, gaz);
    mcxGradient(m, vol, x + (axis == 0), y + (axis == 1), z + (axis == 2), gbx, gby, gbz);
    const float tnum = __fsub_rn(m.iso, fa), tden = __fsub_rn(fb, fa);
    const float t01 = fabsf(tden) > 1e-30f ? __fmul_rn(tnum, rcpApproxF(tden)) : __fdiv_rn(tnum, tden);
    float p[3];
    p[0] = __fadd_rn(__fmul_rn((float)x, m.sd[0]), m.org[0]);
    p[1] = __fadd_rn(__fmul_rn((float)y, m.sd[1]), m.org[1]);
    p[2] = __fadd_rn(__fmul_rn((float)z, m.sd[2]), m.org[2]);
    const int c = axis == 0 ? x : (axis == 1 ? y : z);
    const float pa = p[axis], pb = __fadd_rn(__fmul_rn((float)(c + 1), m.sd[axis]), m.org[axis]);
    p[axis] = __fadd_rn(pa, __fmul_rn(t01, __fsub_rn(pb, pa)));
    const float gx = __fadd_rn(gax, __fmul_rn(t01, __fsub_rn(gbx, gax)));
    const float gy = __fadd_rn(gay, __fmul_rn(t01, __fsub_rn(gby, gay)));
    const float gz = __fadd_rn(gaz, __fmul_rn(t01, __fsub_rn(gbz, gaz)));
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    const float inv = len2 > 0.0f ? -rsqrtApproxF(len2) : 0.0f;
    pos[0] = p[0], pos[1] = p[1], pos[2] = p[2];
    nrm[0] = __fmul_rn(gx, inv), nrm[1] = __fmul_rn(gy, inv), nrm[2] = __fmul_rn(gz, inv);
}

/** Node segments: nsv = ceil(sx / 32) per node row, sy rows, sz planes.  EMIT == false: vcount[s] = crossed edges owned by the nodes
 *  of segment s.  EMIT == true: the vertices, at voff[s] + rank. */
template<bool EMIT>
__global__ void __launch_bounds__(MCX_THREADS) mcx_vertex_kernel(McGeo m, const float* __restrict__ vol, unsigned* __restrict__ vcount,
    const unsigned* __restrict__ voff, float* __restrict__ vpos, float* __restrict__ vnrm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsv = (m.sx + 31) >> 5;
    const int xs = blockIdx.x, y = blockIdx.y * MCX_WARPS + warp, z = blockIdx.z;
    if (y >= m.sy) return;
    const int x = xs * 32 + lane;
    const bool valid = x < m.sx;
    const int xc = min(x, m.sx - 1), x1 = min(x + 1, m.sx - 1), y1 = min(y + 1, m.sy - 1), z1 = min(z + 1, m.szGlobal - 1);
    // clamped neighbours: an edge that leaves the grid compares a node with itself and is never crossed
    const float f0 = mcxSample(m, vol, xc, y, z), fx = mcxSample(m, vol, x1, y, z), fy = mcxSample(m, vol, xc, y1, z), fz = mcxSample(m, vol, xc, y, z1);
    const bool b0 = f0 < m.iso;
    const bool cx = valid && (b0 != (fx < m.iso)), cy = valid && (b0 != (fy < m.iso)), cz = valid && (b0 != (fz < m.iso));
    const unsigned mx = __ballot_sync(0xffffffffu, cx), my = __ballot_sync(0xffffffffu, cy), mz = __ballot_sync(0xffffffffu, cz);
    const size_t seg = xs + static_cast<size_t>(nsv) * (y + static_cast<size_t>(m.sy) * z);
    if (!EMIT) {
        if (lane == 0) vcount[seg] = __popc(mx) + __popc(my) + __popc(mz);
        return;
    }
    if (!(mx | my | mz)) return;
    const unsigned lt = (1u << lane) - 1u;
    size_t id = voff[seg] + __popc(mx & lt) + __popc(my & lt) + __popc(mz & lt);
    if (cx) { mcxEdgeVertex(m, vol, x, y, z, 0, f0, fx, vpos + 3 * id, vnrm + 3 * id); ++id; }
    if (cy) { mcxEdgeVertex(m, vol, x, y, z, 1, f0, fy, vpos + 3 * id, vnrm + 3 * id); ++id; }
    if (cz) mcxEdgeVertex(m, vol, x, y, z, 2, f0, fz, vpos + 3 * id, vnrm + 3 * id);
}

struct McxRow {            // one node row (y + a, z + b) of a cell row's neighbourhood
    unsigned mx, my, mz;   // crossing masks of nodes 0..31
    unsigned off, offNext; // vertOffset of the row's segment and of the next segment (node 32 is its node 0)
    unsigned bits32;       // node 32: bit 0 = its x-edge is crossed, bit 1 = its y-edge
    unsigned pad[2];
};

/** A warp per cell row (32 cells): segOffset = exclusive scan of mc_count_kernel's per-row triangle counts. */
__global__ void __launch_bounds__(MCX_THREADS) mcx_index_kernel(McGeo m, const float* __restrict__ vol, const unsigned* __restrict__ segOffset,
    const unsigned* __restrict__ voff, unsigned* __restrict__ indices) {
    __shared__ McxRow sRow[MCX_WARPS][4];
    __shared__ unsigned short sOwner[MCX_WARPS][E_MAXROWTRIS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xs = blockIdx.x, y = blockIdx.y * MCX_WARPS + warp, z = m.cz0 + blockIdx.z;
    if (y >= m.cy) return;
    const size_t cseg = xs + static_cast<size_t>(m.nsegx) * (y + static_cast<size_t>(m.cy) * (z - m.cz0));
    const unsigned segOff = segOffset[cseg], segTris = segOffset[cseg + 1] - segOff;
    if (segTris == 0) return;
    const int nsv = (m.sx + 31) >> 5;
    const int x0 = xs * 32;
    // "below iso" bits of node rows (y + a, z + b), a, b = 0..2 (not (2, 2)), nodes x0 .. x0 + 33; clamped at the border
    unsigned lo[3][3], hi[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[b][a] = hi[b][a] = 0;
            if (a == 2 && b == 2) continue;
            const int yy = min(y + a, m.sy - 1), zz = min(z + b, m.szGlobal - 1);
            const float v = mcxSample(m, vol, min(x0 + lane, m.sx - 1), yy, zz);
            const float w = lane < 2 ? mcxSample(m, vol, min(x0 + 32 + lane, m.sx - 1), yy, zz) : 0.0f;
            lo[b][a] = __ballot_sync(0xffffffffu, v < m.iso);
            hi[b][a] = __ballot_sync(0xffffffffu, lane < 2 && w < m.iso);
        }
    if (lane < 4) {
        const int a = lane & 1, b = lane >> 1;
        // (runtime indices into the unrolled arrays: select)
        unsigned L = 0, H = 0, Ly = 0, Hy = 0, Lz = 0, Hz = 0;
#pragma unroll
        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
                if (aa == a && bb == b) L = lo[bb][aa], H = hi[bb][aa];
                if (aa == a + 1 && bb == b) Ly = lo[bb][aa], Hy = hi[bb][aa];
                if (aa == a && bb == b + 1) Lz = lo[bb][aa], Hz = hi[bb][aa];
            }
        McxRow r;
        r.mx = L ^ __funnelshift_r(L, H, 1);
        r.my = L ^ Ly;
        r.mz = L ^ Lz;
        r.bits32 = ((H ^ (H >> 1)) & 1u) | (((H ^ Hy) & 1u) << 1);
        const int yy = min(y + a, m.sy - 1), zz = min(z + b, m.szGlobal - 1);
        const size_t s = xs + static_cast<size_t>(nsv) * (yy + static_cast<size_t>(m.sy) * zz);
        r.off = voff[s];
        r.offNext = xs + 1 < nsv ? voff[s + 1] : 0u;
        r.pad[0] = r.pad[1] = 0;
        sRow[warp][lane] = r;
    }
    // cube index (permuted order, as mc_count_kernel / mc_emit_kernel) from the (x, x+1) bit pairs of the four node rows
    const unsigned p00 = __funnelshift_r(lo[0][0], hi[0][0], lane) & 3u, p10 = __funnelshift_r(lo[0][1], hi[0][1], lane) & 3u;
    const unsigned p01 = __funnelshift_r(lo[1][0], hi[1][0], lane) & 3u, p11 = __funnelshift_r(lo[1][1], hi[1][1], lane) & 3u;
    unsigned long long word = 0;
    if (x0 + lane < m.cx) word = __ldg(&kCasePerm.w[p00 | p10 << 2 | p01 << 4 | p11 << 6]);
    const unsigned n = static_cast<unsigned>(word) & 15u;
    unsigned inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    const unsigned first = inc - n;
    unsigned short* owner = sOwner[warp];
#pragma unroll
    for (unsigned q = 0; q < 5; ++q)
        if (q < n) owner[first + q] = static_cast<unsigned short>(lane << 12 | (static_cast<unsigned>(word >> (4 + 12 * q)) & 0xfffu));
    const unsigned tHalf = __shfl_sync(0xffffffffu, first, 16); // the table keeps 4 lane bits: triangles from here on belong to lanes 16..31
    __syncwarp();
    const unsigned ncorn = segTris * 3;
    unsigned* out = indices + static_cast<size_t>(segOff) * 3;
    for (unsigned jc = lane; jc < ncorn; jc += 32) {
        const unsigned t = jc / 3;
        const unsigned ok = owner[t];
        const unsigned L = (ok >> 12) + (t >= tHalf ? 16u : 0u);
        const unsigned e = (ok >> (4 * (jc - 3 * t))) & 15u;
        const unsigned code = edgeCode(e);
        const unsigned i = L + (code & 1u), axis = code >> 3;
        const McxRow& r = sRow[warp][(code >> 1) & 3u]; // row (dy, dz)
        unsigned id;
        if (i < 32u) {
            const unsigned lt = (1u << i) - 1u;
            id = r.off + __popc(r.mx & lt) + __popc(r.my & lt) + __popc(r.mz & lt);
            if (axis >= 1u) id += (r.mx >> i) & 1u;
            if (axis == 2u) id += (r.my >> i) & 1u;
        } else {
            id = r.offNext + (axis >= 1u ? (r.bits32 & 1u) : 0u) + (axis == 2u ? (r.bits32 >> 1) & 1u : 0u);
        }
        out[jc] = id;
    }
}

} // namespace mms
