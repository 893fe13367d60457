// mc_indexed.cuh -- marching cubes with an INDEXED triangle mesh (opt-in; the reference's contract is the unindexed soup of mc.cuh).
//
// trisoup's CallTriMeshData carries indexed meshes as well (SetVertexData + SetTriangleData with 32-bit indices,
// plugins/geometry_calls_gl/include/geometry_calls_gl/CallTriMeshDataGL.h:897-1000): every crossed grid edge becomes ONE vertex
// (24 bytes: position + normal) and a triangle is three indices (12 bytes) -- ~28 bytes per triangle instead of the soup's 72,
// for the HBM writes, the PCIe read-back and the renderer's vertex fetch alike.
//
//   vertex numbering   a crossed edge belongs to its LOW node; the nodes are walked in rows of 32 ("node segments", x fastest, then y,
//                      then z), inside a segment node by node, per node x-edge, y-edge, z-edge.  So
//                          id(node i of segment s, axis a) = vertOffset[s] + popc(mx & lt_i) + popc(my & lt_i) + popc(mz & lt_i)
//                                                            + (a >= 1 ? mx_i : 0) + (a == 2 ? my_i : 0)
//                      with mx/my/mz the segment's crossing masks -- any kernel that knows the "below iso" bits around a cell can name
//                      the vertex of each of its edges without a search or a hash.
//   mask records       per node segment the "below iso" bits and the three crossing masks (one 16-byte record: the later kernels never touch
//                      the volume again to classify) and the segment's crossing count (-> exclusive scan = vertOffset).  mc_count_kernel
//                      has all these bits in its registers and writes the records as a by-product; mcx_mask_kernel (a warp per node
//                      segment) only serves the node rows without a cell row: the last row, the last plane, a lone last segment
//   mcx_vertex_kernel  a warp per node segment: its crossings, compacted in vertex order, are evaluated by full lanes with exactly the
//                      arithmetic of mc_emit_kernel's V stage (gradient normal by central differences clamped at the global border,
//                      SFU reciprocal for the interpolation parameter): position and normal are BIT-IDENTICAL to the soup's corners;
//                      vertex c of the segment goes to slot vertOffset + c: consecutive lanes write consecutive 12-byte pieces
//   mcx_index_kernel   a warp per 32-cell row: cube indices from the mask records of its four node rows, the row's triangle corners
//                      flattened over the lanes (as in mc_emit_kernel), each lane writes one 32-bit index: coalesced 128-byte stores
// Triangle order = the soup's (cell-linear, the table's order inside a cell): expanding the indices reproduces the soup exactly.
#pragma once
#include "mc.cuh"

namespace mms {

constexpr int MCX_THREADS = 256;
constexpr int MCX_WARPS = MCX_THREADS / 32;

/** Node segments: nsv = ceil(sx / 32) per node row, sy rows, sz planes.  rec[s] = {below, mx, my, mz}: bit i <=> node 32 xs + i is below
 *  the iso value / its x-, y-, z-edge (towards the next node) is crossed; vcount[s] = crossed edges owned by the segment's nodes. */
struct McxRange { // node segments [xs0, xs0 + gridDim.x), node rows [y0, y1), planes [z0, z0 + gridDim.z)
    int xs0, y0, y1, z0;
};
/** mc_count_kernel writes the records of every node row that is the low row of a cell row (y < sy-1, z < sz-1, segments that hold a
 *  cell); this kernel serves the rest: the last node row, the last plane and a lone last node segment (sx = 32 k + 1). */
__global__ void __launch_bounds__(MCX_THREADS) mcx_mask_kernel(McGeo m, McxRange rg, const float* __restrict__ vol, uint4* __restrict__ rec,
    unsigned* __restrict__ vcount) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsv = (m.sx + 31) >> 5;
    const int xs = rg.xs0 + blockIdx.x, y = rg.y0 + blockIdx.y * MCX_WARPS + warp, z = rg.z0 + blockIdx.z;
    if (y >= rg.y1) return;
    const int x = xs * 32 + lane;
    const bool valid = x < m.sx;
    const int xc = min(x, m.sx - 1);
    const size_t plane = static_cast<size_t>(m.sx) * m.sy;
    const float* p = vol + xc + static_cast<size_t>(m.sx) * y + plane * (z - m.zPlane0);
    // clamped neighbours: an edge that leaves the grid compares a node with itself and is never crossed
    const float f0 = __ldg(p), fx = __ldg(p + (x + 1 < m.sx ? 1 : 0)), fy = __ldg(p + (y + 1 < m.sy ? m.sx : 0));
    const float fz = __ldg(p + (z + 1 < m.szGlobal ? plane : 0));
    const bool b0 = f0 < m.iso;
    const unsigned below = __ballot_sync(0xffffffffu, b0);
    const unsigned mx = __ballot_sync(0xffffffffu, valid && (b0 != (fx < m.iso)));
    const unsigned my = __ballot_sync(0xffffffffu, valid && (b0 != (fy < m.iso)));
    const unsigned mz = __ballot_sync(0xffffffffu, valid && (b0 != (fz < m.iso)));
    if (lane == 0) {
        const size_t seg = xs + static_cast<size_t>(nsv) * (y + static_cast<size_t>(m.sy) * z);
        rec[seg] = make_uint4(below, mx, my, mz);
        vcount[seg] = __popc(mx) + __popc(my) + __popc(mz);
    }
}

/** The vertex on the crossed edge from node A = (x, y, z) to A + unit(axis).  Operation for operation the V stage of mc_emit_kernel
 *  (mc.cuh): gradient (f(+) - f(-)) * 1/(n*sd) with the samples clamped at the GLOBAL grid border (one-sided there).  Branch-free: the border only changes offsets and the factor. */
__device__ __forceinline__ void mcxEdgeVertex(const McGeo& m, const float* __restrict__ pA, int plane, int x, int y, int z, int axis,
    float* __restrict__ pos, float* __restrict__ nrm) {
    const int sx = m.sx; // (32-bit element offsets: the host admits the indexed mesh only for planes below 2^31 voxels)
    auto gradient = [&](const float* __restrict__ p, int gx_, int gy_, int gz_, float& gx, float& gy, float& gz) {
        const int xm = gx_ > 0, xp = gx_ < m.sx - 1, ym = gy_ > 0, yp = gy_ < m.sy - 1, zm = gz_ > 0, zp = gz_ < m.szGlobal - 1;
        // (no dynamic index into the parameter struct: one-sided and central factors by select; on an axis of a single node both
        //  samples are the node itself and the difference is an exact 0)
        gx = __fmul_rn(__fsub_rn(__ldg(p + xp), __ldg(p - xm)), (xp & xm) ? m.rinv[0][2] : m.rinv[0][1]);
        gy = __fmul_rn(__fsub_rn(__ldg(p + yp * sx), __ldg(p - ym * sx)), (yp & ym) ? m.rinv[1][2] : m.rinv[1][1]);
        gz = __fmul_rn(__fsub_rn(__ldg(p + zp * plane), __ldg(p - zm * plane)), (zp & zm) ? m.rinv[2][2] : m.rinv[2][1]);
    };
    const int ax = axis == 0, ay = axis == 1, az = axis == 2;
    const float* pB = pA + (ax ? 1 : (ay ? sx : plane));
    float gax, gay, gaz, gbx, gby, gbz;
    gradient(pA, x, y, z, gax, gay, gaz);
    gradient(pB, x + ax, y + ay, z + az, gbx, gby, gbz);
    const float fa = __ldg(pA), fb = __ldg(pB);
    const float tnum = __fsub_rn(m.iso, fa), tden = __fsub_rn(fb, fa);
    const float t01 = fabsf(tden) > 1e-30f ? __fmul_rn(tnum, rcpApproxF(tden)) : __fdiv_rn(tnum, tden);
    const float px = __fadd_rn(__fmul_rn((float)x, m.sd[0]), m.org[0]), qx = __fadd_rn(__fmul_rn((float)(x + 1), m.sd[0]), m.org[0]);
    const float py = __fadd_rn(__fmul_rn((float)y, m.sd[1]), m.org[1]), qy = __fadd_rn(__fmul_rn((float)(y + 1), m.sd[1]), m.org[1]);
    const float pz = __fadd_rn(__fmul_rn((float)z, m.sd[2]), m.org[2]), qz = __fadd_rn(__fmul_rn((float)(z + 1), m.sd[2]), m.org[2]);
    const float pa = ax ? px : (ay ? py : pz), pb = ax ? qx : (ay ? qy : qz);
    const float pc = __fadd_rn(pa, __fmul_rn(t01, __fsub_rn(pb, pa)));
    const float gx = __fadd_rn(gax, __fmul_rn(t01, __fsub_rn(gbx, gax)));
    const float gy = __fadd_rn(gay, __fmul_rn(t01, __fsub_rn(gby, gay)));
    const float gz = __fadd_rn(gaz, __fmul_rn(t01, __fsub_rn(gbz, gaz)));
    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
    const float inv = len2 > 0.0f ? -rsqrtApproxF(len2) : 0.0f;
    pos[0] = ax ? pc : px, pos[1] = ay ? pc : py, pos[2] = az ? pc : pz;
    nrm[0] = __fmul_rn(gx, inv), nrm[1] = __fmul_rn(gy, inv), nrm[2] = __fmul_rn(gz, inv);
}

__global__ void __launch_bounds__(MCX_THREADS) mcx_vertex_kernel(McGeo m, const float* __restrict__ vol, const uint4* __restrict__ rec,
    const unsigned* __restrict__ voff, float* __restrict__ vpos, float* __restrict__ vnrm) {
    __shared__ unsigned char sList[MCX_WARPS][96];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsv = (m.sx + 31) >> 5;
    const int xs = blockIdx.x, y = blockIdx.y * MCX_WARPS + warp, z = blockIdx.z;
    if (y >= m.sy) return;
    const size_t seg = xs + static_cast<size_t>(nsv) * (y + static_cast<size_t>(m.sy) * z);
    const uint4 r = __ldg(rec + seg);
    const unsigned n = __popc(r.y) + __popc(r.z) + __popc(r.w);
    if (n == 0) return;
    // the segment's crossings in vertex order (node by node; x-, y-, z-edge): entry = node << 2 | axis
    unsigned char* list = sList[warp];
    const unsigned lt = (1u << lane) - 1u;
    unsigned k = __popc(r.y & lt) + __popc(r.z & lt) + __popc(r.w & lt);
    if ((r.y >> lane) & 1u) list[k++] = static_cast<unsigned char>(lane << 2);
    if ((r.z >> lane) & 1u) list[k++] = static_cast<unsigned char>(lane << 2 | 1);
    if ((r.w >> lane) & 1u) list[k] = static_cast<unsigned char>(lane << 2 | 2);
    __syncwarp();
    const int plane = m.sx * m.sy;
    const float* row = vol + xs * 32 + static_cast<long long>(m.sx) * y + static_cast<long long>(plane) * (z - m.zPlane0);
    const size_t base = voff[seg];
#pragma unroll 1
    for (unsigned c = lane; c < n; c += 32) {
        const unsigned code = list[c];
        const int i = code >> 2, axis = code & 3;
        const size_t id = base + c;
        mcxEdgeVertex(m, row + i, plane, xs * 32 + i, y, z, axis, vpos + 3 * id, vnrm + 3 * id);
    }
}

struct McxRow {                 // one node row (y + a, z + b) of a cell row's neighbourhood
    unsigned below, mx, my, mz; // the row's segment record
    unsigned off, offNext;      // vertOffset of the row's segment and of the next segment (node 32 is its node 0)
    unsigned bits32;            // node 32: bit 0 = below iso, bit 1 = its x-edge is crossed, bit 2 = its y-edge
    unsigned pad;
};

/** A warp per cell row (32 cells): segOffset = exclusive scan of mc_count_kernel's per-row triangle counts. */
__global__ void __launch_bounds__(MCX_THREADS) mcx_index_kernel(McGeo m, const uint4* __restrict__ rec, const unsigned* __restrict__ segOffset,
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
    // lanes 0..3: the segment records of node rows (y + a, z + b), a = lane & 1, b = lane >> 1; lanes 4..7: of the next segment in x
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    unsigned off = 0;
    if (lane < 8) {
        const int a = lane & 1, b = (lane >> 1) & 1, nx = lane >> 2;
        if (xs + nx < nsv) {
            const size_t s = xs + nx + static_cast<size_t>(nsv) * (y + a + static_cast<size_t>(m.sy) * (z + b));
            q = __ldg(rec + s);
            off = voff[s];
        }
    }
    const unsigned nBelow = __shfl_down_sync(0xffffffffu, q.x, 4), nMx = __shfl_down_sync(0xffffffffu, q.y, 4);
    const unsigned nMy = __shfl_down_sync(0xffffffffu, q.z, 4), nOff = __shfl_down_sync(0xffffffffu, off, 4);
    if (lane < 4) {
        McxRow r;
        r.below = q.x, r.mx = q.y, r.my = q.z, r.mz = q.w, r.off = off, r.offNext = nOff;
        r.bits32 = (nBelow & 1u) | ((nMx & 1u) << 1) | ((nMy & 1u) << 2);
        r.pad = 0;
        sRow[warp][lane] = r;
    }
    __syncwarp();
    // cube index (permuted order, as mc_count_kernel / mc_emit_kernel) from the (x, x+1) bit pairs of the four node rows
    const McxRow* R = sRow[warp];
    const unsigned p00 = __funnelshift_r(R[0].below, R[0].bits32 & 1u, lane) & 3u, p10 = __funnelshift_r(R[1].below, R[1].bits32 & 1u, lane) & 3u;
    const unsigned p01 = __funnelshift_r(R[2].below, R[2].bits32 & 1u, lane) & 3u, p11 = __funnelshift_r(R[3].below, R[3].bits32 & 1u, lane) & 3u;
    unsigned long long word = 0;
    if (xs * 32 + lane < m.cx) word = __ldg(&kCasePerm.w[p00 | p10 << 2 | p01 << 4 | p11 << 6]);
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
    for (unsigned k = 0; k < 5; ++k)
        if (k < n) owner[first + k] = static_cast<unsigned short>(lane << 12 | (static_cast<unsigned>(word >> (4 + 12 * k)) & 0xfffu));
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
        const McxRow& r = R[(code >> 1) & 3u]; // row (dy, dz)
        unsigned id;
        if (i < 32u) {
            const unsigned lt = (1u << i) - 1u;
            id = r.off + __popc(r.mx & lt) + __popc(r.my & lt) + __popc(r.mz & lt);
            if (axis >= 1u) id += (r.mx >> i) & 1u;
            if (axis == 2u) id += (r.my >> i) & 1u;
        } else {
            id = r.offNext + (axis >= 1u ? (r.bits32 >> 1) & 1u : 0u) + (axis == 2u ? (r.bits32 >> 2) & 1u : 0u);
        }
        out[jc] = id;
    }
}

} // namespace mms
