// mt.cuh -- marching TETRAHEDRA exactly as trisoup_gl::volumetrics::IsoSurface emits them (compatibility mode).
//
// The reference's isosurface module is not marching cubes: it splits every cell into six tetrahedra (IsoSurface.cpp:30-31),
// finds each vertex by regula falsi on the TRILINEAR interpolant along the tetrahedron edge (:430-465), orients the triangles with
// a half-space test (:606-735) and gives every triangle one flat normal.  For users who need triangle-for-triangle equality with
// the CPU module, these kernels reproduce that output bit for bit: every fp32 operation is an individually rounded intrinsic (the
// reference's baseline x86-64 build has no FMA), the cell frame is the reference's half-voxel shifted one (:238-252), the order is
// its loop order (cells x-fastest, tetrahedra 0..5, tri then tri2).
//   mt_count_kernel   a warp owns a 32-cell x-segment: per-cell triangle count (0..12), warp sum -> segCount (same layout as mc.cuh)
//   mt_emit_kernel    the same mapping; warp scan of the counts, every lane writes its cell's triangles
// It is a compatibility path: correct and parallel, not tuned like the marching-cubes kernels.
#pragma once
#include "common.cuh"

namespace mms {

struct MtGeo {
    int sx, sy;        // volume resolution in x, y
    int zPlane0;       // global z index of plane 0 of the volume (slab)
    int szGlobal;      // global z resolution
    int cx, cy;        // cells in x, y
    int cz0, cnz;      // global cell layers [cz0, cz0 + cnz) of this context
    int nsegx;
    float mn[3];       // osbb Left, Bottom, Back
    float ext[3];      // osbb Width, Height, Depth
    float cell[3];     // ext / float(s)   (IsoSurface.cpp:238-240)
    float iso;
};

constexpr int MT_THREADS = 256;

namespace mt {
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dvd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ bool isEq(float m, float n) { return fabsf(sub(m, n)) < 1e-5f; } // vislib FLOAT_EPSILON (mathfunctions.h:119-136)

struct P3 {
    float x, y, z;
};

// corner j of the cube: a2fVertexOffset (MarchingCubeTables.cpp:11-12) as bits x | y<<1 | z<<2
__device__ __forceinline__ unsigned cornerBits(unsigned j) { return (0x67542310u >> (4 * j)) & 7u; } // corners 0..7 -> 0,1,3,2,4,5,7,6
__device__ __forceinline__ unsigned tetCorner(unsigned tet, unsigned k) { // IsoSurface::tets
    // {0,2,3,7}, {0,2,6,7}, {0,4,6,7}, {0,6,1,2}, {0,6,1,4}, {5,6,1,4}: 4 bits per entry, k-th nibble of the tet's 16-bit word
    const unsigned long long lo = 0x7320ull | (0x7620ull << 16) | (0x7640ull << 32) | (0x2160ull << 48);
    const unsigned hi = 0x4160u | (0x4165u << 16);
    const unsigned w = tet < 4 ? static_cast<unsigned>(lo >> (16 * tet)) & 0xffffu : (hi >> (16 * (tet - 4))) & 0xffffu;
    return (w >> (4 * k)) & 15u;
}

struct Cell {
    float cv[8];
    float p0[3], p1[3]; // corner coordinates for offset 0 / 1 per axis: p + float(off) * cellSize
};

__device__ __forceinline__ P3 cornerPoint(const Cell& c, unsigned j) {
    const unsigned b = cornerBits(j);
    P3 r;
    r.x = (b & 1u) ? c.p1[0] : c.p0[0];
    r.y = (b & 2u) ? c.p1[1] : c.p0[1];
    r.z = (b & 4u) ? c.p1[2] : c.p0[2];
    return r;
}

__device__ __forceinline__ float offsetOf(float v1, float v2, float want) { return dvd(sub(want, v1), sub(v2, v1)); } // IsoSurface::getOffset

__device__ __forceinline__ float valueAt(const float* cv, unsigned i0, unsigned i1, float a) { // getValue (IsoSurface.cpp:405-424)
    const float b = sub(1.0f, a);
    const unsigned c0 = cornerBits(i0), c1 = cornerBits(i1);
    const float x = add(mul(b, (float)(c0 & 1u)), mul(a, (float)(c1 & 1u)));
    const float y = add(mul(b, (float)((c0 >> 1) & 1u)), mul(a, (float)((c1 >> 1) & 1u)));
    const float z = add(mul(b, (float)((c0 >> 2) & 1u)), mul(a, (float)((c1 >> 2) & 1u)));
    const float mx = sub(1.0f, x), my = sub(1.0f, y), mz = sub(1.0f, z);
    float v0 = add(mul(mx, cv[0]), mul(x, cv[1]));
    const float v1 = add(mul(mx, cv[3]), mul(x, cv[2]));
    float v2 = add(mul(mx, cv[4]), mul(x, cv[5]));
    const float v3 = add(mul(mx, cv[7]), mul(x, cv[6]));
    v0 = add(mul(my, v0), mul(y, v1));
    v2 = add(mul(my, v2), mul(y, v3));
    return add(mul(mz, v0), mul(z, v2));
}

__device__ __noinline__ P3 interpolate(const Cell& c, float val, unsigned i0, unsigned i1) { // IsoSurface::interpolate (:430-465)
    float a0 = 0.0f;
    float v0 = valueAt(c.cv, i0, i1, a0);
    if (isEq(v0, val)) return cornerPoint(c, i0);
    float a1 = 1.0f;
    float v1 = valueAt(c.cv, i0, i1, a1);
    if (isEq(v1, val)) return cornerPoint(c, i1);
    float a = offsetOf(c.cv[i0], c.cv[i1], val);
    float v = valueAt(c.cv, i0, i1, a);
    unsigned maxStep = 100;
    const bool flip = c.cv[i0] > c.cv[i1];
    while (maxStep > 0 && !isEq(v, val)) {
        if ((!flip && v > val) || (flip && v < val)) a1 = a, v1 = v;
        else a0 = a, v0 = v;
        a = add(a0, mul(offsetOf(v0, v1, val), sub(a1, a0)));
        v = valueAt(c.cv, i0, i1, a);
        --maxStep;
    }
    const P3 pa = cornerPoint(c, i0), pb = cornerPoint(c, i1);
    const float at = sub(1.0f, a); // AbstractPointImpl::Interpolate: this * (1 - t) + rhs * t
    P3 r;
    r.x = add(mul(pa.x, at), mul(pb.x, a));
    r.y = add(mul(pa.y, at), mul(pb.y, a));
    r.z = add(mul(pa.z, at), mul(pb.z, a));
    return r;
}

__device__ __forceinline__ void cross(const P3& a, const P3& b, P3& r) { // AbstractVector<T,3>::Cross
    r.x = sub(mul(a.y, b.z), mul(a.z, b.y));
    r.y = sub(mul(a.z, b.x), mul(a.x, b.z));
    r.z = sub(mul(a.x, b.y), mul(a.y, b.x));
}
__device__ __forceinline__ void normalise(P3& v) { // AbstractVectorImpl::Normalise / Length
    const float l = __fsqrt_rn(add(add(mul(v.x, v.x), mul(v.y, v.y)), mul(v.z, v.z)));
    if (l != 0.0f) v.x = dvd(v.x, l), v.y = dvd(v.y, l), v.z = dvd(v.z, l);
    else v.x = v.y = v.z = 0.0f;
}
__device__ __forceinline__ P3 diff(const P3& a, const P3& b) { return P3{sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }

/** triangles per tetrahedron case: 0 for 0/15, 2 for 3, 5, 6, 9, 10, 12, else 1 */
__device__ __forceinline__ unsigned tetTriangles(unsigned triIdx) { return (0x16696994u >> (2 * triIdx)) & 3u; }

__device__ __forceinline__ unsigned tetCase(const float* cv, unsigned tet, float val) {
    unsigned triIdx = 0;
#pragma unroll
    for (unsigned k = 0; k < 4; ++k)
        if (cv[tetCorner(tet, k)] < val) triIdx |= 1u << k;
    return triIdx;
}

/** IsoSurface::makeTet(triIdx, tetIdx, ...) (:606-735); returns the number of triangles */
__device__ __noinline__ int makeTet(unsigned triIdx, unsigned tet, const Cell& c, float val, P3 (&tri)[2][3]) {
    const unsigned T0 = tetCorner(tet, 0), T1 = tetCorner(tet, 1), T2 = tetCorner(tet, 2), T3 = tetCorner(tet, 3);
    const P3 p0 = cornerPoint(c, T0), p1 = cornerPoint(c, T1), p2 = cornerPoint(c, T2), p3 = cornerPoint(c, T3);
    P3 n;
    cross(diff(p2, p1), diff(p3, p1), n);
    normalise(n);
    // Plane(p1, norm) (AbstractPlane::Set :572-578), Halfspace(p0) via the normalised parameters (:440-469, :668-679)
    const float pd = mul(-1.0f, add(add(mul(n.x, p1.x), mul(n.y, p1.y)), mul(n.z, p1.z)));
    const float len = __fsqrt_rn(add(add(mul(n.x, n.x), mul(n.y, n.y)), mul(n.z, n.z)));
    float A = 0.0f, B = 0.0f, C = 0.0f, D = 0.0f;
    if (!isEq(len, 0.0f)) A = dvd(n.x, len), B = dvd(n.y, len), C = dvd(n.z, len), D = dvd(pd, len);
    const float dist = add(add(add(mul(A, p0.x), mul(B, p0.y)), mul(C, p0.z)), D);
    bool flip = !isEq(dist, 0.0f) && dist > 0.0f;
    const unsigned T[4] = {T0, T1, T2, T3};
    auto I = [&](int a, int b) { return interpolate(c, val, T[a], T[b]); };
    // every case: (a0,b0) -> tri[0][0]; (a1,b1) -> tri[0][flip?2:1]; (a2,b2) -> tri[0][flip?1:2]; two-triangle cases add one more vertex
    int n1 = 0;
    switch (triIdx) {
    case 0x00: case 0x0F: break;
    case 0x01: flip = !flip; // fall through
    case 0x0E: tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(0, 2); tri[0][flip ? 1 : 2] = I(0, 3); n1 = 1; break;
    case 0x02: flip = !flip; // fall through
    case 0x0D: tri[0][0] = I(1, 0); tri[0][flip ? 2 : 1] = I(1, 3); tri[0][flip ? 1 : 2] = I(1, 2); n1 = 1; break;
    case 0x0C: flip = !flip; // fall through
    case 0x03:
        tri[0][0] = I(0, 3); tri[0][flip ? 2 : 1] = I(0, 2); tri[0][flip ? 1 : 2] = I(1, 3);
        tri[1][0] = tri[0][flip ? 1 : 2]; tri[1][flip ? 1 : 2] = I(1, 2); tri[1][flip ? 2 : 1] = tri[0][flip ? 2 : 1];
        n1 = 2; break;
    case 0x04: flip = !flip; // fall through
    case 0x0B: tri[0][0] = I(2, 0); tri[0][flip ? 2 : 1] = I(2, 1); tri[0][flip ? 1 : 2] = I(2, 3); n1 = 1; break;
    case 0x05: flip = !flip; // fall through
    case 0x0A:
        tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(2, 3); tri[0][flip ? 1 : 2] = I(0, 3);
        tri[1][0] = tri[0][0]; tri[1][flip ? 2 : 1] = I(1, 2); tri[1][flip ? 1 : 2] = tri[0][flip ? 2 : 1];
        n1 = 2; break;
    case 0x06: flip = !flip; // fall through
    case 0x09:
        tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(1, 3); tri[0][flip ? 1 : 2] = I(2, 3);
        tri[1][0] = tri[0][0]; tri[1][flip ? 1 : 2] = I(0, 2); tri[1][flip ? 2 : 1] = tri[0][flip ? 1 : 2];
        n1 = 2; break;
    case 0x08: flip = !flip; // fall through
    case 0x07: tri[0][0] = I(3, 0); tri[0][flip ? 2 : 1] = I(3, 2); tri[0][flip ? 1 : 2] = I(3, 1); n1 = 1; break;
    }
    return n1;
}

/** cube values of cell (x, y, z global) and whether the cell is crossed at all (IsoSurface.cpp:254-265) */
__device__ __forceinline__ bool loadCell(const MtGeo& m, const float* __restrict__ vol, int x, int y, int z, Cell& c) {
    bool bigger = false, smaller = false;
#pragma unroll
    for (unsigned j = 0; j < 8; ++j) {
        const unsigned b = cornerBits(j);
        c.cv[j] = vol[(x + (b & 1u)) + static_cast<size_t>(m.sx) * ((y + ((b >> 1) & 1u)) + static_cast<size_t>(m.sy) * (z + ((b >> 2) & 1u) - m.zPlane0))];
        bigger = bigger || (c.cv[j] >= m.iso);
        smaller = smaller || (c.cv[j] < m.iso);
    }
    return bigger && smaller;
}

__device__ __forceinline__ unsigned cellTriangles(const Cell& c, float val) {
    unsigned n = 0;
#pragma unroll
    for (unsigned tet = 0; tet < 6; ++tet) n += tetTriangles(tetCase(c.cv, tet, val));
    return n;
}
} // namespace mt

__global__ void __launch_bounds__(MT_THREADS) mt_count_kernel(MtGeo m, const float* __restrict__ vol, unsigned* __restrict__ segCount,
    unsigned char* __restrict__ triCount) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xseg = blockIdx.x, y = blockIdx.y * (MT_THREADS / 32) + warp, z = m.cz0 + blockIdx.z;
    if (y >= m.cy) return;
    const int x = xseg * 32 + lane;
    unsigned n = 0;
    if (x < m.cx) {
        mt::Cell c;
        if (mt::loadCell(m, vol, x, y, z, c)) n = mt::cellTriangles(c, m.iso);
        if (triCount) triCount[x + static_cast<size_t>(m.cx) * (y + static_cast<size_t>(m.cy) * (z - m.cz0))] = static_cast<unsigned char>(n);
    }
    const unsigned tot = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0) segCount[xseg + static_cast<size_t>(m.nsegx) * (y + static_cast<size_t>(m.cy) * (z - m.cz0))] = tot;
}

__global__ void __launch_bounds__(MT_THREADS) mt_emit_kernel(MtGeo m, const float* __restrict__ vol, const unsigned* __restrict__ segOffset,
    float* __restrict__ outPos, float* __restrict__ outNrm) {
    using namespace mt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int xseg = blockIdx.x, y = blockIdx.y * (MT_THREADS / 32) + warp, z = m.cz0 + blockIdx.z;
    if (y >= m.cy) return;
    const size_t seg = xseg + static_cast<size_t>(m.nsegx) * (y + static_cast<size_t>(m.cy) * (z - m.cz0));
    const unsigned first = segOffset[seg];
    if (segOffset[seg + 1] == first) return;
    const int x = xseg * 32 + lane;
    Cell c;
    unsigned n = 0;
    if (x < m.cx && loadCell(m, vol, x, y, z, c)) n = cellTriangles(c, m.iso);
    unsigned inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (n == 0) return;
    // cell frame: p = ((idx + 0.5) / s) * extent + min;  corner = p + float(offset) * cellSize   (IsoSurface.cpp:242-252, 293-302)
    const int idx[3] = {x, y, z};
    const int s[3] = {m.sx, m.sy, m.szGlobal};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p = add(mul(dvd(add((float)idx[a], 0.5f), (float)s[a]), m.ext[a]), m.mn[a]);
        c.p0[a] = add(p, mul(0.0f, m.cell[a]));
        c.p1[a] = add(p, mul(1.0f, m.cell[a]));
    }
    size_t o = static_cast<size_t>(first + inc - n) * 9;
    for (unsigned tet = 0; tet < 6; ++tet) {
        const unsigned triIdx = tetCase(c.cv, tet, m.iso);
        if (tetTriangles(triIdx) == 0) continue;
        P3 tri[2][3];
        const int nt = makeTet(triIdx, tet, c, m.iso, tri);
        for (int t = 0; t < nt; ++t) {
            P3 nn;
            cross(diff(tri[t][1], tri[t][0]), diff(tri[t][2], tri[t][0]), nn);
            normalise(nn);
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                outPos[o + 3 * i] = tri[t][i].x, outPos[o + 3 * i + 1] = tri[t][i].y, outPos[o + 3 * i + 2] = tri[t][i].z;
                outNrm[o + 3 * i] = nn.x, outNrm[o + 3 * i + 1] = nn.y, outNrm[o + 3 * i + 2] = nn.z;
            }
            o += 9;
        }
    }
}

} // namespace mms
