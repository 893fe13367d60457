// mc.cuh -- marching cubes on the device-resident density slab.
//
//   mc_count_kernel   classify every cell (cube index bit i <=> corner i < iso, corner order and case table of
//                     trisoup::volumetrics::MarchingCubeTables, MarchingCubeTables.cpp:11-16,58,278) and sum the
//                     triangle counts of each 32-cell x-segment with a warp reduction
//   (exclusive scan of the segment counts: scan.cuh -> triangle offsets in CELL-LINEAR order)
//   mc_emit_kernel    active tiles only: stage density (+1 halo for gradients) in shared memory, build
//                     {f, grad} per node once, classify again, warp-level prefix scan of the per-cell counts,
//                     interpolate vertices/normals, compact them in shared memory and stream them out with
//                     fully coalesced stores.
// Output order = cell-linear (x fastest, then y, then z), inside a cell the table's order: independent of the
// tile shape and of the z-slab decomposition.
#pragma once
#include "common.cuh"

namespace mms {

__constant__ unsigned long long kCaseWords[256] = {
#include "mc_case_words.inc"
};

constexpr int MCX = 32, MCY = 8, MCZ = 4; // cells per block tile; one warp per (y,z) row of 32 cells
constexpr int MC_THREADS = 256;
constexpr int MC_NX = MCX + 1, MC_NY = MCY + 1, MC_NZ = MCZ + 1;       // nodes of the tile's cells
constexpr int MC_HX = MCX + 3, MC_HY = MCY + 3, MC_HZ = MCZ + 3;       // + gradient halo
constexpr int MC_STAGE = 64;                                            // triangles staged per warp round

struct McGeo {
    int sx, sy;        // volume resolution in x, y
    int nzPlanes;      // planes in the slab volume
    int zPlane0;       // global z index of plane 0 of the volume
    int szGlobal;      // global z resolution (gradient clamps at the GLOBAL border only)
    int cx, cy;        // cells in x, y (= s-1)
    int cz0, cnz;      // cell layers [cz0, cz0+cnz) in GLOBAL z handled by this context
    int nsegx;         // ceil(cx / 32)
    float org[3], sd[3];
    float iso;
};

// per edge: low corner (dx,dy,dz) and axis, 5 bits each: dx | dy<<1 | dz<<2 | axis<<3
//  e0 (0,0,0)x  e1 (1,0,0)y  e2 (0,1,0)x  e3 (0,0,0)y  e4 (0,0,1)x  e5 (1,0,1)y  e6 (0,1,1)x  e7 (0,0,1)y
//  e8 (0,0,0)z  e9 (1,0,0)z  e10 (1,1,0)z e11 (0,1,0)z           (MarchingCubeTables.cpp:15-16, low node first)
__device__ __forceinline__ unsigned edgeCode(int e) {
    const unsigned long long codes = (0ull) | (9ull << 5) | (2ull << 10) | (8ull << 15) | (4ull << 20) | (13ull << 25) |
                                     (6ull << 30) | (12ull << 35) | (16ull << 40) | (17ull << 45) | (19ull << 50) | (18ull << 55);
    return static_cast<unsigned>(codes >> (5 * e)) & 31u;
}

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

__global__ void __launch_bounds__(MC_THREADS) mc_count_kernel(McGeo m, const float* __restrict__ vol, unsigned* __restrict__ segCount,
    unsigned char* __restrict__ triCount) {
    __shared__ float f[MC_NZ][MC_NY][MC_NX + 1];
    const int x0 = blockIdx.x * MCX, y0 = blockIdx.y * MCY, zc0 = m.cz0 + blockIdx.z * MCZ; // global cell coords
    // load nodes (clamped; clamped duplicates only feed cells that are masked out below)
    for (int i = threadIdx.x; i < MC_NZ * MC_NY * MC_NX; i += MC_THREADS) {
        const int ix = i % MC_NX, iy = (i / MC_NX) % MC_NY, iz = i / (MC_NX * MC_NY);
        const int x = min(x0 + ix, m.sx - 1), y = min(y0 + iy, m.sy - 1);
        const int zl = min(zc0 + iz - m.zPlane0, m.nzPlanes - 1);
        f[iz][iy][ix] = vol[x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl)];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < MCY * MCZ; r += MC_THREADS / 32) {
        const int ly = r % MCY, lz = r / MCY;
        const int cxi = x0 + lane, cyi = y0 + ly, czi = zc0 + lz;
        if (cyi >= m.cy || czi >= m.cz0 + m.cnz) continue; // warp-uniform
        unsigned n = 0;
        if (cxi < m.cx) {
            const int ci = cubeIndexSmem(&f[lz][ly][lane], MC_NX + 1, (MC_NX + 1) * MC_NY, m.iso);
            n = static_cast<unsigned>(kCaseWords[ci] & 15ull);
            if (triCount) triCount[cxi + static_cast<size_t>(m.cx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0))] = static_cast<unsigned char>(n);
        }
        const unsigned tot = __reduce_add_sync(0xffffffffu, n);
        if (lane == 0) segCount[blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0))] = tot;
    }
}

struct McEmitShared {
    float4 node[MC_NZ][MC_NY][MC_NX];            // f, gx, gy, gz
    float halo[MC_HZ][MC_HY][MC_HX + 1];
    float stagePos[MC_THREADS / 32][MC_STAGE * 9];
    float stageNrm[MC_THREADS / 32][MC_STAGE * 9];
    int anyActive;
};

template<bool COLOUR>
__global__ void __launch_bounds__(MC_THREADS) mc_emit_kernel(McGeo m, const float* __restrict__ vol, const float* __restrict__ rgb,
    const unsigned* __restrict__ segOffset, float* __restrict__ outPos, float* __restrict__ outNrm, float* __restrict__ outCol) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    McEmitShared& sh = *reinterpret_cast<McEmitShared*>(smemRaw);
    const int x0 = blockIdx.x * MCX, y0 = blockIdx.y * MCY, zc0 = m.cz0 + blockIdx.z * MCZ;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // skip tiles without triangles
    if (threadIdx.x == 0) sh.anyActive = 0;
    __syncthreads();
    if (threadIdx.x < MCY * MCZ) {
        const int ly = threadIdx.x % MCY, lz = threadIdx.x / MCY;
        const int cyi = y0 + ly, czi = zc0 + lz;
        if (cyi < m.cy && czi < m.cz0 + m.cnz) {
            const size_t seg = blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0));
            if (segOffset[seg + 1] != segOffset[seg]) sh.anyActive = 1;
        }
    }
    __syncthreads();
    if (!sh.anyActive) return;

    // density with a one-node halo; indices clamped to the GLOBAL grid (one-sided differences at the border)
    for (int i = threadIdx.x; i < MC_HZ * MC_HY * MC_HX; i += MC_THREADS) {
        const int ix = i % MC_HX, iy = (i / MC_HX) % MC_HY, iz = i / (MC_HX * MC_HY);
        const int x = min(max(x0 + ix - 1, 0), m.sx - 1), y = min(max(y0 + iy - 1, 0), m.sy - 1);
        const int zg = min(max(zc0 + iz - 1, 0), m.szGlobal - 1);
        const int zl = min(max(zg - m.zPlane0, 0), m.nzPlanes - 1);
        sh.halo[iz][iy][ix] = vol[x + static_cast<size_t>(m.sx) * (y + static_cast<size_t>(m.sy) * zl)];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MC_NZ * MC_NY * MC_NX; i += MC_THREADS) {
        const int ix = i % MC_NX, iy = (i / MC_NX) % MC_NY, iz = i / (MC_NX * MC_NY);
        const int x = x0 + ix, y = y0 + iy, z = zc0 + iz; // global node index (may exceed the grid for masked cells)
        // distance between the two samples actually used (clamped at the global border)
        const int xm = max(x - 1, 0), xp = min(x + 1, m.sx - 1), ym = max(y - 1, 0), yp = min(y + 1, m.sy - 1);
        const int zm = max(z - 1, 0), zp = min(z + 1, m.szGlobal - 1);
        const float f = sh.halo[iz + 1][iy + 1][ix + 1];
        // halo index of a clamped global coordinate c along x is (c - x0 + 1); out-of-grid nodes are never used
        auto H = [&](int gx, int gy, int gz) -> float {
            const int hx = min(max(gx - x0 + 1, 0), MC_HX - 1), hy = min(max(gy - y0 + 1, 0), MC_HY - 1), hz = min(max(gz - zc0 + 1, 0), MC_HZ - 1);
            return sh.halo[hz][hy][hx];
        };
        float4 n;
        n.x = f;
        n.y = __fdiv_rn(__fsub_rn(H(xp, y, z), H(xm, y, z)), __fmul_rn((float)(xp - xm), m.sd[0]));
        n.z = __fdiv_rn(__fsub_rn(H(x, yp, z), H(x, ym, z)), __fmul_rn((float)(yp - ym), m.sd[1]));
        n.w = __fdiv_rn(__fsub_rn(H(x, y, zp), H(x, y, zm)), __fmul_rn((float)(zp - zm), m.sd[2]));
        sh.node[iz][iy][ix] = n;
    }
    __syncthreads();

    float* sPos = sh.stagePos[warp];
    float* sNrm = sh.stageNrm[warp];
    for (int r = warp; r < MCY * MCZ; r += MC_THREADS / 32) {
        const int ly = r % MCY, lz = r / MCY;
        const int cxi = x0 + lane, cyi = y0 + ly, czi = zc0 + lz;
        if (cyi >= m.cy || czi >= m.cz0 + m.cnz) continue;
        const size_t seg = blockIdx.x + static_cast<size_t>(m.nsegx) * (cyi + static_cast<size_t>(m.cy) * (czi - m.cz0));
        const unsigned segOff = segOffset[seg], segTris = segOffset[seg + 1] - segOff;
        if (segTris == 0) continue;
        unsigned long long word = 0;
        if (cxi < m.cx) {
            int ci = 0;
            ci |= (sh.node[lz][ly][lane].x < m.iso) ? 1 : 0;
            ci |= (sh.node[lz][ly][lane + 1].x < m.iso) ? 2 : 0;
            ci |= (sh.node[lz][ly + 1][lane + 1].x < m.iso) ? 4 : 0;
            ci |= (sh.node[lz][ly + 1][lane].x < m.iso) ? 8 : 0;
            ci |= (sh.node[lz + 1][ly][lane].x < m.iso) ? 16 : 0;
            ci |= (sh.node[lz + 1][ly][lane + 1].x < m.iso) ? 32 : 0;
            ci |= (sh.node[lz + 1][ly + 1][lane + 1].x < m.iso) ? 64 : 0;
            ci |= (sh.node[lz + 1][ly + 1][lane].x < m.iso) ? 128 : 0;
            word = kCaseWords[ci];
        }
        const unsigned n = static_cast<unsigned>(word & 15ull);
        // warp-level exclusive prefix of the triangle counts
        unsigned inc = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const unsigned first = inc - n; // my first triangle within the segment
        for (unsigned win = 0; win < segTris; win += MC_STAGE) {
            // my triangles that fall into [win, win + MC_STAGE)
            for (unsigned k = 0; k < n; ++k) {
                const unsigned t = first + k;
                if (t < win || t >= win + MC_STAGE) continue;
                const unsigned slot = (t - win) * 9;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int e = static_cast<int>((word >> (4 + 4 * (3 * k + c))) & 15ull);
                    const unsigned code = edgeCode(e);
                    const int ax = lane + (code & 1), ay = ly + ((code >> 1) & 1), az = lz + ((code >> 2) & 1);
                    const int axis = code >> 3;
                    const int bx = ax + (axis == 0), by = ay + (axis == 1), bz = az + (axis == 2);
                    const float4 na = sh.node[az][ay][ax], nb = sh.node[bz][by][bx];
                    const float t01 = __fdiv_rn(__fsub_rn(m.iso, na.x), __fsub_rn(nb.x, na.x));
                    // node positions: float(idx)*sd + origin (ParticlesToDensity.cpp:605)
                    const float pax = __fadd_rn(__fmul_rn((float)(x0 + ax), m.sd[0]), m.org[0]);
                    const float pay = __fadd_rn(__fmul_rn((float)(y0 + ay), m.sd[1]), m.org[1]);
                    const float paz = __fadd_rn(__fmul_rn((float)(zc0 + az), m.sd[2]), m.org[2]);
                    const float pbx = __fadd_rn(__fmul_rn((float)(x0 + bx), m.sd[0]), m.org[0]);
                    const float pby = __fadd_rn(__fmul_rn((float)(y0 + by), m.sd[1]), m.org[1]);
                    const float pbz = __fadd_rn(__fmul_rn((float)(zc0 + bz), m.sd[2]), m.org[2]);
                    sPos[slot + 3 * c + 0] = __fadd_rn(pax, __fmul_rn(t01, __fsub_rn(pbx, pax)));
                    sPos[slot + 3 * c + 1] = __fadd_rn(pay, __fmul_rn(t01, __fsub_rn(pby, pay)));
                    sPos[slot + 3 * c + 2] = __fadd_rn(paz, __fmul_rn(t01, __fsub_rn(pbz, paz)));
                    const float gx = __fadd_rn(na.y, __fmul_rn(t01, __fsub_rn(nb.y, na.y)));
                    const float gy = __fadd_rn(na.z, __fmul_rn(t01, __fsub_rn(nb.z, na.z)));
                    const float gz = __fadd_rn(na.w, __fmul_rn(t01, __fsub_rn(nb.w, na.w)));
                    const float len2 = __fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), __fmul_rn(gz, gz));
                    const float inv = len2 > 0.0f ? __fdiv_rn(-1.0f, __fsqrt_rn(len2)) : 0.0f;
                    sNrm[slot + 3 * c + 0] = __fmul_rn(gx, inv);
                    sNrm[slot + 3 * c + 1] = __fmul_rn(gy, inv);
                    sNrm[slot + 3 * c + 2] = __fmul_rn(gz, inv);
                }
            }
            __syncwarp();
            const unsigned cnt = min((unsigned)MC_STAGE, segTris - win) * 9;
            const size_t gbase = (static_cast<size_t>(segOff) + win) * 9;
            for (unsigned j = lane; j < cnt; j += 32) {
                outPos[gbase + j] = sPos[j];
                outNrm[gbase + j] = sNrm[j];
            }
            __syncwarp();
        }
    }
}

} // namespace mms
