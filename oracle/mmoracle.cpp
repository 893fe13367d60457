// oracle/mmoracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// An independent restatement, in plain C++, of the algorithm on MegaMol's particle -> density -> isosurface
// path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load it; the product
// (libmmsurf.so) never links, loads or calls anything in this directory.
//
// What each function follows (paths relative to the reference checkout):
//   home voxel / support box / bump kernel / aggregators 0,1 / reduction order / normalise
//        plugins/datatools/src/ParticlesToDensity.cpp:416-434 (geometry), :458-476 (radius, kernel),
//        :490-529 (aggregators), :561-620 (binning + scatter loop), :669-682 (range, normalise);
//        aggregator 2 (vector field): :493-508, :629-667
//   particle accessors (type -> float conversion, global radius, colour-R as intensity)
//        plugins/geometry_calls/include/geometry_calls/SimpleSphericalParticles.h:86-178,
//        plugins/geometry_calls/include/geometry_calls/Accessor.h:54-146
//   marching-cubes classification + table (bit i <=> value < iso, corner order a2fVertexOffset)
//        plugins/trisoup/src/volumetrics/MarchingCubeTables.cpp:11-16,58,278 and
//        plugins/trisoup_gl/src/volumetrics/IsoSurface.cpp:254-276 (comparison sense)
//   vertex interpolation t=(iso-f0)/(f1-f0), p0+t(p1-p0)
//        plugins/protein_cuda/src/quicksurf/CUDAMarchingCubes.cu:224-227 (semantics reference only)
//   Gaussian (QuickSurf) mode   rho = sum exp2(d^2 * -log2(e)/(2 (r*radscale)^2)), colour = sum w*rgb
//        plugins/protein_cuda/src/quicksurf/CUDAQuickSurf.cu:303-315,1325-1334; QuickSurf.cpp:511-590
//
// PINNING: the reference ships no tests or golden vectors for this path, so this oracle is
// pinned against the compiled reference translation units themselves (oracle/_ref/libmmref.so): see
// tests/test_oracle_vs_reference.py (runs where /root/reference exists) and the fixtures it wrote to
// tests/golden/ with oracle/tools/gen_golden.py (checked everywhere).
//
// Built with -ffp-contract=off: the reference is built for baseline x86-64 without FMA, every fp32
// operation below is individually rounded.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include <omp.h>

namespace {

const uint64_t kCaseWords[256] = {
#include "mc_case_words.inc"
};

// Corner offsets and edge end points of the classic table (MarchingCubeTables.cpp:11-16).
const int kCorner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
const int kEdge[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

enum { VERT_NONE = 0, VERT_FLOAT_XYZ = 1, VERT_FLOAT_XYZR = 2, VERT_SHORT_XYZ = 3, VERT_DOUBLE_XYZ = 4 };
enum {
    COL_NONE = 0, COL_UINT8_RGB = 1, COL_UINT8_RGBA = 2, COL_FLOAT_RGB = 3, COL_FLOAT_RGBA = 4, COL_FLOAT_I = 5,
    COL_USHORT_RGBA = 6, COL_DOUBLE_I = 7
};
const unsigned kVertSize[5] = {0, 12, 16, 6, 24};
const unsigned kColSize[8] = {0, 3, 4, 12, 16, 4, 8, 8};

} // namespace

extern "C" {

struct mmo_list {
    const void* vtx;
    const void* col;
    uint64_t count;
    int32_t vtx_type;
    uint32_t vtx_stride;
    int32_t col_type;
    uint32_t col_stride;
    float global_radius;
    uint8_t global_rgba[4];
    float irange[2];
};

struct mmo_grid {
    float min[3];    // bbox Left/Bottom/Back
    float extent[3]; // bbox Width/Height/Depth
    int32_t res[3];
    int32_t cyclic[3];
};

} // extern "C"

namespace {

struct Particle {
    float x, y, z, r;
};

inline Particle fetch(const mmo_list& l, uint64_t j) {
    const unsigned stride = l.vtx_stride ? l.vtx_stride : kVertSize[l.vtx_type];
    const char* p = static_cast<const char*>(l.vtx) + j * stride;
    Particle q{0, 0, 0, l.global_radius};
    switch (l.vtx_type) {
    case VERT_FLOAT_XYZ: {
        float f[3];
        std::memcpy(f, p, 12);
        q.x = f[0], q.y = f[1], q.z = f[2];
    } break;
    case VERT_FLOAT_XYZR: {
        float f[4];
        std::memcpy(f, p, 16);
        q.x = f[0], q.y = f[1], q.z = f[2], q.r = f[3];
    } break;
    case VERT_SHORT_XYZ: { // raw unsigned short values cast to float, no de-quantisation (SURVEY 8a traps)
        unsigned short s[3];
        std::memcpy(s, p, 6);
        q.x = static_cast<float>(s[0]), q.y = static_cast<float>(s[1]), q.z = static_cast<float>(s[2]);
    } break;
    case VERT_DOUBLE_XYZ: {
        double d[3];
        std::memcpy(d, p, 24);
        q.x = static_cast<float>(d[0]), q.y = static_cast<float>(d[1]), q.z = static_cast<float>(d[2]);
    } break;
    default: break;
    }
    return q;
}

/** Colour accessors as floats (Get_f of cr/cg/cb/ca accessors); out[0] is what aggregator 1 multiplies by. */
inline void fetchColour(const mmo_list& l, uint64_t j, float out[4]) {
    const unsigned stride = l.col_stride ? l.col_stride : kColSize[l.col_type];
    const char* p = l.col ? static_cast<const char*>(l.col) + j * stride : nullptr;
    switch (l.col_type) {
    case COL_UINT8_RGB:
    case COL_UINT8_RGBA: {
        unsigned char c[4] = {0, 0, 0, 255};
        std::memcpy(c, p, l.col_type == COL_UINT8_RGB ? 3 : 4);
        for (int k = 0; k < 4; ++k) out[k] = static_cast<float>(c[k]);
    } break;
    case COL_FLOAT_RGB: {
        std::memcpy(out, p, 12);
        out[3] = 1.0f;
    } break;
    case COL_FLOAT_RGBA: std::memcpy(out, p, 16); break;
    case COL_FLOAT_I: {
        std::memcpy(out, p, 4);
        out[1] = out[2] = out[3] = 0.0f;
    } break;
    case COL_USHORT_RGBA: {
        unsigned short c[4];
        std::memcpy(c, p, 8);
        for (int k = 0; k < 4; ++k) out[k] = static_cast<float>(c[k]);
    } break;
    case COL_DOUBLE_I: {
        double d;
        std::memcpy(&d, p, 8);
        out[0] = static_cast<float>(d);
        out[1] = out[2] = out[3] = 0.0f;
    } break;
    default:
        for (int k = 0; k < 4; ++k) out[k] = static_cast<float>(l.global_rgba[k]) / 255.0f;
    }
}

struct Geometry {
    float min[3], sd[3];
    int s[3];
    bool cyc[3];
};

inline Geometry geometry(const mmo_grid& g) {
    Geometry q;
    for (int a = 0; a < 3; ++a) {
        q.min[a] = g.min[a];
        q.s[a] = g.res[a];
        q.sd[a] = g.extent[a] / static_cast<float>(g.res[a] - 1); // ParticlesToDensity.cpp:430-432
        q.cyc[a] = g.cyclic[a] != 0;
    }
    return q;
}

inline int homeVoxel(float p, float mn, float sd) { // ParticlesToDensity.cpp:564
    return static_cast<int>((p - mn) / sd);
}

inline float bump(float dist, float eps) { // ParticlesToDensity.cpp:472-476
    if (dist >= eps) return 0.0f;
    return std::exp(-1.0f / (1.0f - std::pow((1.0f / eps) * dist, 2.0f)));
}

inline long floorMod(long a, long m) {
    long r = a % m;
    return r < 0 ? r + m : r;
}

} // namespace

extern "C" {

int mmo_version() { return 1; }

/** Home voxel of every particle, list-major: out[3*i + axis].  Bit-exact restatement of :563-568. */
int mmo_home_voxels(int nlists, const mmo_list* lists, const mmo_grid* grid, int32_t* out) {
    const Geometry g = geometry(*grid);
    uint64_t base = 0;
    for (int li = 0; li < nlists; ++li) {
        const mmo_list& l = lists[li];
        if (l.vtx_type == VERT_NONE) continue;
#pragma omp parallel for
        for (int64_t j = 0; j < static_cast<int64_t>(l.count); ++j) {
            const Particle p = fetch(l, j);
            out[3 * (base + j) + 0] = homeVoxel(p.x, g.min[0], g.sd[0]);
            out[3 * (base + j) + 1] = homeVoxel(p.y, g.min[1], g.sd[1]);
            out[3 * (base + j) + 2] = homeVoxel(p.z, g.min[2], g.sd[2]);
        }
        base += l.count;
    }
    return 0;
}

/**
 * ParticlesToDensity density volume, aggregator 0 (sum of bumps) or 1 (bump * colour-R).
 * Accumulation order per voxel = particle order (list-major), i.e. exactly the reference at ONE OpenMP
 * thread; threads here partition the z axis, which leaves every voxel's order untouched, so the result
 * is independent of the thread count.  [z0, z0+nz) restricts the output to a slab (vol has nz planes).
 * minmax (optional) receives min/max BEFORE normalisation; normalize applies (v-min)*(1/(max-min)) (:676-682).
 */
int mmo_density_p2d(int nlists, const mmo_list* lists, const mmo_grid* grid, float sigma, int aggregator,
    int normalize, int z0, int nz, float* vol, float minmax[2]) {
    if (aggregator != 0 && aggregator != 1) return -1;
    const Geometry g = geometry(*grid);
    const int sx = g.s[0], sy = g.s[1], sz = g.s[2];
    if (z0 < 0 || nz < 0 || z0 + nz > sz) return -2;
    const size_t nvox = static_cast<size_t>(sx) * sy * nz;
    std::fill(vol, vol + nvox, 0.0f);
    for (int li = 0; li < nlists; ++li) {
        const mmo_list& l = lists[li];
        if (l.vtx_type == VERT_NONE) continue;
#pragma omp parallel
        {
            const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
            const int zlo = z0 + static_cast<int>(static_cast<long>(nz) * tid / nt);
            const int zhi = z0 + static_cast<int>(static_cast<long>(nz) * (tid + 1) / nt); // exclusive
            for (uint64_t j = 0; j < l.count && zlo < zhi; ++j) {
                const Particle p = fetch(l, j);
                const float rad = p.r;
                if (rad == 0.0f) continue; // volOp early-out (:523)
                const int x = homeVoxel(p.x, g.min[0], g.sd[0]);
                const int y = homeVoxel(p.y, g.min[1], g.sd[1]);
                const int z = homeVoxel(p.z, g.min[2], g.sd[2]);
                const int fx = static_cast<int>(std::ceil(rad / g.sd[0]));
                const int fy = static_cast<int>(std::ceil(rad / g.sd[1]));
                const int fz = static_cast<int>(std::ceil(rad / g.sd[2]));
                float weight = 1.0f;
                if (aggregator == 1) {
                    float c[4];
                    fetchColour(l, j, c);
                    weight = c[0];
                }
                const float eps = sigma * rad;
                for (int hz = z - fz; hz <= z + fz; ++hz) {
                    long tz = hz;
                    if (g.cyc[2]) tz = floorMod(hz, sz); // == (hz + 2*sz) % sz wherever the reference is defined
                    else if (hz < 0 || hz > sz - 1) continue;
                    if (tz < zlo || tz >= zhi) continue;
                    float zd = static_cast<float>(hz) * g.sd[2] + g.min[2];
                    zd = std::fabs(zd - p.z);
                    for (int hy = y - fy; hy <= y + fy; ++hy) {
                        long ty = hy;
                        if (g.cyc[1]) ty = floorMod(hy, sy);
                        else if (hy < 0 || hy > sy - 1) continue;
                        float yd = static_cast<float>(hy) * g.sd[1] + g.min[1];
                        yd = std::fabs(yd - p.y);
                        for (int hx = x - fx; hx <= x + fx; ++hx) {
                            long tx = hx;
                            if (g.cyc[0]) tx = floorMod(hx, sx);
                            else if (hx < 0 || hx > sx - 1) continue;
                            float xd = static_cast<float>(hx) * g.sd[0] + g.min[0];
                            xd = std::fabs(xd - p.x);
                            const float dis = std::sqrt(xd * xd + yd * yd + zd * zd);
                            const float w = bump(dis, eps);
                            float& v = vol[tx + (ty + (tz - z0) * sy) * static_cast<size_t>(sx)];
                            v += (aggregator == 1) ? w * weight : w;
                        }
                    }
                }
            }
        }
    }
    if (nvox == 0) return 0;
    float mx = *std::max_element(vol, vol + nvox), mn = *std::min_element(vol, vol + nvox);
    if (minmax) minmax[0] = mn, minmax[1] = mx;
    if (normalize) {
        const float rcp = 1.0f / (mx - mn);
#pragma omp parallel for
        for (int64_t i = 0; i < static_cast<int64_t>(nvox); ++i) vol[i] = (vol[i] - mn) * rcp;
    }
    return 0;
}

/**
 * ParticlesToDensity aggregator 2 (IVecToSingleCell_Volume): :493-508 (volOp), :629-667 (per-voxel pass), :669-682 (normalise).
 * dirs[li] = DIRDATA_FLOAT_XYZ pointer of list li or NULL (DIRDATA_NONE: the accessors deliver 0), dir_strides[li] bytes (0 = 12).
 * Out: vec = the 3-component volume the module hands to VolumetricDataCall (normalised component-wise with the range of the
 * magnitudes if `normalize`), mag = |v| before normalisation ("densities"), dir = v/|v| ("directions"), minmax = minDens/maxDens
 * (minDens starts at FLT_MAX, maxDens at 0).  Accumulation order = particle order = the reference at ONE OpenMP thread.
 */
int mmo_density_p2d_vector(int nlists, const mmo_list* lists, const void* const* dirs, const uint32_t* dir_strides, const mmo_grid* grid,
    float sigma, int normalize, float* vec, float* mag, float* dir, float minmax[2]) {
    const Geometry g = geometry(*grid);
    const int sx = g.s[0], sy = g.s[1], sz = g.s[2];
    const size_t nvox = static_cast<size_t>(sx) * sy * sz;
    std::vector<float> weights(nvox, 0.0f);
    std::fill(vec, vec + 3 * nvox, 0.0f);
    for (int li = 0; li < nlists; ++li) {
        const mmo_list& l = lists[li];
        if (l.vtx_type == VERT_NONE) continue;
        const char* dptr = dirs ? static_cast<const char*>(dirs[li]) : nullptr;
        const unsigned dstride = (dir_strides && dir_strides[li]) ? dir_strides[li] : 12u;
#pragma omp parallel
        {
            const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
            const int zlo = static_cast<int>(static_cast<long>(sz) * tid / nt);
            const int zhi = static_cast<int>(static_cast<long>(sz) * (tid + 1) / nt); // exclusive
            for (uint64_t j = 0; j < l.count && zlo < zhi; ++j) {
                const Particle p = fetch(l, j);
                const float rad = p.r;
                if (rad == 0.0f) continue;
                float d[3] = {0.0f, 0.0f, 0.0f};
                if (dptr) std::memcpy(d, dptr + j * dstride, 12);
                const int x = homeVoxel(p.x, g.min[0], g.sd[0]);
                const int y = homeVoxel(p.y, g.min[1], g.sd[1]);
                const int z = homeVoxel(p.z, g.min[2], g.sd[2]);
                const int fx = static_cast<int>(std::ceil(rad / g.sd[0]));
                const int fy = static_cast<int>(std::ceil(rad / g.sd[1]));
                const int fz = static_cast<int>(std::ceil(rad / g.sd[2]));
                const float eps = sigma * rad;
                for (int hz = z - fz; hz <= z + fz; ++hz) {
                    long tz = hz;
                    if (g.cyc[2]) tz = floorMod(hz, sz);
                    else if (hz < 0 || hz > sz - 1) continue;
                    if (tz < zlo || tz >= zhi) continue;
                    float zd = static_cast<float>(hz) * g.sd[2] + g.min[2];
                    zd = std::fabs(zd - p.z);
                    for (int hy = y - fy; hy <= y + fy; ++hy) {
                        long ty = hy;
                        if (g.cyc[1]) ty = floorMod(hy, sy);
                        else if (hy < 0 || hy > sy - 1) continue;
                        float yd = static_cast<float>(hy) * g.sd[1] + g.min[1];
                        yd = std::fabs(yd - p.y);
                        for (int hx = x - fx; hx <= x + fx; ++hx) {
                            long tx = hx;
                            if (g.cyc[0]) tx = floorMod(hx, sx);
                            else if (hx < 0 || hx > sx - 1) continue;
                            float xd = static_cast<float>(hx) * g.sd[0] + g.min[0];
                            xd = std::fabs(xd - p.x);
                            const float dis = std::sqrt(xd * xd + yd * yd + zd * zd);
                            const float w = bump(dis, eps);
                            const size_t o = tx + (ty + tz * sy) * static_cast<size_t>(sx);
                            vec[3 * o + 0] += w * d[0];
                            vec[3 * o + 1] += w * d[1];
                            vec[3 * o + 2] += w * d[2];
                            weights[o] += w;
                        }
                    }
                }
            }
        }
    }
    float mx = 0.0f, mn = std::numeric_limits<float>::max();
    for (size_t i = 0; i < nvox; ++i) {
        const float div = weights[i] == 0.0f ? 1.0f : weights[i];
        vec[3 * i + 0] /= div, vec[3 * i + 1] /= div, vec[3 * i + 2] /= div;
        const float den = std::sqrt(vec[3 * i + 0] * vec[3 * i + 0] + vec[3 * i + 1] * vec[3 * i + 1] + vec[3 * i + 2] * vec[3 * i + 2]);
        if (mag) mag[i] = den;
        if (dir)
            for (int k = 0; k < 3; ++k) dir[3 * i + k] = den == 0.0f ? 0.0f : vec[3 * i + k] / den;
        mx = std::max(mx, den);
        mn = std::min(mn, den);
    }
    if (minmax) minmax[0] = mn, minmax[1] = mx;
    if (normalize) {
        const float rcp = 1.0f / (mx - mn);
        for (size_t i = 0; i < 3 * nvox; ++i) vec[i] = (vec[i] - mn) * rcp;
    }
    return 0;
}

/** (v - mn) * (1/(mx-mn)) with caller-supplied range (multi-slab normalisation uses the GLOBAL range). */
int mmo_normalize(float* vol, uint64_t n, float mn, float mx) {
    const float rcp = 1.0f / (mx - mn);
#pragma omp parallel for
    for (int64_t i = 0; i < static_cast<int64_t>(n); ++i) vol[i] = (vol[i] - mn) * rcp;
    return 0;
}

/**
 * QuickSurf-style Gaussian density (+ optional density-weighted RGB volume, 3 floats per voxel, x fastest).
 * Node (i,j,k) sits at origin + (i,j,k)*spacing.  Candidate set = clean radial cutoff
 * d < gausslim*radscale*r_p per particle (spec choice); weight exp2f(d^2 * w_p),
 * w_p = -log2(e) / (2 (r_p*radscale)^2).  Per-voxel order = particle order.  Non-periodic.
 */
int mmo_density_gauss(int nlists, const mmo_list* lists, const float origin[3], const float spacing[3],
    const int32_t res[3], float radscale, float gausslim, int z0, int nz, float* vol, float* rgb) {
    const int sx = res[0], sy = res[1];
    const size_t nvox = static_cast<size_t>(sx) * sy * nz;
    std::fill(vol, vol + nvox, 0.0f);
    if (rgb) std::fill(rgb, rgb + 3 * nvox, 0.0f);
    const float log2e = 1.4426950408889634f;
    for (int li = 0; li < nlists; ++li) {
        const mmo_list& l = lists[li];
        if (l.vtx_type == VERT_NONE) continue;
#pragma omp parallel
        {
            const int nt = omp_get_num_threads(), tid = omp_get_thread_num();
            const int zlo = z0 + static_cast<int>(static_cast<long>(nz) * tid / nt);
            const int zhi = z0 + static_cast<int>(static_cast<long>(nz) * (tid + 1) / nt);
            for (uint64_t j = 0; j < l.count && zlo < zhi; ++j) {
                const Particle p = fetch(l, j);
                const float sr = p.r * radscale;
                if (!(sr > 0.0f)) continue;
                const float w = -log2e / (2.0f * sr * sr);
                const float cut = gausslim * sr;
                const float cut2 = cut * cut;
                float c[4] = {1, 1, 1, 1};
                if (rgb) {
                    fetchColour(l, j, c);
                    if (l.col_type == COL_UINT8_RGB || l.col_type == COL_UINT8_RGBA || l.col_type == COL_USHORT_RGBA)
                        for (int k = 0; k < 3; ++k) c[k] = c[k] / 255.0f; // QuickSurf.cpp:545-577 (USHORT /255 sic)
                    if (l.col_type == COL_FLOAT_I || l.col_type == COL_DOUBLE_I) { // grey, min-max normalised per list
                        const float v = (c[0] - l.irange[0]) / (l.irange[1] - l.irange[0]);
                        c[0] = c[1] = c[2] = v;
                    }
                }
                int lo[3], hi[3];
                const float pp[3] = {p.x, p.y, p.z};
                for (int a = 0; a < 3; ++a) {
                    lo[a] = static_cast<int>(std::floor((pp[a] - cut - origin[a]) / spacing[a])) - 1;
                    hi[a] = static_cast<int>(std::ceil((pp[a] + cut - origin[a]) / spacing[a])) + 1;
                }
                lo[0] = std::max(lo[0], 0), lo[1] = std::max(lo[1], 0), lo[2] = std::max(lo[2], zlo);
                hi[0] = std::min(hi[0], sx - 1), hi[1] = std::min(hi[1], sy - 1), hi[2] = std::min(hi[2], zhi - 1);
                for (int k = lo[2]; k <= hi[2]; ++k) {
                    const float dz = (static_cast<float>(k) * spacing[2] + origin[2]) - p.z;
                    for (int jy = lo[1]; jy <= hi[1]; ++jy) {
                        const float dy = (static_cast<float>(jy) * spacing[1] + origin[1]) - p.y;
                        for (int i = lo[0]; i <= hi[0]; ++i) {
                            const float dx = (static_cast<float>(i) * spacing[0] + origin[0]) - p.x;
                            const float d2 = dx * dx + dy * dy + dz * dz;
                            if (!(d2 < cut2)) continue;
                            const float g = std::exp2(d2 * w);
                            const size_t o = i + (jy + static_cast<size_t>(k - z0) * sy) * sx;
                            vol[o] += g;
                            if (rgb) {
                                rgb[3 * o + 0] += g * c[0];
                                rgb[3 * o + 1] += g * c[1];
                                rgb[3 * o + 2] += g * c[2];
                            }
                        }
                    }
                }
            }
        }
    }
    return 0;
}

/**
 * The REFERENCE's candidate set for the same Gaussian density (protein_cuda QuickSurf): no radial cut-off at all -- a voxel sums
 * exp2f(d^2 w_p) over EVERY atom of the acceleration-grid cells that overlap its 8x8x8 thread-block tile grown by one cell size
 * (CUDAQuickSurf.cu:232-253,293-315; cell size max(gausslim*radscale*rmax, gridspacing), :1259-1264,1279-1282; atoms hashed with
 * min(int(p / cellsize), ncells-1), CUDASpatialSearch.cu:92-94; cells visited z, y, x, atoms of a cell in input order -- the radix
 * sort by cell is stable).  xyzr positions are relative to the grid origin (QuickSurf.cpp:553-556); rgb, if given, becomes the
 * RGB3F volume texture sum(w rgb) * 1/isovalue (:510-513).  Used to pin the restatement above to the compiled reference kernels
 * (tests/test_gpu_quicksurf_ref.py) and to measure what the clean cut-off leaves out.
 */
int mmo_density_gauss_refset(uint64_t natoms, const float* xyzr, const float* rgba, const int32_t res[3], float maxrad, float radscale,
    float gridspacing, float isovalue, float gausslim, float* vol, float* rgb) {
    float acgridspacing = gausslim * radscale * maxrad;
    if (acgridspacing < gridspacing) acgridspacing = gridspacing;
    const int ncx = std::max(int((res[0] * gridspacing) / acgridspacing), 1), ncy = std::max(int((res[1] * gridspacing) / acgridspacing), 1),
              ncz = std::max(int((res[2] * gridspacing) / acgridspacing), 1);
    const float invac = 1.0f / acgridspacing, invisovalue = 1.0f / isovalue;
    const float log2e = std::log2(2.718281828);
    std::vector<std::vector<uint32_t>> cells(static_cast<size_t>(ncx) * ncy * ncz);
    std::vector<float> w(natoms);
    for (uint64_t i = 0; i < natoms; ++i) {
        const int cx = std::min(int(xyzr[4 * i] * invac), ncx - 1), cy = std::min(int(xyzr[4 * i + 1] * invac), ncy - 1),
                  cz = std::min(int(xyzr[4 * i + 2] * invac), ncz - 1);
        if (cx < 0 || cy < 0 || cz < 0) return -1; // the reference would index out of bounds: QuickSurf pads its grid so this cannot happen
        cells[(static_cast<size_t>(cz) * ncy + cy) * ncx + cx].push_back(static_cast<uint32_t>(i));
        const float scaledrad = xyzr[4 * i + 3] * radscale;
        w[i] = -1.0f * log2e / (2.0f * scaledrad * scaledrad);
    }
    const int B = 8; // GBLOCKSZX x GBLOCKSZY x (GBLOCKSZZ * GUNROLL)
    const int nbx = (res[0] + B - 1) / B, nby = (res[1] + B - 1) / B, nbz = (res[2] + B - 1) / B;
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int bz = 0; bz < nbz; ++bz)
        for (int by = 0; by < nby; ++by)
            for (int bx = 0; bx < nbx; ++bx) {
                auto range = [&](int b, int nc, int& lo, int& hi) {
                    lo = int(((b * B) * gridspacing - acgridspacing) * invac);
                    hi = int((((b + 1) * B) * gridspacing + acgridspacing) * invac);
                    lo = lo < 0 ? 0 : lo;
                    hi = hi >= nc - 1 ? nc - 1 : hi;
                };
                int x0, x1, y0, y1, z0, z1;
                range(bx, ncx, x0, x1), range(by, ncy, y0, y1), range(bz, ncz, z0, z1);
                for (int k = bz * B; k < std::min((bz + 1) * B, res[2]); ++k)
                    for (int j = by * B; j < std::min((by + 1) * B, res[1]); ++j)
                        for (int i = bx * B; i < std::min((bx + 1) * B, res[0]); ++i) {
                            const float coorx = gridspacing * i, coory = gridspacing * j, coorz = gridspacing * k;
                            float d = 0.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f;
                            for (int zc = z0; zc <= z1; ++zc)
                                for (int yc = y0; yc <= y1; ++yc)
                                    for (int xc = x0; xc <= x1; ++xc)
                                        for (uint32_t a : cells[(static_cast<size_t>(zc) * ncy + yc) * ncx + xc]) {
                                            const float dx = coorx - xyzr[4 * a], dy = coory - xyzr[4 * a + 1], dz = coorz - xyzr[4 * a + 2];
                                            const float dxy2 = dx * dx + dy * dy;
                                            const float t = std::exp2((dxy2 + dz * dz) * w[a]);
                                            d += t;
                                            if (rgb) cr += t * rgba[4 * a], cg += t * rgba[4 * a + 1], cb += t * rgba[4 * a + 2];
                                        }
                            const size_t o = i + (j + static_cast<size_t>(k) * res[1]) * res[0];
                            vol[o] = d;
                            if (rgb) rgb[3 * o] = cr * invisovalue, rgb[3 * o + 1] = cg * invisovalue, rgb[3 * o + 2] = cb * invisovalue;
                        }
            }
    return 0;
}

/** Cube index of cell (x,y,z): bit i set iff corner i < iso. */
static inline int cubeIndex(const float* vol, int sx, int sy, int x, int y, int z, float iso) {
    int ci = 0;
    for (int c = 0; c < 8; ++c) {
        const float v = vol[(x + kCorner[c][0]) + static_cast<size_t>(sx) * ((y + kCorner[c][1]) + static_cast<size_t>(sy) * (z + kCorner[c][2]))];
        if (v < iso) ci |= 1 << c;
    }
    return ci;
}

/**
 * Per-cell triangle counts, cells x-fastest: out[x + (sx-1)*(y + (sy-1)*z)], (sx-1)(sy-1)(sz-1) entries.
 * cubeidx (optional) receives the 8-bit case.  Returns the total number of triangles.
 */
int64_t mmo_mc_count(const float* vol, const int32_t res[3], float iso, uint8_t* tricount, uint8_t* cubeidx) {
    const int sx = res[0], sy = res[1], sz = res[2];
    const int cx = sx - 1, cy = sy - 1, cz = sz - 1;
    if (cx <= 0 || cy <= 0 || cz <= 0) return 0;
    int64_t total = 0;
#pragma omp parallel for reduction(+ : total)
    for (int z = 0; z < cz; ++z)
        for (int y = 0; y < cy; ++y)
            for (int x = 0; x < cx; ++x) {
                const int ci = cubeIndex(vol, sx, sy, x, y, z, iso);
                const int n = static_cast<int>(kCaseWords[ci] & 15);
                const size_t o = x + static_cast<size_t>(cx) * (y + static_cast<size_t>(cy) * z);
                if (tricount) tricount[o] = static_cast<uint8_t>(n);
                if (cubeidx) cubeidx[o] = static_cast<uint8_t>(ci);
                total += n;
            }
    return total;
}

namespace {
struct Node {
    float f, gx, gy, gz;
};
} // namespace

/**
 * Triangle soup in cell-linear order (x fastest, then y, then z); within a cell the table's order.
 * pos/nrm (and col, if rgb != NULL): 9 floats per triangle.  Node (i,j,k) sits at origin + idx*sd
 * (float(idx)*sd + origin, two roundings, the reference's voxel-position formula :605).
 * Vertex on an edge: always interpolated from the LOWER node to the HIGHER node of the edge's axis
 * (t = (iso - f_lo)/(f_hi - f_lo), p = p_lo + t*(p_hi - p_lo)) so that neighbouring cells produce
 * bit-identical shared vertices.  Normal = -grad(rho) interpolated with the same t and normalised;
 * grad at a node by central differences (one-sided at the border) times 1/(actual node distance).
 * Colour = lerp of the per-node colour rgb/rho (rho == 0 -> 0), same t.
 * z_offset: index of plane 0 of `vol` in the global grid (slabs); positions use global indices.
 * Returns the number of triangles written (<= max_tris), or -1 if max_tris was too small.
 */
int64_t mmo_mc_emit(const float* vol, const float* rgb, const int32_t res[3], const float origin[3], const float sd[3],
    float iso, int z_offset, int64_t max_tris, float* pos, float* nrm, float* col) {
    const int sx = res[0], sy = res[1], sz = res[2];
    const int cx = sx - 1, cy = sy - 1, cz = sz - 1;
    if (cx <= 0 || cy <= 0 || cz <= 0) return 0;
    auto at = [&](int x, int y, int z) -> float { return vol[x + static_cast<size_t>(sx) * (y + static_cast<size_t>(sy) * z)]; };
    // gradient = (f(+) - f(-)) * (1 / (n * sd)) with n = number of grid steps between the two samples (2 inside,
    // 1 at the border); the reciprocal is formed once per axis and n, the product is one rounded multiply
    float rinv[3][3];
    for (int a = 0; a < 3; ++a)
        for (int n = 1; n <= 2; ++n) rinv[a][n] = 1.0f / (static_cast<float>(n) * sd[a]);
    auto node = [&](int x, int y, int z) -> Node {
        Node n;
        n.f = at(x, y, z);
        const int xm = std::max(x - 1, 0), xp = std::min(x + 1, sx - 1);
        const int ym = std::max(y - 1, 0), yp = std::min(y + 1, sy - 1);
        const int zm = std::max(z - 1, 0), zp = std::min(z + 1, sz - 1);
        n.gx = xp > xm ? (at(xp, y, z) - at(xm, y, z)) * rinv[0][xp - xm] : 0.0f;
        n.gy = yp > ym ? (at(x, yp, z) - at(x, ym, z)) * rinv[1][yp - ym] : 0.0f;
        n.gz = zp > zm ? (at(x, y, zp) - at(x, y, zm)) * rinv[2][zp - zm] : 0.0f;
        return n;
    };
    int64_t ntri = 0;
    for (int z = 0; z < cz; ++z)
        for (int y = 0; y < cy; ++y)
            for (int x = 0; x < cx; ++x) {
                const int ci = cubeIndex(vol, sx, sy, x, y, z, iso);
                const uint64_t word = kCaseWords[ci];
                const int n = static_cast<int>(word & 15);
                if (n == 0) continue;
                if (ntri + n > max_tris) return -1;
                for (int k = 0; k < 3 * n; ++k) {
                    const int e = static_cast<int>((word >> (4 + 4 * k)) & 15);
                    int a = kEdge[e][0], b = kEdge[e][1];
                    // canonical direction: lower node first
                    if (kCorner[a][0] + kCorner[a][1] + kCorner[a][2] > kCorner[b][0] + kCorner[b][1] + kCorner[b][2]) std::swap(a, b);
                    const int ax = x + kCorner[a][0], ay = y + kCorner[a][1], az = z + kCorner[a][2];
                    const int bx = x + kCorner[b][0], by = y + kCorner[b][1], bz = z + kCorner[b][2];
                    const Node na = node(ax, ay, az), nb = node(bx, by, bz);
                    const float t = (iso - na.f) / (nb.f - na.f);
                    const float pa[3] = {static_cast<float>(ax) * sd[0] + origin[0], static_cast<float>(ay) * sd[1] + origin[1],
                        static_cast<float>(az + z_offset) * sd[2] + origin[2]};
                    const float pb[3] = {static_cast<float>(bx) * sd[0] + origin[0], static_cast<float>(by) * sd[1] + origin[1],
                        static_cast<float>(bz + z_offset) * sd[2] + origin[2]};
                    float* P = pos + 9 * ntri + 3 * k;
                    for (int c = 0; c < 3; ++c) P[c] = pa[c] + t * (pb[c] - pa[c]);
                    if (nrm) {
                        const float gx = na.gx + t * (nb.gx - na.gx), gy = na.gy + t * (nb.gy - na.gy), gz = na.gz + t * (nb.gz - na.gz);
                        const float len2 = gx * gx + gy * gy + gz * gz;
                        float inv = 0.0f;
                        if (len2 > 0.0f) inv = -1.0f / std::sqrt(len2);
                        float* N = nrm + 9 * ntri + 3 * k;
                        N[0] = gx * inv, N[1] = gy * inv, N[2] = gz * inv;
                    }
                    if (col && rgb) {
                        const size_t oa = ax + static_cast<size_t>(sx) * (ay + static_cast<size_t>(sy) * az);
                        const size_t ob = bx + static_cast<size_t>(sx) * (by + static_cast<size_t>(sy) * bz);
                        float* Cc = col + 9 * ntri + 3 * k;
                        for (int c = 0; c < 3; ++c) {
                            const float ca = na.f > 0.0f ? rgb[3 * oa + c] / na.f : 0.0f;
                            const float cb = nb.f > 0.0f ? rgb[3 * ob + c] / nb.f : 0.0f;
                            Cc[c] = ca + t * (cb - ca);
                        }
                    }
                }
                ntri += n;
            }
    return ntri;
}

// ---------------------------------------------------------------------------------------------------------------
// Marching tetrahedra exactly as trisoup_gl::volumetrics::IsoSurface does it (IsoSurface.cpp:30-31, 229-309, 391-465,
// 606-735): six tetrahedra per cell, vertices by regula falsi on the TRILINEAR interpolant along the tetrahedron edge,
// one flat normal per triangle.  Every operation individually rounded (the reference's baseline x86-64 build has no FMA).
// ---------------------------------------------------------------------------------------------------------------
namespace {
const unsigned kTets[6][4] = {{0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 6, 7}, {0, 6, 1, 2}, {0, 6, 1, 4}, {5, 6, 1, 4}}; // IsoSurface.cpp:30-31
const float kFloatEps = 1e-5f;                                                                                 // vislib FLOAT_EPSILON
inline bool isEq(float m, float n) { return std::fabs(m - n) < kFloatEps; }                                    // mathfunctions.h:119-136

struct P3 {
    float v[3];
};
inline P3 lerpPoint(const P3& a, const P3& b, float t) { // AbstractPointImpl::Interpolate: a*(1-t) + b*t
    const float at = 1.0f - t;
    P3 r;
    for (int d = 0; d < 3; ++d) r.v[d] = a.v[d] * at + b.v[d] * t;
    return r;
}
inline float mtOffset(float v1, float v2, float want) { return (want - v1) / (v2 - v1); } // IsoSurface::getOffset
inline float mtValue(const float* cv, unsigned i0, unsigned i1, float a) {                 // getValue (IsoSurface.cpp:405-424)
    const float b = 1.0f - a;
    const float x = b * static_cast<float>(kCorner[i0][0]) + a * static_cast<float>(kCorner[i1][0]);
    const float y = b * static_cast<float>(kCorner[i0][1]) + a * static_cast<float>(kCorner[i1][1]);
    const float z = b * static_cast<float>(kCorner[i0][2]) + a * static_cast<float>(kCorner[i1][2]);
    float vv[4];
    vv[0] = (1.0f - x) * cv[0] + x * cv[1];
    vv[1] = (1.0f - x) * cv[3] + x * cv[2];
    vv[2] = (1.0f - x) * cv[4] + x * cv[5];
    vv[3] = (1.0f - x) * cv[7] + x * cv[6];
    vv[0] = (1.0f - y) * vv[0] + y * vv[1];
    vv[2] = (1.0f - y) * vv[2] + y * vv[3];
    return (1.0f - z) * vv[0] + z * vv[2];
}
inline P3 mtInterpolate(const P3* pts, const float* cv, float val, unsigned i0, unsigned i1) { // IsoSurface::interpolate (:430-465)
    float a0 = 0.0f;
    float v0 = mtValue(cv, i0, i1, a0);
    if (isEq(v0, val)) return pts[i0];
    float a1 = 1.0f;
    float v1 = mtValue(cv, i0, i1, a1);
    if (isEq(v1, val)) return pts[i1];
    float a = mtOffset(cv[i0], cv[i1], val);
    float v = mtValue(cv, i0, i1, a);
    unsigned maxStep = 100;
    const bool flip = cv[i0] > cv[i1];
    while (maxStep > 0 && !isEq(v, val)) {
        if ((!flip && v > val) || (flip && v < val)) a1 = a, v1 = v;
        else a0 = a, v0 = v;
        a = a0 + mtOffset(v0, v1, val) * (a1 - a0);
        v = mtValue(cv, i0, i1, a);
        --maxStep;
    }
    return lerpPoint(pts[i0], pts[i1], a);
}
inline void cross3(const float* a, const float* b, float* r) { // AbstractVector<T,3>::Cross
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
}
inline float normalise3(float* v) { // AbstractVectorImpl::Normalise / Length
    float l = 0.0f;
    for (int d = 0; d < 3; ++d) l += v[d] * v[d];
    l = std::sqrt(l);
    if (l != 0.0f) for (int d = 0; d < 3; ++d) v[d] /= l;
    else v[0] = v[1] = v[2] = 0.0f;
    return l;
}
/** triangles of one tetrahedron; returns their number (0..2); tri[k] = 3 points */
inline int mtMakeTet(unsigned triIdx, unsigned tet, const P3* pts, const float* cv, float val, P3 tri[2][3]) {
    const unsigned* T = kTets[tet];
    const P3 &p0 = pts[T[0]], &p1 = pts[T[1]], &p2 = pts[T[2]], &p3 = pts[T[3]];
    float e1[3], e2[3], nrm[3];
    for (int d = 0; d < 3; ++d) e1[d] = p2.v[d] - p1.v[d], e2[d] = p3.v[d] - p1.v[d];
    cross3(e1, e2, nrm);
    normalise3(nrm);
    // Plane(p1, norm): d = -1 * (a*x + b*y + c*z);  Halfspace(p0): normalised parameters, IsEqual(dist, 0) -> in plane
    const float pd = -1.0f * (nrm[0] * p1.v[0] + nrm[1] * p1.v[1] + nrm[2] * p1.v[2]);
    const float len = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
    float A = 0, B = 0, C = 0, D = 0;
    if (!isEq(len, 0.0f)) A = nrm[0] / len, B = nrm[1] / len, C = nrm[2] / len, D = pd / len;
    const float dist = A * p0.v[0] + B * p0.v[1] + C * p0.v[2] + D;
    bool flip = !isEq(dist, 0.0f) && dist > 0.0f;
    auto I = [&](int a, int b) { return mtInterpolate(pts, cv, val, T[a], T[b]); };
    int n = 0;
    switch (triIdx) {
    case 0x00: case 0x0F: break;
    case 0x01: flip = !flip; // fall through
    case 0x0E: tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(0, 2); tri[0][flip ? 1 : 2] = I(0, 3); n = 1; break;
    case 0x02: flip = !flip; // fall through
    case 0x0D: tri[0][0] = I(1, 0); tri[0][flip ? 2 : 1] = I(1, 3); tri[0][flip ? 1 : 2] = I(1, 2); n = 1; break;
    case 0x0C: flip = !flip; // fall through
    case 0x03:
        tri[0][0] = I(0, 3); tri[0][flip ? 2 : 1] = I(0, 2); tri[0][flip ? 1 : 2] = I(1, 3);
        tri[1][0] = tri[0][flip ? 1 : 2]; tri[1][flip ? 1 : 2] = I(1, 2); tri[1][flip ? 2 : 1] = tri[0][flip ? 2 : 1];
        n = 2; break;
    case 0x04: flip = !flip; // fall through
    case 0x0B: tri[0][0] = I(2, 0); tri[0][flip ? 2 : 1] = I(2, 1); tri[0][flip ? 1 : 2] = I(2, 3); n = 1; break;
    case 0x05: flip = !flip; // fall through
    case 0x0A:
        tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(2, 3); tri[0][flip ? 1 : 2] = I(0, 3);
        tri[1][0] = tri[0][0]; tri[1][flip ? 2 : 1] = I(1, 2); tri[1][flip ? 1 : 2] = tri[0][flip ? 2 : 1];
        n = 2; break;
    case 0x06: flip = !flip; // fall through
    case 0x09:
        tri[0][0] = I(0, 1); tri[0][flip ? 2 : 1] = I(1, 3); tri[0][flip ? 1 : 2] = I(2, 3);
        tri[1][0] = tri[0][0]; tri[1][flip ? 1 : 2] = I(0, 2); tri[1][flip ? 2 : 1] = tri[0][flip ? 1 : 2];
        n = 2; break;
    case 0x08: flip = !flip; // fall through
    case 0x07: tri[0][0] = I(3, 0); tri[0][flip ? 2 : 1] = I(3, 2); tri[0][flip ? 1 : 2] = I(3, 1); n = 1; break;
    }
    return n;
}
} // namespace

/**
 * The reference IsoSurface's triangle soup (IsoSurface::buildMesh): cells in x-fastest order, tetrahedra 0..5, tri then tri2.
 * bbox = {Left, Bottom, Back, Width, Height, Depth} of the object-space bounding box; cell size = extent / s and cell centres at
 * (idx + 0.5)/s * extent + min -- the reference's own (half-voxel shifted) frame.  pos/nrm may be NULL (count only).
 * Returns the number of triangles, or -1 if max_tris was too small.
 */
int64_t mmo_mt_emit(const float* vol, const int32_t res[3], const float bbox[6], float val, int64_t max_tris, float* pos, float* nrm) {
    const unsigned sx = res[0], sy = res[1], sz = res[2];
    if (sx < 2 || sy < 2 || sz < 2) return 0;
    const float cellX = bbox[3] / static_cast<float>(sx), cellY = bbox[4] / static_cast<float>(sy), cellZ = bbox[5] / static_cast<float>(sz);
    int64_t ntri = 0;
    for (unsigned z = 0; z < sz - 1; ++z) {
        float pz = (static_cast<float>(z) + 0.5f) / static_cast<float>(sz);
        pz = pz * bbox[5] + bbox[2];
        for (unsigned y = 0; y < sy - 1; ++y) {
            float py = (static_cast<float>(y) + 0.5f) / static_cast<float>(sy);
            py = py * bbox[4] + bbox[1];
            for (unsigned x = 0; x < sx - 1; ++x) {
                float px = (static_cast<float>(x) + 0.5f) / static_cast<float>(sx);
                px = px * bbox[3] + bbox[0];
                float cv[8];
                bool bigger = false, smaller = false;
                for (int j = 0; j < 8; ++j) {
                    cv[j] = vol[(x + kCorner[j][0]) + static_cast<size_t>(sx) * ((y + kCorner[j][1]) + static_cast<size_t>(sy) * (z + kCorner[j][2]))];
                    bigger = bigger || (cv[j] >= val);
                    smaller = smaller || (cv[j] < val);
                }
                if (!bigger || !smaller) continue;
                P3 pts[8];
                for (int j = 0; j < 8; ++j) {
                    pts[j].v[0] = px + static_cast<float>(kCorner[j][0]) * cellX;
                    pts[j].v[1] = py + static_cast<float>(kCorner[j][1]) * cellY;
                    pts[j].v[2] = pz + static_cast<float>(kCorner[j][2]) * cellZ;
                }
                for (unsigned tet = 0; tet < 6; ++tet) {
                    unsigned triIdx = 0;
                    for (int k = 0; k < 4; ++k)
                        if (cv[kTets[tet][k]] < val) triIdx |= 1u << k;
                    P3 tri[2][3];
                    const int n = mtMakeTet(triIdx, tet, pts, cv, val, tri);
                    for (int t = 0; t < n; ++t) {
                        if (pos) {
                            if (ntri >= max_tris) return -1;
                            for (int i = 0; i < 3; ++i)
                                for (int d = 0; d < 3; ++d) pos[9 * ntri + 3 * i + d] = tri[t][i].v[d];
                            if (nrm) {
                                float e1[3], e2[3], nn[3];
                                for (int d = 0; d < 3; ++d) e1[d] = tri[t][1].v[d] - tri[t][0].v[d], e2[d] = tri[t][2].v[d] - tri[t][0].v[d];
                                cross3(e1, e2, nn);
                                normalise3(nn);
                                for (int i = 0; i < 3; ++i)
                                    for (int d = 0; d < 3; ++d) nrm[9 * ntri + 3 * i + d] = nn[d];
                            }
                        }
                        ++ntri;
                    }
                }
            }
        }
    }
    return ntri;
}

/** The packed case words (for table tests). */
void mmo_case_words(uint64_t out[256]) { std::memcpy(out, kCaseWords, sizeof(kCaseWords)); }

} // extern "C"
