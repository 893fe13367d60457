/*
 * mmsurf.h -- C ABI of libmmsurf: the B200-native particle -> density volume -> isosurface path of MegaMol.
 *
 * This is the drop-in boundary.  A MegaMol module (plugin/b200surf/src/*.cpp), the pytest suite (ctypes) and
 * bench.py all go through these entry points; signatures carry plain pointers and sizes only.
 *
 * What each entry point replaces in the reference (paths relative to the MegaMol checkout):
 *   mms_set_grid / mms_set_params   the parameter + bounding-box reads at the top of
 *                                   datatools::ParticlesToDensity::createVolumeCPU
 *                                   (plugins/datatools/src/ParticlesToDensity.cpp:395-434, params :78-147)
 *   mms_push_particles              the per-list accessor set-up and particle loop input
 *                                   (ParticlesToDensity.cpp:458-486; SimpleSphericalParticles.h:27-50 enums,
 *                                   :86-178 type -> float conversion)
 *   mms_compute_density             createVolumeCPU's binning + scatter + reduction + range + normalise
 *                                   (ParticlesToDensity.cpp:561-626, :669-682); aggregator 2 also :629-667
 *   mms_push_particles_dir, mms_get_vector_field   the direction accessors and the vector-field outputs of aggregator 2
 *                                   (ParticlesToDensity.cpp:484-508, 629-667)
 *   mms_get_density / _range        what getDataCallback hands to VolumetricDataCall::SetData / metadata
 *                                   Min/MaxValues (ParticlesToDensity.cpp:249-295)
 *   mms_extract_isosurface          trisoup_gl::volumetrics::IsoSurface::buildMesh
 *                                   (plugins/trisoup_gl/src/volumetrics/IsoSurface.cpp:229-309), with the classic
 *                                   marching-cubes table of plugins/trisoup/src/volumetrics/MarchingCubeTables.cpp:58
 *                                   instead of marching tetrahedra, and smooth (gradient) normals
 *   mms_get_mesh                    what outDataCallback passes to Mesh::SetVertexData (IsoSurface.cpp:171-181):
 *                                   unindexed triangle soup, 3 floats position + 3 floats normal per vertex
 *   mms_get_home_voxels, mms_get_cell_tricounts   test hooks for the bit-exact claims (no reference counterpart)
 *
 * Conventions: every function returns MMS_OK (0) or a negative error code and never throws or aborts;
 * mms_last_error() gives the message.  A context is NOT re-entrant; use one per module instance/thread.
 * All pointers returned by mms_get_* stay owned by the library and valid until the next compute on that context.
 * There is no CPU fallback: without a CUDA device mms_create fails with MMS_ERR_CUDA.
 */
#ifndef MMSURF_H
#define MMSURF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMS_OK 0
#define MMS_ERR_INVALID (-1)     /* bad argument / call order */
#define MMS_ERR_CUDA (-2)        /* CUDA runtime error (message has the CUDA string) */
#define MMS_ERR_NOMEM (-3)       /* device or pinned allocation failed */
#define MMS_ERR_UNSUPPORTED (-4) /* valid MegaMol configuration that this path does not implement (yet) */

/* Numerically equal to geocalls::SimpleSphericalParticles::VertexDataType / ColourDataType. */
enum mms_vertex_type { MMS_VERT_NONE = 0, MMS_VERT_FLOAT_XYZ = 1, MMS_VERT_FLOAT_XYZR = 2, MMS_VERT_SHORT_XYZ = 3, MMS_VERT_DOUBLE_XYZ = 4 };
enum mms_colour_type {
    MMS_COL_NONE = 0, MMS_COL_UINT8_RGB = 1, MMS_COL_UINT8_RGBA = 2, MMS_COL_FLOAT_RGB = 3, MMS_COL_FLOAT_RGBA = 4,
    MMS_COL_FLOAT_I = 5, MMS_COL_USHORT_RGBA = 6, MMS_COL_DOUBLE_I = 7
};

/* Density semantics. */
enum mms_mode {
    MMS_MODE_P2D_BUMP = 0, /* datatools::ParticlesToDensity: compact bump RBF, support box from the home voxel */
    MMS_MODE_QS_GAUSS = 1, /* QuickSurf: Gaussian exp2(d^2 w), radial cut-off gausslim*radscale*r, optional colour */
    MMS_MODE_QS_GAUSS_REFCELLS = 2 /* the same Gaussian with the candidate set of protein_cuda's CUDAQuickSurf (CUDAQuickSurf.cu:232-253):
                                      NO radial cut-off, every atom of the acceleration cells (size max(gausslim*radscale*rmax,
                                      spacing)) around the voxel's 8x8x8 block contributes.  Reproduces the reference's density up to
                                      the fp32 summation order -- including its dependence on the block tiling -- at ~15x the pair
                                      evaluations of mode 1; needs a uniform grid spacing and particles inside the grid. */
};

typedef struct mms_ctx mms_ctx;

typedef struct mms_config {
    int32_t device;   /* CUDA device ordinal */
    int32_t reserved; /* must be 0 */
} mms_config;

/* One particle list exactly as a MultiParticleDataCall carries it (borrowed pointers, arbitrary stride;
 * stride 0 = tightly packed).  The pointers may be pageable host, pinned host or device memory. */
typedef struct mms_list {
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
} mms_list;

/* Node-centred grid: voxel (i,j,k) sits at min + (i,j,k) * extent/(res-1)   (docs/volumes.md:14-16). */
typedef struct mms_grid {
    float min[3];      /* object-space bbox Left/Bottom/Back */
    float extent[3];   /* bbox Width/Height/Depth */
    int32_t res[3];    /* sizex/sizey/sizez */
    int32_t cyclic[3]; /* cyclX/cyclY/cyclZ */
} mms_grid;

typedef struct mms_params {
    int32_t mode;            /* mms_mode */
    int32_t aggregator;      /* ParticlesToDensity "aggregator": 0 position, 1 intensity-weighted, 2 direction-weighted vector field
                                (P2D mode, sigma <= 1; directions come in through mms_push_particles_dir, results through mms_get_vector_field) */
    int32_t normalize;       /* ParticlesToDensity "normalize" */
    int32_t defer_normalize; /* 1: compute_density leaves the raw sums; caller normalises with mms_normalize (slabs) */
    float sigma;             /* ParticlesToDensity "sigma" */
    float radscale;          /* QuickSurf "radiusScale" */
    float gausslim;          /* QuickSurf quality -> 2.0/2.5/3.0/4.0 */
    int32_t colour;          /* QS mode: also build the density-weighted RGB volume and coloured mesh */
    int32_t want_home_voxels;   /* test hook: keep per-particle home voxels */
    int32_t want_cell_tricounts;/* test hook: keep per-cell triangle counts */
} mms_params;

typedef struct mms_timings { /* milliseconds, CUDA events on the context's stream, last call of each stage */
    float h2d, bin, density, normalize, mc, d2h_volume, d2h_mesh;
    float mc_emit; /* the emit kernel alone (part of mc) */
} mms_timings;

int mms_create(mms_ctx** out, const mms_config* cfg);
int mms_destroy(mms_ctx* ctx);
const char* mms_last_error(const mms_ctx* ctx); /* ctx may be NULL: error of the last failed mms_create */

int mms_set_grid(mms_ctx* ctx, const mms_grid* grid);
/* z-slab sharding: this context computes density planes [z0, z0+nz) and marching-cubes cell layers
 * [cell_z0, cell_z0+cell_nz) (cell layer k spans planes k and k+1, so cell_z0 >= z0 and cell_z0+cell_nz < z0+nz).
 * Give a slab one extra plane on each interior side of its cell range and its gradient normals are identical to
 * the unsharded ones.  Default (and after every mms_set_grid) = all planes, all cell layers. */
int mms_set_slab(mms_ctx* ctx, int32_t z0, int32_t nz, int32_t cell_z0, int32_t cell_nz);
int mms_set_params(mms_ctx* ctx, const mms_params* params);

int mms_clear_particles(mms_ctx* ctx);
/* Appends lists; data is copied (H2D on a dedicated copy stream, asynchronously for pinned memory, into one of two
 * persistent upload arenas) unless it already lives on the device.  Pinned source buffers must stay valid until the next
 * mms_compute_density returned.  Results of the previous compute stay readable, so frame k+1 can be pushed while frame k
 * is still being read back (streaming). */
int mms_push_particles(mms_ctx* ctx, int32_t nlists, const mms_list* lists);

/* The same with per-list direction data (SimpleSphericalParticles::DIRDATA_FLOAT_XYZ: 3 floats, SimpleSphericalParticles.h:58,179-193,504-510)
 * for aggregator 2: dirs[i] = direction pointer of lists[i] or NULL (DIRDATA_NONE: the reference's accessors then deliver 0),
 * dir_strides[i] = its stride in bytes (0 = 12).  dirs == NULL is mms_push_particles. */
int mms_push_particles_dir(mms_ctx* ctx, int32_t nlists, const mms_list* lists, const void* const* dirs, const uint32_t* dir_strides);

/* Largest radius over the lists pushed so far (global radii and per-particle radii; finite, > 0), found on the device: QuickSurf sizes
 * its grid padding from it before any density is computed (QuickSurf::calculateSurface, plugins/protein_cuda/src/QuickSurf.cpp:419-470).
 * Synchronises. */
int mms_get_max_radius(mms_ctx* ctx, float* rmax);

int mms_compute_density(mms_ctx* ctx);
/* Range of the (un-normalised) sums of the last compute_density: the reference's minDens/maxDens. */
int mms_get_density_range(mms_ctx* ctx, float minmax[2]);
/* v = (v - mn) * (1/(mx - mn)) on the device volume (ParticlesToDensity.cpp:676-682). */
int mms_normalize(mms_ctx* ctx, float mn, float mx);
/* The same without a host round trip (z-slab sharding): the range as two device floats {-min, max} (so that ONE max-all-reduce
 * over the ranks yields the global range in place), and a normalise that reads it from device memory. */
int mms_density_range_device(mms_ctx* ctx, float** dev_negmin_max);
int mms_normalize_device(mms_ctx* ctx, const float* dev_negmin_max);
/* Run all kernels of this context on the caller's CUDA stream (cudaStream_t; NULL = back to the context's own stream), e.g. the
 * stream the caller's collectives are ordered on.  The default stream is named by cudaStreamLegacy ((cudaStream_t)0x1) or
 * cudaStreamPerThread ((cudaStream_t)0x2), never by NULL. */
int mms_set_stream(mms_ctx* ctx, void* cuda_stream);
/* Starts the device-to-host copy of the (final) volume on the library's copy stream and returns: a following mms_extract_isosurface
 * overlaps with it (marching cubes only reads the volume), and mms_get_density then merely waits for the copy.  Optional. */
int mms_prefetch_density(mms_ctx* ctx);
/* Host copy (library-owned pinned memory) of the slab: nz*res[1]*res[0] floats, x fastest. rgb may be NULL. */
int mms_get_density(mms_ctx* ctx, const float** host_volume, const float** host_rgb);
int mms_get_density_device(mms_ctx* ctx, const float** dev_volume, const float** dev_rgb);
/* Aggregator 2 (IVecToSingleCell_Volume, ParticlesToDensity.cpp:493-508,629-667): what createVolumeCPU leaves behind per voxel,
 * x fastest, slab-shaped.  vec = the 3-component volume handed to VolumetricDataCall (sum(w d)/sum(w), after "normalize" each
 * component (c - minDens)/(maxDens - minDens)); magnitude = the un-normalised |v| ("densities"; mms_get_density_range is its
 * range = minDens/maxDens, mms_get_density / the isosurface see this scalar volume); direction = v/|v| (0 where |v| == 0).
 * Any of the three pointers may be NULL.  Host pointers are library-owned pinned memory. */
int mms_get_vector_field(mms_ctx* ctx, const float** host_vec, const float** host_magnitude, const float** host_direction);
int mms_get_vector_field_device(mms_ctx* ctx, const float** dev_vec, const float** dev_magnitude, const float** dev_direction);
/* Use a caller-supplied volume instead of computing one (IsoSurface fed by another VolumetricDataCall source).
 * `volume` may be host or device memory, res-shaped for the current slab. */
int mms_set_density(mms_ctx* ctx, const float* volume);
/* Device-resident hand-off between two contexts on the same device (IsoSurfaceB200 behind a ParticlesToDensityB200, replacing the
 * host array of VolumetricDataCall::GetData, VolumetricDataCall.h:136-144): `ctx` takes over the producer's grid, slab and (colour)
 * volume BY REFERENCE -- nothing is copied -- and can then count / emit / read back isosurfaces with its OWN buffers, so that
 * several consumers of one producer do not overwrite each other's meshes.  The reference lasts until either context computes, sets
 * or adopts another density; the producer's volume must not be recomputed while `ctx` is using it (module callbacks run one at a
 * time).  Stream-ordered after everything the producer has enqueued so far. */
int mms_adopt_density(mms_ctx* ctx, mms_ctx* producer);

int mms_extract_isosurface(mms_ctx* ctx, float isovalue); /* = count + emit into library-owned buffers */

/* Which triangulation mms_count/emit/extract_isosurface produce.
 *   MMS_ISO_MARCHING_CUBES (default)  256-case marching cubes with the reference's table, smooth gradient normals, node-centred frame.
 *   MMS_ISO_MARCHING_TETS             exactly what trisoup_gl::volumetrics::IsoSurface::buildMesh / makeTet / interpolate produce
 *                                     (IsoSurface.cpp:30-31, 229-309, 430-465, 606-735): six tetrahedra per cell, regula falsi on the
 *                                     trilinear interpolant, flat normals, the module's half-voxel shifted frame ((idx+0.5)/s*extent+min
 *                                     with the grid's min/extent = the object-space bounding box) -- triangle for triangle and bit for
 *                                     bit the CPU module's output.  No colour output in this mode. */
enum { MMS_ISO_MARCHING_CUBES = 0, MMS_ISO_MARCHING_TETS = 1 };
int mms_set_isosurface_mode(mms_ctx* ctx, int32_t mode);
/* The two halves of extract, for callers that own the destination: count (classify + scan, one host round trip for the size),
 * then emit into caller-supplied DEVICE memory starting at triangle `first_triangle` (9 floats per triangle and array).
 * The destination may be this GPU's memory, a CUDA-IPC / peer mapping of another GPU's buffer (the z-slab driver lets every
 * rank write its slab's triangles straight into rank 0's mesh over NVLink: compute and gather in ONE kernel), or a
 * graphics-interop mapping of a vertex buffer.  pos == NULL emits into the library's own buffers. */
int mms_count_isosurface(mms_ctx* ctx, float isovalue, uint64_t* ntriangles);
int mms_emit_isosurface(mms_ctx* ctx, float* pos, float* nrm, float* col, uint64_t first_triangle);
/* Triangle soup: nverts = 3 * triangles; pos/nrm/col = 3 floats per vertex (col NULL unless colour mode). */
int mms_get_mesh(mms_ctx* ctx, uint64_t* nverts, const float** pos, const float** nrm, const float** col);
int mms_get_mesh_device(mms_ctx* ctx, uint64_t* nverts, const float** pos, const float** nrm, const float** col);
/* Opt-in INDEXED mesh (the default stays the reference's unindexed soup): one vertex per crossed grid edge + three 32-bit indices
 * per triangle, what CallTriMeshData::Mesh::SetVertexData + SetTriangleData(cnt, unsigned int*) carry
 * (plugins/geometry_calls_gl/include/geometry_calls_gl/CallTriMeshDataGL.h:897-1000) -- about 28 bytes
 * per triangle instead of 72.  Same triangles in the same order as the soup, and pos[idx[k]] / nrm[idx[k]] are bit for bit the soup's
 * k-th vertex.  Marching cubes without colours on a whole-volume context (no z-slabs); mms_set_mesh_indexed invalidates a count.
 * pos, nrm: 3 floats per vertex; idx: 3 indices per triangle.  The soup getters (mms_get_mesh*, mms_share_mesh) refuse an indexed mesh. */
int mms_set_mesh_indexed(mms_ctx* ctx, int32_t on);
int mms_get_mesh_indexed(mms_ctx* ctx, uint64_t* nverts, uint64_t* ntris, const float** pos, const float** nrm, const uint32_t** idx);
int mms_get_mesh_indexed_device(mms_ctx* ctx, uint64_t* nverts, uint64_t* ntris, const float** pos, const float** nrm, const uint32_t** idx);

/* Test hooks (host pointers). home: 3 int32 per particle, list-major input order. tricounts: one byte per
 * cell, x fastest, (res0-1)*(res1-1)*cell_nz entries. */
int mms_get_home_voxels(mms_ctx* ctx, const int32_t** home, uint64_t* nparticles);
int mms_get_cell_tricounts(mms_ctx* ctx, const uint8_t** counts, uint64_t* ncells);

int mms_get_timings(mms_ctx* ctx, mms_timings* out);
/* Stopwatch on the context's own stream (CUDA events): start records, stop records + synchronises. */
int mms_timer_start(mms_ctx* ctx);
int mms_timer_stop(mms_ctx* ctx, float* elapsed_ms);
int mms_synchronize(mms_ctx* ctx);
/* Number of kernel launches issued by this context so far (for bench.py's gpu_launches). */
uint64_t mms_launch_count(const mms_ctx* ctx);

/* Multi-GPU exchange, sender side: stable partition of one DEVICE-resident FLOAT_XYZ/XYZR list by destination z-slab.  Slab i
 * computes density planes [plane_lo[i], plane_hi[i]] (inclusive, halo planes included); a particle is copied to every slab its
 * support box touches (the reference's home voxel +- filter size, periodic images in cyclic z).  send_buf (device) receives the
 * raw records (list stride bytes each) grouped by slab, original order inside a group; counts[i] = records for slab i (host).
 * A slab with plane_lo > plane_hi is switched off (count 0): a rank that keeps its own chunk in place switches its own slab off and
 * extracts only what the OTHER slabs need from it.  Particles that can never contribute (non-finite, radius <= 0) are dropped.
 * Uses the grid and parameters of the context.  Synchronises once (the counts decide the all-to-all split sizes). */
int mms_route_particles(mms_ctx* ctx, const mms_list* list, int32_t nslabs, const int32_t* plane_lo, const int32_t* plane_hi, void* send_buf,
    uint64_t capacity_records, uint64_t* counts);

/* ---- halo exchange without host round trips (replaces mms_route_particles + an all-to-all-v; and, as a whole, the volume-sized
 * MPI_Allreduce of plugins/datatools/src/MPIVolumeAggregator.cpp:94-152) --------------------------------------------------------------
 * Every slab owns a receive buffer (x y z r records) and a record counter in device memory; the other slabs map them (CUDA IPC between
 * processes: mms_ipc_export / mms_ipc_open; peer access inside one process) and APPEND what that slab needs from their share of the
 * frame with one kernel (halo_push_kernel: ballot-aggregated system-scope atomics over NVLink).  Per frame and slab:
 *     mms_clear_particles, mms_push_particles(own share)           (as always)
 *     mms_halo_push(...)                                           one kernel per pushed list + one signal kernel; no host wait
 *     mms_halo_wait(ctx, npeers)                                   stream-ordered: returns at once; before mms_compute_density reads the
 *                                                                  received list, the STREAM waits until `npeers` slabs have signalled that
 *                                                                  their records of this frame have landed (a spinning one-thread kernel
 *                                                                  on a flag in this GPU's memory, placed behind the binning of the
 *                                                                  context's own lists: no collective, no host; inside one process CUDA
 *                                                                  events do the same and the call is not needed)
 *     mms_halo_receive(ctx, radius_bound)                          the received records become one more list; its LENGTH stays on the device
 *     mms_compute_density ...
 * Counters and buffer halves alternate between two sets per frame: a slab clears the counters of the NEXT frame when it signals this one,
 * and frame k+1's records never land in memory frame k is still being binned from. */
/* This context's receive buffer (2 x capacity_records x 16 bytes -- one half per frame parity --, grow-only; growing invalidates earlier
 * mappings) and its counter block (4 x uint32: record counter of even / odd frames, arrival counter of even / odd frames).  Every slab
 * passes the SAME capacity_records. */
int mms_halo_buffers(mms_ctx* ctx, uint64_t capacity_records, void** recv_buf, void** counters);
/* Routes the lists pushed so far: slab i computes planes [plane_lo[i], plane_hi[i]]; peer_bufs[i] / peer_counters[i] are slab i's receive
 * buffer / counter block as mapped into THIS process (ignored for i == my_slab); capacity_records as given to the peers' mms_halo_buffers. */
int mms_halo_push(mms_ctx* ctx, int32_t nslabs, int32_t my_slab, const int32_t* plane_lo, const int32_t* plane_hi, void* const* peer_bufs,
    void* const* peer_counters, uint64_t capacity_records);
/* Adds what the peers have pushed into this context's buffer as a FLOAT_XYZR list (radii <= radius_bound: the largest radius any slab
 * may send -- it sizes the sort cells without a device scan).  Call once every peer's mms_halo_push of this frame has completed. */
int mms_halo_receive(mms_ctx* ctx, float radius_bound);
/* Between processes: makes this context's stream wait for the signals of `npeers` pushing slabs (normally nslabs - 1). */
int mms_halo_wait(mms_ctx* ctx, int32_t npeers);

/* ---- several GPUs behind ONE handle, inside one process (the drop-in modules' `devices` parameter) ---------------------------------
 * The z-slab decomposition of SURVEY 8e driven from C: one context per device, slab g owns cell layers [g(sz-1)/G, (g+1)(sz-1)/G) and
 * computes one extra density plane on each interior side, every list is split into G contiguous shares that travel over each GPU's own
 * PCIe link, the halo particles go through mms_halo_push / mms_halo_receive with peer access (ordered by CUDA events, no host
 * synchronisation), the global density range is combined by a kernel reading the peers' ranges (instead of the reference's
 * volume-sized MPI_Allreduce, plugins/datatools/src/MPIVolumeAggregator.cpp:94-152).  Results are bit-identical to one GPU.
 * A device may be named several times: z-chunks on ONE GPU, which is how a volume of 2^32 voxels or more (more than one context's 32-bit
 * voxel indices reach) is computed on a single device (the reference chunks in z as well, CUDAQuickSurf.cu:1050-1126, 1406-1447).
 * For G > 1 the halo records travel as x y z r (+ RGBA where the QuickSurf colour volume is on): ParticlesToDensity aggregator 0, or the
 * QuickSurf Gaussian with the radial cut-off (MMS_MODE_QS_GAUSS) -- the reference chunks exactly these volumes in z
 * (CUDAQuickSurf.cu:1050-1126). */
typedef struct mms_slabs mms_slabs;
int mms_slabs_create(mms_slabs** out, const int32_t* devices, int32_t ndevices);
int mms_slabs_destroy(mms_slabs* s);
const char* mms_slabs_last_error(const mms_slabs* s); /* s may be NULL: error of the last failed mms_slabs_create */
int32_t mms_slabs_count(const mms_slabs* s);
mms_ctx* mms_slabs_context(mms_slabs* s, int32_t i); /* the i-th device's context (timings, device pointers, mms_adopt_density) */
int mms_slabs_set_grid(mms_slabs* s, const mms_grid* grid);
int mms_slabs_set_params(mms_slabs* s, const mms_params* params);
int mms_slabs_clear_particles(mms_slabs* s);
int mms_slabs_push_particles(mms_slabs* s, int32_t nlists, const mms_list* lists); /* host lists; split into contiguous shares */
int mms_slabs_compute_density(mms_slabs* s);
int mms_slabs_get_density_range(mms_slabs* s, float minmax[2]);
int mms_slabs_get_density(mms_slabs* s, const float** host_volume); /* the WHOLE volume, library-owned pinned memory */
/* mms_adopt_density for groups on the same devices: device-resident hand-off of every slab (IsoSurfaceB200 behind a multi-GPU producer) */
int mms_slabs_adopt_density(mms_slabs* s, mms_slabs* producer);
int mms_slabs_extract_isosurface(mms_slabs* s, float isovalue);
/* The whole mesh in cell-linear order (= the single-GPU order), library-owned pinned memory; every slab is copied over its own link. */
int mms_slabs_get_mesh(mms_slabs* s, uint64_t* nverts, const float** positions, const float** normals);
/* QuickSurf colour outputs of the group (MMS_MODE_QS_GAUSS with params.colour; the halo records then carry their RGBA as well): the
 * density-weighted RGB volume (3 floats per voxel, whole volume) and the per-vertex colours in the order of mms_slabs_get_mesh
 * (QuickSurf.cpp:596-616 is what the reference feeds its colour volume with).  NULL where the mode has no colours. */
int mms_slabs_get_colour_volume(mms_slabs* s, const float** host_rgb);
int mms_slabs_get_mesh_colours(mms_slabs* s, const float** colours);

/* Device buffers that can be shared between the per-GPU processes of one node (CUDA IPC over NVLink / PCIe P2P). */
int mms_device_alloc(int32_t device, size_t bytes, void** ptr);
int mms_device_free(int32_t device, void* ptr);
int mms_ipc_export(int32_t device, const void* devptr, unsigned char handle[64]);
int mms_ipc_open(int32_t device, const unsigned char handle[64], void** ptr);
int mms_ipc_close(int32_t device, void* ptr);

/* ---- Device-resident hand-off to consumers ---------------------------------------------------------------------
 * The reference hands results over in host memory; its one VRAM variant is VolumetricDataCall::SetData(uint32_t texture) with
 * MemLoc = VRAM (geometry_calls/VolumetricDataCall.h:290-292, VolumetricDataCallTypes.h:22,133), and the tri-soup renderers re-buffer
 * the host arrays into GL buffers every frame (trisoup_gl/src/ModernTrisoupRenderer.cpp:308-440).  With mms_share_enable the volume and
 * the mesh live in exportable device memory (CUDA virtual-memory allocations with a POSIX file-descriptor handle): a renderer imports
 * the descriptor once per (re)allocation -- glImportMemoryFdEXT + glNamedBufferStorageMemEXT, VkImportMemoryFdInfoKHR, or
 * cuMemImportFromShareableHandle / mms_share_open for a CUDA consumer -- and draws from it; nothing crosses PCIe.
 * Each mms_share_* call returns NEW descriptors (the caller closes them, close(2); an import consumes nothing of the library's).
 * The calls synchronise the context's stream: the contents are final on return.  A buffer keeps its allocation (and the importer's
 * mapping stays valid) until a later frame outgrows it: alloc_bytes changes then, which is the importer's cue to re-import. */
typedef struct mms_share {
    int32_t fd;           /* POSIX file descriptor of the allocation, -1 = nothing to share (empty mesh, no colours) */
    uint32_t reserved;
    uint64_t alloc_bytes; /* size of the whole allocation = the size to import / map */
    uint64_t offset;      /* first payload byte inside the allocation */
    uint64_t bytes;       /* payload bytes */
} mms_share;
int mms_share_enable(mms_ctx* ctx, int32_t on);
/* Volume (res.x * res.y * slab planes floats) and, in colour / vector mode, the 3-float-per-voxel volume. */
int mms_share_density(mms_ctx* ctx, mms_share* volume, mms_share* rgb);
/* Triangle soup of the last mms_extract_isosurface: 3 floats per vertex each, same layout as mms_get_mesh. */
int mms_share_mesh(mms_ctx* ctx, uint64_t* nverts, mms_share* positions, mms_share* normals, mms_share* colours);
/* The CUDA consumer's side: maps a shared allocation on `device` (may be another process, another device with peer access). */
int mms_share_open(int32_t device, const mms_share* share, void** devptr);
int mms_share_close(int32_t device, void* devptr, const mms_share* share);

/* Pinned host memory for callers that want zero-staging H2D (e.g. an MMPLD reader). */
void* mms_alloc_pinned(size_t bytes);
void mms_free_pinned(void* p);

/* MMPLD frame ingest (replaces the loader of moldyn::MMPLDDataSource, plugins/moldyn/src/io/MMPLDDataSource.cpp:61-217,375-401,
 * for this path): frames are read into pinned, double-buffered host memory and described as mms_list arrays (file colour types
 * are translated to the in-memory enum).  Lists returned by read_frame stay valid until the second-next read_frame. */
typedef struct mms_mmpld mms_mmpld;
int mms_mmpld_open(mms_mmpld** out, const char* path);
int mms_mmpld_close(mms_mmpld* reader);
const char* mms_mmpld_last_error(const mms_mmpld* reader);
int mms_mmpld_info(const mms_mmpld* reader, uint32_t* frames, uint32_t* version, float bbox[6], float clipbox[6]);
int mms_mmpld_prefetch(mms_mmpld* reader, uint32_t frame); /* background thread, into the other buffer */
int mms_mmpld_read_frame(mms_mmpld* reader, uint32_t frame, int32_t* nlists, const mms_list** lists, float* timestamp);

int mms_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MMSURF_H */
