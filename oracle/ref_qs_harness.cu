// ref_qs_harness.cu -- TEST INFRASTRUCTURE: runs the UNMODIFIED QuickSurf density kernels of the reference on the GPU box.
//
// The reference's Gaussian density (plugins/protein_cuda/src/quicksurf/CUDAQuickSurf.cu:219-670) and its spatial hashing
// (CUDASpatialSearch.cu:80-251, CUDASort.cu) are compiled from the sources where they lie; this file #includes CUDAQuickSurf.cu so
// that its file-static kernels are visible and repeats, call for call, the density part of CUDAQuickSurf::calc_surf (:1258-1470): the
// acceleration grid, the per-atom exponent factor, vmd_cuda_build_density_atom_grid, one launch of gaussdensity_fast /
// gaussdensity_fast_tex3f over the whole volume.  Nothing of calc_surf's marching cubes / GL hand-off is executed; the
// CUDAMarchingCubes member functions it references are link-time stubs below (CUDAMarchingCubes.cu itself does not compile with
// CUDA 12: texture references).  Built by oracle/Makefile.ref into oracle/_ref/libmmrefqs.so; used by tests/test_gpu_quicksurf_ref.py.
#include "quicksurf/CUDAQuickSurf.cu"

// ---- link-time stubs (never called) ---------------------------------------------------------------------------------------------
CUDAMarchingCubes::CUDAMarchingCubes() {}
CUDAMarchingCubes::~CUDAMarchingCubes() {}
bool CUDAMarchingCubes::Initialize(uint3) { return false; }
void CUDAMarchingCubes::SetSubVolume(uint3, uint3) {}
bool CUDAMarchingCubes::SetVolumeData(float*, float3*, uint3, float3, float3, bool) { return false; }
bool CUDAMarchingCubes::SetVolumeData(float*, uchar4*, uint3, float3, float3, bool) { return false; }
void CUDAMarchingCubes::computeIsosurface(float3*, float3*, float3*, unsigned int) {}
void CUDAMarchingCubes::computeIsosurface(float3*, float3*, uchar4*, unsigned int) {}
void CUDAMarchingCubes::computeIsosurface(float3*, char3*, uchar4*, unsigned int) {}

#define MMQ_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "ref_qs_harness: %s: %s\n", #call, cudaGetErrorString(e_));            \
            return -1;                                                                             \
        }                                                                                          \
    } while (0)

extern "C" {

/**
 * Density (and, with colours, the RGB3F volume texture) of the reference's QuickSurf for `natoms` atoms.
 * xyzr: x y z radius per atom, positions RELATIVE TO THE GRID ORIGIN (QuickSurf.cpp:553-556); rgba: 4 floats per atom or NULL.
 * out_density: numvoxels[0]*[1]*[2] floats, x fastest; out_rgb: 3 floats per voxel (already scaled by 1/isovalue, :510-513) or NULL.
 * accel[3] receives the acceleration-grid size (informational).
 */
int mmq_density(long natoms, const float* xyzr_f, const float* rgba, const int numvoxels[3], float maxrad, float radscale, float gridspacing,
    float isovalue, float gausslim, float* out_density, float* out_rgb, int accel[3]) {
    const bool colorperatom = rgba != nullptr && out_rgb != nullptr;
    // ---- CUDAQuickSurf::calc_surf, :1258-1282 ------------------------------------------------------------------------------------
    float acgridspacing = gausslim * radscale * maxrad;
    if (acgridspacing < gridspacing)
        acgridspacing = gridspacing;
    int3 volsz = make_int3(numvoxels[0], numvoxels[1], numvoxels[2]);
    int3 accelcells;
    accelcells.x = max(int((volsz.x * gridspacing) / acgridspacing), 1);
    accelcells.y = max(int((volsz.y * gridspacing) / acgridspacing), 1);
    accelcells.z = max(int((volsz.z * gridspacing) / acgridspacing), 1);
    if (accel) accel[0] = accelcells.x, accel[1] = accelcells.y, accel[2] = accelcells.z;
    dim3 Bsz(GBLOCKSZX, GBLOCKSZY, GBLOCKSZZ);
    if (colorperatom)
        Bsz.z = GTEXBLOCKSZZ;
    // ---- :1323-1338: the per-atom exponent factor ------------------------------------------------------------------------------------
    float4* xyzr = (float4*)malloc(natoms * sizeof(float4));
    float log2e = log2(2.718281828);
    for (long i = 0, i4 = 0; i < natoms; i++, i4 += 4) {
        xyzr[i].x = xyzr_f[i4];
        xyzr[i].y = xyzr_f[i4 + 1];
        xyzr[i].z = xyzr_f[i4 + 2];
        float scaledrad = xyzr_f[i4 + 3] * radscale;
        float arinv = -1.0f * log2e / (2.0f * scaledrad * scaledrad);
        xyzr[i].w = arinv;
    }
    // ---- buffers (alloc_bufs, :855-960) ----------------------------------------------------------------------------------------------
    const long ncells = (long)volsz.x * volsz.y * volsz.z, acncells = (long)accelcells.x * accelcells.y * accelcells.z;
    float4 *xyzr_d = nullptr, *sorted_xyzr_d = nullptr, *color_d = nullptr, *sorted_color_d = nullptr;
    unsigned int *atomIndex_d = nullptr, *sorted_atomIndex_d = nullptr, *atomHash_d = nullptr;
    uint2* cellStartEnd_d = nullptr;
    float* devdensity = nullptr;
    float3* devvoltexmap = nullptr;
    // The density kernels test only the FIRST of their GUNROLL / GTEXUNROLL unrolled planes against volsz.z and store the others
    // unconditionally (CUDAQuickSurf.cu:645-657): a z extent that is no multiple of 8 is written up to 7 planes past its end.  In the
    // reference those stores land in the rest of the caller's allocation; here the buffers carry the slack so that they land in ours.
    const long padcells = (long)volsz.x * volsz.y * (((long)volsz.z + 7) / 8 * 8);
    MMQ_CUDA(cudaMalloc((void**)&devdensity, padcells * sizeof(float)));
    MMQ_CUDA(cudaMalloc((void**)&xyzr_d, natoms * sizeof(float4)));
    MMQ_CUDA(cudaMalloc((void**)&sorted_xyzr_d, natoms * sizeof(float4)));
    MMQ_CUDA(cudaMalloc((void**)&atomIndex_d, natoms * sizeof(unsigned int)));
    MMQ_CUDA(cudaMalloc((void**)&sorted_atomIndex_d, natoms * sizeof(unsigned int)));
    MMQ_CUDA(cudaMalloc((void**)&atomHash_d, natoms * sizeof(unsigned int)));
    MMQ_CUDA(cudaMalloc((void**)&cellStartEnd_d, acncells * sizeof(uint2)));
    // vmd_cuda_build_density_atom_grid sorts the colour array unconditionally (sortAtomsColorsGenCellLists, CUDASpatialSearch.cu:237-239):
    // the buffers exist with or without colours (MegaMol's QuickSurf module always passes colours, QuickSurf.cpp:511-577)
    MMQ_CUDA(cudaMalloc((void**)&color_d, natoms * sizeof(float4)));
    MMQ_CUDA(cudaMalloc((void**)&sorted_color_d, natoms * sizeof(float4)));
    MMQ_CUDA(cudaMemset(color_d, 0, natoms * sizeof(float4)));
    if (colorperatom) {
        MMQ_CUDA(cudaMalloc((void**)&devvoltexmap, padcells * sizeof(float3)));
        MMQ_CUDA(cudaMemcpy(color_d, rgba, natoms * sizeof(float4), cudaMemcpyHostToDevice));
    }
    MMQ_CUDA(cudaMemcpy(xyzr_d, xyzr, natoms * sizeof(float4), cudaMemcpyHostToDevice));
    free(xyzr);
    // ---- :1340-1348: uniform grid acceleration structure -----------------------------------------------------------------------------
    if (vmd_cuda_build_density_atom_grid(natoms, xyzr_d, color_d, sorted_xyzr_d, sorted_color_d, atomIndex_d, sorted_atomIndex_d, atomHash_d,
            cellStartEnd_d, accelcells, 1.0f / acgridspacing) != 0)
        return -2;
    // ---- :1404-1470: one slab = the whole volume --------------------------------------------------------------------------------------
    float invacgridspacing = 1.0f / acgridspacing;
    float invisovalue = 1.0f / isovalue;
    int3 curslab = volsz;
    dim3 Gsz((curslab.x + Bsz.x - 1) / Bsz.x, (curslab.y + Bsz.y - 1) / Bsz.y, (curslab.z + (Bsz.z * GUNROLL) - 1) / (Bsz.z * GUNROLL));
    if (colorperatom)
        Gsz.z = (curslab.z + (Bsz.z * GTEXUNROLL) - 1) / (Bsz.z * GTEXUNROLL);
    if (colorperatom)
        gaussdensity_fast_tex3f<<<Gsz, Bsz, 0>>>(natoms, sorted_xyzr_d, sorted_color_d, curslab, accelcells, acgridspacing, invacgridspacing,
            cellStartEnd_d, gridspacing, 0, devdensity, devvoltexmap, invisovalue, false, 0, 0);
    else
        gaussdensity_fast<<<Gsz, Bsz, 0>>>(natoms, sorted_xyzr_d, curslab, accelcells, acgridspacing, invacgridspacing, cellStartEnd_d, gridspacing,
            0, devdensity, false, 0, 0);
    MMQ_CUDA(cudaDeviceSynchronize());
    MMQ_CUDA(cudaMemcpy(out_density, devdensity, ncells * sizeof(float), cudaMemcpyDeviceToHost));
    if (colorperatom) MMQ_CUDA(cudaMemcpy(out_rgb, devvoltexmap, ncells * sizeof(float3), cudaMemcpyDeviceToHost));
    cudaFree(devdensity), cudaFree(xyzr_d), cudaFree(sorted_xyzr_d), cudaFree(atomIndex_d), cudaFree(sorted_atomIndex_d), cudaFree(atomHash_d);
    cudaFree(cellStartEnd_d), cudaFree(devvoltexmap), cudaFree(color_d), cudaFree(sorted_color_d);
    return 0;
}

} // extern "C"
