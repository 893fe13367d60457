// Stand-alone probe of the 3-D TMA plane load used by mc_emit_kernel (debug helper).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smemAddr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
template<int MODE>
__global__ void k(const __grid_constant__ CUtensorMap map, float* out, int bx, int by, int x, int y, int z) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* dst = reinterpret_cast<float*>(raw);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(raw + 8192);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bx * by * 4) : "memory");
        if (MODE == 0)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smemAddr(dst)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(z), "r"(smemAddr(bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smemAddr(dst)), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(x), "r"(y), "r"(z), "r"(smemAddr(bar)) : "memory");
    }
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smemAddr(bar)), "r"(0) : "memory");
    } while (!done);
    for (int i = threadIdx.x; i < bx * by; i += blockDim.x) out[i] = dst[i];
}
int main(int argc, char** argv) {
    const int sx = atoi(argv[1]), sy = atoi(argv[2]), sz = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]);
    const int x = atoi(argv[6]), y = atoi(argv[7]), z = atoi(argv[8]), mode = atoi(argv[9]), l2 = atoi(argv[10]);
    std::vector<float> h(size_t(sx) * sy * sz);
    for (size_t i = 0; i < h.size(); ++i) h[i] = float(i);
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, bx * by * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    using Encode = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q{};
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: err %d q %d fn %p\n", (int)e, (int)q, fn);
    CUtensorMap map{};
    const cuuint64_t dims[3] = {(cuuint64_t)sx, (cuuint64_t)sy, (cuuint64_t)sz};
    const cuuint64_t strides[2] = {(cuuint64_t)sx * 4, (cuuint64_t)sx * sy * 4};
    const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<Encode>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    if (mode == 0) k<0><<<1, 128, 16384>>>(map, o, bx, by, x, y, z);
    else k<1><<<1, 128, 16384>>>(map, o, bx, by, x, y, z);
    e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e) return 1;
    std::vector<float> got(bx * by);
    cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < by; ++j)
        for (int i = 0; i < bx; ++i) {
            const int gx = x + i, gy = y + j;
            const bool in = gx >= 0 && gx < sx && gy >= 0 && gy < sy && z >= 0 && z < sz;
            const float want = in ? h[gx + size_t(sx) * (gy + size_t(sy) * z)] : 0.0f;
            if (got[j * bx + i] != want) ++bad;
        }
    printf("mismatches: %d of %d (first %g %g %g)\n", bad, bx * by, got[0], got[1], got[bx]);
    return 0;
}
