// TMA probe using libcu++'s own wrappers (cuda/barrier)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
#include <cstdlib>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BX = 38, BY = 14;
__constant__ int g_bx;
__global__ void probe(const __grid_constant__ CUtensorMap tm, double* out, int x, int y, int z, int mode, const double* src) {
    __shared__ alignas(128) double s[BX * BY + 64];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (mode == 0) {
            cde::cp_async_bulk_tensor_3d_global_to_shared(s, &tm, x, y, z, bar);
            token = cuda::device::barrier_arrive_tx(bar, 1, g_bx * BY * 8);
        } else {
            cde::cp_async_bulk_global_to_shared(s, src, 4096, bar);
            token = cuda::device::barrier_arrive_tx(bar, 1, 4096);
        }
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    int S[3] = {264, 263, 40};
    size_t n = (size_t)S[0] * S[1] * S[2];
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, BX * BY * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)S[0], (cuuint64_t)S[1], (cuuint64_t)S[2]};
    cuuint64_t st[2] = {(cuuint64_t)S[0] * 8, (cuuint64_t)S[0] * S[1] * 8};
    int bx = argc > 2 ? atoi(argv[2]) : BX; int l2 = argc > 3 ? atoi(argv[3]) : 0;
    cudaMemcpyToSymbol(g_bx, &bx, 4);
    cuuint32_t box[3] = {(cuuint32_t)bx, BY, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    probe<<<1, 128>>>(tm, o, 1, 5, 7, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d kernel -> %s\n", mode, cudaGetErrorString(e));
    std::vector<double> ho(BX * BY);
    cudaMemcpy(ho.data(), o, BX * BY * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BY; ++j) for (int i = 0; i < BX; ++i) {
        double want = mode == 0 ? (double)((1 + i) + (size_t)S[0] * ((5 + j) + (size_t)S[1] * 7)) : (double)(j * BX + i);
        if (mode == 0 ? (i < bx && ho[j * bx + i] != want) : (j * BX + i < 512 && ho[j * BX + i] != want)) ++bad;
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
