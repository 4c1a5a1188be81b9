// CUDA programming-guide style 2-D TMA example (int32, 64x64 box of a 256x256 array)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int GW = 256, GH = 256, SW = 64, SH = 64;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
    __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SH * SW; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
int main() {
    std::vector<int> h(GW * GH);
    for (int i = 0; i < GW * GH; ++i) h[i] = i;
    int *d, *o;
    cudaMalloc(&d, GW * GH * 4); cudaMalloc(&o, SW * SH * 4);
    cudaMemcpy(d, h.data(), GW * GH * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    printf("entry point: %s qr=%d p=%p\n", cudaGetErrorString(ce), (int)qr, p);
    CUtensorMap tm{};
    cuuint64_t size[2] = {GW, GH};
    cuuint64_t stride[1] = {GW * sizeof(int)};
    cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d; tensormap words:", (int)r);
    for (int i = 0; i < 16; ++i) printf(" %016llx", ((unsigned long long*)&tm)[i]);
    printf("\n");
    kernel<<<1, 128>>>(tm, 64, 32, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    std::vector<int> ho(SW * SH);
    cudaMemcpy(ho.data(), o, SW * SH * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < SH; ++j) for (int i = 0; i < SW; ++i) if (ho[j * SW + i] != (64 + i) + GW * (32 + j)) ++bad;
    printf("mismatches: %d\n", bad);
}
