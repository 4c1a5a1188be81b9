// parametrised tensor-TMA probe: ./tma_probe4 rank esize GW GH box_w box_h
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void kernel(const __grid_constant__ CUtensorMap tm, int rank, int x, int y, int z, int bytes, unsigned char* out) {
    extern __shared__ __align__(128) unsigned char smem[];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        if (rank == 2) cde::cp_async_bulk_tensor_2d_global_to_shared(smem, &tm, x, y, bar);
        else cde::cp_async_bulk_tensor_3d_global_to_shared(smem, &tm, x, y, z, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
int main(int argc, char** argv) {
    int rank = atoi(argv[1]), es = atoi(argv[2]), GW = atoi(argv[3]), GH = atoi(argv[4]), bw = atoi(argv[5]), bh = atoi(argv[6]);
    int GD = 8;
    size_t n = (size_t)GW * GH * GD * es;
    std::vector<unsigned char> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (unsigned char)(i * 2654435761u >> 13);
    unsigned char *d, *o;
    cudaMalloc(&d, n); cudaMalloc(&o, bw * bh * es);
    cudaMemcpy(d, h.data(), n, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    CUtensorMap tm{};
    cuuint64_t size[3] = {(cuuint64_t)GW, (cuuint64_t)GH, (cuuint64_t)GD};
    cuuint64_t stride[2] = {(cuuint64_t)GW * es, (cuuint64_t)GW * GH * es};
    cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, est[3] = {1, 1, 1};
    CUtensorMapDataType dt = es == 4 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : (es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    CUresult r = ((EncodeFn)p)(&tm, dt, rank, d, size, stride, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes = bw * bh * es;
    int x = argc > 7 ? atoi(argv[7]) : 4, y = 3, z = 2;
    kernel<<<1, 128, bytes + 256>>>(tm, rank, x, y, z, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<unsigned char> ho(bytes);
    cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < bh; ++j) for (int i = 0; i < bw * es; ++i) {
        size_t src = (size_t)(x * es + i) + (size_t)GW * es * ((y + j) + (size_t)GH * (rank == 3 ? z : 0));
        if (ho[(size_t)j * bw * es + i] != h[src]) ++bad;
    }
    printf("rank %d es %d G %dx%d box %dx%d: encode %d kernel '%s' mismatches %d\n", rank, es, GW, GH, bw, bh, (int)r, cudaGetErrorString(e), bad);
}
