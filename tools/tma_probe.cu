// minimal TMA probe: load a (BX, BY, 1) box of doubles from a 3-D tensor into shared memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <cstdlib>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int BX = 38, BY = 14;
__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
struct Args { CUtensorMap tm; double* out; int x, y, z, variant, bx; const CUtensorMap* gtm; };
__global__ void probe(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ __align__(8) unsigned long long bar;
    double* s = (double*)raw;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&bar)), "r"(1));
        if (!(a.variant & 1)) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar)), "r"(a.bx * BY * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(su32(s)), "l"((a.variant & 8) ? a.gtm : &a.tm), "r"(a.x), "r"(a.y), "r"(a.z), "r"(su32(&bar)) : "memory");
    }
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(su32(&bar)), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) a.out[i] = s[i];
}
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    int S[3] = {264, 263, 40};
    size_t n = (size_t)S[0] * S[1] * S[2];
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double)i;
    double *d, *o;
    cudaMalloc(&d, n * 8); cudaMalloc(&o, BX * BY * 8);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    Args a;
    cuuint64_t dims[3] = {(cuuint64_t)S[0], (cuuint64_t)S[1], (cuuint64_t)S[2]};
    cuuint64_t st[2] = {(cuuint64_t)S[0] * 8, (cuuint64_t)S[0] * S[1] * 8};
    int bx = (variant & 4) ? 32 : BX;
    cuuint32_t box[3] = {(cuuint32_t)bx, BY, 1}, es[3] = {1, 1, 1};
    CUtensorMapDataType dt = (variant & 2) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    if (variant & 2) { dims[0] *= 2; box[0] *= 2; }
    a.variant = variant; a.bx = bx;
    CUresult r = ((EncodeFn)p)(&a.tm, dt, 3, d, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d\n", (int)r);
    CUtensorMap* gtm; cudaMalloc(&gtm, sizeof(CUtensorMap)); cudaMemcpy(gtm, &a.tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice); a.gtm = gtm;
    a.out = o; a.x = (variant & 2) ? 2 : 1; a.y = 5; a.z = 7;
    probe<<<1, 128, BX * BY * 8 + 128>>>(a);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel -> %s\n", cudaGetErrorString(e));
    std::vector<double> ho(BX * BY);
    cudaMemcpy(ho.data(), o, BX * BY * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int j = 0; j < BY; ++j) for (int i = 0; i < BX; ++i) {
        double want = (double)((1 + i) + (size_t)S[0] * ((a.y + j) + (size_t)S[1] * a.z));
        if (i < bx && ho[j * bx + i] != want) ++bad;
    }
    printf("mismatches: %d\n", bad);
    return 0;
}
