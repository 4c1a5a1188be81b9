// tma_util.cuh -- PTX helpers (mbarrier, cp.async.bulk.tensor) and the tensor-map cache shared by the TMA-staged
// tendency kernels (tendency_tma.cu, tendency_fused.cu, tendency_bz.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace ob {
namespace tmau {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// 3-D tiled map over a field in the internal layout (common.cuh), box (box_x, box_y, 1), no swizzle.  Encodings are
// cached per (address, extents, element size, box); evict_maps() drops the entries of a freed allocation
// (called by the field destructor, so a recycled address can never meet a stale entry and the cache stays bounded).
template <class FT> CUtensorMap make_map(const GridD<FT>& g, const FT* base, int box_x, int box_y);
void evict_maps(const void* base, size_t bytes);
size_t cached_maps();

}  // namespace tmau
}  // namespace ob
