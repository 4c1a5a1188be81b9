// common.cuh -- device-side grid descriptor, field layout and shared helpers.
//
// Internal field layout (all fields of one grid share it, whatever their location):
//   x fastest; Julia index (i, j, k) lives at  base + (i-1+O[0]) + (j-1+O[1])*sy + (k-1+O[2])*sz
//   O[0] = H rounded up so that interior row starts are 32-byte aligned (sector aligned),
//   S[d] = O[d] + N[d] + H[d] + 1 (room for the extra Face point of Bounded dims), S[0] rounded
//   up to the same alignment.  Flat dimensions: O = 0, S = 1.
// The reference's parent layout (Grids/new_data.jl:16-22) is converted at upload/download.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <stdexcept>

#define OB_PERIODIC 0
#define OB_BOUNDED 1
#define OB_FLAT 2
#define OB_COMM 3      // FullyConnected: periodic-like, halos filled by neighbour exchange
#define OB_C 0
#define OB_F 1

namespace ob {

template <class FT>
struct GridD {
    int N[3], H[3], topo[3], O[3], S[3];
    long long st[3];        // element strides (1, sy, sz)
    long long off0;         // offset of Julia index (0,0,0): p0 = base + off0, then p0[i + j*sy + k*sz]
    long long total;        // allocated elements per field
    int regular[3];
    FT d[3];                // spacing of regular (and Flat: 1) dimensions
    FT invd[3];             // 1 / d[] (regular dimensions: derivatives multiply instead of dividing)
    const FT* dC[3];        // stretched: Δ at centers, pre-offset: dC[d][i] with Julia index i
    const FT* dF[3];        // stretched: Δ at faces
    const FT* izC;          // stretched z: 1 / dC[2][k], 1 / dF[2][k] (same indexing; tendency_fused.cu)
    const FT* izF;
    FT L[3];
};

template <class FT>
__host__ __device__ inline FT spacing(const GridD<FT>& g, int d, int loc, int idx) {
    if (g.regular[d]) return g.d[d];
    return loc == OB_F ? g.dF[d][idx] : g.dC[d][idx];
}

// host-side error plumbing -----------------------------------------------------------------
struct Error : std::runtime_error {
    explicit Error(const std::string& s) : std::runtime_error(s) {}
};

void count_launch(int n = 1);
cudaStream_t stream();
// CUDA-event phase timer (capi.cu: ob200_profile_*); a no-op unless profiling is enabled
struct PhaseScope {
    void* impl;
    explicit PhaseScope(const char* name);
    ~PhaseScope();
    PhaseScope(const PhaseScope&) = delete;
    PhaseScope& operator=(const PhaseScope&) = delete;
};

#define OB_CUDA(x)                                                                        \
    do {                                                                                  \
        cudaError_t e__ = (x);                                                            \
        if (e__ != cudaSuccess)                                                           \
            throw ob::Error(std::string(#x) + ": " + cudaGetErrorString(e__));            \
    } while (0)

#define OB_LAUNCH_CHECK()                                                                 \
    do {                                                                                  \
        ob::count_launch();                                                               \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess)                                                           \
            throw ob::Error(std::string("kernel launch: ") + cudaGetErrorString(e__));    \
    } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace ob
