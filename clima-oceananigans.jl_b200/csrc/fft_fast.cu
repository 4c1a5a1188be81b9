// fft_fast.cu -- fast path of the FFT-based Poisson solver for Periodic power-of-two dimensions:
// real-to-complex in x (half spectrum, Nx/2+1 modes), register-radix passes, fused
// divergence -> first pass, eigenvalue divide between the forward and backward transform of the
// last axis, and real part -> pressure field (+ periodic x halo) on the last pass.
//
// Reference semantics: Solvers/fft_based_poisson_solver.jl:93-125 with the pressure source term of
// Models/NonhydrostaticModels/solve_for_pressure.jl:15-18.  The reference runs an in-place c2c
// transform over a complex Nx*Ny*Nz array (6 full passes, 24 words per point); for a real source
// the half spectrum carries the same information (10-12 words per point).
//
// In-block engine: n = R1*R2(*R3) with radices <= 16; each thread does a radix-R butterfly in
// registers, data is exchanged through shared memory once per pass (2 round trips for n <= 256
// instead of log2 n for a radix-2 kernel).  Forward passes are decimation in frequency, backward
// passes decimation in time, so spectral data stays in digit-reversed order along y and z and no
// reordering pass exists; eigenvalue tables are uploaded in that order.
#include "internal.h"
#include <cuda.h>
#include <map>
#include <vector>
#include <cmath>
#include <algorithm>
#include <cstdlib>

namespace ob {
namespace comm {
bool active(); int rank(); int size(); bool peer_access_ok();
void group_start(); void group_end();
void send(const void*, size_t, int); void recv(void*, size_t, int);
void allgather_bytes(const void*, void*, size_t); void barrier(); void fast_barrier();
}
namespace ff {

template <class FT> struct Cx;
template <> struct Cx<float> { using T = float2; };
template <> struct Cx<double> { using T = double2; };

template <int LOG2N> struct Rad;
template <> struct Rad<4> { static constexpr int R1 = 16, R2 = 1, R3 = 1; };
template <> struct Rad<5> { static constexpr int R1 = 8, R2 = 4, R3 = 1; };
template <> struct Rad<6> { static constexpr int R1 = 8, R2 = 8, R3 = 1; };
template <> struct Rad<7> { static constexpr int R1 = 16, R2 = 8, R3 = 1; };
template <> struct Rad<8> { static constexpr int R1 = 16, R2 = 16, R3 = 1; };
template <> struct Rad<9> { static constexpr int R1 = 8, R2 = 8, R3 = 8; };
template <> struct Rad<10> { static constexpr int R1 = 16, R2 = 8, R3 = 8; };
template <> struct Rad<11> { static constexpr int R1 = 16, R2 = 16, R3 = 8; };

template <int LOG2N> struct Geo {
    static constexpr int N = 1 << LOG2N;
    static constexpr int RL = Rad<LOG2N>::R3 > 1 ? Rad<LOG2N>::R3 : (Rad<LOG2N>::R2 > 1 ? Rad<LOG2N>::R2 : Rad<LOG2N>::R1);
    static constexpr int LS = N + N / RL + 1;          // padded, odd line stride (in complex elements)
    __host__ __device__ static constexpr int pos(int idx) { return idx + idx / RL; }
};

template <class CT> __device__ __forceinline__ CT cadd(CT a, CT b) { CT r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <class CT> __device__ __forceinline__ CT csub(CT a, CT b) { CT r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <class CT> __device__ __forceinline__ CT cmul(CT a, CT b) { CT r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <class CT> __device__ __forceinline__ CT cmulc(CT a, CT b) { CT r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r; }

// cos/sin(2 pi k / 16), k = 0..7
__device__ constexpr double C16[8] = {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
                                      0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848};
__device__ constexpr double S16[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848,
                                      1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173};

__host__ __device__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) if (x & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int ilog2c(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

// in-register DFT of length R (2,4,8,16), DIF; result y[s] = x[brev(s)].  INV uses conjugate twiddles.
template <int R, bool INV, class CT>
__device__ __forceinline__ void dft_reg(CT* x) {
    using FT = decltype(x[0].x);
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int blk = 0; blk < R / len; ++blk) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                CT a = x[blk * len + j], b = x[blk * len + j + half];
                x[blk * len + j] = cadd(a, b);
                CT d = csub(a, b);
                const int tk = j * (16 / len);            // twiddle exp(-+ 2 pi i tk / 16)
                if (tk == 0) {
                    x[blk * len + j + half] = d;
                } else if (tk == 4) {                     // -i (forward) / +i (inverse)
                    CT r;
                    if (!INV) { r.x = d.y; r.y = -d.x; } else { r.x = -d.y; r.y = d.x; }
                    x[blk * len + j + half] = r;
                } else {
                    CT w;
                    w.x = (FT)C16[tk]; w.y = (FT)(INV ? S16[tk] : -S16[tk]);
                    x[blk * len + j + half] = cmul(d, w);
                }
            }
        }
    }
}

// one forward (DIF) pass of radix R on blocks of size m = R*S of every line; tw = exp(-2 pi i t / n)
template <int R, int LOG2N, class CT>
__device__ __forceinline__ void pass_fwd(CT* s, const CT* tw, int S, int nlines) {
    using G = Geo<LOG2N>;
    constexpr int nb = G::N / R, LB = ilog2c(R);
    const int m = R * S;
    for (int w = threadIdx.x; w < nlines * nb; w += blockDim.x) {
        int t = w / nb, b = w - t * nb;
        int blk = b / S, j = b - blk * S;
        CT* line = s + t * G::LS;
        int base = blk * m + j;
        CT x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = line[G::pos(base + q * S)];
        dft_reg<R, false>(x);
        const int tstep = j * (G::N / m);
#pragma unroll
        for (int sidx = 0; sidx < R; ++sidx) {
            CT v = x[brev(sidx, LB)];
            if (sidx > 0 && S > 1) v = cmul(v, tw[tstep * sidx]);
            line[G::pos(base + sidx * S)] = v;
        }
    }
}
// the exact inverse of pass_fwd (up to the factor R): conjugate twiddles first, then conjugate DFT
template <int R, int LOG2N, class CT>
__device__ __forceinline__ void pass_inv(CT* s, const CT* tw, int S, int nlines) {
    using G = Geo<LOG2N>;
    constexpr int nb = G::N / R, LB = ilog2c(R);
    const int m = R * S;
    for (int w = threadIdx.x; w < nlines * nb; w += blockDim.x) {
        int t = w / nb, b = w - t * nb;
        int blk = b / S, j = b - blk * S;
        CT* line = s + t * G::LS;
        int base = blk * m + j;
        const int tstep = j * (G::N / m);
        CT x[R];
#pragma unroll
        for (int sidx = 0; sidx < R; ++sidx) {
            CT v = line[G::pos(base + sidx * S)];
            if (sidx > 0 && S > 1) v = cmulc(v, tw[tstep * sidx]);
            x[sidx] = v;
        }
        dft_reg<R, true>(x);
#pragma unroll
        for (int q = 0; q < R; ++q) line[G::pos(base + q * S)] = x[brev(q, LB)];
    }
}

template <int LOG2N, class CT>
__device__ __forceinline__ void fft_fwd(CT* s, const CT* tw, int nlines) {
    using R = Rad<LOG2N>;
    constexpr int N = 1 << LOG2N;
    pass_fwd<R::R1, LOG2N>(s, tw, N / R::R1, nlines);
    __syncthreads();
    if constexpr (R::R2 > 1) { pass_fwd<R::R2, LOG2N>(s, tw, N / (R::R1 * R::R2), nlines); __syncthreads(); }
    if constexpr (R::R3 > 1) { pass_fwd<R::R3, LOG2N>(s, tw, 1, nlines); __syncthreads(); }
}
template <int LOG2N, class CT>
__device__ __forceinline__ void fft_inv(CT* s, const CT* tw, int nlines) {
    using R = Rad<LOG2N>;
    constexpr int N = 1 << LOG2N;
    if constexpr (R::R3 > 1) { pass_inv<R::R3, LOG2N>(s, tw, 1, nlines); __syncthreads(); }
    if constexpr (R::R2 > 1) { pass_inv<R::R2, LOG2N>(s, tw, N / (R::R1 * R::R2), nlines); __syncthreads(); }
    pass_inv<R::R1, LOG2N>(s, tw, N / R::R1, nlines);
    __syncthreads();
}
// frequency index held at position P after fft_fwd
static int freq_of_pos(int log2n, int P) {
    int R1, R2, R3;
    switch (log2n) {
        case 4: R1 = 16; R2 = 1; R3 = 1; break;
        case 5: R1 = 8; R2 = 4; R3 = 1; break;
        case 6: R1 = 8; R2 = 8; R3 = 1; break;
        case 7: R1 = 16; R2 = 8; R3 = 1; break;
        case 8: R1 = 16; R2 = 16; R3 = 1; break;
        case 9: R1 = 8; R2 = 8; R3 = 8; break;
        case 10: R1 = 16; R2 = 8; R3 = 8; break;
        default: R1 = 16; R2 = 16; R3 = 8; break;
    }
    int n = 1 << log2n, S1 = n / R1, S2 = S1 / R2;
    int s1 = P / S1, rem = P % S1, s2 = rem / S2, s3 = rem % S2;
    (void)R3;
    return s1 + R1 * (s2 + R2 * s3);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <class FT>
struct XArgs {
    typename Cx<FT>::T* spec;       // half spectrum, [Nz][Ny][NXP]
    int Nx, Ny, Nz, NXP, T;
    const typename Cx<FT>::T* twM;  // exp(-2 pi i t / M), M = Nx/2
    const typename Cx<FT>::T* twN;  // exp(-2 pi i k / Nx), k = 0..M
    const int* kpos;                // position (in the M-point engine) of frequency k
    // forward input: divergence of (u, v, w) / dt  (Julia-(0,0,0) pointers, strides) or a real array
    const FT* u; const FT* v; const FT* w; const FT* real_in;
    long long st[3];
    FT ax, ay, az, invV, dt;
    int has_z;
    long long wrap[3];              // Periodic dims: N * stride, so that index N+1 is read as index 1 (no halo needed)
    long long x0;                   // offset of the first element of a padded row from its Julia index 0 (1 - Ox)
    // Fourier-tridiagonal solve (Bounded, possibly stretched z): the source term is Δzᶜ_k div(U*) / Δt
    // (solve_for_pressure.jl:28-33) with the level's own areas and volume; dzC is indexed by the Julia level
    int tri;
    const FT* dzC;                  // nullptr: regular z (spacing dz)
    FT dx, dy, dz;
    // backward output
    FT* phi_p0;
    int Hx;
    FT scale;
};

// two consecutive reals as one aligned vector load / store (interior rows start 32-byte aligned, common.cuh)
template <class FT> struct Pair;
template <> struct Pair<double> { using T = double2; };
template <> struct Pair<float> { using T = float2; };

// forward x: real line -> half spectrum (natural kx order).
// Thread -> (line t = tid / tpl, lane l = tid % tpl): the row pointers of a thread are fixed and the x loops advance
// by tpl pairs, so the loops carry no index arithmetic (the first version spent half of its instructions on it:
// ncu r1g, IMAD + LEA + IADD3 + SHF = 43 % of 75 M warp instructions).
template <class FT, int LOG2M>
__global__ void __launch_bounds__(256) x_r2c_kernel(const __grid_constant__ XArgs<FT> A) {
    using CT = typename Cx<FT>::T;
    using P2 = typename Pair<FT>::T;
    using G = Geo<LOG2M>;
    constexpr int M = 1 << LOG2M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CT* s = reinterpret_cast<CT*>(smem_raw);
    CT* stw = s + A.T * G::LS;
    for (int w = threadIdx.x; w < M; w += blockDim.x) stw[w] = A.twM[w];
    const int j0 = blockIdx.x * A.T, k = blockIdx.y;
    const int nl = min(A.T, A.Ny - j0);
    const int Nx = A.Nx;
    const int tpl = blockDim.x / A.T, t = threadIdx.x / tpl, l = threadIdx.x - t * tpl;
    CT* const sl = s + t * G::LS;
    if (t < nl) {
        const int j = j0 + t;
        if (A.real_in) {
            const FT* row = A.real_in + (long long)Nx * (j + (long long)A.Ny * k);
            const FT lev = A.tri ? (A.dzC ? A.dzC[k + 1] : A.dz) : FT(1);       // set_source_term!: times Δzᶜ
            for (int m = l; m < M; m += tpl) {
                CT c; c.x = row[2 * m]; c.y = row[2 * m + 1];
                if (A.tri) { c.x *= lev; c.y *= lev; }
                sl[G::pos(m)] = c;
            }
        } else {
            // divᶜᶜᶜ on a regular grid: 1/V (Ax δx u + Ay δy v + Az δz w), then / Δt (solve_for_pressure.jl:15-18);
            // index N+1 of a Periodic dimension is read as index 1 (A.wrap), so the velocities' halos need not be valid
            const long long p = A.st[0] + (j + 1) * A.st[1] + (k + 1) * A.st[2];     // Julia (1, j+1, k+1)
            const FT* u0 = A.u + p;
            const FT* v0 = A.v + p;
            const FT* v1 = v0 + A.st[1] - (j + 1 == A.Ny ? A.wrap[1] : 0);
            const FT* w0 = A.w + p;
            const FT* w1 = w0 + A.st[2] - (k + 1 == A.Nz ? A.wrap[2] : 0);
            const FT inv_dt = FT(1) / A.dt;
            // level metrics (regular solve: the constants of the plan; tridiagonal solve: this level's Δzᶜ)
            FT ax = A.ax, ay = A.ay, invV = A.invV, lev = FT(1);
            if (A.tri) {
                const FT dzk = A.dzC ? A.dzC[k + 1] : A.dz;
                ax = A.dy * dzk; ay = A.dx * dzk; invV = 1 / ((A.dx * A.dy) * dzk); lev = dzk;
            }
            constexpr int XU = 4;
            for (int mb = l; mb < M; mb += XU * tpl) {
                P2 ua[XU], va[XU], vb[XU], wa[XU], wb[XU];
                FT un[XU];
#pragma unroll
                for (int e = 0; e < XU; ++e) {
                    const int m = mb + e * tpl;
                    if (m < M) {
                        ua[e] = *reinterpret_cast<const P2*>(u0 + 2 * m);
                        un[e] = u0[2 * m + 2 - (2 * m + 2 == Nx ? A.wrap[0] : 0)];
                        va[e] = *reinterpret_cast<const P2*>(v0 + 2 * m);
                        vb[e] = *reinterpret_cast<const P2*>(v1 + 2 * m);
                        if (A.has_z) {
                            wa[e] = *reinterpret_cast<const P2*>(w0 + 2 * m);
                            wb[e] = *reinterpret_cast<const P2*>(w1 + 2 * m);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < XU; ++e) {
                    const int m = mb + e * tpl;
                    if (m < M) {
                        FT tx0 = ax * ua[e].y - ax * ua[e].x, tx1 = ax * un[e] - ax * ua[e].y;
                        FT ty0 = ay * vb[e].x - ay * va[e].x, ty1 = ay * vb[e].y - ay * va[e].y;
                        FT tz0 = FT(0), tz1 = FT(0);
                        if (A.has_z) { tz0 = A.az * wb[e].x - A.az * wa[e].x; tz1 = A.az * wb[e].y - A.az * wa[e].y; }
                        CT c;
                        c.x = (lev * (invV * ((tx0 + ty0) + tz0))) * inv_dt;
                        c.y = (lev * (invV * ((tx1 + ty1) + tz1))) * inv_dt;
                        sl[G::pos(m)] = c;
                    }
                }
            }
        }
    }
    __syncthreads();
    fft_fwd<LOG2M>(s, stw, nl);
    // untangle: X[k] = E[k] + w_N^k O[k], E = (Z[k] + conj Z[M-k]) / 2, O = (Z[k] - conj Z[M-k]) / (2i)
    if (t < nl) {
        CT* out = A.spec + (long long)A.NXP * ((j0 + t) + (long long)A.Ny * k);
        for (int kk = l; kk <= M; kk += tpl) {
            CT a = sl[G::pos(A.kpos[kk & (M - 1)])];
            CT b = sl[G::pos(A.kpos[(M - kk) & (M - 1)])];
            CT E, O;
            E.x = FT(0.5) * (a.x + b.x); E.y = FT(0.5) * (a.y - b.y);
            O.x = FT(0.5) * (a.y + b.y); O.y = FT(-0.5) * (a.x - b.x);
            out[kk] = cadd(E, cmul(O, A.twN[kk]));
        }
    }
}

// backward x: half spectrum -> real line, written into the haloed field (+ periodic x halos)
template <class FT, int LOG2M>
__global__ void __launch_bounds__(256) x_c2r_kernel(const __grid_constant__ XArgs<FT> A) {
    using CT = typename Cx<FT>::T;
    using P2 = typename Pair<FT>::T;
    using G = Geo<LOG2M>;
    constexpr int M = 1 << LOG2M;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CT* s = reinterpret_cast<CT*>(smem_raw);
    CT* stw = s + A.T * G::LS;
    for (int w = threadIdx.x; w < M; w += blockDim.x) stw[w] = A.twM[w];
    const int j0 = blockIdx.x * A.T, k = blockIdx.y;
    const int nl = min(A.T, A.Ny - j0);
    const int tpl = blockDim.x / A.T, t = threadIdx.x / tpl, l = threadIdx.x - t * tpl;
    CT* const sl = s + t * G::LS;
    // tangle: Z[k] = E[k] + i O[k], E = (X[k] + conj X[M-k]) / 2, O = conj(w_N^k) (X[k] - conj X[M-k]) / 2
    if (t < nl) {
        const CT* row = A.spec + (long long)A.NXP * ((j0 + t) + (long long)A.Ny * k);
        constexpr int XU = 4;
        for (int kb = l; kb < M; kb += XU * tpl) {
            CT av[XU], bv[XU];
#pragma unroll
            for (int e = 0; e < XU; ++e) {
                const int kk = kb + e * tpl;
                if (kk < M) { av[e] = row[kk]; bv[e] = row[M - kk]; }
            }
#pragma unroll
            for (int e = 0; e < XU; ++e) {
                const int kk = kb + e * tpl;
                if (kk < M) {
                    CT a = av[e], b = bv[e];
                    CT E, D;
                    E.x = FT(0.5) * (a.x + b.x); E.y = FT(0.5) * (a.y - b.y);
                    D.x = FT(0.5) * (a.x - b.x); D.y = FT(0.5) * (a.y + b.y);
                    CT O = cmulc(D, A.twN[kk]);
                    CT Z; Z.x = E.x - O.y; Z.y = E.y + O.x;
                    sl[G::pos(A.kpos[kk])] = Z;
                }
            }
        }
    }
    __syncthreads();
    fft_inv<LOG2M>(s, stw, nl);
    const int Nx = A.Nx, H = A.Hx;
    if (t < nl) {
        FT* row = A.phi_p0 + (j0 + t + 1) * A.st[1] + (k + 1) * A.st[2];      // Julia (0, j, k)
        for (int m = l; m < M; m += tpl) {
            CT z = sl[G::pos(m)];
            P2 r; r.x = z.x * A.scale; r.y = z.y * A.scale;
            const int i = 2 * m + 1;                 // Julia index of the first of the two reals
            *reinterpret_cast<P2*>(row + i) = r;
            // periodic halos in x (fill_halo_regions_periodic.jl:37-46)
            if (i > Nx - H) row[i - Nx] = r.x;
            if (i + 1 > Nx - H) row[i + 1 - Nx] = r.y;
            if (i <= H) row[i + Nx] = r.x;
            if (i + 1 <= H) row[i + 1 + Nx] = r.y;
        }
    }
}

// address of spectral element (kx, m, o): kx blocked by kxb (for the all-to-all chunks), the line
// index m optionally split in (m / split, m % split) (y lines gathered from R ranks)
struct Lay {
    int kxb, split;
    long long s_blk, s_ml, s_mh, s_o;
    // optional per-block base pointers (peer memory of the other ranks, mapped with CUDA IPC): the block id is
    // the kx block (pmode 1) or the upper part of the split line index (pmode 2); the block stride is then unused
    int pmode;
    float inv_kxb, inv_split;
    void* ptr[8];
    template <class CT> __device__ __forceinline__ CT* addr(CT* base, int kx, int m, int o) const {
        if (pmode == 0 && kxb >= (1 << 30) && split >= (1 << 30)) return base + (kx + m * s_ml + o * s_o);   // natural
        // exact small-integer division by multiplication with a float reciprocal ((k + 1/2) / d is never
        // within rounding distance of an integer for k < 2^12)
        int kb = kxb >= (1 << 30) ? 0 : __float2int_rz(((float)kx + 0.5f) * inv_kxb);
        int mh = split >= (1 << 30) ? 0 : __float2int_rz(((float)m + 0.5f) * inv_split);
        int kl = kx - kb * kxb, ml = m - mh * split;
        if (pmode == 1) return (CT*)ptr[kb] + (kl + ml * s_ml + mh * s_mh + o * s_o);
        if (pmode == 2) return (CT*)ptr[mh] + (kb * s_blk + kl + ml * s_ml + o * s_o);
        return base + (kb * s_blk + kl + ml * s_ml + mh * s_mh + o * s_o);
    }
};

template <class FT>
struct LArgs {
    const typename Cx<FT>::T* in;
    typename Cx<FT>::T* out;
    Lay lin, lout;
    int n, NXH, nOther;            // lines: kx in [0, NXH), other index in [0, nOther)
    int kx0;                       // global kx of local kx 0 (slab-decomposed spectral space)
    int T;
    const typename Cx<FT>::T* tw;
    FT scale;
    const double* lamx;            // natural kx (global)
    const double* lamL;            // along the line, position order
    const double* lamO;            // along the other (non-x) dimension, its storage order
    int line_is_y;                 // 1: line along y (other = z) ; 0: line along z (other = y)
};

enum { LM_FWD = 0, LM_INV = 1, LM_FWD_DIV_INV = 2, LM_COPY = 3 };       // LM_COPY: change of layout only

// pointer to element (kx, m = 0, o) and a functor for the m-th element of that line: the natural layout is an
// affine function of m; the blocked / split / peer-memory layouts of the slab-decomposed solve go through Lay::addr
template <class CT>
struct LinePtr {
    CT* p0; long long sm; const Lay* lay; CT* base; int kx, o; bool natural;
    __device__ __forceinline__ LinePtr(const Lay& l, CT* b, int kx_, int o_) : lay(&l), base(b), kx(kx_), o(o_) {
        natural = l.pmode == 0 && l.kxb >= (1 << 30) && l.split >= (1 << 30);
        p0 = b + (kx_ + o_ * l.s_o); sm = l.s_ml;
    }
    __device__ __forceinline__ CT* at(int m) const { return natural ? p0 + m * sm : lay->addr(base, kx, m, o); }
};

__device__ __forceinline__ double rcp_full(double x) {       // 1/x to double precision: MUFU seed + one cubic step + one Newton step
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, fma(e, e, e), r);
    return fma(r, fma(-x, r, 1.0), r);
}

template <class FT, int LOG2N, int MODE>
__global__ void __launch_bounds__(256) line_kernel(const __grid_constant__ LArgs<FT> A) {
    using CT = typename Cx<FT>::T;
    using G = Geo<LOG2N>;
    using R = Rad<LOG2N>;
    constexpr int N = 1 << LOG2N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CT* s = reinterpret_cast<CT*>(smem_raw);
    CT* stw = s + A.T * G::LS;
    if constexpr (MODE != LM_COPY)
        for (int w = threadIdx.x; w < N; w += blockDim.x) stw[w] = A.tw[w];
    const int x0 = blockIdx.x * A.T, o = blockIdx.y;
    const int nl = min(A.T, A.NXH - x0);
    // thread -> (line t, first element m0); blockDim is a multiple of T, so t is fixed and m advances by mstep:
    // no integer division in the load / store loops (they were ~2/3 of the instructions of the first version)
    const int t = threadIdx.x % A.T, m0 = threadIdx.x / A.T, mstep = blockDim.x / A.T;
    CT* const sl = s + t * G::LS;
    if (t < nl) {
        LinePtr<const CT> in(A.lin, A.in, x0 + t, o);
        constexpr int LU = 8;
        for (int m = m0; m < N; m += LU * mstep) {
            CT v[LU];
#pragma unroll
            for (int e = 0; e < LU; ++e) if (m + e * mstep < N) v[e] = *in.at(m + e * mstep);
#pragma unroll
            for (int e = 0; e < LU; ++e) if (m + e * mstep < N) sl[G::pos(m + e * mstep)] = v[e];
        }
    }
    __syncthreads();
    if constexpr (MODE == LM_FWD) fft_fwd<LOG2N>(s, stw, nl);
    if constexpr (MODE == LM_INV) fft_inv<LOG2N>(s, stw, nl);
    if constexpr (MODE == LM_FWD_DIV_INV) {
        // all forward passes but the last one
        constexpr int RLAST = R::R3 > 1 ? R::R3 : (R::R2 > 1 ? R::R2 : R::R1);
        if constexpr (R::R2 > 1) { pass_fwd<R::R1, LOG2N>(s, stw, N / R::R1, nl); __syncthreads(); }
        if constexpr (R::R3 > 1) { pass_fwd<R::R2, LOG2N>(s, stw, N / (R::R1 * R::R2), nl); __syncthreads(); }
        // last forward pass (stride 1, no twiddles), eigenvalue divide and first backward pass, all in registers:
        // phi_hat = -b_hat / (lx + ly + lz), zero mode = 0 (fft_based_poisson_solver.jl:106-111)
        constexpr int nb = N / RLAST, LB = ilog2c(RLAST);
        for (int w = threadIdx.x; w < nl * nb; w += blockDim.x) {
            const int tt = w / nb, blk = w - tt * nb, base = blk * RLAST;
            CT* line = s + tt * G::LS;
            CT x[RLAST], y[RLAST];
#pragma unroll
            for (int q = 0; q < RLAST; ++q) x[q] = line[G::pos(base + q)];
            dft_reg<RLAST, false>(x);
            const int kx = A.kx0 + x0 + tt;
            const double lxo = A.line_is_y ? A.lamx[kx] : (A.lamx[kx] + (A.lamO ? A.lamO[o] : 0.0));
            const double lO = A.lamO ? A.lamO[o] : 0.0;
#pragma unroll
            for (int sidx = 0; sidx < RLAST; ++sidx) {
                const int m = base + sidx;
                const double lL = A.lamL[m];
                const double lam = A.line_is_y ? ((lxo + lL) + lO) : (lxo + lL);
                const CT v = x[brev(sidx, LB)];
                const double r = (kx == 0 && m == 0 && o == 0) ? 0.0 : -rcp_full(lam);
                y[sidx].x = (FT)((double)v.x * r); y[sidx].y = (FT)((double)v.y * r);
            }
            dft_reg<RLAST, true>(y);
#pragma unroll
            for (int q = 0; q < RLAST; ++q) line[G::pos(base + q)] = y[brev(q, LB)];
        }
        __syncthreads();
        if constexpr (R::R3 > 1) { pass_inv<R::R2, LOG2N>(s, stw, N / (R::R1 * R::R2), nl); __syncthreads(); }
        if constexpr (R::R2 > 1) { pass_inv<R::R1, LOG2N>(s, stw, N / R::R1, nl); __syncthreads(); }
    }
    if (t < nl) {
        LinePtr<CT> out(A.lout, A.out, x0 + t, o);
        constexpr int LU = 8;
        for (int m = m0; m < N; m += LU * mstep) {
            CT v[LU];
#pragma unroll
            for (int e = 0; e < LU; ++e) if (m + e * mstep < N) {
                v[e] = sl[G::pos(m + e * mstep)];
                if (MODE != LM_FWD && MODE != LM_COPY) { v[e].x *= A.scale; v[e].y *= A.scale; }
            }
#pragma unroll
            for (int e = 0; e < LU; ++e) if (m + e * mstep < N) *out.at(m + e * mstep) = v[e];
        }
    }
}

#include "fft_tma.cuh"
#include "fft_xtma.cuh"

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
template <class FT>
struct FastPoisson {
    using CT = typename Cx<FT>::T;
    int N[3], log2[3], has_z, NXH, NXP;
    int R = 1, rank = 0, NyG = 0, KXB = 0;      // slab decomposition in y: NyG = R * N[1], KXB = NXP / R
    CT* bufA = nullptr; CT* bufB = nullptr;     // all-to-all staging (distributed only)
    CT* peerA[8] = {}; CT* peerB[8] = {};       // the same buffers of every rank, mapped through CUDA IPC
    bool p2p = false;
    CT* spec = nullptr;
    CT* twM = nullptr; CT* twN = nullptr; CT* twY = nullptr; CT* twZ = nullptr;
    int* kpos = nullptr;
    double *lamx = nullptr, *lamy = nullptr, *lamz = nullptr;
    std::vector<void*> owned;
    int tri = 0;                                // Fourier-tridiagonal solve: z is Bounded and solved by Thomas
    std::function<void()> zhook;                // ... or transformed by the caller's DCT pass (regular z, FFT-based solver)
    std::function<void(int)> yhook;             // Bounded y: the caller's forward (0) / backward (1) DCT pass over the y lines
    const double *dzF = nullptr, *dzC = nullptr; // device, Julia-indexed 0..Nz+1 (owned by the PoissonPlan)
    FT* tsc = nullptr;                          // Thomas scratch t, [Nz][Ny][NXP]
    bool tma_ok = false;                        // persistent TMA-pipelined y / z passes (fft_tma.cuh)
    CUtensorMap tm_y, tm_z;
    // slab-decomposed solve through TMA (peer-memory tensor maps): see distributed_middle_tma
    bool dtma_ok = false;
    // bulk = true: the transform kernels store into LOCAL chunk buffers and the chunks travel to their ranks as large
    // contiguous copies (copy engines over NVLink); false: the kernels store straight into the peers' buffers
    bool bulk_a2a = false;
    bool ysplit = false;                        // gathered y lines as R-point butterflies across chunks + NyL-point lines
    CUtensorMap tm4_ys, tmr_ys[8];
    double* lamy2 = nullptr;                    // [R][NyL]: eigenvalue of frequency r + R freq_of_pos(m)
    typename Cx<FT>::T* twYL = nullptr;         // exp(-2 pi i t / NyL): twiddles of the NyL-point lines
    double Ly_global = 0;
    // copy streams (one per peer: copies to different peers run on different copy engines; a single stream serialised
    // them at ~310 GB/s) + events of the pipelined bulk transposes
    cudaStream_t cpy = nullptr, cpys[8] = {};
    cudaEvent_t ev_piece[16] = {}, ev_done = nullptr, ev_dones[8] = {};
    CT* bufC = nullptr;
    int dtk_y = 8;                              // columns per y tile (8, 4, 2 for gathered lines of <= 512, 1024, 2048)
    CUtensorMap tm4_zi, tm4_y, tmr_zf[8], tmr_zi[8], tmr_y[8];
};

namespace cm = ::ob::comm;
template <class T> static T* up(const std::vector<T>& h, std::vector<void*>& owned) {
    T* d = nullptr;
    OB_CUDA(cudaMalloc(&d, std::max<size_t>(1, h.size()) * sizeof(T)));
    OB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    owned.push_back(d);
    return d;
}
static bool pow2(int n) { return n >= 2 && (n & (n - 1)) == 0; }

template <class FT>
bool fast_poisson_supported(const GridD<FT>& g) {
    if (g.topo[0] != OB_PERIODIC || (g.topo[1] != OB_PERIODIC && g.topo[1] != OB_COMM)) return false;
    if (g.topo[2] == OB_BOUNDED) return false;
    int R = g.topo[1] == OB_COMM ? comm::size() : 1;
    if (g.topo[1] == OB_COMM && (g.topo[2] != OB_PERIODIC || !pow2(R) || R > 8)) return false;
    if (!pow2(g.N[0]) || g.N[0] < 32 || g.N[0] > 2048) return false;
    if (!pow2(g.N[1]) || g.N[1] * R < 16 || g.N[1] * R > 2048) return false;
    if (g.topo[2] == OB_PERIODIC && (!pow2(g.N[2]) || g.N[2] < 16 || g.N[2] > 1024)) return false;
    for (int d = 0; d < 3; ++d) if (!g.regular[d]) return false;
    return true;
}

template <class FT> static void setup_tma(FastPoisson<FT>* p);

// Fourier-tridiagonal fast path: x, y Periodic power-of-two regular; z Bounded with any spacing; y may be slab-decomposed
template <class FT>
bool fast_ft_supported(const GridD<FT>& g) {
    if (g.topo[0] != OB_PERIODIC || (g.topo[1] != OB_PERIODIC && g.topo[1] != OB_COMM) || g.topo[2] != OB_BOUNDED) return false;
    const int R = g.topo[1] == OB_COMM ? comm::size() : 1;
    if (!pow2(g.N[0]) || g.N[0] < 32 || g.N[0] > 2048) return false;
    if (!pow2(g.N[1]) || g.N[1] * R < 16 || g.N[1] * R > 2048) return false;
    if (!g.regular[0] || !g.regular[1] || g.N[2] < 2) return false;
    // slab-decomposed in y: the z lines change layout through the power-of-two line kernel (distributed_middle_tri)
    if (R > 1 && (!pow2(R) || R > 8 || !pow2(g.N[2]) || g.N[2] < 16 || g.N[2] > 1024)) return false;
    return true;
}

// x Periodic power-of-two, y Bounded power-of-two (transformed by the caller's hook), z Periodic power-of-two or Bounded;
// regular x, y; one GPU
template <class FT>
bool fast_bounded_y_supported(const GridD<FT>& g) {
    if (g.topo[0] != OB_PERIODIC || g.topo[1] != OB_BOUNDED || g.topo[2] == OB_FLAT || g.topo[2] == OB_COMM) return false;
    if (!pow2(g.N[0]) || g.N[0] < 32 || g.N[0] > 2048 || !pow2(g.N[1]) || g.N[1] < 16 || g.N[1] > 2048) return false;
    if (!g.regular[0] || !g.regular[1]) return false;
    if (g.topo[2] == OB_PERIODIC && (!pow2(g.N[2]) || g.N[2] < 16 || g.N[2] > 1024 || !g.regular[2])) return false;
    return g.N[2] >= 2;
}
template bool fast_bounded_y_supported<float>(const GridD<float>&);
template bool fast_bounded_y_supported<double>(const GridD<double>&);

template <class FT>
FastPoisson<FT>* fast_poisson_create(const GridD<FT>& g) {
    using CT = typename Cx<FT>::T;
    auto* p = new FastPoisson<FT>();
    const long double PI = 3.14159265358979323846264338327950288L;
    for (int d = 0; d < 3; ++d) { p->N[d] = g.N[d]; p->log2[d] = ilog2c(g.N[d]); }
    p->has_z = g.topo[2] != OB_FLAT;
    p->tri = g.topo[2] == OB_BOUNDED;           // only reached through fast_ft_supported
    if (g.topo[1] == OB_COMM) { p->R = cm::size(); p->rank = cm::rank(); }
    p->NyG = p->R * g.N[1];
    p->log2[1] = ilog2c(p->NyG);
    int Nx = g.N[0], M = Nx / 2, lm = p->log2[0] - 1;
    p->NXH = M + 1;
    p->NXP = ((p->NXH + 7) / 8) * 8;           // multiple of 8 hence of R (R in {1,2,4,8})
    p->KXB = p->NXP / p->R;
    size_t tot = (size_t)p->NXP * g.N[1] * g.N[2];
    OB_CUDA(cudaMalloc(&p->spec, tot * sizeof(CT)));
    OB_CUDA(cudaMemset(p->spec, 0, tot * sizeof(CT)));
    if (p->R > 1) {
        OB_CUDA(cudaMalloc(&p->bufA, tot * sizeof(CT)));
        OB_CUDA(cudaMalloc(&p->bufB, tot * sizeof(CT)));
        OB_CUDA(cudaMemset(p->bufA, 0, tot * sizeof(CT)));
        OB_CUDA(cudaMemset(p->bufB, 0, tot * sizeof(CT)));
        if (cm::peer_access_ok()) {            // collective probe (comm.cu); otherwise NCCL all-to-all
            // exchange CUDA IPC handles of bufA / bufB so that the transform kernels can store straight into the
            // destination rank's buffer over NVLink (the transfer is part of the kernel, not a separate collective)
            struct H2 { cudaIpcMemHandle_t a, b; };
            H2 mine;
            OB_CUDA(cudaIpcGetMemHandle(&mine.a, p->bufA));
            OB_CUDA(cudaIpcGetMemHandle(&mine.b, p->bufB));
            H2 *dsend, *drecv;
            OB_CUDA(cudaMalloc(&dsend, sizeof(H2)));
            OB_CUDA(cudaMalloc(&drecv, sizeof(H2) * p->R));
            OB_CUDA(cudaMemcpyAsync(dsend, &mine, sizeof(H2), cudaMemcpyHostToDevice, stream()));
            cm::allgather_bytes(dsend, drecv, sizeof(H2));
            std::vector<H2> all(p->R);
            OB_CUDA(cudaMemcpyAsync(all.data(), drecv, sizeof(H2) * p->R, cudaMemcpyDeviceToHost, stream()));
            OB_CUDA(cudaStreamSynchronize(stream()));
            cudaFree(dsend); cudaFree(drecv);
            for (int r = 0; r < p->R; ++r) {
                if (r == p->rank) { p->peerA[r] = p->bufA; p->peerB[r] = p->bufB; continue; }
                OB_CUDA(cudaIpcOpenMemHandle((void**)&p->peerA[r], all[r].a, cudaIpcMemLazyEnablePeerAccess));
                OB_CUDA(cudaIpcOpenMemHandle((void**)&p->peerB[r], all[r].b, cudaIpcMemLazyEnablePeerAccess));
            }
            p->p2p = true;
        }
    }
    auto twid = [&](int n, int count) {
        std::vector<CT> t(count);
        for (int k = 0; k < count; ++k) {
            long double a = -2.0L * PI * k / n;
            t[k].x = (FT)cosl(a); t[k].y = (FT)sinl(a);
        }
        return t;
    };
    p->twM = up(twid(M, M), p->owned);
    p->twN = up(twid(Nx, M + 1), p->owned);
    p->twY = up(twid(p->NyG, p->NyG), p->owned);
    if (p->has_z && !p->tri) p->twZ = up(twid(g.N[2], g.N[2]), p->owned);
    if (p->tri) {
        OB_CUDA(cudaMalloc(&p->tsc, tot * sizeof(FT)));
        p->owned.push_back(p->tsc);
    }
    std::vector<int> kp(M);
    for (int P = 0; P < M; ++P) kp[freq_of_pos(lm, P)] = P;
    p->kpos = up(kp, p->owned);
    // eigenvalues (poisson_eigenvalues.jl:8-11), Float64
    auto lam = [&](int d, int i) {
        // global extent and size of the dimension (the local slab holds 1/R of y)
        double L = (double)g.L[d] * (d == 1 ? p->R : 1);
        int n = d == 1 ? p->NyG : g.N[d];
        double v = 2 * sin(i * (double)PI / n) / (L / n);
        return v * v;
    };
    std::vector<double> lx(p->NXP, 1.0), ly(p->NyG), lz(std::max(1, g.N[2]), 0.0);
    for (int k = 0; k <= M; ++k) lx[k] = lam(0, k);
    for (int P = 0; P < p->NyG; ++P) ly[P] = lam(1, freq_of_pos(p->log2[1], P));
    if (p->has_z && !p->tri) for (int P = 0; P < g.N[2]; ++P) lz[P] = lam(2, freq_of_pos(p->log2[2], P));
    p->lamx = up(lx, p->owned); p->lamy = up(ly, p->owned); p->lamz = up(lz, p->owned);
    p->Ly_global = (double)g.L[1] * p->R;
    setup_tma(p);
    return p;
}
// the Δzᶠ / Δzᶜ tables of the tridiagonal solve (device, Julia-indexed; owned by the caller's plan)
template <class FT> void fast_poisson_set_tridiagonal(FastPoisson<FT>* p, const double* dzF_dev, const double* dzC_dev) {
    p->dzF = dzF_dev; p->dzC = dzC_dev;
}
template <class FT> FastSpecInfo fast_poisson_spec_info(FastPoisson<FT>* p) {
    return FastSpecInfo{(void*)p->spec, p->NXH, p->NXP, p->N[1], p->N[2], p->lamx, p->lamy};
}
template <class FT> void fast_poisson_set_zhook(FastPoisson<FT>* p, std::function<void()> hook) { p->zhook = std::move(hook); }
// lamy_dev: the eigenvalues of the caller's y transform in ITS storage order (they replace the Periodic table of the plan)
template <class FT> void fast_poisson_set_yhook(FastPoisson<FT>* p, std::function<void(int)> hook, const double* lamy_dev) {
    p->yhook = std::move(hook);
    p->lamy = const_cast<double*>(lamy_dev);
}
template <class FT> void fast_poisson_destroy(FastPoisson<FT>* p) {
    if (!p) return;
    cudaFree(p->spec);
    for (int r = 0; r < p->R; ++r)
        if (p->p2p && r != p->rank) { cudaIpcCloseMemHandle(p->peerA[r]); cudaIpcCloseMemHandle(p->peerB[r]); }
    if (p->bufA) cudaFree(p->bufA);
    if (p->bufB) cudaFree(p->bufB);
    for (void* q : p->owned) cudaFree(q);
    if (p->cpy) cudaStreamDestroy(p->cpy);
    for (cudaStream_t s : p->cpys) if (s) cudaStreamDestroy(s);
    for (cudaEvent_t e : p->ev_dones) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : p->ev_piece) if (e) cudaEventDestroy(e);
    if (p->ev_done) cudaEventDestroy(p->ev_done);
    delete p;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <class FT, int MODE>
static void launch_line(const LArgs<FT>& A, int log2n, dim3 grd, size_t smem) {
    // one radix-16 butterfly per thread and pass
    int threads = std::min(256, std::max(64, A.T * (1 << log2n) / 16));
    auto go = [&](auto kern) {
        OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        kern<<<grd, threads, smem, stream()>>>(A);
        OB_LAUNCH_CHECK();
    };
    switch (log2n) {
        case 4: go(line_kernel<FT, 4, MODE>); break;
        case 5: go(line_kernel<FT, 5, MODE>); break;
        case 6: go(line_kernel<FT, 6, MODE>); break;
        case 7: go(line_kernel<FT, 7, MODE>); break;
        case 8: go(line_kernel<FT, 8, MODE>); break;
        case 9: go(line_kernel<FT, 9, MODE>); break;
        case 10: go(line_kernel<FT, 10, MODE>); break;
        default: go(line_kernel<FT, 11, MODE>); break;
    }
}
static size_t line_smem(int log2n, int T, size_t csize) {
    int n = 1 << log2n;
    int RL = log2n == 4 ? 16 : (log2n == 5 ? 4 : (log2n == 8 ? 16 : 8));
    int LS = n + n / RL + 1;
    return ((size_t)T * LS + n) * csize;
}

static Lay natural_lay(int NXP, long long s_m, long long s_o) {
    Lay l{}; l.kxb = 1 << 30; l.split = 1 << 30; l.s_blk = 0; l.s_ml = s_m; l.s_mh = 0; l.s_o = s_o; l.pmode = 0;
    return l;
}

template <class FT>
static void launch_line_any(FastPoisson<FT>* p, LArgs<FT>& A, int log2n, int mode) {
    using CT = typename Cx<FT>::T;
    static const int T0 = env_int("OB200_FFT_TL", 8);
    int T = T0;
    while (T > 1 && line_smem(log2n, T, sizeof(CT)) > 100 * 1024) T >>= 1;
    A.T = T;
    A.scale = (FT)(1.0 / A.n);
    A.lamx = p->lamx;
    dim3 grd(cdiv(A.NXH, T), A.nOther);
    size_t smem = line_smem(log2n, T, sizeof(CT));
    if (mode == LM_FWD) launch_line<FT, LM_FWD>(A, log2n, grd, smem);
    else if (mode == LM_INV) launch_line<FT, LM_INV>(A, log2n, grd, smem);
    else if (mode == LM_COPY) launch_line<FT, LM_COPY>(A, log2n, grd, smem);
    else launch_line<FT, LM_FWD_DIV_INV>(A, log2n, grd, smem);
}

// ---- TMA-pipelined y / z passes (fft_tma.cuh) --------------------------------------------------------------
constexpr int TMA_TK = 8;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <class FT>
static bool make_spec_map(FastPoisson<FT>* p, bool along_y, CUtensorMap* out) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return false;
        fn = (EncodeFn)q;
    }
    const int Ny = p->N[1], Nz = p->N[2];
    cuuint64_t dims[3] = {(cuuint64_t)2 * p->NXP, (cuuint64_t)Ny, (cuuint64_t)Nz};
    cuuint64_t strides[2] = {(cuuint64_t)2 * p->NXP * sizeof(FT), (cuuint64_t)2 * p->NXP * Ny * sizeof(FT)};
    cuuint32_t box[3] = {2 * TMA_TK, (cuuint32_t)(along_y ? std::min(Ny, 256) : 1), (cuuint32_t)(along_y ? 1 : std::min(Nz, 256))};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(out, sizeof(FT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)p->spec,
                    dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
// tensor map over `base` viewed as reals: rank-nd dims / strides (strides in BYTES for dims 1..), box
template <class FT>
static bool encode_map(CUtensorMap* out, void* base, int nd, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return false;
        fn = (EncodeFn)q;
    }
    cuuint32_t es[4] = {1, 1, 1, 1};
    return fn(out, sizeof(FT) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, nd, base, dims,
              strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// maps of the slab-decomposed solve: chunk layout [R][Nz][NyL][KXB] in bufA / bufB of every rank (peer mapped)
template <class FT>
static void setup_dist_tma(FastPoisson<FT>* p) {
    using CT = typename Cx<FT>::T;
    p->dtma_ok = false;
    if (getenv("OB200_NO_FFT_TMA") != nullptr || getenv("OB200_NO_FFT_DTMA") != nullptr) return;
    if (p->R < 2 || !p->p2p || !p->has_z) return;
    const int NyL = p->N[1], Nz = p->N[2], KXB = p->KXB, R = p->R;
    if (NyL > 256 || Nz > 256 || p->log2[2] < 4 || p->log2[2] > 8 || p->log2[1] < 4 || p->log2[1] > 11) return;
    p->dtk_y = p->log2[1] <= 9 ? 8 : (p->log2[1] == 10 ? 4 : 2);
    const size_t W = sizeof(FT);
    const long long chunk = (long long)KXB * NyL * Nz;
    if (!make_spec_map(p, false, &p->tm_z)) return;                       // natural half spectrum, z lines
    cuuint64_t d3[3] = {(cuuint64_t)2 * KXB, (cuuint64_t)NyL, (cuuint64_t)Nz};
    cuuint64_t sc[2] = {(cuuint64_t)2 * KXB * W, (cuuint64_t)2 * KXB * NyL * W};                  // chunk strides
    cuuint64_t sn[2] = {(cuuint64_t)2 * p->NXP * W, (cuuint64_t)2 * p->NXP * NyL * W};           // natural strides
    cuuint32_t bz[3] = {2 * TMA_TK, 1, (cuuint32_t)Nz}, by[3] = {(cuuint32_t)(2 * p->dtk_y), (cuuint32_t)NyL, 1};
    // default: peer stores for 2 ranks (half of the data stays local and the rest overlaps the transform), bulk
    // copies beyond (measured: the 128-byte rows of the fused stores reach ~440 GB/s of NVLink, contiguous copies more)
    const char* eb = getenv("OB200_FFT_BULK_A2A");
    p->bulk_a2a = eb ? atoi(eb) != 0 : true;
    if (p->bulk_a2a && !p->bufC) {
        OB_CUDA(cudaMalloc(&p->bufC, (size_t)R * chunk * sizeof(CT)));
        p->owned.push_back(p->bufC);
    }
    for (int r = 0; r < R; ++r) {
        // own chunk: always straight into its final place; other ranks: peer buffer, or local staging (bufA / bufC)
        CT* zf = (!p->bulk_a2a || r == p->rank) ? p->peerB[r] + (long long)p->rank * chunk : p->bufA + (long long)r * chunk;
        CT* yy = (!p->bulk_a2a || r == p->rank) ? p->peerA[r] + (long long)p->rank * chunk : p->bufC + (long long)r * chunk;
        if (!encode_map<FT>(&p->tmr_zf[r], zf, 3, d3, sc, bz)) return;
        if (!encode_map<FT>(&p->tmr_zi[r], p->spec + (long long)r * KXB, 3, d3, sn, bz)) return;
        if (!encode_map<FT>(&p->tmr_y[r], yy, 3, d3, sc, by)) return;
    }
    cuuint64_t d4[4] = {(cuuint64_t)2 * KXB, (cuuint64_t)NyL, (cuuint64_t)Nz, (cuuint64_t)R};
    cuuint64_t s4[3] = {sc[0], sc[1], (cuuint64_t)2 * chunk * W};
    cuuint32_t b4z[4] = {2 * TMA_TK, 1, (cuuint32_t)Nz, 1}, b4y[4] = {(cuuint32_t)(2 * p->dtk_y), (cuuint32_t)NyL, 1, (cuuint32_t)R};
    if (!encode_map<FT>(&p->tm4_zi, p->bufA, 4, d4, s4, b4z)) return;
    if (!encode_map<FT>(&p->tm4_y, p->bufB, 4, d4, s4, b4y)) return;
    // split y lines (ADDR_YS), OB200_FFT_YSPLIT=1: measured at 8 ranks (2048-point lines) the split transform needs 0.20 ms of
    // kernels per solve against ~0.3 ms for the gathered lines, but the phase is bound by the NVLink all-to-all behind it
    // (125 MB per rank and transpose, 0.23 ms at the ~550 GB/s the 8-rank exchange reaches) and the twelve small launches
    // of the pipelined form cost what they save: 1.29 (split) vs 1.26 ms per step (gathered), so it stays off by default
    const char* ys = getenv("OB200_FFT_YSPLIT");
    p->ysplit = p->bulk_a2a && p->log2[1] - ilog2c(R) >= 4 && (ys ? atoi(ys) != 0 : false);
    if (p->ysplit) {
        cuuint32_t b4s[4] = {2 * TMA_TK, (cuuint32_t)NyL, 1, 1}, b3s[3] = {2 * TMA_TK, (cuuint32_t)NyL, 1};
        if (!encode_map<FT>(&p->tm4_ys, p->bufB, 4, d4, s4, b4s)) return;
        for (int r = 0; r < R; ++r)
            if (!encode_map<FT>(&p->tmr_ys[r], p->bufB + (long long)r * chunk, 3, d3, sc, b3s)) return;
        const int l2 = ilog2c(NyL);
        std::vector<double> t((size_t)R * NyL);
        const double PI = 3.14159265358979323846;
        const double dy = (double)p->Ly_global / p->NyG;
        for (int r = 0; r < R; ++r)
            for (int P = 0; P < NyL; ++P) {
                const int k = r + R * freq_of_pos(l2, P);
                const double v = 2 * sin(k * PI / p->NyG) / dy;
                t[(size_t)r * NyL + P] = v * v;
            }
        p->lamy2 = up(t, p->owned);
        std::vector<CT> twl(NyL);
        const long double PIl = 3.14159265358979323846264338327950288L;
        for (int k = 0; k < NyL; ++k) {
            const long double a = -2.0L * PIl * k / NyL;
            twl[k].x = (FT)cosl(a); twl[k].y = (FT)sinl(a);
        }
        p->twYL = up(twl, p->owned);
    }
    p->dtma_ok = true;
}
template <class FT>
static void setup_tma(FastPoisson<FT>* p) {
    p->tma_ok = false;
    if (p->R > 1) { setup_dist_tma(p); return; }
    if (getenv("OB200_NO_FFT_TMA") != nullptr || p->NXP % TMA_TK) return;
    const bool zfft = p->has_z && !p->tri;
    if (p->log2[1] < 4 || p->log2[1] > 9 || (zfft && (p->log2[2] < 4 || p->log2[2] > 9))) return;
    if (!make_spec_map(p, true, &p->tm_y)) return;
    if (zfft && !make_spec_map(p, false, &p->tm_z)) return;
    p->tma_ok = true;
}

template <class FT, int MODE, int STAGES, int TK = TMA_TK>
static void launch_line_tma(const tl::TArgs<FT>& A, int log2n) {
    using CT = typename Cx<FT>::T;
    const int n = 1 << log2n;
    const size_t smem = (size_t)STAGES * n * TK * sizeof(CT) + (size_t)n * sizeof(CT);
    const int threads = std::min(256, std::max(64, TK * n / 16));
    auto go = [&](auto kern) {
        // all instantiations share one function-pointer type: key the per-kernel set-up on the pointer
        static std::map<const void*, int> occ;
        int& blocks_per_sm = occ[(const void*)kern];
        if (!blocks_per_sm) {
            OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
            blocks_per_sm = std::max(1, blocks_per_sm);
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = std::min(A.nkx * A.nOther, sms * blocks_per_sm);
        kern<<<grid, threads, smem, stream()>>>(A);
        OB_LAUNCH_CHECK();
    };
    if constexpr (TK == TMA_TK) {
        switch (log2n) {
            case 4: go(tl::line_tma_kernel<FT, 4, MODE, TK, STAGES>); break;
            case 5: go(tl::line_tma_kernel<FT, 5, MODE, TK, STAGES>); break;
            case 6: go(tl::line_tma_kernel<FT, 6, MODE, TK, STAGES>); break;
            case 7: go(tl::line_tma_kernel<FT, 7, MODE, TK, STAGES>); break;
            case 8: go(tl::line_tma_kernel<FT, 8, MODE, TK, STAGES>); break;
            default:       // 512-point lines: 64 KB tiles, two stages
                if constexpr (STAGES == 2) go(tl::line_tma_kernel<FT, 9, MODE, TK, STAGES>);
                else throw Error("line_tma: unsupported length");
                break;
        }
    } else if constexpr (TK == 4) {
        go(tl::line_tma_kernel<FT, 10, MODE, TK, STAGES>);
    } else {
        go(tl::line_tma_kernel<FT, 11, MODE, TK, STAGES>);
    }
}
template <class FT>
static void run_line_tma(FastPoisson<FT>* p, int dim, int mode) {
    tl::TArgs<FT> A;
    A.tm = dim == 1 ? p->tm_y : p->tm_z;
    A.addr = tl::ADDR_NAT; A.R = 1; A.KXB = p->NXP; A.tpc = p->NXP / TMA_TK; A.NyL = p->N[1]; A.kx_base = 0; A.NXP = p->NXP;
    A.Nz = p->N[2];
    A.r_only = -1; A.o_first = 0;
    A.line_is_y = dim == 1;
    A.nkx = p->NXP / TMA_TK;
    A.nOther = dim == 1 ? p->N[2] : p->N[1];
    A.tw = dim == 1 ? p->twY : p->twZ;
    A.scale = (FT)(1.0 / p->N[dim]);
    A.lamx = p->lamx;
    A.lamL = dim == 1 ? p->lamy : p->lamz;
    A.lamO = dim == 1 ? p->lamz : p->lamy;
    static const int stages = env_int("OB200_FFT_STAGES", 3);
    const int l = p->log2[dim];
#define GO(M)  { if (stages == 2 || l >= 9) launch_line_tma<FT, M, 2>(A, l); else launch_line_tma<FT, M, 3>(A, l); }
    if (mode == LM_FWD) GO(LM_FWD) else if (mode == LM_INV) GO(LM_INV) else GO(LM_FWD_DIV_INV)
#undef GO
}

// single-GPU passes on the natural [Nz][Ny][NXP] layout, in place
template <class FT>
static void run_line(FastPoisson<FT>* p, int dim, int mode) {
    if (p->tma_ok) { run_line_tma(p, dim, mode); return; }
    LArgs<FT> A;
    int Ny = p->N[1], Nz = p->N[2];
    A.in = p->spec; A.out = p->spec;
    A.NXH = p->NXH; A.kx0 = 0;
    A.n = p->N[dim];
    A.line_is_y = dim == 1;
    long long s_m = dim == 1 ? p->NXP : (long long)p->NXP * Ny;
    long long s_o = dim == 1 ? (long long)p->NXP * Ny : p->NXP;
    A.lin = A.lout = natural_lay(p->NXP, s_m, s_o);
    A.nOther = dim == 1 ? Nz : Ny;
    A.tw = dim == 1 ? p->twY : p->twZ;
    A.lamL = dim == 1 ? p->lamy : p->lamz;
    A.lamO = dim == 1 ? p->lamz : p->lamy;
    launch_line_any(p, A, p->log2[dim], mode);
}

// slab-decomposed solve (y split over R ranks): x and z transforms are local; the y transform needs
// the lines gathered, i.e. one all-to-all that turns y-slabs into kx-slabs and one that turns them back
// (reference: Distributed/distributed_fft_based_poisson_solver.jl:50-93,146-196 via PencilFFTs).
// The z passes write / read the all-to-all chunk layout directly, so no pack or unpack kernel exists.
template <class FT>
static void all_to_all(FastPoisson<FT>* p, const typename Cx<FT>::T* src, typename Cx<FT>::T* dst) {
    using CT = typename Cx<FT>::T;
    size_t chunk = (size_t)p->KXB * p->N[1] * p->N[2];
    cm::group_start();
    for (int r = 0; r < p->R; ++r) {
        cm::send(src + r * chunk, chunk * sizeof(CT), r);
        cm::recv(dst + r * chunk, chunk * sizeof(CT), r);
    }
    cm::group_end();
}

// Split y transform, the part ACROSS the chunks.  With y = s NyL + yl (s = source rank) and k = k1 + R k2,
//   X[k1 + R k2] = sum_yl W_NyL^(yl k2) [ W_NyG^(yl k1) sum_s W_R^(s k1) x[s NyL + yl] ],
// so the NyG-point line is an R-point butterfly over the R chunk slabs (pointwise in (z, yl, kx)), a twiddle, and
// NyL-point lines inside slab k1 (line_tma_kernel, ADDR_YS).  INV is the exact inverse (conjugate twiddle, conjugate
// butterfly; the 1 / NyG is applied by the line kernel).  outs[s] lets the backward pass write the own rank's chunk
// straight into its final place and the others into the staging buffer of the bulk transposes.
template <class FT> struct YSplitPtrs { typename Cx<FT>::T* p[8]; };
template <class FT, int RR, bool INV>
__global__ void __launch_bounds__(256) ysplit_kernel(const typename Cx<FT>::T* __restrict__ in, long long chunk, YSplitPtrs<FT> outs,
                                                     int NyL, int KXB, int z0, int nz, const typename Cx<FT>::T* __restrict__ tw) {
    using CT = typename Cx<FT>::T;
    constexpr int LB = ilog2c(RR);
    const long long per = (long long)nz * NyL * KXB;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < per; e += (long long)gridDim.x * blockDim.x) {
        const long long off = (long long)z0 * NyL * KXB + e;
        const int yl = (int)((off / KXB) % NyL);
        CT x[RR];
        if (!INV) {
#pragma unroll
            for (int s = 0; s < RR; ++s) x[s] = in[s * chunk + off];
            dft_reg<RR, false>(x);
#pragma unroll
            for (int k1 = 0; k1 < RR; ++k1) {
                CT v = x[brev(k1, LB)];
                if (k1 > 0) v = cmul(v, tw[yl * k1]);
                outs.p[k1][off] = v;
            }
        } else {
#pragma unroll
            for (int k1 = 0; k1 < RR; ++k1) {
                CT v = in[k1 * chunk + off];
                if (k1 > 0) v = cmulc(v, tw[yl * k1]);
                x[k1] = v;
            }
            dft_reg<RR, true>(x);
#pragma unroll
            for (int s = 0; s < RR; ++s) outs.p[s][off] = x[brev(s, LB)];
        }
    }
}
template <class FT, bool INV>
static void launch_ysplit(FastPoisson<FT>* p, const typename Cx<FT>::T* in, const YSplitPtrs<FT>& outs, int z0, int nz) {
    const long long chunk = (long long)p->KXB * p->N[1] * p->N[2];
    const long long per = (long long)nz * p->N[1] * p->KXB;
    const int blocks = (int)std::min<long long>(148 * 8, (per + 255) / 256);
#define YS(RV) ysplit_kernel<FT, RV, INV><<<blocks, 256, 0, stream()>>>(in, chunk, outs, p->N[1], p->KXB, z0, nz, p->twY)
    if (p->R == 2) YS(2); else if (p->R == 4) YS(4); else YS(8);
#undef YS
    OB_LAUNCH_CHECK();
}

// the slab-decomposed middle of the solve through the TMA-pipelined kernel: z forward (stores into the peers' bufB),
// gathered y lines forward / divide / backward (stores back into the peers' bufA), z backward
template <class FT>
static void distributed_middle_tma(FastPoisson<FT>* p) {
    const int R = p->R, KXB = p->KXB, NyL = p->N[1], Nz = p->N[2];
    tl::TArgs<FT> A;
    A.R = R; A.KXB = KXB; A.NyL = NyL; A.NXP = p->NXP; A.kx_base = 0; A.Nz = Nz;
    A.lamx = p->lamx;
    // z forward
    A.addr = tl::ADDR_ZF; A.tm = p->tm_z;
    for (int r = 0; r < R; ++r) A.tmr[r] = p->tmr_zf[r];
    A.tpc = cdiv(KXB, TMA_TK); A.nkx = R * A.tpc; A.nOther = NyL; A.line_is_y = 0;
    A.tw = p->twZ; A.scale = (FT)(1.0 / Nz); A.lamL = p->lamz; A.lamO = nullptr;
    using CTt = typename Cx<FT>::T;
    const size_t chunk_bytes = (size_t)KXB * NyL * Nz * sizeof(CTt);
    const long long chunk_el = (long long)KXB * NyL * Nz;
    // bulk transposes are pipelined with the transform: the launch is split (z forward: one launch per destination
    // chunk; y lines: one launch per block of z levels, whose slice of every chunk is contiguous) and each finished
    // piece is copied on a second stream while the next piece is transformed
    cudaStream_t& cpy = p->cpy;                       // owned by the plan (released in fast_poisson_destroy)
    cudaEvent_t* ev_piece = p->ev_piece;
    cudaEvent_t& ev_done = p->ev_done;
    if (p->bulk_a2a && !cpy) {
        OB_CUDA(cudaStreamCreateWithFlags(&cpy, cudaStreamNonBlocking));
        for (int q = 1; q < R; ++q) {
            OB_CUDA(cudaStreamCreateWithFlags(&p->cpys[q], cudaStreamNonBlocking));
            OB_CUDA(cudaEventCreateWithFlags(&p->ev_dones[q], cudaEventDisableTiming));
        }
        for (int q = 0; q < 16; ++q) OB_CUDA(cudaEventCreateWithFlags(&ev_piece[q], cudaEventDisableTiming));
        OB_CUDA(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    }
    A.r_only = -1; A.o_first = 0;
    {
        PhaseScope ph("fft_z_fwd");
        if (!p->bulk_a2a) {
            launch_line_tma<FT, LM_FWD, 3>(A, p->log2[2]);
        } else {
            A.nkx = A.tpc;
            for (int q = 1; q <= R; ++q) {          // the other ranks' chunks first, the own chunk (no copy) last
                const int r = (p->rank + q) % R;
                A.r_only = r;
                launch_line_tma<FT, LM_FWD, 3>(A, p->log2[2]);
                if (r == p->rank) continue;
                OB_CUDA(cudaEventRecord(ev_piece[q], stream()));
                OB_CUDA(cudaStreamWaitEvent(p->cpys[q], ev_piece[q], 0));
                OB_CUDA(cudaMemcpyAsync(p->peerB[r] + (long long)p->rank * chunk_el, p->bufA + (long long)r * chunk_el,
                                        chunk_bytes, cudaMemcpyDefault, p->cpys[q]));
            }
            A.r_only = -1;
            for (int q = 1; q < R; ++q) {
                OB_CUDA(cudaEventRecord(p->ev_dones[q], p->cpys[q]));
                OB_CUDA(cudaStreamWaitEvent(stream(), p->ev_dones[q], 0));
            }
        }
    }
    { PhaseScope ph("fft_sync"); cm::fast_barrier(); }
    // gathered y lines: forward, eigenvalue divide, backward
    A.addr = tl::ADDR_Y; A.tm4 = p->tm4_y;
    for (int r = 0; r < R; ++r) A.tmr[r] = p->tmr_y[r];
    A.tpc = cdiv(KXB, p->dtk_y); A.nkx = A.tpc; A.nOther = Nz; A.line_is_y = 1; A.kx_base = p->rank * KXB;
    A.tw = p->twY; A.scale = (FT)(1.0 / p->NyG); A.lamL = p->lamy; A.lamO = p->lamz;
    {
        PhaseScope ph("fft_y");
        const int l = p->log2[1];
        auto launch_y = [&]() {
            if (l <= 8) launch_line_tma<FT, LM_FWD_DIV_INV, 3>(A, l);
            else if (l == 9) launch_line_tma<FT, LM_FWD_DIV_INV, 2>(A, l);
            else if (l == 10) launch_line_tma<FT, LM_FWD_DIV_INV, 2, 4>(A, l);
            else launch_line_tma<FT, LM_FWD_DIV_INV, 2, 2>(A, l);
        };
        if (p->ysplit) {
            // butterfly across the chunks, NyL-point lines (forward, divide, backward) inside each chunk, inverse butterfly
            // per block of z levels with the bulk copies of the finished block behind it
            YSplitPtrs<FT> inplace, back;
            for (int r = 0; r < R; ++r) {
                inplace.p[r] = p->bufB + (long long)r * chunk_el;
                back.p[r] = r == p->rank ? p->bufA + (long long)p->rank * chunk_el : p->bufC + (long long)r * chunk_el;
            }
            tl::TArgs<FT> S = A;
            S.addr = tl::ADDR_YS; S.tm4 = p->tm4_ys; S.Nz = Nz;
            for (int r = 0; r < R; ++r) S.tmr[r] = p->tmr_ys[r];
            S.tpc = cdiv(KXB, TMA_TK); S.nkx = S.tpc; S.r_only = -1;
            S.lamL = p->lamy2; S.tw = p->twYL;
            // all three passes per block of z levels: the bulk copies of block b run while block b+1 is transformed
            static const int nb_env = env_int("OB200_FFT_YBLOCKS", 4);
            const int NB = Nz >= 64 ? std::max(1, std::min(nb_env, 16)) : 1;
            const int zb = cdiv(Nz, NB);
            for (int b = 0; b < NB; ++b) {
                const int z0 = b * zb, nz = std::min(zb, Nz - z0);
                if (nz <= 0) break;
                { PhaseScope pa("fft_y_butterfly"); launch_ysplit<FT, false>(p, p->bufB, inplace, z0, nz); }
                S.o_first = z0 * R; S.nOther = nz * R;
                { PhaseScope pb("fft_y_lines"); launch_line_tma<FT, LM_FWD_DIV_INV, 3>(S, p->log2[1] - ilog2c(R)); }
                { PhaseScope pc("fft_y_butterfly"); launch_ysplit<FT, true>(p, p->bufB, back, z0, nz); }
                OB_CUDA(cudaEventRecord(ev_piece[b], stream()));
                for (int q = 1; q < R; ++q) OB_CUDA(cudaStreamWaitEvent(p->cpys[q], ev_piece[b], 0));
                const long long off = (long long)z0 * NyL * KXB;
                const size_t bytes = (size_t)nz * NyL * KXB * sizeof(CTt);
                for (int q = 1; q < R; ++q) {
                    const int r = (p->rank + q) % R;
                    OB_CUDA(cudaMemcpyAsync(p->peerA[r] + (long long)p->rank * chunk_el + off, p->bufC + (long long)r * chunk_el + off,
                                            bytes, cudaMemcpyDefault, p->cpys[q]));
                }
            }
            {
                PhaseScope pw("fft_y_copywait");
                for (int q = 1; q < R; ++q) {
                    OB_CUDA(cudaEventRecord(p->ev_dones[q], p->cpys[q]));
                    OB_CUDA(cudaStreamWaitEvent(stream(), p->ev_dones[q], 0));
                }
            }
        } else if (!p->bulk_a2a) {
            launch_y();
        } else {
            const int NB = Nz >= 64 ? 4 : 1;                 // blocks of z levels
            const int zb = cdiv(Nz, NB);
            for (int b = 0; b < NB; ++b) {
                const int z0 = b * zb, nz = std::min(zb, Nz - z0);
                if (nz <= 0) break;
                A.o_first = z0; A.nOther = nz;
                launch_y();
                OB_CUDA(cudaEventRecord(ev_piece[b], stream()));
                for (int q = 1; q < R; ++q) OB_CUDA(cudaStreamWaitEvent(p->cpys[q], ev_piece[b], 0));
                const long long off = (long long)z0 * NyL * KXB;       // chunk layout [z][yl][kx]: a z block is contiguous
                const size_t bytes = (size_t)nz * NyL * KXB * sizeof(CTt);
                for (int q = 1; q < R; ++q) {
                    const int r = (p->rank + q) % R;
                    OB_CUDA(cudaMemcpyAsync(p->peerA[r] + (long long)p->rank * chunk_el + off, p->bufC + (long long)r * chunk_el + off,
                                            bytes, cudaMemcpyDefault, p->cpys[q]));
                }
            }
            A.o_first = 0; A.nOther = Nz;
            for (int q = 1; q < R; ++q) {
                OB_CUDA(cudaEventRecord(p->ev_dones[q], p->cpys[q]));
                OB_CUDA(cudaStreamWaitEvent(stream(), p->ev_dones[q], 0));
            }
        }
    }
    { PhaseScope ph("fft_sync"); cm::fast_barrier(); }
    // z backward
    A.addr = tl::ADDR_ZI; A.tm4 = p->tm4_zi; A.kx_base = 0; A.r_only = -1; A.o_first = 0;
    for (int r = 0; r < R; ++r) A.tmr[r] = p->tmr_zi[r];
    A.tpc = cdiv(KXB, TMA_TK); A.nkx = R * A.tpc; A.nOther = NyL; A.line_is_y = 0;
    A.tw = p->twZ; A.scale = (FT)(1.0 / Nz); A.lamL = p->lamz; A.lamO = nullptr;
    { PhaseScope ph("fft_z_inv"); launch_line_tma<FT, LM_INV, 3>(A, p->log2[2]); }
}

template <class FT>
static void distributed_middle(FastPoisson<FT>* p) {
    int NyL = p->N[1], Nz = p->N[2], KXB = p->KXB;
    long long chunk = (long long)KXB * NyL * Nz;
    Lay nat = natural_lay(p->NXP, (long long)p->NXP * NyL, p->NXP);          // z lines on [Nz][NyL][NXP]
    Lay blk{};                                                                // [R][Nz][NyL][KXB]
    blk.inv_kxb = 1.0f / KXB; blk.inv_split = 0.f;
    blk.kxb = KXB; blk.split = 1 << 30; blk.s_blk = chunk; blk.s_ml = (long long)NyL * KXB; blk.s_mh = 0; blk.s_o = KXB;
    // y lines gathered from the R source ranks: m = s * NyL + yl, other = z, kx local to this rank
    Lay gy{};
    gy.inv_kxb = 0.f; gy.inv_split = 1.0f / NyL;
    gy.kxb = 1 << 30; gy.split = NyL; gy.s_blk = 0; gy.s_ml = KXB; gy.s_mh = chunk; gy.s_o = (long long)NyL * KXB;
    LArgs<FT> A;
    // forward z: natural -> chunk layout.  With peer memory the chunk for rank r is stored directly into rank r's
    // bufB (slot = this rank), so the all-to-all IS the store phase of this kernel; otherwise NCCL moves bufA -> bufB.
    A.in = p->spec; A.out = p->bufA; A.lin = nat; A.lout = blk;
    if (p->p2p) {
        A.lout.pmode = 1;
        for (int r = 0; r < p->R; ++r) A.lout.ptr[r] = p->peerB[r] + (long long)p->rank * chunk;
    }
    A.NXH = p->NXP; A.kx0 = 0; A.n = Nz; A.line_is_y = 0; A.nOther = NyL;
    A.tw = p->twZ; A.lamL = p->lamz; A.lamO = nullptr;
    { PhaseScope ph("fft_z_fwd"); launch_line_any(p, A, p->log2[2], LM_FWD); }
    { PhaseScope ph("fft_sync"); if (p->p2p) cm::fast_barrier(); else all_to_all(p, p->bufA, p->bufB); }
    A.in = p->bufB; A.out = p->bufB; A.lin = A.lout = gy;
    if (p->p2p) {      // transposed back on the fly: the part of the line that came from rank s returns to rank s's bufA
        A.out = p->bufA;
        A.lout.pmode = 2;
        for (int r = 0; r < p->R; ++r) A.lout.ptr[r] = p->peerA[r] + (long long)p->rank * chunk;
    }
    A.NXH = KXB; A.kx0 = p->rank * KXB; A.n = p->NyG; A.line_is_y = 1; A.nOther = Nz;
    A.tw = p->twY; A.lamL = p->lamy; A.lamO = p->lamz;
    { PhaseScope ph("fft_y"); launch_line_any(p, A, p->log2[1], LM_FWD_DIV_INV); }
    { PhaseScope ph("fft_sync"); if (p->p2p) cm::fast_barrier(); else all_to_all(p, p->bufB, p->bufA); }
    // backward z: chunk layout -> natural
    A.in = p->bufA; A.out = p->spec; A.lin = blk; A.lout = nat;
    A.NXH = p->NXP; A.kx0 = 0; A.n = Nz; A.line_is_y = 0; A.nOther = NyL;
    A.tw = p->twZ; A.lamL = p->lamz; A.lamO = nullptr;
    { PhaseScope ph("fft_z_inv"); launch_line_any(p, A, p->log2[2], LM_INV); }
}

// ---- slab-decomposed Fourier-tridiagonal solve -------------------------------------------------------------------------------
// Thomas sweep on the gathered layout: this rank holds its KXB wavenumbers kx of ALL NyG y modes and all Nz levels, element
// (kx, m = s NyL + yl, z) at s * chunk + (z NyL + yl) KXB + kx.  Same arithmetic as thomas_half_kernel.
template <class FT>
__global__ void __launch_bounds__(128) thomas_gathered_kernel(typename Cx<FT>::T* spec, FT* tsc, int KXB, int kx0, int NyL, int NyG,
                                                               int Nz, long long chunk, const double* lamx, const double* lamy,
                                                               const double* dzF, const double* dzC) {
    using CT = typename Cx<FT>::T;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (kx >= KXB) return;
    const long long col = (long long)(m / NyL) * chunk + (long long)(m % NyL) * KXB + kx, pl = (long long)NyL * KXB;
    const double lam = lamx[kx0 + kx] + lamy[m];
    auto diag = [&](int k) -> double {     // k is 1-based
        if (k == 1) return -1 / dzF[2] - dzC[1] * lam;
        if (k == Nz) return -1 / dzF[Nz] - dzC[Nz] * lam;
        return -(1 / dzF[k + 1] + 1 / dzF[k]) - dzC[k] * lam;
    };
    const double eps10 = 10 * (sizeof(FT) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16);
    double beta = diag(1);
    CT f1 = spec[col], prev;
    prev.x = (FT)((double)f1.x / beta); prev.y = (FT)((double)f1.y / beta);
    spec[col] = prev;
    for (int k = 2; k <= Nz; ++k) {
        const double ak = 1 / dzF[k];
        const FT t = (FT)(ak / beta);
        tsc[col + (k - 1) * pl] = t;
        beta = diag(k) - ak * (double)t;
        if (!(fabs(beta) > eps10)) break;
        // the horizontal-mean column of a Neumann problem is singular: its last pivot is zero in exact arithmetic and rounding
        // noise otherwise, and noise / noise would add an arbitrary, input-bit-sensitive constant (it is removed with the mean
        // below, but costs eps * |constant| of accuracy and bit reproducibility).  The reference's `break` is the same pin
        // whenever the noise happens to be below its threshold; here the column is always pinned.
        if (k == Nz && lam == 0.0) break;
        const CT fk = spec[col + (k - 1) * pl];
        CT r;
        r.x = (FT)(((double)fk.x - ak * (double)prev.x) / beta);
        r.y = (FT)(((double)fk.y - ak * (double)prev.y) / beta);
        spec[col + (k - 1) * pl] = r;
        prev = r;
    }
    CT nxt = spec[col + (long long)(Nz - 1) * pl];
    for (int k = Nz - 1; k >= 1; --k) {
        const FT t = tsc[col + k * pl];
        CT v = spec[col + (k - 1) * pl];
        v.x -= t * nxt.x; v.y -= t * nxt.y;
        spec[col + (k - 1) * pl] = v;
        nxt = v;
    }
    if (kx0 + kx == 0 && m == 0) {         // the horizontal-mean column: phi .-= mean(phi)
        double sum = 0;
        for (int k = 0; k < Nz; ++k) sum += (double)spec[col + k * pl].x;
        const FT mean = (FT)(sum / Nz);
        for (int k = 0; k < Nz; ++k) { CT v = spec[col + k * pl]; v.x -= mean; v.y = 0; spec[col + k * pl] = v; }
    }
}

// x is transformed locally (run_x); then: the z lines move to the chunk layout of their destination ranks (no transform: z is
// solved, not transformed), all-to-all, forward y lines over the gathered y, Thomas in z, backward y lines stored transposed
// back, all-to-all, chunk layout -> natural.  fourier_tridiagonal_poisson_solver.jl:74-101 on a y-slab decomposition.
template <class FT>
static void distributed_middle_tri(FastPoisson<FT>* p) {
    int NyL = p->N[1], Nz = p->N[2], KXB = p->KXB;
    long long chunk = (long long)KXB * NyL * Nz;
    Lay nat = natural_lay(p->NXP, (long long)p->NXP * NyL, p->NXP);
    Lay blk{};
    blk.inv_kxb = 1.0f / KXB; blk.inv_split = 0.f;
    blk.kxb = KXB; blk.split = 1 << 30; blk.s_blk = chunk; blk.s_ml = (long long)NyL * KXB; blk.s_mh = 0; blk.s_o = KXB;
    Lay gy{};
    gy.inv_kxb = 0.f; gy.inv_split = 1.0f / NyL;
    gy.kxb = 1 << 30; gy.split = NyL; gy.s_blk = 0; gy.s_ml = KXB; gy.s_mh = chunk; gy.s_o = (long long)NyL * KXB;
    LArgs<FT> A;
    A.in = p->spec; A.out = p->bufA; A.lin = nat; A.lout = blk;
    if (p->p2p) {
        A.lout.pmode = 1;
        for (int r = 0; r < p->R; ++r) A.lout.ptr[r] = p->peerB[r] + (long long)p->rank * chunk;
    }
    A.NXH = p->NXP; A.kx0 = 0; A.n = Nz; A.line_is_y = 0; A.nOther = NyL;
    A.tw = p->twN; A.lamL = nullptr; A.lamO = nullptr;
    { PhaseScope ph("fft_z_fwd"); launch_line_any(p, A, p->log2[2], LM_COPY); }
    { PhaseScope ph("fft_sync"); if (p->p2p) cm::fast_barrier(); else all_to_all(p, p->bufA, p->bufB); }
    A.in = p->bufB; A.out = p->bufB; A.lin = A.lout = gy;
    A.NXH = KXB; A.kx0 = p->rank * KXB; A.n = p->NyG; A.line_is_y = 1; A.nOther = Nz;
    A.tw = p->twY; A.lamL = p->lamy; A.lamO = nullptr;
    { PhaseScope ph("fft_y"); launch_line_any(p, A, p->log2[1], LM_FWD); }
    {
        PhaseScope ph("fft_z");
        dim3 blk3(64), grd3(cdiv(KXB, 64), p->NyG);
        thomas_gathered_kernel<FT><<<grd3, blk3, 0, stream()>>>(p->bufB, p->tsc, KXB, p->rank * KXB, NyL, p->NyG, Nz, chunk, p->lamx,
                                                                p->lamy, p->dzF, p->dzC);
        OB_LAUNCH_CHECK();
    }
    if (p->p2p) {
        A.out = p->bufA;
        A.lout.pmode = 2;
        for (int r = 0; r < p->R; ++r) A.lout.ptr[r] = p->peerA[r] + (long long)p->rank * chunk;
    }
    { PhaseScope ph("fft_y"); launch_line_any(p, A, p->log2[1], LM_INV); }
    { PhaseScope ph("fft_sync"); if (p->p2p) cm::fast_barrier(); else all_to_all(p, p->bufB, p->bufA); }
    A.in = p->bufA; A.out = p->spec; A.lin = blk; A.lout = nat;
    A.NXH = p->NXP; A.kx0 = 0; A.n = Nz; A.line_is_y = 0; A.nOther = NyL;
    A.tw = p->twN; A.lamL = nullptr; A.lamO = nullptr;
    { PhaseScope ph("fft_z_inv"); launch_line_any(p, A, p->log2[2], LM_COPY); }
}

// persistent bulk-copy-pipelined x passes (fft_xtma.cuh); false if the configuration is not covered
template <class FT, bool FWD>
static bool run_x_tma(FastPoisson<FT>* p, XArgs<FT>& A) {
    using CT = typename Cx<FT>::T;
    static const bool off = getenv("OB200_NO_FFT_XTMA") != nullptr;
    const int lm = p->log2[0] - 1;
    if (off || !p->tma_ok || lm < 4 || lm > 9) return false;
    // measured at 256^3 (profiles/r1_summary.md): the staged backward pass is faster than the direct-load one
    // (84 vs 93 us), the staged forward pass is not (194 vs 156 us: one 160 KB block per SM leaves 8 warps for the
    // divergence + transform + untangle chain), so the forward pass keeps the direct-load kernel by default
    static const bool fwd_on = env_int("OB200_FFT_XTMA_FWD", 0) != 0;
    if (FWD && (!fwd_on || A.real_in || A.tri)) return false;
    static const int T0 = env_int("OB200_FFT_TXT", 8);
    int T = T0;
    while (T > 1 && (A.Ny % T)) T >>= 1;
    if (T < 2) return false;
    if (((size_t)A.st[1] * sizeof(FT)) % 16 || ((size_t)T * A.NXP * sizeof(CT)) % 16) return false;
    A.T = T;
    const int M = 1 << lm;
    const int RL = lm == 4 ? 16 : (lm == 5 ? 4 : (lm == 8 ? 16 : 8));
    const size_t work = ((size_t)T * (M + M / RL + 1) + M) * sizeof(CT);
    static const int stages_f = env_int("OB200_FFT_XSTAGES_F", 2), stages_b = env_int("OB200_FFT_XSTAGES_B", 3);
    const int stages = FWD ? stages_f : stages_b;
    const int nrows = A.has_z ? 4 * T + 1 : 2 * T + 1;
    const size_t stage = FWD ? ((size_t)nrows * A.st[1] * sizeof(FT) + 127) / 128 * 128
                             : ((size_t)T * A.NXP * sizeof(CT) + 127) / 128 * 128;
    const size_t smem = stages * stage + work;
    if (smem > 200 * 1024) return false;
    static const int threads = env_int("OB200_FFT_XTHREADS", 256);
    auto go = [&](auto kern) {
        static std::map<const void*, int> occ;
        int& bps = occ[(const void*)kern];
        if (!bps) {
            OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            OB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, threads, smem));
            bps = std::max(1, bps);
        }
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int grid = std::min((A.Ny / T) * A.Nz, sms * bps);
        kern<<<grid, threads, smem, stream()>>>(A);
        OB_LAUNCH_CHECK();
    };
#define XT(L)                                                                                        \
    case L:                                                                                          \
        if (FWD) { if (stages == 2) go(tx::x_r2c_tma_kernel<FT, L, 2>); else go(tx::x_r2c_tma_kernel<FT, L, 3>); } \
        else { if (stages == 2) go(tx::x_c2r_tma_kernel<FT, L, 2>); else go(tx::x_c2r_tma_kernel<FT, L, 3>); }     \
        break;
    switch (lm) { XT(4) XT(5) XT(6) XT(7) XT(8) default: XT(9) }
#undef XT
    return true;
}

template <class FT, bool FWD>
static void run_x(FastPoisson<FT>* p, XArgs<FT>& A) {
    using CT = typename Cx<FT>::T;
    int lm = p->log2[0] - 1;
    A.spec = p->spec; A.Nx = p->N[0]; A.Ny = p->N[1]; A.Nz = p->N[2]; A.NXP = p->NXP;
    A.twM = p->twM; A.twN = p->twN; A.kpos = p->kpos;
    A.scale = (FT)(1.0 / (p->N[0] / 2));
    if (run_x_tma<FT, FWD>(p, A)) return;
    static const int T0 = env_int("OB200_FFT_TX", 8);
    int T = T0;
    while (T > 1 && line_smem(lm, T, sizeof(CT)) > 100 * 1024) T >>= 1;
    A.T = T;
    dim3 grd(cdiv(A.Ny, T), A.Nz);
    size_t smem = line_smem(lm, T, sizeof(CT));
    int threads = std::min(256, std::max(64, T * (1 << lm) / 16));
    auto go = [&](auto kern) {
        OB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        kern<<<grd, threads, smem, stream()>>>(A);
        OB_LAUNCH_CHECK();
    };
#define XCASE(L)                                                             \
    case L: if (FWD) go(x_r2c_kernel<FT, L>); else go(x_c2r_kernel<FT, L>); break;
    switch (lm) { XCASE(4) XCASE(5) XCASE(6) XCASE(7) XCASE(8) XCASE(9) default: if (FWD) go(x_r2c_kernel<FT, 10>); else go(x_c2r_kernel<FT, 10>); break; }
#undef XCASE
}

// Thomas sweep of the Fourier-tridiagonal solver on the half spectrum [Nz][Ny][NXP]: one (kx, y) column per thread,
// coalesced in kx; same arithmetic as thomas_kernel (fft.cu) -- batched_tridiagonal_solver.jl:91-122 with the main
// diagonal of fourier_tridiagonal_poisson_solver.jl:16-28 generated on the fly and the `|beta| <= 10 eps` break.
// The column (kx, ky) = (0, 0) carries the horizontal means: removing ITS vertical mean afterwards is the
// `phi .-= mean(phi)` of :93-99 done in spectral space.
template <class FT>
__global__ void __launch_bounds__(128) thomas_half_kernel(typename Cx<FT>::T* spec, FT* tsc, int NXH, int NXP, int Ny, int Nz,
                                                           const double* lamx, const double* lamy, const double* dzF,
                                                           const double* dzC) {
    using CT = typename Cx<FT>::T;
    const int kx = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (kx >= NXH) return;
    const long long col = kx + (long long)NXP * y, pl = (long long)NXP * Ny;
    const double lam = lamx[kx] + lamy[y];
    auto diag = [&](int k) -> double {     // k is 1-based
        if (k == 1) return -1 / dzF[2] - dzC[1] * lam;
        if (k == Nz) return -1 / dzF[Nz] - dzC[Nz] * lam;
        return -(1 / dzF[k + 1] + 1 / dzF[k]) - dzC[k] * lam;
    };
    const double eps10 = 10 * (sizeof(FT) == 4 ? 1.1920928955078125e-07 : 2.220446049250313e-16);
    double beta = diag(1);
    CT f1 = spec[col], prev;
    prev.x = (FT)((double)f1.x / beta); prev.y = (FT)((double)f1.y / beta);
    spec[col] = prev;
    for (int k = 2; k <= Nz; ++k) {
        const double ak = 1 / dzF[k];            // lower = upper diagonal: 1/Δzᶠ_k
        const FT t = (FT)(ak / beta);
        tsc[col + (k - 1) * pl] = t;
        beta = diag(k) - ak * (double)t;
        if (!(fabs(beta) > eps10)) break;
        // the horizontal-mean column of a Neumann problem is singular: its last pivot is zero in exact arithmetic and rounding
        // noise otherwise, and noise / noise would add an arbitrary, input-bit-sensitive constant (it is removed with the mean
        // below, but costs eps * |constant| of accuracy and bit reproducibility).  The reference's `break` is the same pin
        // whenever the noise happens to be below its threshold; here the column is always pinned.
        if (k == Nz && lam == 0.0) break;
        const CT fk = spec[col + (k - 1) * pl];
        CT r;
        r.x = (FT)(((double)fk.x - ak * (double)prev.x) / beta);
        r.y = (FT)(((double)fk.y - ak * (double)prev.y) / beta);
        spec[col + (k - 1) * pl] = r;
        prev = r;
    }
    CT nxt = spec[col + (long long)(Nz - 1) * pl];
    for (int k = Nz - 1; k >= 1; --k) {
        const FT t = tsc[col + k * pl];
        CT v = spec[col + (k - 1) * pl];
        v.x -= t * nxt.x; v.y -= t * nxt.y;
        spec[col + (k - 1) * pl] = v;
        nxt = v;
    }
    if (kx == 0 && y == 0) {
        double sum = 0;
        for (int k = 0; k < Nz; ++k) sum += (double)spec[col + k * pl].x;
        const FT mean = (FT)(sum / Nz);
        for (int k = 0; k < Nz; ++k) { CT v = spec[col + k * pl]; v.x -= mean; v.y = 0; spec[col + k * pl] = v; }
    }
}
template <class FT>
static void run_thomas(FastPoisson<FT>* p) {
    dim3 blk(64), grd(cdiv(p->NXH, 64), p->N[1]);
    thomas_half_kernel<FT><<<grd, blk, 0, stream()>>>(p->spec, p->tsc, p->NXH, p->NXP, p->N[1], p->N[2], p->lamx, p->lamy,
                                                      p->dzF, p->dzC);
    OB_LAUNCH_CHECK();
}

// source term from the velocities (u, v, w Julia-(0,0,0) pointers) or from a real array; result in phi
template <class FT>
void fast_poisson_solve(FastPoisson<FT>* p, const GridD<FT>& g, const FT* u, const FT* v, const FT* w, FT dt,
                        const FT* real_in, FT* phi_p0) {
    XArgs<FT> A{};
    A.u = u; A.v = v; A.w = w; A.real_in = real_in;
    for (int d = 0; d < 3; ++d) {
        A.st[d] = g.st[d];
        // a Periodic (not slab-decomposed) dimension is read with wrap-around: the velocities' halos need not be valid
        A.wrap[d] = g.topo[d] == OB_PERIODIC ? (long long)g.N[d] * g.st[d] : 0;
    }
    A.x0 = 1 - g.O[0];
    A.tri = p->zhook ? 0 : p->tri; A.dzC = g.regular[2] ? nullptr : g.dC[2]; A.dx = g.d[0]; A.dy = g.d[1]; A.dz = g.d[2];
    // divᶜᶜᶜ (divergence_operators.jl:16-19): 1/V * (δx(Ax u) + δy(Ay v) + δz(Az w)), then / Δt
    A.ax = g.d[1] * g.d[2]; A.ay = g.d[0] * g.d[2]; A.az = g.d[0] * g.d[1];
    A.invV = 1 / ((g.d[0] * g.d[1]) * g.d[2]);
    A.dt = dt;
    A.has_z = p->has_z;
    A.phi_p0 = phi_p0; A.Hx = g.H[0];
    { PhaseScope ph("fft_x_fwd"); run_x<FT, true>(p, A); }
    if (p->R > 1 && p->tri) {
        distributed_middle_tri(p);
    } else if (p->R > 1) {
        if (p->dtma_ok) distributed_middle_tma(p); else distributed_middle(p);
    } else if (p->has_z) {
        // y: Periodic lines of this file, or the caller's DCT pass (Bounded y); z: the caller's DCT pass (regular Bounded z),
        // the Thomas sweep (Fourier-tridiagonal solve) or Periodic lines with the eigenvalue divide
        { PhaseScope ph("fft_y"); if (p->yhook) p->yhook(0); else run_line(p, 1, LM_FWD); }
        { PhaseScope ph("fft_z"); if (p->zhook) p->zhook(); else if (p->tri) run_thomas(p); else run_line(p, 2, LM_FWD_DIV_INV); }
        { PhaseScope ph("fft_y"); if (p->yhook) p->yhook(1); else run_line(p, 1, LM_INV); }
    } else {
        PhaseScope ph("fft_y");
        run_line(p, 1, LM_FWD_DIV_INV);
    }
    { PhaseScope ph("fft_x_inv"); run_x<FT, false>(p, A); }
}

#define INST(FT)                                                                                   \
    template bool fast_ft_supported<FT>(const GridD<FT>&);                                          \
    template void fast_poisson_set_tridiagonal<FT>(FastPoisson<FT>*, const double*, const double*); \
    template FastSpecInfo fast_poisson_spec_info<FT>(FastPoisson<FT>*);                             \
    template void fast_poisson_set_zhook<FT>(FastPoisson<FT>*, std::function<void()>);              \
    template void fast_poisson_set_yhook<FT>(FastPoisson<FT>*, std::function<void(int)>, const double*); \
    template bool fast_poisson_supported<FT>(const GridD<FT>&);                                     \
    template FastPoisson<FT>* fast_poisson_create<FT>(const GridD<FT>&);                            \
    template void fast_poisson_destroy<FT>(FastPoisson<FT>*);                                       \
    template void fast_poisson_solve<FT>(FastPoisson<FT>*, const GridD<FT>&, const FT*, const FT*, \
                                         const FT*, FT, const FT*, FT*);
INST(float)
INST(double)
}  // namespace ff
}  // namespace ob
