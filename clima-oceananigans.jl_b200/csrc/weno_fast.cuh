// weno_fast.cuh -- arithmetic shared by the specialised tendency kernels (tendency_fast.cu, tendency_tma.cu):
// one-sided WENO5 reconstruction on a uniform grid, organised for the FP64 pipe of sm_100a.
//
// Reference quantities (src/Advection/weno_fifth_order.jl): smoothness indicators :311-317, Z / JS weights
// :380-403, candidate polynomials :518-524, optimal weights :12-14, eps = 1e-6 and exponent 2 :18-19.
// What is restated here is the same rational function of the five window values, evaluated with fewer
// operations (differences of order 1e-16 relative; parity <= 1e-12 per step is tested):
//   * second differences t_k of the three sub-stencils are shared by the smoothness indicators
//     beta_k = 13/12 t_k^2 + 1/4 s_k^2 and by the candidate values: with m = (c + d)/2,
//     p0 = m - t0/6, p1 = m - t1/6, p2 = m - (3 t1 - 2 t2)/6, so
//     sum_k w_k p_k = m - (g0 t0 + g1 t1 + g2 (3 t1 - 2 t2)) / (6 sum_k g_k)   for any unnormalised weights g_k;
//   * the first-difference parts are s2 = t2 + 2 (x2 - b), s1 = b - d, s0 = t0 + 2 (x0 - d); on the left-biased side
//     x2 = x0 = c, on the right-biased side (window mirrored) x2 = a, x0 = e, which reproduces the reference's
//     non-mirrored right-biased indicators (:315-317);
//   * everything is homogeneous in beta, so 4 beta (and 4 eps) is used and the factor 1/4 disappears;
//   * Float64: the six divisions of the weights are folded into ONE reciprocal,
//     g_k = C_k (E_k + tau^2) prod_{j != k} E_j = C_k (E0 E1 E2 + tau^2 prod_{j != k} E_j),  E_k = (beta_k + eps)^2,
//     and the reciprocal is a MUFU seed (relative error 2^-23) plus one cubic (Halley) correction, 2^-69;
//   * constants live in __constant__ memory so that they are instruction operands (no UMOV pairs in the loop).
#pragma once
#include <cuda_runtime.h>

namespace ob {
namespace wf {

// 0: 13/3   1: 4 eps   2: -1/6   3: 7/12   4: -1/12   5: -1/3
__constant__ double kd[6] = {13.0 / 3.0, 4.0e-6, -1.0 / 6.0, 7.0 / 12.0, -1.0 / 12.0, -1.0 / 3.0};
__constant__ float kf[6] = {(float)(13.0 / 3.0), 4.0e-6f, (float)(-1.0 / 6.0), (float)(7.0 / 12.0), (float)(-1.0 / 12.0), (float)(-1.0 / 3.0)};
template <class FT> __device__ __forceinline__ FT K(int i) {
    if constexpr (sizeof(FT) == 8) return kd[i]; else return kf[i];
}

__device__ __forceinline__ double rcp_seed(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    return r;
}

// One-sided reconstruction at a face.  (a, b, c, d, e) is the five-point window ordered TOWARDS the face:
// psi[f-3 .. f+1] for the left-biased side, psi[f+2 .. f-2] for the right-biased side;
// (x2, x0) = (c, c) on the left-biased side and (a, e) on the right-biased side.
template <class FT, bool ZW>
__device__ __forceinline__ FT weno_face(FT a, FT b, FT c, FT d, FT e, FT x2, FT x0) {
    const FT t2 = fma(FT(-2), b, a) + c, t1 = fma(FT(-2), c, b) + d, t0 = fma(FT(-2), d, c) + e;
    const FT s2 = fma(FT(2), x2 - b, t2), s1 = b - d, s0 = fma(FT(2), x0 - d, t0);
    const FT k133 = K<FT>(0), eps4 = K<FT>(1);
    const FT B2 = fma(s2, s2, k133 * (t2 * t2));      // 4 beta_k
    const FT B1 = fma(s1, s1, k133 * (t1 * t1));
    const FT B0 = fma(s0, s0, k133 * (t0 * t0));
    const FT r2 = fma(FT(3), t1, FT(-2) * t2);
    const FT m = FT(0.5) * (c + d);
    if constexpr (sizeof(FT) == 8) {
        const FT D0 = B0 + eps4, D1 = B1 + eps4, D2 = B2 + eps4;
        const FT E0 = D0 * D0, E1 = D1 * D1, E2 = D2 * D2;
        const FT P12 = E1 * E2, P02 = E0 * E2, P01 = E0 * E1;
        FT g0, g1, g2;       // 10 x the unnormalised weights
        if (ZW) {
            const FT tau = B2 - B0, tt = tau * tau, PI = E0 * P12;
            g0 = FT(3) * fma(tt, P12, PI);
            g1 = FT(6) * fma(tt, P02, PI);
            g2 = fma(tt, P01, PI);
        } else {
            g0 = FT(3) * P12; g1 = FT(6) * P02; g2 = P01;
        }
        const FT den = (g0 + g1) + g2;
        const FT S = fma(g0, t0, fma(g1, t1, g2 * r2));
        const FT r0 = rcp_seed(den);
        const FT er = fma(-den, r0, FT(1));
        const FT q0 = S * r0;
        const FT q = fma(q0, fma(er, er, er), q0);     // S / den
        return fma(q, K<FT>(2), m);
    } else {
        // Float32: the product form above would overflow for large-amplitude fields, so the weights keep the
        // reference's quotient form, with fast (2 ulp) divisions -- the tolerance of this precision is 1e-5
        const FT D0 = B0 + eps4, D1 = B1 + eps4, D2 = B2 + eps4;
        FT a0, a1, a2;
        if (ZW) {
            const FT tau = B2 - B0;
            const FT q0 = __fdividef(tau, D0), q1 = __fdividef(tau, D1), q2 = __fdividef(tau, D2);
            a0 = FT(3) * fma(q0, q0, FT(1)); a1 = FT(6) * fma(q1, q1, FT(1)); a2 = fma(q2, q2, FT(1));
        } else {
            a0 = __fdividef(FT(3), D0 * D0); a1 = __fdividef(FT(6), D1 * D1); a2 = __fdividef(FT(1), D2 * D2);
        }
        const FT S = fma(a0, t0, fma(a1, t1, a2 * r2));
        return fma(__fdividef(S, (a0 + a1) + a2), K<FT>(2), m);
    }
}

// weno_face with the constant factors folded out (tendency_fused.cu): returns TWICE the reconstruction,
// (c + d) - S / (3 den), and folds 4 eps into the smoothness-indicator FMA (D_k = s_k^2 + (13/3 t_k) t_k + 4 eps):
// 47 FP64 instructions against 51.
template <class FT, bool ZW>
__device__ __forceinline__ FT weno_face2(FT a, FT b, FT c, FT d, FT e, FT x2, FT x0) {
    const FT t2 = fma(FT(-2), b, a) + c, t1 = fma(FT(-2), c, b) + d, t0 = fma(FT(-2), d, c) + e;
    const FT s2 = fma(FT(2), x2 - b, t2), s1 = b - d, s0 = fma(FT(2), x0 - d, t0);
    const FT k133 = K<FT>(0), eps4 = K<FT>(1);
    const FT D2 = fma(s2, s2, fma(k133 * t2, t2, eps4));      // 4 (beta_k + eps)
    const FT D1 = fma(s1, s1, fma(k133 * t1, t1, eps4));
    const FT D0 = fma(s0, s0, fma(k133 * t0, t0, eps4));
    const FT r2 = fma(FT(3), t1, FT(-2) * t2);
    const FT m2 = c + d;
    if constexpr (sizeof(FT) == 8) {
        const FT E0 = D0 * D0, E1 = D1 * D1, E2 = D2 * D2;
        const FT P12 = E1 * E2, P02 = E0 * E2, P01 = E0 * E1;
        FT g0, g1, g2;       // 10 x the unnormalised weights
        if (ZW) {
            const FT tau = D2 - D0, tt = tau * tau, PI = E0 * P12;
            g0 = FT(3) * fma(tt, P12, PI);
            g1 = FT(6) * fma(tt, P02, PI);
            g2 = fma(tt, P01, PI);
        } else {
            g0 = FT(3) * P12; g1 = FT(6) * P02; g2 = P01;
        }
        const FT den = (g0 + g1) + g2;
        const FT S = fma(g0, t0, fma(g1, t1, g2 * r2));
        const FT r0 = rcp_seed(den);
        const FT er = fma(-den, r0, FT(1));
        const FT q0 = S * r0;
        const FT q = fma(q0, fma(er, er, er), q0);     // S / den
        return fma(q, K<FT>(5), m2);
    } else {
        FT a0, a1, a2;
        if (ZW) {
            const FT tau = D2 - D0;
            const FT q0 = __fdividef(tau, D0), q1 = __fdividef(tau, D1), q2 = __fdividef(tau, D2);
            a0 = FT(3) * fma(q0, q0, FT(1)); a1 = FT(6) * fma(q1, q1, FT(1)); a2 = fma(q2, q2, FT(1));
        } else {
            a0 = __fdividef(FT(3), D0 * D0); a1 = __fdividef(FT(6), D1 * D1); a2 = __fdividef(FT(1), D2 * D2);
        }
        const FT S = fma(a0, t0, fma(a1, t1, a2 * r2));
        return fma(__fdividef(S, (a0 + a1) + a2), K<FT>(5), m2);
    }
}

// weno_face2 on a STRETCHED axis (weno_fifth_order.jl:526-553: the candidate polynomials take their coefficients from
// per-index tables, the smoothness indicators stay the uniform ones): cf = the nine coefficients of the three candidates
// for the window ordered TOWARDS the face, pre-multiplied by 2 -- p0 (c, d, e), p1 (b, c, d), p2 (a, b, c) -- so that the
// right-biased side is the same arithmetic on the mirrored window with the mirrored table row.
template <class FT, bool ZW>
__device__ __forceinline__ FT weno_face_tab2(FT a, FT b, FT c, FT d, FT e, FT x2, FT x0, const FT (&cf)[10]) {
    const FT t2 = fma(FT(-2), b, a) + c, t1 = fma(FT(-2), c, b) + d, t0 = fma(FT(-2), d, c) + e;
    const FT s2 = fma(FT(2), x2 - b, t2), s1 = b - d, s0 = fma(FT(2), x0 - d, t0);
    const FT k133 = K<FT>(0), eps4 = K<FT>(1);
    const FT D2 = fma(s2, s2, fma(k133 * t2, t2, eps4));      // 4 (beta_k + eps)
    const FT D1 = fma(s1, s1, fma(k133 * t1, t1, eps4));
    const FT D0 = fma(s0, s0, fma(k133 * t0, t0, eps4));
    const FT p0 = fma(cf[0], c, fma(cf[1], d, cf[2] * e));
    const FT p1 = fma(cf[3], b, fma(cf[4], c, cf[5] * d));
    const FT p2 = fma(cf[6], a, fma(cf[7], b, cf[8] * c));
    if constexpr (sizeof(FT) == 8) {
        const FT E0 = D0 * D0, E1 = D1 * D1, E2 = D2 * D2;
        const FT P12 = E1 * E2, P02 = E0 * E2, P01 = E0 * E1;
        FT g0, g1, g2;       // 10 x the unnormalised weights
        if (ZW) {
            const FT tau = D2 - D0, tt = tau * tau, PI = E0 * P12;
            g0 = FT(3) * fma(tt, P12, PI);
            g1 = FT(6) * fma(tt, P02, PI);
            g2 = fma(tt, P01, PI);
        } else {
            g0 = FT(3) * P12; g1 = FT(6) * P02; g2 = P01;
        }
        const FT den = (g0 + g1) + g2;
        const FT S = fma(g0, p0, fma(g1, p1, g2 * p2));
        const FT r0 = rcp_seed(den);
        const FT er = fma(-den, r0, FT(1));
        const FT q0 = S * r0;
        return fma(q0, fma(er, er, er), q0);     // S / den
    } else {
        FT a0, a1, a2;
        if (ZW) {
            const FT tau = D2 - D0;
            const FT q0 = __fdividef(tau, D0), q1 = __fdividef(tau, D1), q2 = __fdividef(tau, D2);
            a0 = FT(3) * fma(q0, q0, FT(1)); a1 = FT(6) * fma(q1, q1, FT(1)); a2 = fma(q2, q2, FT(1));
        } else {
            a0 = __fdividef(FT(3), D0 * D0); a1 = __fdividef(FT(6), D1 * D1); a2 = __fdividef(FT(1), D2 * D2);
        }
        return __fdividef(fma(a0, p0, fma(a1, p1, a2 * p2)), (a0 + a1) + a2);
    }
}

// upwind selection of the window out of the six values psi[f-3 .. f+2] around face f
template <class FT, bool ZW>
__device__ __forceinline__ FT weno_upwind(bool pos, FT w0, FT w1, FT w2, FT w3, FT w4, FT w5) {
    const FT c = pos ? w2 : w3;
    return weno_face<FT, ZW>(pos ? w0 : w5, pos ? w1 : w4, c, pos ? w3 : w2, pos ? w4 : w1,
                             pos ? c : w5, pos ? c : w1);
}

// fourth-order centred interpolation of the advecting velocity, folded with the two-point average that follows it
// (centered_fourth_order.jl:17-33 then the 1/2 (. + .) of momentum_advection_operators.jl): with
// I(c) = c0 - (c+ - 2 c0 + c-)/6,   (I(c0) + I(c1))/2 = 7/12 (c0 + c1) - 1/12 (c- + c2).
template <class FT>
__device__ __forceinline__ FT interp4(FT cm, FT c0, FT c1, FT c2) {
    return fma(K<FT>(4), cm + c2, K<FT>(3) * (c0 + c1));
}

}  // namespace wf
}  // namespace ob
