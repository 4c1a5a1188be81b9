/*
 * oracle_cpu.c -- compiled twin of the NumPy oracle (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * Plain C + OpenMP restatement of the reference's CPU algorithm for the headline configuration:
 * triply periodic (or Flat-z) regular RectilinearGrid, WENO5 (Z or JS weights) advection, one
 * buoyancy tracer with BuoyancyTracer, closure = nothing, FFT-based pressure solve, RK3.
 * Like the reference's CPU kernels, every cell evaluates both faces of every flux and both the
 * left- and right-biased reconstructions (src/Advection/momentum_advection_operators.jl:52-56,
 * upwind_biased_advective_fluxes.jl:10-128) -- this is the reference's work per point, which is
 * what the CPU baseline is meant to time.  Compiled with -ffp-contract=off so the arithmetic is
 * the same IEEE sequence as the NumPy oracle's (no FMA).  Used only by tests/ (cross-check against
 * the NumPy oracle) and by bench.py's cpu_baseline / --impl reference legs.
 *
 * Reference files followed (paths relative to /root/reference/src):
 *   TimeSteppers/runge_kutta_3.jl:81-218, Models/NonhydrostaticModels/*.jl,
 *   Advection/weno_fifth_order.jl:266-317,380-403,489-524, centered_fourth_order.jl:17-33,
 *   BoundaryConditions/fill_halo_regions_periodic.jl:37-105, Solvers/fft_based_poisson_solver.jl:93-125,
 *   Solvers/poisson_eigenvalues.jl:8-11.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define H 3

typedef struct {
    int N[3], S[3], flatz;
    long st[3];
    double d[3], L[3];
    int zweno;
} G;

static inline long IDX(const G* g, int i, int j, int k) {      /* Julia indices */
    return (i - 1 + H) * g->st[0] + (j - 1 + H) * g->st[1] + (g->flatz ? 0 : (k - 1 + H) * g->st[2]);
}

/* fill_periodic_*_halo!: x then y then z, each over the full extent of the other dims */
static void fill_halo(const G* g, double* f) {
    int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
    long sx = g->st[0], sy = g->st[1], sz = g->st[2];
    int Sy = g->S[1], Sz = g->S[2];
#pragma omp parallel for collapse(2)
    for (int c = 0; c < Sz; ++c)
        for (int b = 0; b < Sy; ++b) {
            double* r = f + b * sy + c * sz;
            for (int h = 0; h < H; ++h) { r[h * sx] = r[(Nx + h) * sx]; r[(Nx + H + h) * sx] = r[(H + h) * sx]; }
        }
    int Sx = g->S[0];
#pragma omp parallel for collapse(2)
    for (int c = 0; c < Sz; ++c)
        for (int a = 0; a < Sx; ++a) {
            double* r = f + a * sx + c * sz;
            for (int h = 0; h < H; ++h) { r[h * sy] = r[(Ny + h) * sy]; r[(Ny + H + h) * sy] = r[(H + h) * sy]; }
        }
    if (g->flatz) return;
#pragma omp parallel for collapse(2)
    for (int b = 0; b < Sy; ++b)
        for (int a = 0; a < Sx; ++a) {
            double* r = f + a * sx + b * sy;
            for (int h = 0; h < H; ++h) { r[h * sz] = r[(Nz + h) * sz]; r[(Nz + H + h) * sz] = r[(H + h) * sz]; }
        }
}

/* weno_{left,right}_biased_interpolate (weno_fifth_order.jl:489-498); window = the 5 values of the side */
static inline double weno(int right, int zweno, double a, double b, double c, double d, double e) {
    /* psi2 = (a,b,c), psi1 = (b,c,d), psi0 = (c,d,e) */
    double t0 = c - 2 * d + e, t1 = b - 2 * c + d, t2 = a - 2 * b + c;
    double s0, s1 = b - d, s2, C0, C1 = 3.0 / 5.0, C2;
    double p0, p1, p2;
    if (!right) {
        s0 = 3 * c - 4 * d + e; s2 = a - 4 * b + 3 * c; C0 = 3.0 / 10.0; C2 = 1.0 / 10.0;
        p0 = (1.0 / 3.0) * c + (5.0 / 6.0) * d + (-(1.0 / 6.0)) * e;
        p1 = (-(1.0 / 6.0)) * b + (5.0 / 6.0) * c + (1.0 / 3.0) * d;
        p2 = (1.0 / 3.0) * a + (-(7.0 / 6.0)) * b + (11.0 / 6.0) * c;
    } else {
        s0 = c - 4 * d + 3 * e; s2 = 3 * a - 4 * b + c; C0 = 1.0 / 10.0; C2 = 3.0 / 10.0;
        p0 = (11.0 / 6.0) * c + (-(7.0 / 6.0)) * d + (1.0 / 3.0) * e;
        p1 = (1.0 / 3.0) * b + (5.0 / 6.0) * c + (-(1.0 / 6.0)) * d;
        p2 = (-(1.0 / 6.0)) * a + (5.0 / 6.0) * b + (1.0 / 3.0) * c;
    }
    double b0 = (13.0 / 12.0) * (t0 * t0) + 0.25 * (s0 * s0);
    double b1 = (13.0 / 12.0) * (t1 * t1) + 0.25 * (s1 * s1);
    double b2 = (13.0 / 12.0) * (t2 * t2) + 0.25 * (s2 * s2);
    const double eps = 1e-6;
    double a0, a1, a2;
    if (zweno) {
        double tau = fabs(b2 - b0);
        double q0 = tau / (b0 + eps), q1 = tau / (b1 + eps), q2 = tau / (b2 + eps);
        a0 = C0 * (1 + q0 * q0); a1 = C1 * (1 + q1 * q1); a2 = C2 * (1 + q2 * q2);
    } else {
        double d0 = b0 + eps, d1 = b1 + eps, d2 = b2 + eps;
        a0 = C0 / (d0 * d0); a1 = C1 / (d1 * d1); a2 = C2 / (d2 * d2);
    }
    double sa = a0 + a1 + a2;
    double w0 = a0 / sa, w1 = a1 / sa, w2 = a2 / sa;
    return w0 * p0 + w1 * p1 + w2 * p2;
}
static inline double wenoL(const G* g, const double* f, long p, long s) {
    return weno(0, g->zweno, f[p - 3 * s], f[p - 2 * s], f[p - s], f[p], f[p + s]);
}
static inline double wenoR(const G* g, const double* f, long p, long s) {
    return weno(1, g->zweno, f[p - 2 * s], f[p - s], f[p], f[p + s], f[p + 2 * s]);
}
static inline double I3(const double* c, long p, long s) {
    return c[p] - ((c[p + s] - c[p]) - (c[p] - c[p - s])) / 6;
}
static inline double upw(double u, double l, double r) { return ((u + fabs(u)) * l + (u - fabs(u)) * r) / 2; }

/* advective_momentum_flux_{A}{B}: advection OF component B BY component A at linear index p */
static inline double mflux(const G* g, int A, int B, const double* Ua, const double* psi, long p) {
    long sA = g->st[A], sB = g->st[B];
    double ar = A == 0 ? g->d[1] * g->d[2] : (A == 1 ? g->d[0] * g->d[2] : g->d[0] * g->d[1]);
    double ut;
    long pf = p;
    if (A == B) { ut = 0.5 * (I3(Ua, p, sA) + I3(Ua, p + sA, sA)); pf = p + sA; }
    else if (B == 2 && g->flatz) ut = Ua[p];
    else ut = 0.5 * (I3(Ua, p - sB, sB) + I3(Ua, p, sB));
    return ar * upw(ut, wenoL(g, psi, pf, sA), wenoR(g, psi, pf, sA));
}
static inline double tflux(const G* g, int A, const double* Ua, const double* c, long p) {
    long sA = g->st[A];
    double ar = A == 0 ? g->d[1] * g->d[2] : (A == 1 ? g->d[0] * g->d[2] : g->d[0] * g->d[1]);
    return ar * upw(Ua[p], wenoL(g, c, p, sA), wenoR(g, c, p, sA));
}

static void tendencies(const G* g, double* const* F, const double* pHY, double** Gn) {
    int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
    double V = (g->d[0] * g->d[1]) * g->d[2];
    int nd = g->flatz ? 2 : 3;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; ++k)
        for (int j = 1; j <= Ny; ++j)
            for (int i = 1; i <= Nx; ++i) {
                long p = IDX(g, i, j, k);
                for (int B = 0; B < 3; ++B) {          /* div_𝐯u, div_𝐯v, div_𝐯w */
                    double t[3] = {0, 0, 0};
                    for (int A = 0; A < nd; ++A) {
                        long s = g->st[A];
                        if (A == B) t[A] = mflux(g, A, B, F[A], F[B], p) - mflux(g, A, B, F[A], F[B], p - s);
                        else t[A] = mflux(g, A, B, F[A], F[B], p + s) - mflux(g, A, B, F[A], F[B], p);
                    }
                    double Gv = -(1 / V * (t[0] + t[1] + t[2]));
                    if (pHY && B < 2) Gv = Gv - (pHY[p] - pHY[p - g->st[B]]) / g->d[B];
                    Gn[B][p] = Gv;
                }
                double t[3] = {0, 0, 0};
                for (int A = 0; A < nd; ++A) t[A] = tflux(g, A, F[A], F[3], p + g->st[A]) - tflux(g, A, F[A], F[3], p);
                Gn[3][p] = -(1 / V * (t[0] + t[1] + t[2]));
            }
}

/* _update_hydrostatic_pressure! (update_hydrostatic_pressure.jl:10-18) */
static void hydrostatic(const G* g, const double* b, double* pHY) {
    int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
#pragma omp parallel for collapse(2)
    for (int j = 1; j <= Ny; ++j)
        for (int i = 1; i <= Nx; ++i) {
            double acc = -(0.5 * (b[IDX(g, i, j, Nz)] + b[IDX(g, i, j, Nz + 1)])) * g->d[2];
            pHY[IDX(g, i, j, Nz)] = acc;
            for (int k = Nz - 1; k >= 1; --k) {
                acc = acc - (0.5 * (b[IDX(g, i, j, k)] + b[IDX(g, i, j, k + 1)])) * g->d[2];
                pHY[IDX(g, i, j, k)] = acc;
            }
        }
}

/* ---- in-place complex FFT (iterative radix-2, power-of-two lengths), stride access ---------- */
static void fft1d(double* re, double* im, int n, long s, int inverse, const double* cs, const double* sn) {
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { double t = re[i * s]; re[i * s] = re[j * s]; re[j * s] = t; t = im[i * s]; im[i * s] = im[j * s]; im[j * s] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len)
            for (int k = 0; k < half; ++k) {
                double wr = cs[k * step], wi = inverse ? sn[k * step] : -sn[k * step];
                double* ar = re + (i + k) * s; double* ai = im + (i + k) * s;
                double* br = re + (i + k + half) * s; double* bi = im + (i + k + half) * s;
                double xr = *br * wr - *bi * wi, xi = *br * wi + *bi * wr;
                *br = *ar - xr; *bi = *ai - xi; *ar = *ar + xr; *ai = *ai + xi;
            }
    }
}
static void fft3d(const G* g, double* re, double* im, int inverse) {
    int N[3] = {g->N[0], g->N[1], g->N[2]};
    long st[3] = {1, N[0], (long)N[0] * N[1]};
    for (int d = 0; d < 3; ++d) {
        int n = N[d];
        if (n == 1) continue;
        double* cs = malloc(sizeof(double) * n), *sn = malloc(sizeof(double) * n);
        for (int k = 0; k < n; ++k) { cs[k] = cos(2 * M_PI * k / n); sn[k] = sin(2 * M_PI * k / n); }
        int a = d == 0 ? 1 : 0, b = d == 2 ? 1 : 2;
#pragma omp parallel for collapse(2)
        for (int ib = 0; ib < N[b]; ++ib)
            for (int ia = 0; ia < N[a]; ++ia) {
                long off = ia * st[a] + ib * st[b];
                fft1d(re + off, im + off, n, st[d], inverse, cs, sn);
            }
        free(cs); free(sn);
    }
    if (inverse) {
        double sc = 1.0 / ((double)N[0] * N[1] * N[2]);
        long tot = (long)N[0] * N[1] * N[2];
#pragma omp parallel for
        for (long q = 0; q < tot; ++q) { re[q] *= sc; im[q] *= sc; }
    }
}

/* calculate_pressure_correction! + pressure_correct_velocities! (pressure_correction.jl:10-56) */
static void pressure_step(const G* g, double** F, double* pN, double dt, double* re, double* im) {
    int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
    for (int q = 0; q < 3; ++q) fill_halo(g, F[q]);
    double ax = g->d[1] * g->d[2], ay = g->d[0] * g->d[2], az = g->d[0] * g->d[1], V = (g->d[0] * g->d[1]) * g->d[2];
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; ++k)
        for (int j = 1; j <= Ny; ++j)
            for (int i = 1; i <= Nx; ++i) {
                long p = IDX(g, i, j, k);
                double tx = ax * F[0][p + g->st[0]] - ax * F[0][p];
                double ty = ay * F[1][p + g->st[1]] - ay * F[1][p];
                double tz = g->flatz ? 0.0 : az * F[2][p + g->st[2]] - az * F[2][p];
                long q = (i - 1) + (long)Nx * ((j - 1) + (long)Ny * (k - 1));
                re[q] = (1 / V * (tx + ty + tz)) / dt;
                im[q] = 0;
            }
    fft3d(g, re, im, 0);
#pragma omp parallel for collapse(2)
    for (int k = 0; k < Nz; ++k)
        for (int j = 0; j < Ny; ++j)
            for (int i = 0; i < Nx; ++i) {
                double lx = 2 * sin(i * M_PI / Nx) / (g->L[0] / Nx), ly = 2 * sin(j * M_PI / Ny) / (g->L[1] / Ny);
                double lz = g->flatz ? 0.0 : 2 * sin(k * M_PI / Nz) / (g->L[2] / Nz);
                double lam = lx * lx + ly * ly + lz * lz;
                long q = i + (long)Nx * (j + (long)Ny * k);
                if (q == 0) { re[q] = 0; im[q] = 0; }
                else { re[q] = -re[q] / lam; im[q] = -im[q] / lam; }
            }
    fft3d(g, re, im, 1);
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; ++k)
        for (int j = 1; j <= Ny; ++j)
            for (int i = 1; i <= Nx; ++i)
                pN[IDX(g, i, j, k)] = re[(i - 1) + (long)Nx * ((j - 1) + (long)Ny * (k - 1))];
    fill_halo(g, pN);
#pragma omp parallel for collapse(2)
    for (int k = 1; k <= Nz; ++k)
        for (int j = 1; j <= Ny; ++j)
            for (int i = 1; i <= Nx; ++i) {
                long p = IDX(g, i, j, k);
                F[0][p] -= (pN[p] - pN[p - g->st[0]]) / g->d[0] * dt;
                F[1][p] -= (pN[p] - pN[p - g->st[1]]) / g->d[1] * dt;
                if (!g->flatz) F[2][p] -= (pN[p] - pN[p - g->st[2]]) / g->d[2] * dt;
            }
}

static void update_state(const G* g, double** F, double* pHY) {
    for (int q = 0; q < 4; ++q) fill_halo(g, F[q]);
    if (!g->flatz) { hydrostatic(g, F[3], pHY); fill_halo(g, pHY); }
}

/*
 * Run `nsteps` RK3 steps.  u, v, w, b: interior arrays Nx*Ny*Nz (x fastest), updated in place.
 * project != 0 applies the set! projection (set_nonhydrostatic_model.jl:51-56) first.
 */
int oc_rk3_run(const int* N, const double* L, int zweno, double* u, double* v, double* w, double* b,
               int nsteps, double dt, int project, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    G g;
    g.flatz = N[2] == 1;
    for (int d = 0; d < 3; ++d) {
        g.N[d] = N[d]; g.L[d] = L[d];
        g.S[d] = (d == 2 && g.flatz) ? 1 : N[d] + 2 * H;
        g.d[d] = (d == 2 && g.flatz) ? 1.0 : L[d] / N[d];
    }
    if (g.flatz) g.L[2] = 1.0;
    g.st[0] = 1; g.st[1] = g.S[0]; g.st[2] = (long)g.S[0] * g.S[1];
    g.zweno = zweno;
    long tot = (long)g.S[0] * g.S[1] * g.S[2], ni = (long)N[0] * N[1] * N[2];
    double *F[4], *Gn[4], *Gm[4], *in[4] = {u, v, w, b};
    for (int q = 0; q < 4; ++q) { F[q] = calloc(tot, 8); Gn[q] = calloc(tot, 8); Gm[q] = calloc(tot, 8); }
    double* pN = calloc(tot, 8), *pHY = calloc(tot, 8), *re = malloc(ni * 8), *im = malloc(ni * 8);
    for (int q = 0; q < 4; ++q)
        for (int k = 1; k <= N[2]; ++k) for (int j = 1; j <= N[1]; ++j) for (int i = 1; i <= N[0]; ++i)
            F[q][IDX(&g, i, j, k)] = in[q][(i - 1) + (long)N[0] * ((j - 1) + (long)N[1] * (k - 1))];
    update_state(&g, F, pHY);
    if (project) { pressure_step(&g, F, pN, 1.0, re, im); update_state(&g, F, pHY); }
    const double g1 = 8.0 / 15.0, g2 = 5.0 / 12.0, g3 = 3.0 / 4.0, z2 = -17.0 / 60.0, z3 = -5.0 / 12.0;
    const double gam[3] = {g1, g2, g3}, zet[3] = {0, z2, z3};
    const double sdt[3] = {g1 * dt, (g2 + z2) * dt, (g3 + z3) * dt};
    for (int n = 0; n < nsteps; ++n)
        for (int s = 0; s < 3; ++s) {
            tendencies(&g, F, g.flatz ? NULL : pHY, Gn);
#pragma omp parallel for collapse(2)
            for (int k = 1; k <= N[2]; ++k)
                for (int j = 1; j <= N[1]; ++j)
                    for (int i = 1; i <= N[0]; ++i) {
                        long p = IDX(&g, i, j, k);
                        for (int q = 0; q < 4; ++q) {
                            if (s == 0) F[q][p] += dt * gam[0] * Gn[q][p];
                            else F[q][p] += dt * (gam[s] * Gn[q][p] + zet[s] * Gm[q][p]);
                        }
                    }
            pressure_step(&g, F, pN, sdt[s], re, im);
            if (s < 2) for (int q = 0; q < 4; ++q) { double* t = Gn[q]; Gn[q] = Gm[q]; Gm[q] = t; }
            update_state(&g, F, pHY);
        }
    for (int q = 0; q < 4; ++q)
        for (int k = 1; k <= N[2]; ++k) for (int j = 1; j <= N[1]; ++j) for (int i = 1; i <= N[0]; ++i)
            in[q][(i - 1) + (long)N[0] * ((j - 1) + (long)N[1] * (k - 1))] = F[q][IDX(&g, i, j, k)];
    for (int q = 0; q < 4; ++q) { free(F[q]); free(Gn[q]); free(Gm[q]); }
    free(pN); free(pHY); free(re); free(im);
    return 0;
}

int oc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
