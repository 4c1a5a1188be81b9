#!/usr/bin/env python
"""BASELINE.json config 4: standalone Poisson solve sweep (FFT-based, triply periodic, Float64), N^3 for
N in 128..1024.  Prints solves/s, points/s and achieved GB/s against the 12-words-per-point algorithmic
traffic of SURVEY.md 8(d), plus the residual max|lap(phi) - R| / max|R| of the reference's own check."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clima-oceananigans.jl_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import ocean_b200 as ob
    from ocean_b200._lib import lib
    arch = ob.B200(0)
    stream = torch.cuda.current_stream()
    lib.ob200_set_stream(C.c_void_p(stream.cuda_stream))
    out = []
    for N in [int(x) for x in (sys.argv[1:] or ["128", "256", "512"])]:
        g = ob.RectilinearGrid(arch, np.float64, size=(N, N, N), extent=(1, 1, 1), topology=("Periodic",) * 3)
        rng = np.random.default_rng(4)
        U = {}
        for n, f in (("u", ob.XFaceField), ("v", ob.YFaceField), ("w", ob.ZFaceField)):
            U[n] = f(g)
            U[n].set(rng.uniform(-1, 1, (N, N, N)))
        ob.fill_halo_regions(list(U.values()))
        s = ob.FFTBasedPoissonSolver(g)
        phi = ob.CenterField(g)
        for _ in range(3):
            ob.solve_for_pressure(phi, s, 1.0, U)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            ob.solve_for_pressure(phi, s, 1.0, U)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res = None
        if N <= 256:     # the reference's own check: lap(phi) == div(U) (test/dependencies_for_poisson_solvers.jl:86-104)
            ob.fill_halo_regions(phi)
            p = phi.parent()
            d, H = 1.0 / N, 3
            I = slice(H, N + H)
            c = p[I, I, I]
            lap = ((p[H + 1:N + H + 1, I, I] - 2 * c + p[H - 1:N + H - 1, I, I]) +
                   (p[I, H + 1:N + H + 1, I] - 2 * c + p[I, H - 1:N + H - 1, I]) +
                   (p[I, I, H + 1:N + H + 1] - 2 * c + p[I, I, H - 1:N + H - 1])) / d ** 2
            u, v, w = (U[n].parent() for n in "uvw")
            div = ((u[H + 1:N + H + 1, I, I] - u[I, I, I]) + (v[I, H + 1:N + H + 1, I] - v[I, I, I]) +
                   (w[I, I, H + 1:N + H + 1] - w[I, I, I])) / d
            res = float(np.max(np.abs(lap - div)) / np.max(np.abs(div)))
        out.append({"N": N, "ms_per_solve": ms, "solves_per_s": 1e3 / ms, "points_per_s": N ** 3 / (ms * 1e-3),
                    "algorithmic_GBps": 96.0 * N ** 3 / (ms * 1e-3) / 1e9, "residual": res})
        del s, phi, U, g
    print(json.dumps(out))


if __name__ == "__main__":
    main()
