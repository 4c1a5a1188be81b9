#!/usr/bin/env python
"""fast FFT Poisson path on a list of sizes: residual of lap(phi) = rhs against numpy (debug helper)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clima-oceananigans.jl_b200")):
    sys.path.insert(0, p)
import ocean_b200 as ob
arch = ob.B200(0)
sizes = [(32, 16, 16), (64, 32, 128), (64, 64, 64), (128, 64, 32), (32, 256, 16), (64, 512, 16), (64, 16, 512), (128, 128, 128), (256, 256, 256)]
for N in sizes:
    try:
        g = ob.RectilinearGrid(arch, np.float64, size=N, extent=(1, 1, 1), topology=("Periodic",) * 3)
        rng = np.random.default_rng(1)
        rhs = rng.uniform(-1, 1, N); rhs -= rhs.mean()
        phi = ob.CenterField(g)
        ob.solve(phi, ob.FFTBasedPoissonSolver(g), rhs)
        ob.sync()
        p = phi.interior()
        d = [1.0 / n for n in N]
        lap = sum((np.roll(p, -1, a) - 2 * p + np.roll(p, 1, a)) / d[a] ** 2 for a in range(3))
        print(N, "residual", float(np.max(np.abs(lap - rhs)) / np.max(np.abs(rhs))))
    except Exception as e:
        print(N, "ERROR", e)
