#!/usr/bin/env python
"""BASELINE.json configs[2]: 3-D Bounded-z vertically stretched 512^2 x 256 (ocean wind-mixing / convection set-up),
WENO5(grid), buoyancy tracer, FPlane, ScalarDiffusivity, top flux / bottom gradient BCs, Fourier-tridiagonal solver, RK3.
Prints ms per step and grid-point updates/s (CUDA events).  usage: tools/c3_bench.py [Nx Ny Nz] [steps]"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clima-oceananigans.jl_b200")):
    sys.path.insert(0, p)


def z_faces(Nz, Lz=32.0, refinement=1.2, stretching=12.0):
    """examples/ocean_wind_mixing_and_convection.jl:41-54, scaled to Nz levels"""
    k = np.arange(1, Nz + 2)
    h = (Nz + 1 - k) / Nz
    zeta0 = 1 + (h - 1) / refinement
    Sigma = (1 - np.exp(-stretching * h)) / (1 - np.exp(-stretching))
    return Lz * (zeta0 * Sigma - 1)


def main():
    import ctypes as C
    import torch
    import ocean_b200 as ob
    from ocean_b200._lib import lib
    a = [int(x) for x in sys.argv[1:]]
    Nx, Ny, Nz = (a + [512, 512, 256])[:3] if len(a) >= 3 else (512, 512, 256)
    steps = a[3] if len(a) > 3 else 3
    arch = ob.B200(0)
    stream = torch.cuda.current_stream()
    lib.ob200_set_stream(C.c_void_p(stream.cuda_stream))
    g = ob.RectilinearGrid(arch, np.float64, size=(Nx, Ny, Nz), x=(0, 64), y=(0, 64), z=z_faces(Nz),
                           topology=("Periodic", "Periodic", "Bounded"))
    bcs = {"u": {"top": ob.BoundaryCondition("Flux", -1e-4)},
           "b": {"top": ob.BoundaryCondition("Flux", 1e-8), "bottom": ob.BoundaryCondition("Gradient", 1e-5)}}
    m = ob.NonhydrostaticModel(g, advection=ob.WENO5(grid=g), tracers=("b",), buoyancy=ob.Buoyancy(ob.BuoyancyTracer(), None),
                               coriolis=ob.FPlane(1e-4), closure=ob.ScalarDiffusivity("ThreeDimensional", ν=1e-4, κ=1e-4),
                               timestepper="RungeKutta3", boundary_conditions=bcs)
    rng = np.random.default_rng(3)
    vals = {n: 1e-2 * rng.uniform(-1, 1, m.fields[n].size()) for n in "uvw"}
    zc = 0.5 * (z_faces(Nz)[1:] + z_faces(Nz)[:-1])
    vals["b"] = 1e-5 * zc.reshape(1, 1, Nz) + 1e-7 * rng.uniform(-1, 1, (Nx, Ny, Nz))
    ob.set_model(m, **vals)
    dt = 0.05
    ob.time_step(m, dt)
    lib.ob200_profile_reset(); lib.ob200_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(stream)
    for _ in range(steps):
        ob.time_step(m, dt)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.ob200_profile_enable(0)
    ph = {}
    for name in ("tendency", "poisson", "halo", "pressure_correct", "hydrostatic"):
        t, c = C.c_double(), C.c_int64()
        lib.ob200_profile_query(name.encode(), C.byref(t), C.byref(c))
        ph[name] = round(t.value / steps, 3)
    d = m.diagnostics()
    print({"workload": f"C3 {Nx}x{Ny}x{Nz} stretched Bounded z", "ms_per_step": ms, "points_per_s": Nx * Ny * Nz / (ms * 1e-3),
           "phases_ms": ph, "max_abs_div": d["max_abs_div"], "ke": d["kinetic_energy"]})


if __name__ == "__main__":
    main()
