"""Long-run trace parity (BASELINE.json north_star: global kinetic-energy / tracer-variance traces must
agree to 1e-9 relative after 1000 steps).  Config 1 physics (README example): 2-D periodic turbulence,
WENO5, default AB2 time stepper, Flat z -- at 64^2 for 1000 steps and at BASELINE configs[0]'s own 128^2 for 400 steps
(t = 4); the headline physics (3-D, RK3) at 32^3 for 200 steps against the NumPy oracle and at 64^3 for 1000 steps
against its compiled twin."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ob():
    import ocean_b200 as ob
    ob.arch = ob.B200()
    return ob


def test_c1_2d_turbulence_1000_step_traces(ob):
    N = 64
    kw = dict(size=(N, N), extent=(2 * np.pi, 2 * np.pi), topology=("Periodic", "Periodic", "Flat"))
    go = O.RectilinearGrid(np.float64, **kw)
    gb = ob.RectilinearGrid(ob.arch, np.float64, **kw)
    mo = O.NonhydrostaticModel(go, advection=O.WENO5(), tracers=("c",))
    mb = ob.NonhydrostaticModel(gb, advection=ob.WENO5(), tracers=("c",))
    rng = np.random.default_rng(1)
    vals = {n: rng.uniform(-1, 1, (N, N, 1)) for n in "uv"}          # README.md:99-100
    vals["c"] = np.sin(go.nodes(("c", "c", "c"))[0]) + 0 * vals["u"]
    mo.set(**vals)
    ob.set_model(mb, **vals)
    dt = 0.01
    ke_o, ke_b, var_o, var_b = [], [], [], []
    for step in range(1000):
        mo.time_step(dt)
        ob.time_step(mb, dt)
        if (step + 1) % 100 == 0:
            ke_o.append(mo.kinetic_energy())
            ke_b.append(mb.diagnostics()["kinetic_energy"])
            var_o.append(float(np.sum(mo.tracers["c"].interior ** 2)))
            var_b.append(mb.tracers["c"].reduce()["sumsq"])
    ke_o, ke_b, var_o, var_b = map(np.array, (ke_o, ke_b, var_o, var_b))
    assert np.max(np.abs(ke_b - ke_o) / ke_o) < 1e-9, np.abs(ke_b - ke_o) / ke_o
    assert np.max(np.abs(var_b - var_o) / var_o) < 1e-9, np.abs(var_b - var_o) / var_o
    # the fields themselves stay close too (chaotic growth of rounding differences stays small over t = 10)
    u_o, u_b = mo.velocities["u"].interior, mb.velocities["u"].interior()
    assert np.max(np.abs(u_b - u_o)) / np.max(np.abs(u_o)) < 1e-8


def test_c2_3d_rk3_200_step_traces(ob):
    N = (32, 32, 32)
    kw = dict(size=N, extent=(1, 1, 1), topology=("Periodic",) * 3)
    go = O.RectilinearGrid(np.float64, **kw)
    gb = ob.RectilinearGrid(ob.arch, np.float64, **kw)
    mo = O.NonhydrostaticModel(go, advection=O.WENO5(), tracers=("b",), buoyancy=O.BuoyancyTracer(),
                               timestepper="RungeKutta3")
    mb = ob.NonhydrostaticModel(gb, advection=ob.WENO5(), tracers=("b",), buoyancy=ob.BuoyancyTracer(),
                                timestepper="RungeKutta3")
    rng = np.random.default_rng(2)
    vals = {}
    for n in "uvw":
        a = rng.uniform(-1, 1, N)
        vals[n] = a - a.mean()
    vals["b"] = 1e-1 * go.nodes(("c", "c", "c"))[2] + 1e-2 * rng.uniform(-1, 1, N)
    mo.set(**vals)
    ob.set_model(mb, **vals)
    for _ in range(200):
        mo.time_step(2e-3)
        ob.time_step(mb, 2e-3)
    ke_o, ke_b = mo.kinetic_energy(), mb.diagnostics()["kinetic_energy"]
    var_o, var_b = float(np.sum(mo.tracers["b"].interior ** 2)), mb.tracers["b"].reduce()["sumsq"]
    assert abs(ke_b - ke_o) / ke_o < 1e-9
    assert abs(var_b - var_o) / var_o < 1e-9


def test_c1_128x128_400_step_traces(ob):
    """BASELINE.json configs[0] at its own size: 128^2, 400 AB2 steps of 0.01 (t = 4), traces sampled every 50 steps"""
    N = 128
    kw = dict(size=(N, N), extent=(2 * np.pi, 2 * np.pi), topology=("Periodic", "Periodic", "Flat"))
    go = O.RectilinearGrid(np.float64, **kw)
    gb = ob.RectilinearGrid(ob.arch, np.float64, **kw)
    mo = O.NonhydrostaticModel(go, advection=O.WENO5(), tracers=("c",))
    mb = ob.NonhydrostaticModel(gb, advection=ob.WENO5(), tracers=("c",))
    rng = np.random.default_rng(5)
    vals = {n: rng.uniform(-1, 1, (N, N, 1)) for n in "uv"}
    vals["c"] = np.sin(go.nodes(("c", "c", "c"))[0]) + 0 * vals["u"]
    mo.set(**vals)
    ob.set_model(mb, **vals)
    worst_ke = worst_var = 0.0
    for step in range(400):
        mo.time_step(0.01)
        ob.time_step(mb, 0.01)
        if (step + 1) % 50 == 0:
            ke_o, var_o = mo.kinetic_energy(), float(np.sum(mo.tracers["c"].interior ** 2))
            worst_ke = max(worst_ke, abs(mb.diagnostics()["kinetic_energy"] - ke_o) / ke_o)
            worst_var = max(worst_var, abs(mb.tracers["c"].reduce()["sumsq"] - var_o) / var_o)
    assert worst_ke < 1e-9 and worst_var < 1e-9, (worst_ke, worst_var)
    assert abs(mb.clock.time - 4.0) < 1e-12


def test_c2_64cubed_1000_step_traces_against_compiled_oracle(ob):
    """headline physics (triply periodic, WENO5 + b, FFT solve, RK3) at 64^3 for 1000 steps: kinetic energy and buoyancy
    variance of the CUDA path against the compiled OpenMP twin of the oracle (oracle/oracle_cpu.c), sampled every 250 steps"""
    import os
    from oracle import cpu_twin
    N, dt, chunk = 64, 1e-3, 250
    rng = np.random.default_rng(6)
    vals = {}
    for n in "uvw":
        a = rng.uniform(-1, 1, (N, N, N))
        vals[n] = a - a.mean()
    z = (np.arange(N) + 0.5) / N
    vals["b"] = 1e-1 * z.reshape(1, 1, N) + 1e-2 * rng.uniform(-1, 1, (N, N, N))
    gb = ob.RectilinearGrid(ob.arch, np.float64, size=(N, N, N), extent=(1, 1, 1), topology=("Periodic",) * 3)
    mb = ob.NonhydrostaticModel(gb, advection=ob.WENO5(), tracers=("b",), buoyancy=ob.BuoyancyTracer(), timestepper="RungeKutta3")
    ob.set_model(mb, **vals)
    ref = [vals[n] for n in "uvwb"]
    nth = len(os.sched_getaffinity(0))
    worst_ke = worst_var = 0.0
    for it in range(1000 // chunk):
        # the twin projects its input first; the state it returns is divergence free, so restarting it is a no-op projection
        ref = cpu_twin.rk3_run((N, N, N), (1.0, 1.0, 1.0), ref[0], ref[1], ref[2], ref[3], chunk, dt, project=(it == 0), nthreads=nth)
        for _ in range(chunk):
            ob.time_step(mb, dt)
        u, v, w, b = ref
        ke_o = 0.5 * float(np.sum(u ** 2) + np.sum(v ** 2) + np.sum(w ** 2))      # oracle.model.kinetic_energy
        var_o = float(np.sum(b ** 2))
        d = mb.diagnostics()
        worst_ke = max(worst_ke, abs(d["kinetic_energy"] - ke_o) / ke_o)
        worst_var = max(worst_var, abs(mb.tracers["b"].reduce()["sumsq"] - var_o) / var_o)
    assert worst_ke < 1e-9 and worst_var < 1e-9, (worst_ke, worst_var)
