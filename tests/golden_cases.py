"""Definitions of the golden cases (shared by tests/golden/make_golden.py and the tests that read the fixtures)."""
import numpy as np


def _zf(n, p=1.5):
    return -np.linspace(1, 0, n + 1) ** p


MODEL_CASES = {
    # headline physics (BASELINE.json configs[1]) at a size that takes the specialised TMA tendency kernels and the
    # fast FFT path: triply periodic, WENO5, buoyancy tracer, RK3
    "c2_periodic_weno_rk3": dict(grid=dict(size=(32, 16, 16), extent=(1, 1, 1), topology=("Periodic",) * 3),
                                 adv="WENO5", tracers=("b",), buoyancy=True, ts="RungeKutta3", dt=2e-3, steps=3, seed=101),
    # configs[0] physics (README example): 2-D periodic turbulence, Flat z, default AB2
    "c1_2d_flat_weno_ab2": dict(grid=dict(size=(32, 32), extent=(2 * np.pi, 2 * np.pi), topology=("Periodic", "Periodic", "Flat")),
                                adv="WENO5", tracers=(), buoyancy=False, ts="QuasiAdamsBashforth2", dt=5e-3, steps=5, seed=102),
    # configs[2] physics: Bounded stretched z, WENO5(grid), Fourier-tridiagonal solver, flux / gradient BCs, FPlane
    "c3_stretched_weno_rk3": dict(grid=dict(size=(12, 8, 14), x=(0, 1), y=(0, 1), z=_zf(14),
                                            topology=("Periodic", "Periodic", "Bounded")),
                                  adv="WENO5grid", tracers=("b",), buoyancy=True, closure=("ThreeDimensional", 1e-4, 1e-4),
                                  f=1e-2, ts="RungeKutta3", dt=5e-3, steps=3, seed=103,
                                  bcs={"u": {"top": ("Flux", -1e-3)}, "b": {"top": ("Flux", 1e-4), "bottom": ("Gradient", 1e-2)}}),
}

POISSON_CASES = {
    "fft_ppp": dict(grid=dict(size=(32, 16, 16), extent=(1, 2, 1.5), topology=("Periodic",) * 3), solver="fft", seed=201),
    "fft_ppb": dict(grid=dict(size=(12, 10, 9), extent=(1, 1, 1), topology=("Periodic", "Periodic", "Bounded")), solver="fft", seed=202),
    "ft_ppb_stretched": dict(grid=dict(size=(12, 8, 14), x=(0, 1), y=(0, 1), z=_zf(14),
                                       topology=("Periodic", "Periodic", "Bounded")), solver="ft", seed=203),
}


def build_model(M, grid, cfg):
    """M is the module providing the reference-named constructors: `oracle` or `ocean_b200`."""
    FT = np.float64
    adv = cfg["adv"]
    a = M.WENO5(grid=grid) if adv == "WENO5grid" else (M.WENO5(FT) if adv == "WENO5" else getattr(M, adv)())
    clo = M.ScalarDiffusivity(cfg["closure"][0], ν=cfg["closure"][1], κ=cfg["closure"][2]) if cfg.get("closure") else None
    cor = M.FPlane(cfg["f"]) if cfg.get("f") else None
    bu = M.Buoyancy(M.BuoyancyTracer(), None) if cfg.get("buoyancy") else None
    bcs = None
    if cfg.get("bcs"):
        bcs = {n: {s: M.BoundaryCondition(*kv) for s, kv in d.items()} for n, d in cfg["bcs"].items()}
    return M.NonhydrostaticModel(grid, advection=a, closure=clo, coriolis=cor, buoyancy=bu, tracers=cfg["tracers"],
                                 timestepper=cfg["ts"], boundary_conditions=bcs)


def build_oracle_model(O, cfg):
    return build_model(O, O.RectilinearGrid(np.float64, **cfg["grid"]), cfg)


def model_initial_values(oracle_model, seed):
    """seeded initial state: uniform(-1,1) velocities with the mean removed, tracers = 0.5 z + 0.1 noise"""
    rng = np.random.default_rng(seed)
    vals = {}
    for n in oracle_model.names:
        f = oracle_model.fields[n]
        a = rng.uniform(-1, 1, f.size())
        if n in "uvw":
            a = a - a.mean()
        else:
            a = 0.5 * oracle_model.grid.nodes(f.loc)[2] + 0.1 * a
        vals[n] = a
    return vals


def poisson_rhs(oracle_grid, seed):
    """random source term satisfying the solvability condition of the all-Neumann / periodic problem: its
    VOLUME-weighted mean is zero (on a stretched z the plain mean is not enough: the kx = ky = 0 tridiagonal system
    is singular and an incompatible right-hand side makes the answer round-off noise)"""
    from oracle.grids import Center
    from oracle.fields import R
    g = oracle_grid
    rng = np.random.default_rng(seed)
    r = rng.uniform(-1, 1, (g.Nx, g.Ny, g.Nz))
    dz = np.broadcast_to(np.asarray(g.Δz(Center, R(1, g.Nz)), dtype=np.float64), (1, 1, g.Nz) if not g.regular[2] else ())
    dz = np.ones((1, 1, g.Nz)) * dz
    return r - np.sum(r * dz) / (np.sum(dz) * g.Nx * g.Ny)
