"""Ports of the reference's own analytic / self-consistency tests for the in-scope physics, run on BOTH back ends:
the oracle (CPU, pins the restatement independently of the CUDA code) and the CUDA path through the C ABI (`-m gpu`,
pins the product to the same analytic answers, independently of the oracle).

  * internal-wave packet, error < 1e-4 after 10 steps        test/test_internal_wave_dynamics.jl:1-74, test_dynamics.jl:625-681
  * inertial oscillations with FPlane                        test/test_dynamics.jl:354-396 (z rotation; other axes are out of scope)
  * stratified fluid at rest under tilted gravity            test/test_dynamics.jl:260-352 (buoyancy tracer and temperature tracer)
  * budgets under isotropic / vertical / horizontal diffusion test/test_dynamics.jl:31-60, 410-455
  * operator known answers (oracle operators)                test/test_operators.jl:9-108,115-196
  * the grid constructor examples of the docstring           src/Grids/rectilinear_grid.jl:154-240
"""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "clima-oceananigans.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle as O  # noqa: E402
from oracle.grids import Center as C, Face as F  # noqa: E402

LOCS = {"u": (F, C, C), "v": (C, F, C), "w": (C, C, F)}


class OracleBackend:
    name = "oracle"
    M = O

    def grid(self, FT=np.float64, **kw):
        return O.RectilinearGrid(FT, **kw)

    def nodes(self, model, name):
        return model.grid.nodes(LOCS.get(name, (C, C, C)))

    def set(self, model, **vals):
        model.set(**vals)

    def step(self, model, dt):
        model.time_step(dt)

    def interior(self, model, name):
        return np.array(model.fields[name].interior)

    def time(self, model):
        return model.clock.time


class CudaBackend:
    name = "cuda"

    def __init__(self):
        import ocean_b200 as ob
        self.M = ob
        self.arch = ob.B200(0)

    def grid(self, FT=np.float64, **kw):
        return self.M.RectilinearGrid(self.arch, FT, **kw)

    def nodes(self, model, name):
        loc = {"u": ("Face", "Center", "Center"), "v": ("Center", "Face", "Center"), "w": ("Center", "Center", "Face")}.get(
            name, ("Center",) * 3)
        return model.grid.nodes(loc)

    def set(self, model, **vals):
        self.M.set_model(model, **vals)

    def step(self, model, dt):
        self.M.time_step(model, dt)

    def interior(self, model, name):
        return model.fields[name].interior()

    def time(self, model):
        return model.clock.time


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def be(request):
    return OracleBackend() if request.param == "oracle" else CudaBackend()


def evaluate(be, model, name, fn):
    x, y, z = be.nodes(model, name)
    shape = np.broadcast_shapes(np.shape(x), np.shape(y), np.shape(z))
    return np.broadcast_to(fn(np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), np.asarray(z, dtype=np.float64)),
                           shape).astype(np.float64).copy()


# ---- internal wave ------------------------------------------------------------------------------------------------
def internal_wave_solution(L):
    """test/test_internal_wave_dynamics.jl:1-55"""
    nu = kap = 1e-9
    z0, delta, a0, m, k, f, NN = -L / 3, L / 20, 1e-3, 16, 1, 0.2, 1.0
    sig = math.sqrt((NN ** 2 * k ** 2 + f ** 2 * m ** 2) / (k ** 2 + m ** 2))
    dt = 0.01 / sig
    cg = m * sig / (k ** 2 + m ** 2) * (f ** 2 / sig ** 2 - 1)
    U = a0 * k * sig / (sig ** 2 - f ** 2)
    V = a0 * k * f / (sig ** 2 - f ** 2)
    W = a0 * m * sig / (sig ** 2 - NN ** 2)
    B = a0 * m * NN ** 2 / (sig ** 2 - NN ** 2)
    a = lambda x, y, z, t: np.exp(-(z - cg * t - z0) ** 2 / (2 * delta) ** 2)
    sol = dict(u=lambda x, y, z, t: a(x, y, z, t) * U * np.cos(k * x + m * z - sig * t) + 0 * y,
               v=lambda x, y, z, t: a(x, y, z, t) * V * np.sin(k * x + m * z - sig * t) + 0 * y,
               w=lambda x, y, z, t: a(x, y, z, t) * W * np.cos(k * x + m * z - sig * t) + 0 * y,
               b=lambda x, y, z, t: a(x, y, z, t) * B * np.sin(k * x + m * z - sig * t) + NN ** 2 * z + 0 * y)
    return sol, dict(nu=nu, kappa=kap, f=f), dt


@pytest.mark.parametrize("variant", ["y_periodic_regular", "y_flat_regular", "y_periodic_stretched_faces"])
def test_internal_wave_packet(be, variant):
    M = be.M
    N, L = 128, 2 * math.pi
    sol, par, dt = internal_wave_solution(L)
    zf = np.linspace(-L, 0, N + 1)
    if variant == "y_periodic_regular":
        g = be.grid(size=(N, 1, N), x=(0, L), y=(0, L), z=(-L, 0), topology=("Periodic", "Periodic", "Bounded"))
    elif variant == "y_flat_regular":
        g = be.grid(size=(N, N), x=(0, L), z=(-L, 0), topology=("Periodic", "Flat", "Bounded"))
    else:   # the same faces passed as an array: the stretched code path and the Fourier-tridiagonal solver
        g = be.grid(size=(N, 1, N), x=(0, L), y=(0, L), z=zf, topology=("Periodic", "Periodic", "Bounded"))
    m = M.NonhydrostaticModel(g, advection=M.CenteredSecondOrder(), tracers=("b",), buoyancy=M.BuoyancyTracer(),
                              coriolis=M.FPlane(par["f"]), closure=M.ScalarDiffusivity("ThreeDimensional", ν=par["nu"], κ=par["kappa"]))
    be.set(m, **{n: evaluate(be, m, n, lambda x, y, z, n=n: sol[n](x, y, z, 0.0)) for n in ("u", "v", "w", "b")})
    for _ in range(10):
        be.step(m, dt)
    t = be.time(m)
    u_num = be.interior(m, "u")
    u_ans = evaluate(be, m, "u", lambda x, y, z: sol["u"](x, y, z, t))
    err = np.mean((u_num - u_ans) ** 2) / np.mean(u_ans ** 2)
    assert err < 1e-4, err                       # the reference's tolerance


# ---- inertial oscillations ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ts", ["RungeKutta3", "QuasiAdamsBashforth2"])
def test_inertial_oscillations(be, ts):
    """uniform flow on an f-plane turns in inertial circles: (u, v)(t) = (cos f t, -sin f t), |U| = 1, w = 0"""
    M = be.M
    g = be.grid(size=(4, 4, 4), extent=(1, 1, 1), topology=("Periodic",) * 3)
    m = M.NonhydrostaticModel(g, advection=M.CenteredSecondOrder(), coriolis=M.FPlane(1.0), timestepper=ts)
    be.set(m, u=np.ones((4, 4, 4)))
    dt, nsteps = 1e-2, 157                       # ~ a quarter of the inertial period
    for _ in range(nsteps):
        be.step(m, dt)
    t = be.time(m)
    u, v, w = (be.interior(m, n) for n in "uvw")
    assert np.all(w == 0)
    assert np.allclose(np.sqrt(u ** 2 + v ** 2), 1, atol=(1e-6 if ts == "RungeKutta3" else 2e-3))
    tol = 1e-6 if ts == "RungeKutta3" else 2e-3
    assert np.allclose(u, math.cos(t), atol=tol) and np.allclose(v, -math.sin(t), atol=tol)


# ---- stratified fluid at rest under tilted gravity ---------------------------------------------------------------------
@pytest.mark.parametrize("tracer", ["b", "T"])
def test_stratified_fluid_remains_at_rest_with_tilted_gravity(be, tracer):
    M = be.M
    N, L, theta, N2 = 16, 2000.0, 60.0, 1e-5
    g = be.grid(size=(1, N, N), extent=(L, L, L), topology=("Periodic", "Bounded", "Bounded"))
    gt = (0.0, math.sin(math.radians(theta)), math.cos(math.radians(theta)))
    if tracer == "b":
        buoy, tracers, grad = M.Buoyancy(M.BuoyancyTracer(), gt), ("b",), N2
    else:
        sw = M.SeawaterBuoyancy()
        buoy, tracers = M.Buoyancy(sw, gt), ("T", "S")
        grad = N2 / (sw.gravitational_acceleration * sw.equation_of_state.thermal_expansion)
    ybc, zbc = M.BoundaryCondition("Gradient", grad * gt[1]), M.BoundaryCondition("Gradient", grad * gt[2])
    bcs = {tracer: dict(bottom=zbc, top=zbc, south=ybc, north=ybc)}
    m = M.NonhydrostaticModel(g, advection=M.CenteredSecondOrder(), buoyancy=buoy, tracers=tracers, boundary_conditions=bcs)
    be.set(m, **{tracer: evaluate(be, m, tracer, lambda x, y, z: grad * (x * gt[0] + y * gt[1] + z * gt[2]))})
    for _ in range(6):
        be.step(m, 600.0)
    c = be.interior(m, tracer)
    dy = dz = L / N
    assert np.allclose(np.diff(c, axis=1) / dy, grad * gt[1], rtol=1.5e-8, atol=0)
    assert np.allclose(np.diff(c, axis=2) / dz, grad * gt[2], rtol=1.5e-8, atol=0)
    for n in "uvw":
        assert np.max(np.abs(be.interior(m, n))) < 1e-8        # round-off of the pressure / buoyancy balance only


# ---- diffusion budgets ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ts", ["QuasiAdamsBashforth2", "RungeKutta3"])
@pytest.mark.parametrize("topo", [("Periodic",) * 3, ("Periodic", "Periodic", "Bounded"), ("Periodic", "Bounded", "Bounded"),
                                  ("Bounded",) * 3])
@pytest.mark.parametrize("form,td", [("ThreeDimensional", "Explicit"), ("Vertical", "Explicit"), ("Horizontal", "Explicit"),
                                     ("ThreeDimensional", "VerticallyImplicit"), ("Vertical", "VerticallyImplicit")])
def test_scalar_diffusivity_budget(be, ts, topo, form, td):
    """the mean of a diffusing field is conserved (no-flux walls, periodic wrap); explicit and vertically implicit time
    discretisations (the latter on vertically Bounded domains only, test_dynamics.jl:419-424)"""
    M = be.M
    if td == "VerticallyImplicit" and topo[2] == "Periodic":
        pytest.skip("VerticallyImplicitTimeDiscretization needs a Bounded z")
    g = be.grid(size=(4, 4, 4), extent=(1, 1, 1), topology=topo)
    names = ["c"] + [n for n, t in zip("uvw", topo) if t == "Periodic"]
    rng = np.random.default_rng(7)
    for name in names:
        m = M.NonhydrostaticModel(g, advection=M.CenteredSecondOrder(),
                                  closure=M.ScalarDiffusivity(form, ν=1.0, κ=1.0, time_discretization=td),
                                  tracers=("c",), timestepper=ts)
        shape = be.interior(m, name).shape
        f0 = rng.uniform(0, 1, shape)
        if name in "uvw":
            # a velocity component with a non-zero mean over a Periodic direction is divergence free only if uniform along
            # itself; the reference sets rand() and lets the projection act -- the budget is checked after that
            pass
        be.set(m, **{name: f0})
        init = float(np.mean(be.interior(m, name)))
        dt = 1e-4 * 0.25 ** 2 / 1.0
        for _ in range(10):
            be.step(m, dt)
        final = float(np.mean(be.interior(m, name)))
        assert math.isclose(init, final, rel_tol=1.5e-8, abs_tol=1e-14), (name, init, final)


# ---- operator known answers (oracle operators; the CUDA operators are exercised through the tendencies) ---------------------
def test_grid_lengths_areas_volumes():
    """test/test_operators.jl:115-177"""
    g = O.RectilinearGrid(np.float64, size=(1, 1, 1), extent=(math.pi, 2 * math.pi, 3 * math.pi), topology=("Periodic", "Periodic", "Bounded"))
    R = O.R
    one = R(1)
    for lx in (C, F):
        for ly in (C, F):
            for lz in (C, F):
                assert g.Δx(lx, one) == np.float64(math.pi)
                assert g.Δy(ly, one) == np.float64(2 * math.pi)
                assert g.Δz(lz, one) == np.float64(3 * math.pi)
                assert np.isclose(g.Ax(one, one, one, lx, ly, lz), 6 * math.pi ** 2, rtol=1e-15)
                assert np.isclose(g.Ay(one, one, one, lx, ly, lz), 3 * math.pi ** 2, rtol=1e-15)
                assert np.isclose(g.Az(one, one, one, lx, ly, lz), 2 * math.pi ** 2, rtol=1e-15)
                assert np.isclose(g.V(one, one, one, lx, ly, lz), 6 * math.pi ** 3, rtol=1e-15)


def test_function_differentiation_and_interpolation():
    """test/test_operators.jl:9-108: derivatives and interpolations of f(i, j, k) = phi[i, j, k]^2 at (2, 2, 2), regular and
    stretched grids"""
    from oracle.operators import deriv, INTERP
    R = O.R
    rng = np.random.default_rng(3)
    phi = rng.uniform(0, 1, (3, 3, 3))
    phi2 = phi ** 2
    P = lambda i, j, k: phi2[i - 1, j - 1, k - 1]
    two = R(2)

    def f(i, j, k, grid):
        # i, j, k are index ranges R: evaluate phi^2 on them (1-based)
        return phi2[i.lo - 1:i.hi, j.lo - 1:j.hi, k.lo - 1:k.hi]
    for stretched in (False, True):
        if not stretched:
            g = O.RectilinearGrid(np.float64, size=(3, 3, 3), extent=(3, 3, 3), topology=("Periodic", "Periodic", "Bounded"))
            dc = lambda i: 1.0
            df = lambda i: 1.0
        else:
            sf = np.array([0.0, 1.0, 3.0, 6.0])
            sc = {0: -0.5, 1: 0.5, 2: 2.0, 3: 4.5, 4: 7.5}
            g = O.RectilinearGrid(np.float64, size=(3, 3, 3), x=sf, y=sf, z=sf, topology=("Bounded",) * 3)
            dc = lambda i: sf[i] - sf[i - 1]
            df = lambda i: sc[i] - sc[i - 1]
        sh = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
        for d in range(3):
            e = sh[d]
            want_c = (P(2 + e[0], 2 + e[1], 2 + e[2]) - P(2, 2, 2)) / dc(2)
            want_f = (P(2, 2, 2) - P(2 - e[0], 2 - e[1], 2 - e[2])) / df(2)
            for others in ((C, C), (C, F), (F, C), (F, F)):
                for lc, want in ((C, want_c), (F, want_f)):
                    loc = list(others)
                    loc.insert(d, lc)
                    got = deriv(d, *loc)(two, two, two, g, f)
                    assert np.asarray(got).ravel()[0] == want, (stretched, d, loc)
            if not stretched:
                ic = INTERP[C][d](two, two, two, g, f)
                iff = INTERP[F][d](two, two, two, g, f)
                assert np.asarray(ic).ravel()[0] == (P(2 + e[0], 2 + e[1], 2 + e[2]) + P(2, 2, 2)) / 2
                assert np.asarray(iff).ravel()[0] == (P(2, 2, 2) + P(2 - e[0], 2 - e[1], 2 - e[2])) / 2


# ---- the grid constructor examples of the reference's docstring --------------------------------------------------------------
def _grid_summary(g, Face_, Bounded_):
    """what `show(grid)` prints: extents of the face nodes and the spacings (rectilinear_grid.jl:154-240)"""
    out = []
    for d in range(3):
        if g.topology[d] == "Flat" or str(g.topology[d]).endswith("Flat"):
            out.append(None)
            continue
        Fd = g.nodesF[d] if hasattr(g, "nodesF") else g.F[d]
        dC = g.dC[d]
        lo = float(Fd[1])
        hi = float(Fd[g.N[d] + 1])
        if g.regular[d]:
            out.append((lo, hi, float(dC), float(dC)))
        else:
            a = np.asarray(dC.slice(1, g.N[d]) if hasattr(dC, "slice") else dC.span(1, g.N[d]))
            out.append((lo, hi, float(a.min()), float(a.max())))
    return out


@pytest.mark.parametrize("which", ["oracle", "host"])
def test_grid_docstring_examples(which):
    """RectilinearGrid docstring examples: the numbers `show` prints (6 significant digits)"""
    if which == "oracle":
        mk = lambda FT=np.float64, **kw: O.RectilinearGrid(FT, **kw)
    else:
        pytest.importorskip("ocean_b200")
        import ocean_b200 as ob
        try:
            arch = ob.B200(0)
        except Exception:
            pytest.skip("the host grid mirror creates a device handle: needs a GPU")
        mk = lambda FT=np.float64, **kw: ob.RectilinearGrid(arch, FT, **kw)
    close = lambda a, b: math.isclose(a, b, rel_tol=5e-6, abs_tol=1e-15)     # `show` prints 6 significant digits
    ppb = ("Periodic", "Periodic", "Bounded")
    s = _grid_summary(mk(size=(32, 32, 32), extent=(1, 2, 3), topology=ppb), None, None)
    assert s[0][:3] == (0.0, 1.0, 0.03125) and s[1][:3] == (0.0, 2.0, 0.0625) and s[2][:3] == (-3.0, 0.0, 0.09375)
    s = _grid_summary(mk(np.float32, size=(32, 32, 16), x=(0, 8), y=(-10, 10), z=(-math.pi, math.pi), topology=ppb), None, None)
    assert s[0][:3] == (0.0, 8.0, 0.25) and s[1][:3] == (-10.0, 10.0, 0.625)
    assert close(s[2][0], -3.14159) and close(s[2][1], 3.14159) and close(s[2][2], 0.392699)
    s = _grid_summary(mk(size=(32, 32), extent=(2 * math.pi, 4 * math.pi), topology=("Periodic", "Periodic", "Flat")), None, None)
    assert close(s[0][1], 6.28319) and close(s[0][2], 0.19635) and close(s[1][1], 12.5664) and close(s[1][2], 0.392699) and s[2] is None
    assert abs(s[0][0]) < 1e-15 and abs(s[1][0]) < 1e-15          # the reference prints 3.6e-17 / 7.2e-17: rounding of the range
    s = _grid_summary(mk(size=256, z=(-128, 0), topology=("Flat", "Flat", "Bounded")), None, None)
    assert s[0] is None and s[1] is None and s[2][:3] == (-128.0, 0.0, 0.5)
    sig, Nz, Lz = 1.1, 24, 32
    hyper = lambda k: -Lz * (1 - math.tanh(sig * (k - 1) / Nz) / math.tanh(sig))
    s = _grid_summary(mk(size=(32, 32, Nz), x=(0, 64), y=(0, 64), z=hyper, topology=ppb), None, None)
    assert s[0][:3] == (0.0, 64.0, 2.0) and s[2][0] == -32.0 and abs(s[2][1]) < 1e-14
    assert close(s[2][2], 0.682695) and close(s[2][3], 1.83091)
    Ny, Ly = 30, 100
    cheb = lambda j: -Ly / 2 * math.cos(math.pi * (j - 1) / Ny)
    s = _grid_summary(mk(size=(32, Ny, Nz), x=(0, 200), y=cheb, z=hyper, topology=("Periodic", "Bounded", "Bounded")), None, None)
    assert s[0][:3] == (0.0, 200.0, 6.25) and close(s[1][0], -50.0) and close(s[1][1], 50.0)
    assert close(s[1][2], 0.273905) and close(s[1][3], 5.22642)


# ---- vertically implicit diffusion: cosine decay with a time step far above the explicit limit ---------------------------------
@pytest.mark.parametrize("name", ["c", "u"])
def test_vertically_implicit_cosine_diffusion(be, name):
    """test_dynamics.jl:62-80 (test_diffusion_cosine) with VerticallyImplicitTimeDiscretization: cos(m z) decays as
    exp(-kappa m^2 t); backward Euler in time, so the check uses the discrete amplification factor 1 / (1 + dt kappa lam) with the
    second-order eigenvalue lam of the grid, which the scheme reproduces to round-off; dt = 20 x the explicit stability limit"""
    M = be.M
    N, kap = 32, 1.0
    g = be.grid(size=(4, 4, N), x=(0, 1), y=(0, 1), z=(-1, 0), topology=("Periodic", "Periodic", "Bounded"))
    m = M.NonhydrostaticModel(g, advection=None, closure=M.ScalarDiffusivity("Vertical", ν=kap, κ=kap, time_discretization="VerticallyImplicit"),
                              tracers=("c",), timestepper="QuasiAdamsBashforth2")
    mz = 2 * math.pi
    prof = lambda x, y, z: np.cos(mz * z) + 0 * x + 0 * y
    be.set(m, **{name: evaluate(be, m, name, prof)})
    dz = 1.0 / N
    dt = 20 * dz * dz / (2 * kap)
    lam = (2 * math.sin(mz * dz / 2) / dz) ** 2
    nsteps = 5
    for _ in range(nsteps):
        be.step(m, dt)
    want = evaluate(be, m, name, prof) / (1 + dt * kap * lam) ** nsteps
    got = be.interior(m, name)
    assert np.max(np.abs(got - want)) < 1e-12
