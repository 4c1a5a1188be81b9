"""Host-side mirror of the reference's operator interface for the NonhydrostaticModel path on the
`B200()` architecture: Field, fill_halo_regions!, FFTBasedPoissonSolver /
FourierTridiagonalPoissonSolver / BatchedTridiagonalSolver + solve!, NonhydrostaticModel, set!,
update_state!, time_step!.  Names, argument meaning and error behaviour follow the reference
(files cited per class); all arithmetic happens in libocean_b200.so."""
import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import lib, check, B200Error
from .grids import RectilinearGrid, Periodic, Bounded, Flat, Center, Face

_LOC = {Center: L.CENTER, Face: L.FACE}
_BCK = {"Periodic": L.BC_PERIODIC, "Flux": L.BC_FLUX, "Value": L.BC_VALUE, "Gradient": L.BC_GRADIENT,
        "Open": L.BC_OPEN, None: L.BC_NONE}


# ---- boundary conditions (src/BoundaryConditions/boundary_condition.jl) ----------------------
class BoundaryCondition:
    def __init__(self, kind, condition=None):
        if callable(condition):
            raise ValueError("function-valued boundary conditions cannot cross the C ABI of the B200 "
                             "architecture (SURVEY.md 8(b)); use constants")
        self.kind, self.condition = kind, condition


def FluxBoundaryCondition(v):
    return BoundaryCondition("Flux", v)


def ValueBoundaryCondition(v):
    return BoundaryCondition("Value", v)


def GradientBoundaryCondition(v):
    return BoundaryCondition("Gradient", v)


_SIDES = ("west", "east", "south", "north", "bottom", "top")


def _default_bc(topo, loc, auxiliary=False):
    """default_prognostic_bc / default_auxiliary_bc (field_boundary_conditions.jl:13-34)."""
    if topo in (Periodic, "FullyConnected"):
        return BoundaryCondition("Periodic")
    if topo == Flat:
        return BoundaryCondition(None)
    if loc == Center:
        return BoundaryCondition("Flux", None)
    return BoundaryCondition(None) if auxiliary else BoundaryCondition("Open", None)


def _bc_array(grid, loc, user=None, auxiliary=False):
    arr = (L.BC * 6)()
    user = user or {}
    for s, name in enumerate(_SIDES):
        bc = user.get(name) or _default_bc(grid.topology[s // 2], loc[s // 2], auxiliary)
        arr[s].kind = _BCK[bc.kind]
        arr[s].value = 0.0 if bc.condition is None else float(bc.condition)
    return arr


# ---- Field (src/Fields/field.jl:16-31) ---------------------------------------------------------
class Field:
    def __init__(self, loc, grid, boundary_conditions=None, _handle=None, auxiliary=False):
        self.grid, self.loc = grid, tuple(loc)
        self._owned = _handle is None
        if _handle is None:
            h = C.c_void_p()
            locs = (C.c_int32 * 3)(*[_LOC[l] for l in self.loc])
            check(lib.ob200_field_create(grid.handle, locs, _bc_array(grid, self.loc, boundary_conditions, auxiliary),
                                         C.byref(h)))
            _handle = h
        self.handle = _handle
        ps = (C.c_int32 * 3)()
        check(lib.ob200_field_parent_size(self.handle, ps))
        self.parent_size = tuple(ps)

    def size(self):
        g = self.grid
        return tuple(g.N[d] + (1 if (self.loc[d] == Face and g.topology[d] == Bounded) else 0) for d in range(3))

    def parent(self):
        """Array(parent(field)) -- reference layout incl. halos, column-major."""
        a = np.zeros(self.parent_size, dtype=self.grid.FT, order="F")
        check(lib.ob200_field_get_parent(self.handle, a.ctypes.data_as(C.c_void_p)))
        return a

    def set_parent(self, a):
        a = np.asfortranarray(a, dtype=self.grid.FT)
        if a.shape != self.parent_size:
            raise ValueError(f"parent array has shape {a.shape}, expected {self.parent_size}")
        check(lib.ob200_field_set_parent(self.handle, a.ctypes.data_as(C.c_void_p)))

    def _islice(self):
        n, H = self.size(), self.grid.H
        return tuple(slice(H[d], H[d] + n[d]) for d in range(3))

    def interior(self):
        return self.parent()[self._islice()]

    def set(self, value):
        """set!(field, value) (src/Fields/set!.jl:20-65): array or function of (x, y, z)."""
        n = self.size()
        if callable(value):
            x, y, z = self.grid.nodes(self.loc)
            value = value(x, y, z) + np.zeros(n)
        p = self.parent()
        p[self._islice()] = np.asarray(value, dtype=self.grid.FT).reshape(n)
        self.set_parent(p)

    def slice(self, lo, hi, out=None, sync=True):
        """dense copy of the index box lo..hi (Julia indices, 1-based, inclusive; halo indices allowed), gathered on the device:
        only the box travels to the host (fetch_output with a FieldSlicer, OutputWriters/fetch_output.jl:24-36).  With
        sync=False the array is valid after sync() / ob200_sync_downloads."""
        shape = tuple(int(h) - int(l) + 1 for l, h in zip(lo, hi))
        if out is None:
            out = np.empty(shape, dtype=self.grid.FT, order="F")
        check(lib.ob200_field_slice_async(self.handle, (C.c_int32 * 3)(*[int(x) for x in lo]),
                                          (C.c_int32 * 3)(*[int(x) for x in hi]), out.ctypes.data_as(C.c_void_p)))
        if sync:
            check(lib.ob200_sync())
        return out

    def average(self, dims, out=None, sync=True):
        """mean(field, dims=...) over the interior, reduced on the device (AveragedField); dims: 1-based dimension numbers as in
        the reference, e.g. (1, 2) for a horizontal average.  Averaged dimensions keep extent 1."""
        flags = [1 if (d + 1) in tuple(dims) else 0 for d in range(3)]
        n = self.size()
        shape = tuple(1 if flags[d] else n[d] for d in range(3))
        if out is None:
            out = np.empty(shape, dtype=self.grid.FT, order="F")
        check(lib.ob200_field_average_async(self.handle, (C.c_int32 * 3)(*flags), out.ctypes.data_as(C.c_void_p)))
        if sync:
            check(lib.ob200_sync())
        return out

    def reduce(self):
        s, s2, mx, nan = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
        check(lib.ob200_field_reduce(self.handle, C.byref(s), C.byref(s2), C.byref(mx), C.byref(nan)))
        return dict(sum=s.value, sumsq=s2.value, maxabs=mx.value, has_nan=bool(nan.value))

    def __del__(self):
        try:
            if self._owned and self.handle:
                lib.ob200_field_destroy(self.handle)
        except Exception:
            pass


def CenterField(grid, boundary_conditions=None):
    return Field((Center, Center, Center), grid, boundary_conditions)


def XFaceField(grid, boundary_conditions=None):
    return Field((Face, Center, Center), grid, boundary_conditions)


def YFaceField(grid, boundary_conditions=None):
    return Field((Center, Face, Center), grid, boundary_conditions)


def ZFaceField(grid, boundary_conditions=None):
    return Field((Center, Center, Face), grid, boundary_conditions)


def fill_halo_regions(fields):
    """fill_halo_regions!(fields) (src/BoundaryConditions/fill_halo_regions.jl:34-82,
    src/Fields/field_tuples.jl:51-77)."""
    if isinstance(fields, Field):
        fields = [fields]
    fields = list(fields.values()) if isinstance(fields, dict) else list(fields)
    arr = (C.c_void_p * len(fields))(*[f.handle for f in fields])
    check(lib.ob200_fill_halo_regions(arr, len(fields)))


# ---- solvers (src/Solvers) -----------------------------------------------------------------------
class _PoissonSolver:
    KIND = L.SOLVER_AUTO

    def __init__(self, grid):
        self.grid = grid
        h = C.c_void_p()
        check(lib.ob200_poisson_create(grid.handle, self.KIND, C.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            lib.ob200_poisson_destroy(self.handle)
        except Exception:
            pass


class FFTBasedPoissonSolver(_PoissonSolver):
    """src/Solvers/fft_based_poisson_solver.jl:50-72."""
    KIND = L.SOLVER_FFT


class FourierTridiagonalPoissonSolver(_PoissonSolver):
    """src/Solvers/fourier_tridiagonal_poisson_solver.jl:30-72."""
    KIND = L.SOLVER_FT


def solve(phi, solver, rhs):
    """solve!(ϕ, solver, rhs): rhs is the real source term, an (Nx, Ny, Nz) array
    (fft_based_poisson_solver.jl:93-120; fourier_tridiagonal_poisson_solver.jl:74-118)."""
    g = solver.grid
    rhs = np.asfortranarray(rhs, dtype=g.FT)
    if rhs.shape != g.N:
        raise ValueError("rhs must have the grid's interior size")
    check(lib.ob200_poisson_solve(solver.handle, phi.handle, rhs.ctypes.data_as(C.c_void_p)))
    return phi


def solve_for_pressure(pressure, solver, dt, U):
    """solve_for_pressure!(pressure, solver, Δt, U★) (solve_for_pressure.jl:55-89)."""
    check(lib.ob200_solve_for_pressure(solver.handle, pressure.handle, float(dt), U["u"].handle, U["v"].handle,
                                       U["w"].handle))


class BatchedTridiagonalSolver:
    """src/Solvers/batched_tridiagonal_solver.jl:10-72 (array coefficients)."""

    def __init__(self, grid, lower_diagonal, diagonal, upper_diagonal):
        self.grid = grid
        Nx, Ny, Nz = grid.N
        self.a = np.ascontiguousarray(lower_diagonal, dtype=np.float64)
        self.c = np.ascontiguousarray(upper_diagonal, dtype=np.float64)
        b = np.asarray(diagonal, dtype=np.float64)
        if b.ndim == 1:
            b = np.broadcast_to(b.reshape(1, 1, Nz), (Nx, Ny, Nz))
        self.b = np.asfortranarray(b)

    def solve(self, rhs):
        g = self.grid
        Nx, Ny, Nz = g.N
        cplx = np.iscomplexobj(rhs)
        dt = (np.complex64 if g.FT == np.float32 else np.complex128) if cplx else g.FT
        rhs = np.asfortranarray(np.broadcast_to(rhs, (Nx, Ny, Nz)) if np.ndim(rhs) == 3 else
                                np.broadcast_to(np.asarray(rhs).reshape(1, 1, Nz), (Nx, Ny, Nz)), dtype=dt)
        out = np.zeros((Nx, Ny, Nz), dtype=dt, order="F")
        P = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib.ob200_batched_tridiagonal_solve(L.F32 if g.FT == np.float32 else L.F64, int(cplx), Nx, Ny, Nz,
                                                  P(self.a), P(self.b), P(self.c), P(rhs), P(out)))
        return out


# ---- model components -----------------------------------------------------------------------------
class _Adv:
    name = "none"
    required_halo = 1


class CenteredSecondOrder(_Adv):
    name, required_halo = "CenteredSecondOrder", 1


class CenteredFourthOrder(_Adv):
    name, required_halo = "CenteredFourthOrder", 2


class UpwindBiasedFirstOrder(_Adv):
    name, required_halo = "UpwindBiasedFirstOrder", 2


class UpwindBiasedThirdOrder(_Adv):
    name, required_halo = "UpwindBiasedThirdOrder", 2


class UpwindBiasedFifthOrder(_Adv):
    name, required_halo = "UpwindBiasedFifthOrder", 3


def _eno_weights(r, x, i):
    """interp_weights(r, coord, i, 0, -) (src/Advection/weno_fifth_order.jl:740-772)."""
    out = []
    for j in range(3):
        c = 0
        for m in range(j + 1, 4):
            num = 0
            for l in range(4):
                if l == m:
                    continue
                pr = 1
                for q in range(4):
                    if q != m and q != l:
                        pr *= x[i] - x[i - (r - q + 1)]
                num += pr
            den = 1
            for l in range(4):
                if l != m:
                    den *= x[i - (r - m + 1)] - x[i - (r - l + 1)]
            c += num / den
        out.append(c * (x[i - (r - j)] - x[i - (r - j + 1)]))
    return out


class WENO5(_Adv):
    """WENO5(FT; grid=nothing, zweno=true) (src/Advection/weno_fifth_order.jl:164-180).  With
    `grid`, stretched dimensions get the ENO coefficient tables of :182-209,562-584."""
    name, required_halo = "WENO5", 3

    def __init__(self, FT=None, grid=None, zweno=True, stretched_smoothness=False):
        if stretched_smoothness:
            raise ValueError("stretched_smoothness=true is outside the B200 path's supported set")
        self.FT = grid.FT if grid is not None else (np.dtype(FT).type if FT is not None else np.float64)
        self.zweno = bool(zweno)
        self.tables = {}
        if grid is not None:
            g4 = grid.with_halo((4, 4, 4))
            for d in range(3):
                if g4.regular[d]:
                    continue
                for li, nodes in ((0, g4.F[d]), (1, g4.C[d])):
                    n2 = g4.N[d] + 2
                    t = np.zeros((4, n2, 3))
                    for ri, r in enumerate((-1, 0, 1, 2)):
                        for i in range(n2):
                            t[ri, i] = [float(self.FT(w)) for w in _eno_weights(r, nodes, i)]
                    self.tables[(d, li)] = np.ascontiguousarray(t)


class ScalarDiffusivity:
    """ScalarDiffusivity(formulation; ν, κ) (scalar_diffusivity.jl:60-76), constants only."""
    required_halo = 1

    def __init__(self, formulation="ThreeDimensional", ν=0.0, κ=0.0, nu=None, kappa=None, time_discretization="Explicit"):
        ν = nu if nu is not None else ν
        κ = kappa if kappa is not None else κ
        if time_discretization not in ("Explicit", "VerticallyImplicit"):
            raise ValueError("time_discretization must be 'Explicit' or 'VerticallyImplicit'")
        if time_discretization == "VerticallyImplicit" and formulation == "Horizontal":
            raise ValueError("VerticallyImplicitTimeDiscretization is not supported for HorizontalScalarDiffusivity")
        self.time_discretization = time_discretization
        if callable(ν) or callable(κ) or isinstance(ν, np.ndarray):
            raise ValueError("only constant ν, κ are supported on the B200 architecture")
        if formulation not in L.CLOSURE:
            raise ValueError(f"unsupported formulation {formulation}")
        self.formulation, self.ν, self.κ = formulation, ν, κ


class SmagorinskyLilly:
    """SmagorinskyLilly(FT; C=0.16, Cb=1.0, Pr=1.0) (src/TurbulenceClosures/turbulence_closure_implementations/
    smagorinsky_lilly.jl:6-72), explicit time discretisation; Pr is a number or a dict {tracer name: number}."""
    required_halo = 1
    formulation = "ThreeDimensional"

    def __init__(self, C=0.16, Cb=1.0, Pr=1.0):
        self.C, self.Cb, self.Pr = C, Cb, Pr

    def prandtl(self, name):
        return self.Pr[name] if isinstance(self.Pr, dict) else self.Pr


class AnisotropicMinimumDissipation:
    """AnisotropicMinimumDissipation(FT; C=1/12, Cν=nothing, Cκ=nothing, Cb=nothing) (src/TurbulenceClosures/
    turbulence_closure_implementations/anisotropic_minimum_dissipation.jl:96-105); constant Poincaré constants (Cκ a
    number or a dict per tracer), explicit time discretisation."""
    required_halo = 1
    formulation = "ThreeDimensional"

    def __init__(self, C=1 / 12, Cν=None, Cκ=None, Cb=None):
        self.Cν = C if Cν is None else Cν
        self.Cκ = C if Cκ is None else Cκ
        self.Cb = Cb
        for v in (self.Cν, self.Cκ):
            if callable(v):
                raise ValueError("only constant Poincaré constants are supported on the B200 architecture")

    def Ck(self, name):
        return self.Cκ[name] if isinstance(self.Cκ, dict) else self.Cκ


def VerticalScalarDiffusivity(**kw):
    return ScalarDiffusivity("Vertical", **kw)


def HorizontalScalarDiffusivity(**kw):
    return ScalarDiffusivity("Horizontal", **kw)


class FPlane:
    def __init__(self, f):
        self.f = f


class BuoyancyTracer:
    pass


class LinearEquationOfState:
    """LinearEquationOfState(FT; thermal_expansion=1.67e-4, haline_contraction=7.80e-4)
    (src/BuoyancyModels/linear_equation_of_state.jl:6-30)"""

    def __init__(self, thermal_expansion=1.67e-4, haline_contraction=7.80e-4):
        self.thermal_expansion, self.haline_contraction = thermal_expansion, haline_contraction


class SeawaterBuoyancy:
    """SeawaterBuoyancy(FT; gravitational_acceleration=g_Earth, equation_of_state=LinearEquationOfState(FT),
    constant_temperature=nothing, constant_salinity=nothing) (src/BuoyancyModels/seawater_buoyancy.jl:61-73).
    Only the linear equation of state crosses the C ABI."""
    g_Earth = 9.80665            # BuoyancyModels.jl:20

    def __init__(self, gravitational_acceleration=None, equation_of_state=None, constant_temperature=None,
                 constant_salinity=None):
        self.gravitational_acceleration = self.g_Earth if gravitational_acceleration is None else gravitational_acceleration
        self.equation_of_state = equation_of_state or LinearEquationOfState()
        if not isinstance(self.equation_of_state, LinearEquationOfState):
            raise ValueError("only LinearEquationOfState is supported on the B200 architecture")
        self.constant_temperature = 0.0 if constant_temperature is True else constant_temperature
        self.constant_salinity = 0.0 if constant_salinity is True else constant_salinity
        if self.constant_temperature is not None and self.constant_salinity is not None:
            raise ValueError("temperature and salinity cannot both be constant")

    def required_tracers(self):
        if self.constant_salinity is not None:
            return ("T",)
        if self.constant_temperature is not None:
            return ("S",)
        return ("T", "S")


class Buoyancy:
    def __init__(self, model=None, gravity_unit_vector=None):
        self.model = model or BuoyancyTracer()
        if not isinstance(self.model, (BuoyancyTracer, SeawaterBuoyancy)):
            raise ValueError("only BuoyancyTracer and SeawaterBuoyancy(LinearEquationOfState) are supported on the B200 architecture")
        self.g = gravity_unit_vector


class Clock:
    def __init__(self, model):
        self._m = model

    @property
    def time(self):
        t = C.c_double()
        check(lib.ob200_model_clock(self._m.handle, C.byref(t), None))
        return t.value

    @property
    def iteration(self):
        it = C.c_int64()
        check(lib.ob200_model_clock(self._m.handle, None, C.byref(it)))
        return it.value


class NonhydrostaticModel:
    """NonhydrostaticModel(; grid, advection, closure, coriolis, buoyancy, tracers, timestepper,
    boundary_conditions) (src/Models/NonhydrostaticModels/nonhydrostatic_model.jl:102-203).
    Defaults as in the reference: advection = CenteredSecondOrder(), timestepper =
    :QuasiAdamsBashforth2.  Unsupported pieces (forcings, function BCs, immersed boundaries, background
    fields, particles, closures other than ScalarDiffusivity / SmagorinskyLilly / AnisotropicMinimumDissipation)
    raise ArgumentError-like ValueErrors."""

    def __init__(self, grid, advection="default", closure=None, coriolis=None, buoyancy=None, tracers=(),
                 timestepper="QuasiAdamsBashforth2", boundary_conditions=None, forcing=None,
                 background_fields=None, particles=None, immersed_boundary=None, stokes_drift=None,
                 pressure_solver=None, chi=0.1):
        for name, v in (("forcing", forcing), ("background_fields", background_fields), ("particles", particles),
                        ("immersed_boundary", immersed_boundary), ("stokes_drift", stokes_drift)):
            if v:
                raise ValueError(f"`{name}` is not supported on the B200 architecture (SURVEY.md 8(b))")
        if not isinstance(grid, RectilinearGrid):
            raise ValueError("the B200 architecture supports RectilinearGrid only")
        if advection == "default":
            advection = CenteredSecondOrder()
        if timestepper not in L.TS:
            raise ValueError(f"unknown timestepper {timestepper}")
        if isinstance(buoyancy, (BuoyancyTracer, SeawaterBuoyancy)):
            buoyancy = Buoyancy(buoyancy)
        tracers = (tracers,) if isinstance(tracers, str) else tuple(tracers or ())
        if buoyancy is not None:          # validate_buoyancy (BuoyancyModels.jl:37-46)
            req = ("b",) if isinstance(buoyancy.model, BuoyancyTracer) else buoyancy.model.required_tracers()
            for n in req:
                if n not in tracers:
                    raise ValueError(f"{type(buoyancy.model).__name__} requires a tracer named :{n}")
        if len(tracers) > L.MAX_TRACERS:
            raise ValueError("too many tracers")
        if isinstance(advection, WENO5) and advection.FT != grid.FT:
            raise ValueError("WENO5 float type differs from the grid's; construct it as WENO5(grid.FT) or WENO5(grid=grid)")
        # halo inflation (nonhydrostatic_model.jl:140-148)
        H = list(grid.H)
        for term in (advection, closure):
            req = 1 if term is None else term.required_halo
            for d in range(3):
                H[d] = 0 if grid.topology[d] == Flat else max(req, H[d])
        if tuple(H) != grid.H:
            grid = grid.with_halo(H)
        self.grid, self.advection, self.closure, self.coriolis, self.buoyancy = grid, advection, closure, coriolis, buoyancy
        self.tracer_names, self.timestepper = tracers, timestepper
        d = L.ModelDesc()
        d.grid = grid.handle
        d.timestepper, d.chi = L.TS[timestepper], float(chi)
        d.advection = L.ADV[advection.name] if advection is not None else 0
        d.weno_zweno = int(getattr(advection, "zweno", True))
        self._keep = []
        if isinstance(advection, WENO5):
            for (dim, li), t in advection.tables.items():
                self._keep.append(t)
                d.weno_coeff[dim][li] = t.ctypes.data_as(C.POINTER(C.c_double))
        if isinstance(closure, AnisotropicMinimumDissipation):
            d.closure = L.CLOSURE["AnisotropicMinimumDissipation"]
            d.amd_Cnu = float(closure.Cν)
            d.amd_has_Cb, d.amd_Cb = int(closure.Cb is not None), float(closure.Cb or 0.0)
            for k, name in enumerate(tracers):
                d.amd_Ckappa[k] = float(closure.Ck(name))
        elif isinstance(closure, SmagorinskyLilly):
            d.closure = L.CLOSURE["SmagorinskyLilly"]
            d.smagorinsky_C, d.smagorinsky_Cb = float(closure.C), float(closure.Cb)
            for k, name in enumerate(tracers):
                d.prandtl[k] = float(closure.prandtl(name))
        else:
            d.closure = L.CLOSURE[closure.formulation] if closure is not None else 0
            if closure is not None:
                d.closure_vertically_implicit = int(closure.time_discretization == "VerticallyImplicit")
                d.nu = float(closure.ν)
                for k, name in enumerate(tracers):
                    d.kappa[k] = float(closure.κ[name] if isinstance(closure.κ, dict) else closure.κ)
        d.coriolis_fplane = int(coriolis is not None)
        if coriolis is not None:
            if not isinstance(coriolis, FPlane):
                raise ValueError("only FPlane is supported on the B200 architecture")
            d.f = float(coriolis.f)
        d.buoyancy_tracer, d.buoyancy_kind, d.temperature_tracer, d.salinity_tracer = -1, 0, -1, -1
        if buoyancy is not None and isinstance(buoyancy.model, SeawaterBuoyancy):
            sw = buoyancy.model
            d.buoyancy_kind = 1
            req = sw.required_tracers()
            d.temperature_tracer = tracers.index("T") if "T" in req else -1
            d.salinity_tracer = tracers.index("S") if "S" in req else -1
            d.gravitational_acceleration = float(sw.gravitational_acceleration)
            d.thermal_expansion = float(sw.equation_of_state.thermal_expansion)
            d.haline_contraction = float(sw.equation_of_state.haline_contraction)
        elif buoyancy is not None:
            d.buoyancy_tracer = tracers.index("b")
        d.gravity_tilted = int(buoyancy is not None and buoyancy.g is not None)
        g_hat = buoyancy.g if (buoyancy is not None and buoyancy.g is not None) else (0.0, 0.0, 1.0)
        for k in range(3):
            d.g_hat[k] = float(g_hat[k])
        d.ntracers = len(tracers)
        bcs = boundary_conditions or {}
        names = ("u", "v", "w") + tracers
        locs = [(Face, Center, Center), (Center, Face, Center), (Center, Center, Face)] + [(Center,) * 3] * len(tracers)
        for q, (n, loc) in enumerate(zip(names, locs)):
            arr = _bc_array(grid, loc, bcs.get(n))
            for s in range(6):
                d.bcs[q][s].kind, d.bcs[q][s].value = arr[s].kind, arr[s].value
        d.pressure_solver = L.SOLVER_AUTO if pressure_solver is None else pressure_solver
        h = C.c_void_p()
        check(lib.ob200_model_create(C.byref(d), C.byref(h)))
        self.handle = h
        self.names = names
        self.velocities = {n: self._field(n, n, locs[i]) for i, n in enumerate("uvw")}
        self.tracers = {n: self._field(f"c{k}", n, (Center,) * 3) for k, n in enumerate(tracers)}
        self.fields = {**self.velocities, **self.tracers}
        self.pressures = {"pNHS": self._field("pNHS", "pNHS", (Center,) * 3)}
        if grid.topology[2] != Flat:
            self.pressures["pHY′"] = self._field("pHY", "pHY", (Center,) * 3)
        self.diffusivity_fields = {}
        if isinstance(closure, (SmagorinskyLilly, AnisotropicMinimumDissipation)):
            self.diffusivity_fields["νₑ"] = self._field("nu_e", "nu_e", (Center,) * 3)
        if isinstance(closure, AnisotropicMinimumDissipation):
            self.diffusivity_fields["κₑ"] = {n: self._field(f"kappa_e{k}", f"kappa_e_{n}", (Center,) * 3) for k, n in enumerate(tracers)}
        self.Gn = {n: self._field("Gn_" + (n if n in "uvw" else f"c{tracers.index(n)}"), n, locs[i])
                   for i, n in enumerate(names)}
        self.Gm = {n: self._field("Gm_" + (n if n in "uvw" else f"c{tracers.index(n)}"), n, locs[i])
                   for i, n in enumerate(names)}
        self.clock = Clock(self)

    def _field(self, cname, name, loc):
        h = C.c_void_p()
        check(lib.ob200_model_field(self.handle, cname.encode(), C.byref(h)))
        return Field(loc, self.grid, _handle=h)

    def use_fast_kernels(self, on=True):
        check(lib.ob200_model_use_fast_kernels(self.handle, int(on)))

    def diagnostics(self):
        a, b = C.c_double(), C.c_double()
        check(lib.ob200_model_diagnostics(self.handle, C.byref(a), C.byref(b)))
        return dict(max_abs_div=a.value, kinetic_energy=b.value)

    def set_clock(self, time=0.0, iteration=0, previous_Δt=float("inf")):
        """model.clock.time / .iteration = ...; previous_Δt = inf makes the next QuasiAdamsBashforth2 step a forward-Euler step
        with G^- cleared, as for a new model (quasi_adams_bashforth_2.jl:76-81)"""
        check(lib.ob200_model_set_clock(self.handle, float(time), int(iteration), float(previous_Δt)))

    def destroy(self):
        """release the library model now (device memory, streams, events, tensor maps); also done by the finalizer"""
        if getattr(self, "handle", None) is not None:
            lib.ob200_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def update_state(model):
    """update_state!(model) (update_nonhydrostatic_model_state.jl:14-37)."""
    check(lib.ob200_model_update_state(model.handle))


def calculate_tendencies(model):
    """calculate_tendencies!(model) (calculate_nonhydrostatic_tendencies.jl:12-36)."""
    check(lib.ob200_model_calculate_tendencies(model.handle))


def set_model(model, enforce_incompressibility=True, **kw):
    """set!(model; enforce_incompressibility=true, kwargs...) (set_nonhydrostatic_model.jl:32-59)."""
    for name, value in kw.items():
        if name not in model.fields:
            raise ValueError(f"name {name} not found in model.velocities or model.tracers.")
        model.fields[name].set(value)
    update_state(model)
    if enforce_incompressibility:
        check(lib.ob200_model_pressure_project(model.handle, 1.0))
        update_state(model)


def time_step(model, dt, euler=False):
    """time_step!(model, Δt; euler=false) (runge_kutta_3.jl:81-152, quasi_adams_bashforth_2.jl:70-104).
    Enqueues the whole step on the stream; does not synchronise."""
    check(lib.ob200_model_time_step(model.handle, float(dt), int(euler)))


def sync():
    check(lib.ob200_sync())
