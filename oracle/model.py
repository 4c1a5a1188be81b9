"""
NonhydrostaticModel and its RK3 / AB2 time steppers (test infrastructure -- see
oracle/__init__.py).

Follows Models/NonhydrostaticModels/nonhydrostatic_model.jl:102-203 (constructor: halo
inflation, fields, solver choice, initial update_state!), calculate_nonhydrostatic_tendencies.jl:12-200,
nonhydrostatic_tendency_kernel_functions.jl:44-232 (term order), solve_for_pressure.jl:15-89,
pressure_correction.jl:10-56, update_hydrostatic_pressure.jl:10-40,
update_nonhydrostatic_model_state.jl:14-37, set_nonhydrostatic_model.jl:32-59,
TimeSteppers/runge_kutta_3.jl:48-218, quasi_adams_bashforth_2.jl:70-166, store_tendencies.jl:8-36,
clock.jl:48-60, Coriolis/f_plane.jl:42-44 and BuoyancyModels/{buoyancy.jl,buoyancy_tracer.jl,g_dot_b.jl}.
"""
import numpy as np

from .grids import Periodic, Bounded, Flat, Center as C, Face as F
from .fields import Field, FieldBoundaryConditions, fill_halo_regions, apply_flux_bcs, R
from .operators import deriv, div_ccc, Ixy_fca, Ixy_cfa, Izf
from .advection import div_Uu, div_Uc
from .closures import (div_τ, div_q, SmagorinskyLilly, smagorinsky_viscosity, AnisotropicMinimumDissipation,
                       amd_viscosity, amd_diffusivity, ivd_coefficients)
from .solvers import FFTBasedPoissonSolver, FourierTridiagonalPoissonSolver, BatchedTridiagonalSolver


class FPlane:
    def __init__(self, f):
        self.f = f


class BuoyancyTracer:
    pass


class LinearEquationOfState:
    """LinearEquationOfState(FT; thermal_expansion=1.67e-4, haline_contraction=7.80e-4) (linear_equation_of_state.jl:6-30)."""

    def __init__(self, thermal_expansion=1.67e-4, haline_contraction=7.80e-4):
        self.thermal_expansion, self.haline_contraction = thermal_expansion, haline_contraction


class SeawaterBuoyancy:
    """SeawaterBuoyancy(FT; gravitational_acceleration=g_Earth, equation_of_state=LinearEquationOfState(FT),
    constant_temperature=nothing, constant_salinity=nothing) (seawater_buoyancy.jl:61-73); only the linear equation
    of state.  `constant_temperature=True` becomes 0 as in :67-68; for a linear EOS the constant value is irrelevant."""
    g_Earth = 9.80665            # Oceananigans.jl:178 via BuoyancyModels: g_Earth

    def __init__(self, gravitational_acceleration=None, equation_of_state=None, constant_temperature=None,
                 constant_salinity=None):
        self.gravitational_acceleration = self.g_Earth if gravitational_acceleration is None else gravitational_acceleration
        self.equation_of_state = equation_of_state or LinearEquationOfState()
        self.constant_temperature = 0.0 if constant_temperature is True else constant_temperature
        self.constant_salinity = 0.0 if constant_salinity is True else constant_salinity
        assert not (self.constant_temperature is not None and self.constant_salinity is not None)

    def required_tracers(self):
        """seawater_buoyancy.jl:18-20"""
        if self.constant_salinity is not None:
            return ("T",)
        if self.constant_temperature is not None:
            return ("S",)
        return ("T", "S")


def buoyancy_perturbation(model, tracers, FT):
    """buoyancy_perturbation(i, j, k, grid, b, C) as a function (i, j, k, grid) -> array:
    BuoyancyTracer: C.b (buoyancy_tracer.jl:12); linear SeawaterBuoyancy (linear_equation_of_state.jl:69-77, evaluation order
    kept): g * (α T - β S), g * α * T (active temperature only), - g * β * S (active salinity only); parameters are FT."""
    if isinstance(model, BuoyancyTracer):
        b = tracers["b"]
        return lambda i, j, k, grid: b[i, j, k]
    g = FT(model.gravitational_acceleration)
    al, be = FT(model.equation_of_state.thermal_expansion), FT(model.equation_of_state.haline_contraction)
    req = model.required_tracers()
    if req == ("T", "S"):
        T, S = tracers["T"], tracers["S"]
        return lambda i, j, k, grid: g * (al * T[i, j, k] - be * S[i, j, k])
    if req == ("T",):
        T = tracers["T"]
        return lambda i, j, k, grid: g * al * T[i, j, k]
    S = tracers["S"]
    return lambda i, j, k, grid: -g * be * S[i, j, k]


class Buoyancy:
    """Buoyancy(model=BuoyancyTracer(), gravity_unit_vector=ZDirection()) (buoyancy.jl:3-42)."""

    def __init__(self, model=None, gravity_unit_vector=None):
        self.model = model if model is not None else BuoyancyTracer()
        self.g = gravity_unit_vector        # None = ZDirection


LOCS = {"u": (F, C, C), "v": (C, F, C), "w": (C, C, F)}


class Clock:
    def __init__(self):
        self.time = 0.0
        self.iteration = 0
        self.stage = 1


class NonhydrostaticModel:
    def __init__(self, grid, advection=None, closure=None, coriolis=None, buoyancy=None,
                 tracers=(), timestepper="QuasiAdamsBashforth2", boundary_conditions=None,
                 chi=0.1):
        # halo inflation, nonhydrostatic_model.jl:140-148 + Grids/automatic_halo_sizing.jl:27-36
        H = list(grid.H)
        for term in (advection, closure):
            req = 1 if term is None else term.required_halo
            for d in range(3):
                H[d] = 0 if grid.topology[d] == Flat else max(req, H[d])
        if tuple(H) != grid.H:
            grid = grid.with_halo(H)
        self.grid = grid
        FT = grid.FT
        self.advection, self.closure, self.coriolis = advection, closure, coriolis
        if isinstance(buoyancy, (BuoyancyTracer, SeawaterBuoyancy)):
            buoyancy = Buoyancy(buoyancy)              # regularize_buoyancy
        self.buoyancy = buoyancy
        self.tracer_names = tuple(tracers)
        if buoyancy is not None:                       # validate_buoyancy (BuoyancyModels.jl:37-46)
            req = ("b",) if isinstance(buoyancy.model, BuoyancyTracer) else buoyancy.model.required_tracers()
            assert all(n in self.tracer_names for n in req), f"buoyancy requires tracers {req}"
        bcs = boundary_conditions or {}

        def mk(name, loc):
            return Field(grid, loc, FieldBoundaryConditions(grid, loc, **bcs.get(name, {})))
        self.velocities = {n: mk(n, LOCS[n]) for n in "uvw"}
        self.tracers = {n: mk(n, (C, C, C)) for n in self.tracer_names}
        self.names = ("u", "v", "w") + self.tracer_names
        self.fields = {**self.velocities, **self.tracers}
        # PressureFields (Fields/field_tuples.jl:199-222): pHY′ absent when z is Flat
        self.pNHS = Field(grid, (C, C, C), auxiliary=True)
        self.pHY = None if grid.topology[2] == Flat else Field(grid, (C, C, C), auxiliary=True)
        # PressureSolver (NonhydrostaticModels.jl:18-27)
        if all(grid.regular):
            self.pressure_solver = FFTBasedPoissonSolver(grid)
        else:
            self.pressure_solver = FourierTridiagonalPoissonSolver(grid)
        # DiffusivityFields (smagorinsky_lilly.jl:205-221): νₑ at ccc with default boundary conditions
        les = isinstance(closure, (SmagorinskyLilly, AnisotropicMinimumDissipation))
        self.νe = Field(grid, (C, C, C), FieldBoundaryConditions(grid, (C, C, C))) if les else None
        # AMD: one eddy diffusivity field per tracer (anisotropic_minimum_dissipation.jl:378-389)
        self.κe = ({n: Field(grid, (C, C, C), FieldBoundaryConditions(grid, (C, C, C))) for n in self.tracer_names}
                   if isinstance(closure, AnisotropicMinimumDissipation) else None)
        self.timestepper = timestepper
        self.Gn = {n: Field(grid, self.fields[n].loc) for n in self.names}
        self.Gm = {n: Field(grid, self.fields[n].loc) for n in self.names}
        # RK3 coefficients stored as FT (runge_kutta_3.jl:57-66)
        self.γ1, self.γ2, self.γ3 = FT(8 / 15), FT(5 / 12), FT(3 / 4)
        self.ζ2, self.ζ3 = FT(-17 / 60), FT(-5 / 12)
        self.χ = FT(chi)
        self.previous_Δt = np.inf
        self.clock = Clock()
        self.update_state()

    # ---- helpers ---------------------------------------------------------------------
    def _box(self):
        g = self.grid
        return R(1, g.Nx), R(1, g.Ny), R(1, g.Nz)

    def set(self, enforce_incompressibility=True, **kw):
        """set!(model; kwargs...) set_nonhydrostatic_model.jl:32-59."""
        for name, value in kw.items():
            self.fields[name].set(value)
        self.update_state()
        if enforce_incompressibility:
            one = self.grid.FT(1)
            self.calculate_pressure_correction(one)
            self.pressure_correct_velocities(one)
            self.update_state()

    # ---- update_state! ---------------------------------------------------------------
    def update_state(self):
        fill_halo_regions([self.fields[n] for n in self.names])
        self.calculate_diffusivities()
        self.update_hydrostatic_pressure()
        if self.pHY is not None:
            fill_halo_regions(self.pHY)

    def calculate_diffusivities(self):
        """calculate_diffusivities! + fill_halo_regions!(diffusivity_fields) (update_nonhydrostatic_model_state.jl:29-31,
        smagorinsky_lilly.jl:109-127)"""
        if self.νe is None:
            return
        g = self.grid
        i, j, k = self._box()
        u, v, w = (self.velocities[n] for n in "uvw")
        dz_b = None
        if self.buoyancy is not None:
            mdl = self.buoyancy.model
            dz = deriv(2, C, C, F)
            if isinstance(mdl, BuoyancyTracer):
                b = self.tracers["b"]
                dz_b = lambda i, j, k, grid: dz(i, j, k, grid, b)                   # ∂z_b buoyancy_tracer.jl:16
            else:                                                                    # seawater_buoyancy.jl:166-171
                FT = g.FT
                gr, al, be = FT(mdl.gravitational_acceleration), FT(mdl.equation_of_state.thermal_expansion), FT(mdl.equation_of_state.haline_contraction)
                req = mdl.required_tracers()
                zero = lambda i, j, k, grid: FT(0)
                dT = (lambda i, j, k, grid: dz(i, j, k, grid, self.tracers["T"])) if "T" in req else zero
                dS = (lambda i, j, k, grid: dz(i, j, k, grid, self.tracers["S"])) if "S" in req else zero
                dz_b = lambda i, j, k, grid: gr * (al * dT(i, j, k, grid) - be * dS(i, j, k, grid))
        if isinstance(self.closure, AnisotropicMinimumDissipation):
            bp = buoyancy_perturbation(self.buoyancy.model, self.tracers, g.FT) if self.buoyancy is not None else None
            self.νe[i, j, k] = amd_viscosity(i, j, k, g, self.closure, bp, u, v, w)
            for n in self.tracer_names:
                self.κe[n][i, j, k] = amd_diffusivity(i, j, k, g, self.closure.Ck(n), u, v, w, self.tracers[n])
            fill_halo_regions([self.νe] + [self.κe[n] for n in self.tracer_names])
            return
        self.νe[i, j, k] = smagorinsky_viscosity(i, j, k, g, self.closure, dz_b, u, v, w)
        fill_halo_regions(self.νe)

    def update_hydrostatic_pressure(self):
        """_update_hydrostatic_pressure! update_hydrostatic_pressure.jl:10-18."""
        g = self.grid
        if g.topology[2] == Flat or self.pHY is None:
            return
        i, j, _ = self._box()
        Nz = g.Nz
        p = self.pHY
        if self.buoyancy is None:
            zb = lambda i, j, k, grid: grid.FT(0) * np.zeros((i.n, j.n, k.n), dtype=grid.FT)
        else:
            bp = buoyancy_perturbation(self.buoyancy.model, self.tracers, g.FT)
            gz = 1 if self.buoyancy.g is None else g.FT(self.buoyancy.g[2])
            if self.buoyancy.g is None:
                zb = bp
            else:
                zb = lambda i, j, k, grid: gz * bp(i, j, k, grid)
        k = R(Nz + 1)
        p[i, j, R(Nz)] = -Izf(i, j, k, g, zb) * g.Δz(F, k)
        for kk in range(Nz - 1, 0, -1):
            k = R(kk + 1)
            p[i, j, R(kk)] = p[i, j, R(kk + 1)] - Izf(i, j, k, g, zb) * g.Δz(F, k)

    # ---- tendencies --------------------------------------------------------------------
    def calculate_tendencies(self):
        g = self.grid
        FT = g.FT
        i, j, k = self._box()
        u, v, w = (self.velocities[n] for n in "uvw")
        U = (u, v, w)
        adv, clo, cor, buoy = self.advection, self.closure, self.coriolis, self.buoyancy
        pHY = self.pHY
        zero = FT(0)

        def gb(d):
            """x/y_dot_g_b (g_dot_b.jl:1-7): ĝ * b[i,j,k], zero for ZDirection."""
            if buoy is None or buoy.g is None:
                return 0
            return FT(buoy.g[d]) * buoyancy_perturbation(buoy.model, self.tracers, FT)(i, j, k, g)
        # u: nonhydrostatic_tendency_kernel_functions.jl:61-71
        corx = (-FT(cor.f) * Ixy_fca(i, j, k, g, v)) if cor is not None else zero     # x_f_cross_U
        px = deriv(0, F, C, C)(i, j, k, g, pHY) if pHY is not None else zero
        Gu = (- div_Uu(0, i, j, k, g, adv, U, u) - zero - zero
              - corx
              - px
              - div_τ(0, i, j, k, g, clo, u, v, w, self.νe)
              - zero + zero + zero
              + gb(0)
              + zero)
        # v: :118-128
        cory = (FT(cor.f) * Ixy_cfa(i, j, k, g, u)) if cor is not None else zero     # y_f_cross_U
        py = deriv(1, C, F, C)(i, j, k, g, pHY) if pHY is not None else zero
        Gv = (- div_Uu(1, i, j, k, g, adv, U, v) - zero - zero
              - cory
              - py
              - div_τ(1, i, j, k, g, clo, u, v, w, self.νe)
              - zero + zero + zero
              + gb(1)
              + zero)
        # w: :172-180 (no buoyancy and no pHY′ term)
        Gw = (- div_Uu(2, i, j, k, g, adv, U, w) - zero - zero
              - zero
              - div_τ(2, i, j, k, g, clo, u, v, w, self.νe)
              - zero + zero + zero
              + zero)
        self.Gn["u"][i, j, k] = Gu
        self.Gn["v"][i, j, k] = Gv
        self.Gn["w"][i, j, k] = Gw
        for name in self.tracer_names:
            c = self.tracers[name]
            amd = isinstance(clo, AnisotropicMinimumDissipation)
            κ = 0 if (clo is None or amd) else (clo.prandtl(name) if isinstance(clo, SmagorinskyLilly) else clo.kappa(name))
            # :225-231
            Gc = (- div_Uc(i, j, k, g, adv, U, c) - zero - zero
                  - div_q(i, j, k, g, clo, κ, c, self.νe, self.κe[name] if amd else None)
                  - zero
                  + zero)
            self.Gn[name][i, j, k] = Gc
        # calculate_boundary_tendency_contributions! :187-200
        for name in self.names:
            apply_flux_bcs(self.Gn[name], self.fields[name])

    # ---- pressure ------------------------------------------------------------------------
    def calculate_pressure_correction(self, Δt):
        """pressure_correction.jl:10-23 + solve_for_pressure.jl:15-89."""
        g = self.grid
        u, v, w = (self.velocities[n] for n in "uvw")
        fill_halo_regions([u, v, w])
        i, j, k = self._box()
        s = self.pressure_solver
        if isinstance(s, FFTBasedPoissonSolver):
            s.storage[...] = div_ccc(i, j, k, g, u, v, w) / Δt
            s.solve(self.pNHS)
        else:
            s.source_term[...] = g.Δz(C, k) * div_ccc(i, j, k, g, u, v, w) / Δt
            s.solve(self.pNHS)
        fill_halo_regions(self.pNHS)

    def pressure_correct_velocities(self, Δt):
        """_pressure_correct_velocities! pressure_correction.jl:34-40."""
        g = self.grid
        i, j, k = self._box()
        p = self.pNHS
        u, v, w = (self.velocities[n] for n in "uvw")
        u[i, j, k] = u[i, j, k] - deriv(0, F, C, C)(i, j, k, g, p) * Δt
        v[i, j, k] = v[i, j, k] - deriv(1, C, F, C)(i, j, k, g, p) * Δt
        w[i, j, k] = w[i, j, k] - deriv(2, C, C, F)(i, j, k, g, p) * Δt

    def store_tendencies(self):
        i, j, k = self._box()
        for n in self.names:
            self.Gm[n][i, j, k] = self.Gn[n][i, j, k]

    # ---- time stepping ---------------------------------------------------------------------
    def time_step(self, Δt, euler=False):
        FT = self.grid.FT
        Δt = FT(Δt)
        if self.timestepper == "RungeKutta3":
            self._rk3(Δt)
        else:
            self._ab2(Δt, euler)

    def _rk3(self, Δt):
        """time_step!(::RungeKutta3TimeStepper) runge_kutta_3.jl:81-152."""
        if self.clock.iteration == 0:
            self.update_state()
        γ1, γ2, γ3, ζ2, ζ3 = self.γ1, self.γ2, self.γ3, self.ζ2, self.ζ3
        stages = ((γ1, None, γ1 * Δt), (γ2, ζ2, (γ2 + ζ2) * Δt), (γ3, ζ3, (γ3 + ζ3) * Δt))
        i, j, k = self._box()
        for m, (γ, ζ, sΔt) in enumerate(stages):
            self.calculate_tendencies()
            for n in self.names:                 # rk3_substep_field! :204-218
                f, Gn, Gm = self.fields[n], self.Gn[n], self.Gm[n]
                if ζ is None:
                    f[i, j, k] = f[i, j, k] + Δt * γ * Gn[i, j, k]
                else:
                    f[i, j, k] = f[i, j, k] + Δt * (γ * Gn[i, j, k] + ζ * Gm[i, j, k])
            self.implicit_step(sΔt)              # stage_Δt(Δt, γ, ζ) = Δt (γ + ζ)
            self.calculate_pressure_correction(sΔt)
            self.pressure_correct_velocities(sΔt)
            self.clock.time += float(sΔt)
            if m < 2:
                self.clock.stage += 1
                self.store_tendencies()
            else:
                self.clock.iteration += 1
                self.clock.stage = 1
            self.update_state()

    def _ab2(self, Δt, euler=False):
        """time_step!(::QuasiAdamsBashforth2TimeStepper) quasi_adams_bashforth_2.jl:70-104."""
        FT = self.grid.FT
        euler = euler or (Δt != self.previous_Δt)
        χ = FT(-0.5) if euler else self.χ
        i, j, k = self._box()
        if euler:
            for n in self.names:
                self.Gm[n].parent[...] = 0
        self.previous_Δt = Δt
        if self.clock.iteration == 0:
            self.update_state()
        self.calculate_tendencies()
        for n in self.names:                     # ab2_step_field! :158-166
            f, Gn, Gm = self.fields[n], self.Gn[n], self.Gm[n]
            f[i, j, k] = f[i, j, k] + Δt * ((FT(1.5) + χ) * Gn[i, j, k] - (FT(0.5) + χ) * Gm[i, j, k])
        self.implicit_step(Δt)
        self.calculate_pressure_correction(Δt)
        self.pressure_correct_velocities(Δt)
        self.store_tendencies()
        self.clock.time += float(Δt)
        self.clock.iteration += 1
        self.update_state()

    def implicit_step(self, Δt):
        """implicit_step!(field, implicit_solver, closure, ...) for every prognostic field right after its substep
        (runge_kutta_3.jl:178-185, quasi_adams_bashforth_2.jl:137-144, vertically_implicit_diffusion_solver.jl:153-195):
        in-place tridiagonal solve over k = 1..Nz of every column"""
        clo = self.closure
        if clo is None or not getattr(clo, "vitd", False):
            return
        g = self.grid
        assert g.topology[2] == Bounded, "VerticallyImplicitTimeDiscretization needs a Bounded z"
        for n in self.names:
            f = self.fields[n]
            κ = clo.ν if n in "uvw" else clo.kappa(n)
            a, b, c = ivd_coefficients(g, Δt, κ, z_face_field=(n == "w"))
            solver = BatchedTridiagonalSolver(g, a, b, c)
            ϕ = f.interior[:, :, :g.Nz]            # a view: the solve is in place, rhs = the field itself
            solver.t = np.zeros(ϕ.shape, dtype=g.FT)
            solver.solve(ϕ, ϕ)

    # ---- diagnostics ------------------------------------------------------------------------
    def max_divergence(self):
        i, j, k = self._box()
        u, v, w = (self.velocities[n] for n in "uvw")
        return float(np.max(np.abs(div_ccc(i, j, k, self.grid, u, v, w))))

    def kinetic_energy(self):
        i, j, k = self._box()
        return float(sum(np.sum(self.velocities[n][i, j, k].astype(np.float64) ** 2) for n in "uvw")) * 0.5
