"""Simulation / run! (reference src/Simulations/simulation.jl:44-85, run.jl:42-140): unchanged
host logic driving `time_step!` on the B200 model; NaNChecker (nan_checker.jl:33-52) uses the
library's device reduction."""
import ctypes as C
import math

import numpy as np

from ._lib import lib, check
from .grids import Flat
from .model import time_step, sync


def min_spacings(grid):
    """min_Δx, min_Δy, min_Δz (src/Grids/rectilinear_grid.jl:429-454): Inf along Flat dimensions, the scalar spacing of a
    regular dimension, the minimum of the Δᶜ vector (halos included) of a stretched one."""
    out = []
    for d in range(3):
        if grid.topology[d] == Flat:
            out.append(math.inf)
        elif grid.regular[d]:
            out.append(float(grid.dC[d]))
        else:
            out.append(float(np.min(grid.dC[d].a)))
    return tuple(out)


def max_abs_velocities(model):
    """maximum(abs, parent(u / v / w)) by three device reductions in one library call"""
    out = (C.c_double * 3)()
    check(lib.ob200_model_max_abs_velocities(model.handle, out))
    return tuple(out)


def cell_advection_timescale(model):
    """cell_advection_timescale(model) (src/Utils/cell_advection_timescale.jl:4-21): min(Δx/umax, Δy/vmax, Δz/wmax)"""
    umax, vmax, wmax = max_abs_velocities(model)
    dx, dy, dz = min_spacings(model.grid)
    div = lambda a, b: (math.inf if b == 0 else a / b) if math.isfinite(a) else math.inf
    return min(div(dx, umax), div(dy, vmax), div(dz, wmax))


def cell_diffusion_timescale(model):
    """cell_diffusion_timescale (src/TurbulenceClosures/turbulence_closure_diagnostics.jl:23-39) for nothing / ScalarDiffusivity"""
    clo = model.closure
    if clo is None:
        return math.inf
    if type(clo).__name__ == "SmagorinskyLilly":      # turbulence_closure_diagnostics.jl:48-53: Δ² / (max νₑ max(1, 1/min Pr))
        dx, dy, dz = min_spacings(model.grid)
        Δ = min(dx, dy, dz)
        prs = list(clo.Pr.values()) if isinstance(clo.Pr, dict) else [clo.Pr]
        min_pr = min(prs) if model.tracers else 1
        max_ν = model.diffusivity_fields["νₑ"].reduce()["maxabs"] * max(1, 1 / min_pr)
        return math.inf if max_ν == 0 else Δ ** 2 / max_ν
    dx, dy, dz = min_spacings(model.grid)
    Δ = {"ThreeDimensional": min(dx, dy, dz), "Horizontal": min(dx, dy), "Vertical": dz}[clo.formulation]
    κs = list(clo.κ.values()) if isinstance(clo.κ, dict) else [clo.κ]
    max_κ = max(κs) if κs else 0
    div = lambda a, b: math.inf if b == 0 else a / b
    return min(div(Δ ** 2, clo.ν), div(Δ ** 2, max_κ))


class TimeStepWizard:
    """TimeStepWizard(; cfl=0.2, diffusive_cfl=Inf, max_change=1.1, min_change=0.5, max_Δt=Inf, min_Δt=0)
    (src/Simulations/time_step_wizard.jl:17-95): a callback that resets simulation.Δt from the CFL numbers; the only device
    work is the velocity max-reduction."""

    def __init__(self, cfl=0.2, diffusive_cfl=math.inf, max_change=1.1, min_change=0.5, max_Δt=math.inf, min_Δt=0.0):
        self.cfl, self.diffusive_cfl = cfl, diffusive_cfl
        self.max_change, self.min_change, self.max_Δt, self.min_Δt = max_change, min_change, max_Δt, min_Δt

    def new_time_step(self, old_Δt, model):
        advective = self.cfl * cell_advection_timescale(model)
        diffusive = self.diffusive_cfl * cell_diffusion_timescale(model) if math.isfinite(self.diffusive_cfl) else math.inf
        new = min(advective, diffusive)
        new = min(self.max_change * old_Δt, new)
        new = max(self.min_change * old_Δt, new)
        return min(max(new, self.min_Δt), self.max_Δt)

    def __call__(self, simulation):
        simulation.Δt = self.new_time_step(simulation.Δt, simulation.model)


class Simulation:
    def __init__(self, model, Δt, stop_iteration=math.inf, stop_time=math.inf, nan_check_interval=100):
        self.model, self.Δt = model, Δt
        self.stop_iteration, self.stop_time = stop_iteration, stop_time
        self.nan_check_interval = nan_check_interval
        self.running = True
        self.callbacks = []

    def aligned_time_step(self):
        """run.jl:42-57: clip the step to land on stop_time; fall back to Δt if that is <= 0."""
        t = self.model.clock.time
        aligned = min(self.Δt, self.stop_time - t)
        return self.Δt if aligned <= 0 else aligned

    def stop_criteria(self):
        c = self.model.clock
        if c.iteration >= self.stop_iteration or c.time >= self.stop_time:
            self.running = False


def run(sim):
    """run!(simulation) (run.jl:86-140)."""
    sim.running = True
    sim.stop_criteria()
    while sim.running:
        time_step(sim.model, sim.aligned_time_step())
        it = sim.model.clock.iteration
        if sim.nan_check_interval and it % sim.nan_check_interval == 0:
            if sim.model.velocities["u"].reduce()["has_nan"]:
                sim.running = False
                raise FloatingPointError(f"NaN found in u at iteration {it}")
        for cb in sim.callbacks:
            cb(sim)
        sim.stop_criteria()
    sync()
