"""Simulation / run! (reference src/Simulations/simulation.jl:44-85, run.jl:42-140): unchanged
host logic driving `time_step!` on the B200 model; NaNChecker (nan_checker.jl:33-52) uses the
library's device reduction."""
import math

from .model import time_step, sync


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
