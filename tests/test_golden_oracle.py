"""CPU: the oracle reproduces the committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py).
The vectors are oracle outputs frozen at commit time -- the Julia reference cannot run here -- so this
pins the oracle against later edits; the GPU tests check the CUDA path against the same stored numbers."""
import os

import numpy as np
import pytest

import oracle as O
from golden_cases import MODEL_CASES, POISSON_CASES, build_oracle_model, model_initial_values, poisson_rhs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_oracle_reproduces_golden_model_steps(name):
    cfg = MODEL_CASES[name]
    gold = np.load(os.path.join(GOLD, f"model_{name}.npz"))
    m = build_oracle_model(O, cfg)
    m.set(**model_initial_values(m, cfg["seed"]))
    for step in range(cfg["steps"]):
        m.time_step(cfg["dt"])
        if step == 0:
            for n in m.names:
                assert rel(m.fields[n].interior, gold[f"step1_{n}"]) < 1e-13, (name, n)
    for n in m.names:
        assert rel(m.fields[n].interior, gold[f"final_{n}"]) < 1e-13, (name, n)
    assert abs(m.kinetic_energy() - float(gold["kinetic_energy"])) <= 1e-13 * abs(float(gold["kinetic_energy"]))


@pytest.mark.parametrize("name", list(POISSON_CASES))
def test_oracle_reproduces_golden_poisson(name):
    cfg = POISSON_CASES[name]
    gold = np.load(os.path.join(GOLD, f"poisson_{name}.npz"))["phi"]
    g = O.RectilinearGrid(np.float64, **cfg["grid"])
    rhs = poisson_rhs(g, cfg["seed"])
    phi = O.Field(g, auxiliary=True)
    if cfg["solver"] == "ft":
        O.FourierTridiagonalPoissonSolver(g).solve(phi, rhs)
    else:
        s = O.FFTBasedPoissonSolver(g)
        s.storage[...] = rhs
        s.solve(phi)
    assert rel(phi.interior, gold) < 1e-13
    assert np.max(np.abs(gold)) > 0


def test_compiled_twin_matches_golden_c2():
    """oracle/oracle_cpu.c (the OpenMP twin used for the CPU baseline) against the same stored C2 vectors"""
    from oracle import cpu_twin
    cfg = MODEL_CASES["c2_periodic_weno_rk3"]
    gold = np.load(os.path.join(GOLD, "model_c2_periodic_weno_rk3.npz"))
    m = build_oracle_model(O, cfg)
    vals = model_initial_values(m, cfg["seed"])
    m.set(**vals)                      # set! projects the initial velocities; the twin starts from the projected state
    st = {n: np.array(m.fields[n].interior) for n in m.names}
    out = cpu_twin.rk3_run(cfg["grid"]["size"], cfg["grid"]["extent"], st["u"], st["v"], st["w"], st["b"],
                           cfg["steps"], cfg["dt"], project=False)
    for n, a in zip("uvwb", out):
        assert rel(a, gold[f"final_{n}"]) < 1e-11, n
