#!/usr/bin/env python
"""Generates tests/golden/*.npz.

The reference (Oceananigans.jl v0.76.8) is pure Julia and cannot run in this image (no Julia toolchain, SURVEY.md
8(c)), so these vectors are NOT outputs of the reference itself: they are outputs of the oracle (oracle/, the NumPy
restatement whose fidelity is pinned by the reference's own analytic tests, tests/test_oracle_known_answers.py),
frozen so that (a) any later edit of the oracle that changes its results is caught on CPU, and (b) the CUDA path is
checked against stored numbers as well as against the live oracle.  Inputs are regenerated from the seeds by
golden_cases.py, only the outputs are stored.

    python tests/golden/make_golden.py          # rewrites the .npz files
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O                                    # noqa: E402
from golden_cases import MODEL_CASES, POISSON_CASES, build_oracle_model, model_initial_values, poisson_rhs  # noqa: E402


def main():
    for name, cfg in MODEL_CASES.items():
        m = build_oracle_model(O, cfg)
        m.set(**model_initial_values(m, cfg["seed"]))
        out = {}
        for step in range(cfg["steps"]):
            m.time_step(cfg["dt"])
            if step == 0:
                out.update({f"step1_{n}": np.array(m.fields[n].interior) for n in m.names})
        out.update({f"final_{n}": np.array(m.fields[n].interior) for n in m.names})
        out["kinetic_energy"] = np.float64(m.kinetic_energy())
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.shape})
    for name, cfg in POISSON_CASES.items():
        g = O.RectilinearGrid(np.float64, **cfg["grid"])
        rhs = poisson_rhs(g, cfg["seed"])
        phi = O.Field(g, auxiliary=True)
        if cfg["solver"] == "ft":
            O.FourierTridiagonalPoissonSolver(g).solve(phi, rhs)
        else:
            s = O.FFTBasedPoissonSolver(g)
            s.storage[...] = rhs
            s.solve(phi)
        np.savez_compressed(os.path.join(HERE, f"poisson_{name}.npz"), phi=np.array(phi.interior))
        print(name, phi.interior.shape)


if __name__ == "__main__":
    main()
