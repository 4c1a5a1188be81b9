"""
oracle -- CPU restatement of the Oceananigans.jl v0.76.8 NonhydrostaticModel time step.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is the checker for the CUDA path, never the
product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it.  The product (`libocean_b200.so` and the
`ocean_b200` host package) never imports, links or executes anything in here.

What it is: a vectorised NumPy restatement, in the reference's own operation order, of the
in-scope subset named in SURVEY.md section 8(a):  RectilinearGrid metrics, the staggered
operators, WENO5 / centred / upwind advection, ScalarDiffusivity, FPlane, BuoyancyTracer,
halo filling and flux boundary conditions, the RK3 and AB2 time steppers, the FFT-based
and Fourier-tridiagonal Poisson solvers, and the pressure projection.  Every function
cites the reference file:line it follows (paths relative to /root/reference/src).

Parity pinning: the reference is pure Julia and no Julia toolchain exists in this image
(nor on the GPU box), and the reference's stored regression data lives in an external
artifact repository that cannot be fetched offline.  STEP-LEVEL PARITY AGAINST THE REAL
JULIA REFERENCE IS THEREFORE UNPINNED BY STORED DATA ("parity unpinned").  What pins
the oracle instead are the reference's own analytic / self-consistency tests, ported in
tests/test_oracle_*.py: Poisson `lap(phi) == R` for all topologies, Thomas vs dense
solve, halo identities, Taylor-Green, Gaussian advection, cosine diffusion decay,
incompressibility, AB2-first-step-is-Euler, tracer conservation, flux-BC budgets and
WENO5 fifth-order convergence.

A compiled twin with identical arithmetic (oracle/oracle_cpu.c, OpenMP) exists for the
CPU-baseline timing and as a second, independent restatement.
"""
from .grids import RectilinearGrid, Periodic, Bounded, Flat, Center, Face  # noqa: F401
from .fields import (Field, FieldBoundaryConditions, BoundaryCondition,  # noqa: F401
                     fill_halo_regions, R)
from .advection import (WENO5, CenteredSecondOrder, CenteredFourthOrder,  # noqa: F401
                        UpwindBiasedFirstOrder, UpwindBiasedThirdOrder,
                        UpwindBiasedFifthOrder)
from .closures import ScalarDiffusivity, SmagorinskyLilly, AnisotropicMinimumDissipation  # noqa: F401
from .solvers import (FFTBasedPoissonSolver, FourierTridiagonalPoissonSolver,  # noqa: F401
                      BatchedTridiagonalSolver, poisson_eigenvalues)
from .model import (NonhydrostaticModel, FPlane, BuoyancyTracer, Buoyancy, SeawaterBuoyancy,  # noqa: F401
                    LinearEquationOfState)
