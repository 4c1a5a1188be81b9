"""
Poisson solvers (test infrastructure -- see oracle/__init__.py).

Follows Solvers/poisson_eigenvalues.jl:8-31, fft_based_poisson_solver.jl:50-125,
plan_transforms.jl:21-39,124-146 (CPU path: r2r REDFT10/REDFT01 over the Bounded dims, c2c FFT
over the Periodic dims), discrete_transforms.jl:25-33 (REDFT01 normalised by 1/2N),
fourier_tridiagonal_poisson_solver.jl:16-123 and batched_tridiagonal_solver.jl:91-122.

The third-party arithmetic at this boundary is FFTW (FFTW.jl ^1, binary FFTW 3.3.x), absent
here.  Its published definitions are restated with scipy.fft (pocketfft):
  FFTW forward c2c      = unnormalised DFT           = scipy.fft.fft
  FFTW backward c2c /N  = plan_ifft!                 = scipy.fft.ifft
  REDFT10               = 2 sum x_n cos(pi (n+1/2) k / N) = scipy.fft.dct(type=2, norm=None)
  REDFT01 / (2N)        = inverse of the above       = scipy.fft.idct(type=2, norm=None)
Bit patterns differ from FFTW's (different factorisation); parity at this boundary is
anchored on the reference's own identity test lap(phi) == rhs
(test/dependencies_for_poisson_solvers.jl:86-104) which holds for any correct DFT/DCT.
"""
import numpy as np
import scipy.fft as sfft

from .grids import Periodic, Bounded, Flat, Center, Face
from .fields import R


def poisson_eigenvalues(N, L, dim, topo):
    """poisson_eigenvalues.jl:8-31 (always Float64)."""
    shape = [1, 1, 1]
    shape[dim] = N
    if topo == Flat:
        return np.zeros(N).reshape(shape)
    inds = np.arange(1, N + 1, dtype=np.float64)
    L = float(L)
    if topo == Periodic:
        lam = (2 * np.sin((inds - 1) * np.pi / N) / (L / N)) ** 2
    else:
        lam = (2 * np.sin((inds - 1) * np.pi / (2 * N)) / (L / N)) ** 2
    return lam.reshape(shape)


def _forward(b, topo):
    bd = [d for d in range(3) if topo[d] == Bounded]
    pd = [d for d in range(3) if topo[d] == Periodic]
    for d in bd:     # r2r acts on real and imaginary parts separately
        b = sfft.dct(b.real, type=2, axis=d) + 1j * sfft.dct(b.imag, type=2, axis=d)
    if pd:
        b = sfft.fftn(b, axes=pd)
    return b


def _backward(b, topo):
    bd = [d for d in range(3) if topo[d] == Bounded]
    pd = [d for d in range(3) if topo[d] == Periodic]
    if pd:
        b = sfft.ifftn(b, axes=pd)
    for d in bd:
        b = sfft.idct(b.real, type=2, axis=d) + 1j * sfft.idct(b.imag, type=2, axis=d)
    return b


class FFTBasedPoissonSolver:
    def __init__(self, grid):
        assert all(grid.regular), "FFTBasedPoissonSolver needs a regular grid"
        self.grid = grid
        t = grid.topology
        self.λx = poisson_eigenvalues(grid.Nx, grid.Lx, 0, t[0])
        self.λy = poisson_eigenvalues(grid.Ny, grid.Ly, 1, t[1])
        self.λz = poisson_eigenvalues(grid.Nz, grid.Lz, 2, t[2])
        self.ctype = np.complex64 if grid.FT == np.float32 else np.complex128
        self.storage = np.zeros(grid.N, dtype=self.ctype, order="F")

    def solve(self, ϕ, b=None, m=0):
        """solve!(ϕ, solver, b, m) fft_based_poisson_solver.jl:93-120; ϕ is a Field."""
        g = self.grid
        b = self.storage if b is None else b
        b = _forward(b.astype(self.ctype), g.topology).astype(self.ctype)
        with np.errstate(divide="ignore", invalid="ignore"):
            ϕc = (-b / (self.λx + self.λy + self.λz - m)).astype(self.ctype)
        if m == 0:
            ϕc[0, 0, 0] = 0
        ϕc = _backward(ϕc, g.topology).astype(self.ctype)
        self.storage[...] = ϕc
        ϕ[R(1, g.Nx), R(1, g.Ny), R(1, g.Nz)] = ϕc.real      # copy_real_component!
        return ϕ


class BatchedTridiagonalSolver:
    """batched_tridiagonal_solver.jl:10-122; a, c are vectors (length Nz-1), b is a 3-D array
    (or a vector), one (i, j) column per 'thread', serial in k."""

    def __init__(self, grid, lower_diagonal, diagonal, upper_diagonal):
        self.grid = grid
        self.a, self.b, self.c = lower_diagonal, diagonal, upper_diagonal
        self.t = np.zeros(grid.N, dtype=grid.FT)

    @staticmethod
    def _coef(x, k):
        """get_coefficient: 1-D arrays indexed by k, 3-D arrays by (i, j, k); k is 1-based."""
        x = np.asarray(x)
        return x[k - 1] if x.ndim == 1 else x[:, :, k - 1]

    def solve(self, ϕ, f):
        Nz = self.grid.Nz
        a, b, c, t = self.a, self.b, self.c, self.t
        eps = np.finfo(np.float32 if ϕ.real.dtype == np.float32 else np.float64).eps
        β = self._coef(b, 1) + np.zeros(ϕ.shape[:2])
        ϕ[:, :, 0] = self._coef(f, 1) / β
        active = np.ones(ϕ.shape[:2], dtype=bool)          # columns that have not hit `break`
        for k in range(2, Nz + 1):
            ck, bk, ak = self._coef(c, k - 1), self._coef(b, k), self._coef(a, k - 1)
            tk = ck / β
            t[:, :, k - 1] = np.where(active, tk, t[:, :, k - 1])
            βn = bk - ak * tk
            β = np.where(active, βn, β)
            dd = np.abs(β) > 10 * eps
            active = active & dd
            new = (self._coef(f, k) - ak * ϕ[:, :, k - 2]) / np.where(active, β, 1)
            ϕ[:, :, k - 1] = np.where(active, new, ϕ[:, :, k - 1])
        for k in range(Nz - 1, 0, -1):
            ϕ[:, :, k - 1] = ϕ[:, :, k - 1] - t[:, :, k] * ϕ[:, :, k]
        return ϕ


class FourierTridiagonalPoissonSolver:
    def __init__(self, grid):
        assert grid.topology[2] == Bounded
        assert grid.regular[0] and grid.regular[1]
        self.grid = g = grid
        t = g.topology
        self.λx = poisson_eigenvalues(g.Nx, g.Lx, 0, t[0])
        self.λy = poisson_eigenvalues(g.Ny, g.Ly, 1, t[1])
        Nz = g.Nz
        ΔzF = lambda k: float(g.dF[2]) if g.regular[2] else float(g.dF[2][k])
        ΔzC = lambda k: float(g.dC[2]) if g.regular[2] else float(g.dC[2][k])
        lower = np.array([1 / ΔzF(k) for k in range(2, Nz + 1)])
        # compute_main_diagonals! :16-28 (Float64 array)
        λ = (self.λx + self.λy)[:, :, 0]
        D = np.zeros(g.N)
        D[:, :, 0] = -1 / ΔzF(2) - ΔzC(1) * λ
        for k in range(2, Nz):
            D[:, :, k - 1] = -(1 / ΔzF(k + 1) + 1 / ΔzF(k)) - ΔzC(k) * λ
        D[:, :, Nz - 1] = -1 / ΔzF(Nz) - ΔzC(Nz) * λ
        self.bt = BatchedTridiagonalSolver(g, lower, D, lower)
        self.ctype = np.complex64 if g.FT == np.float32 else np.complex128
        self.source_term = np.zeros(g.N, dtype=self.ctype, order="F")
        self.storage = np.zeros(g.N, dtype=self.ctype, order="F")
        self.ΔzC = np.array([ΔzC(k) for k in range(1, Nz + 1)], dtype=g.FT).reshape(1, 1, Nz)

    def set_source_term(self, src):
        """:109-123: source_term .= src ; source_term *= Δzᵃᵃᶜ."""
        self.source_term[...] = src
        self.source_term *= self.ΔzC

    def solve(self, x, b=None):
        """:74-101; x is a Field."""
        g = self.grid
        if b is not None:
            self.set_source_term(b)
        topo_xy = (g.topology[0], g.topology[1], Flat)
        st = _forward(self.source_term, topo_xy).astype(self.ctype)
        self.source_term[...] = st
        ϕ = self.storage
        self.bt.t = np.zeros(g.N, dtype=g.FT)
        self.bt.solve(ϕ, st)
        ϕ[...] = _backward(ϕ, topo_xy).astype(self.ctype)
        ϕ[...] = ϕ.real
        ϕ[...] = ϕ - np.mean(ϕ)
        x[R(1, g.Nx), R(1, g.Ny), R(1, g.Nz)] = ϕ.real
        return x
