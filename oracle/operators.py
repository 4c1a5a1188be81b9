"""
Staggered-grid operators (test infrastructure -- see oracle/__init__.py).

Follows Operators/difference_operators.jl:7-49, interpolation_operators.jl:20-114,
derivative_operators.jl:6-29, spacings_and_areas_and_volumes.jl, divergence_operators.jl:16-19,
laplacian_operators.jl:36-40 and products_between_fields_and_grid_metrics.jl.

Naming: the reference's  δxᶜᵃᵃ -> dxc,  δxᶠᵃᵃ -> dxf,  ℑxᶜᵃᵃ -> Ixc,  ℑxᶠᵃᵃ -> Ixf  (likewise
y, z).  Every operator takes (i, j, k, grid, f, *args) where i, j, k are `R` index ranges
and `f` is a Field or a function f(i, j, k, grid, *args), exactly as in the reference.
Flat dimensions: differences are zero and interpolations are the identity
(difference_operators.jl:27-49, interpolation_operators.jl:92-114).
"""
from functools import partial

from .grids import Flat, Center, Face


def val(f, i, j, k, grid, *args):
    if callable(f):
        return f(i, j, k, grid, *args)
    return f[i, j, k]


def _sh(ijk, d, n):
    ijk = list(ijk)
    ijk[d] = ijk[d] + n
    return ijk


def _delta_c(d, i, j, k, grid, f, *args):
    """δxᶜᵃᵃ: f(i+1) - f(i)."""
    if grid.topology[d] == Flat:
        return grid.FT(0)
    return val(f, *_sh((i, j, k), d, 1), grid, *args) - val(f, i, j, k, grid, *args)


def _delta_f(d, i, j, k, grid, f, *args):
    """δxᶠᵃᵃ: f(i) - f(i-1)."""
    if grid.topology[d] == Flat:
        return grid.FT(0)
    return val(f, i, j, k, grid, *args) - val(f, *_sh((i, j, k), d, -1), grid, *args)


def _interp_c(d, i, j, k, grid, f, *args):
    """ℑxᶜᵃᵃ: FT(0.5) * (f(i) + f(i+1))."""
    if grid.topology[d] == Flat:
        return val(f, i, j, k, grid, *args)
    return grid.FT(0.5) * (val(f, i, j, k, grid, *args) + val(f, *_sh((i, j, k), d, 1), grid, *args))


def _interp_f(d, i, j, k, grid, f, *args):
    """ℑxᶠᵃᵃ: FT(0.5) * (f(i-1) + f(i))."""
    if grid.topology[d] == Flat:
        return val(f, i, j, k, grid, *args)
    return grid.FT(0.5) * (val(f, *_sh((i, j, k), d, -1), grid, *args) + val(f, i, j, k, grid, *args))


dxc, dyc, dzc = (partial(_delta_c, d) for d in range(3))
dxf, dyf, dzf = (partial(_delta_f, d) for d in range(3))
Ixc, Iyc, Izc = (partial(_interp_c, d) for d in range(3))
Ixf, Iyf, Izf = (partial(_interp_f, d) for d in range(3))
DELTA = {Center: (dxc, dyc, dzc), Face: (dxf, dyf, dzf)}
INTERP = {Center: (Ixc, Iyc, Izc), Face: (Ixf, Iyf, Izf)}


# two-dimensional interpolations (interpolation_operators.jl:60-75): outer op applied to inner op
def Ixy_fca(i, j, k, grid, f, *a):      # ℑxyᶠᶜᵃ = ℑyᵃᶜᵃ(ℑxᶠᵃᵃ)
    return Iyc(i, j, k, grid, Ixf, f, *a)


def Ixy_cfa(i, j, k, grid, f, *a):      # ℑxyᶜᶠᵃ = ℑyᵃᶠᵃ(ℑxᶜᵃᵃ)
    return Iyf(i, j, k, grid, Ixc, f, *a)


# ---- metrics as functions of (i, j, k, grid) at a named location -------------------------
def area(d, lx, ly, lz):
    def A(i, j, k, grid):
        return (grid.Ax, grid.Ay, grid.Az)[d](i, j, k, lx, ly, lz)
    return A


def volume(lx, ly, lz):
    def V(i, j, k, grid):
        return grid.V(i, j, k, lx, ly, lz)
    return V


def A_q(d, lx, ly, lz):
    """Ax_qᶠᶜᶜ etc. (products_between_fields_and_grid_metrics.jl): metric * q."""
    A = area(d, lx, ly, lz)

    def Aq(i, j, k, grid, q, *args):
        return A(i, j, k, grid) * val(q, i, j, k, grid, *args)
    return Aq


def deriv(d, lx, ly, lz):
    """∂xᶠᶜᶜ etc. (derivative_operators.jl:6-29): δ / Δ at the result location."""
    loc = (lx, ly, lz)
    delta = DELTA[loc[d]][d]

    def D(i, j, k, grid, f, *args):
        return delta(i, j, k, grid, f, *args) / grid.spacing(d, loc[d], (i, j, k)[d])
    return D


def div_ccc(i, j, k, grid, u, v, w):
    """divᶜᶜᶜ, divergence_operators.jl:16-19."""
    return 1 / grid.V(i, j, k, Center, Center, Center) * (
        dxc(i, j, k, grid, A_q(0, Face, Center, Center), u) +
        dyc(i, j, k, grid, A_q(1, Center, Face, Center), v) +
        dzc(i, j, k, grid, A_q(2, Center, Center, Face), w))


def laplacian_ccc(i, j, k, grid, c):
    """∇²ᶜᶜᶜ, laplacian_operators.jl:36-40:
    1/V * (δxᶜ(Ax_∂xᶠᶜᶜ c) + δyᶜ(Ay_∂yᶜᶠᶜ c) + δzᶜ(Az_∂zᶜᶜᶠ c))."""
    def A_d(d, lx, ly, lz):
        A, D = area(d, lx, ly, lz), deriv(d, lx, ly, lz)
        return lambda i, j, k, grid, c: A(i, j, k, grid) * D(i, j, k, grid, c)
    return 1 / grid.V(i, j, k, Center, Center, Center) * (
        dxc(i, j, k, grid, A_d(0, Face, Center, Center), c) +
        dyc(i, j, k, grid, A_d(1, Center, Face, Center), c) +
        dzc(i, j, k, grid, A_d(2, Center, Center, Face), c))
