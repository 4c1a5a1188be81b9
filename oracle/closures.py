"""
ScalarDiffusivity closure (test infrastructure -- see oracle/__init__.py).

Follows TurbulenceClosures/closure_kernel_operators.jl:22-48 (flux divergences),
abstract_scalar_diffusivity_closure.jl:172-207 (viscous and diffusive fluxes for the
ThreeDimensional / Horizontal / Vertical formulations), velocity_tracer_gradients.jl:25-43
(strain-rate components), Operators/divergence_operators.jl:35-37 (div_xyᶜᶜᶜ) and
Operators/vorticity_operators.jl:2-5 (ζ₃ᶠᶠᶜ).  Constant ν, κ only; explicit time
discretisation only.
"""
import numpy as np

from .grids import Center as C, Face as F
from .operators import DELTA, A_q, area, deriv, dxc, dyc, dxf, dyf, val, Ixc, Iyc, Izc, Ixf, Iyf, Izf

THREE_D, HORIZONTAL, VERTICAL = "ThreeDimensional", "Horizontal", "Vertical"


class ScalarDiffusivity:
    """ScalarDiffusivity(formulation; ν=0, κ=0) (scalar_diffusivity.jl:60-76); κ is a number
    or a dict {tracer name: number}."""

    def __init__(self, formulation=THREE_D, ν=0.0, κ=0.0):
        self.formulation = formulation
        self.ν = ν
        self.κ = κ
        self.required_halo = 1

    def kappa(self, name):
        return self.κ[name] if isinstance(self.κ, dict) else self.κ


class SmagorinskyLilly:
    """SmagorinskyLilly(FT; C=0.16, Cb=1.0, Pr=1.0) (turbulence_closure_implementations/smagorinsky_lilly.jl:6-72):
    eddy viscosity nu_e = (C Delta)^2 sqrt(2 Sigma^2) sqrt(1 - min(1, Cb N^2 / Sigma^2)) at cell centres, kappa_e = nu_e / Pr;
    ThreeDimensionalFormulation, explicit time discretisation.  Pr is a number or a dict {tracer name: number}."""
    formulation = THREE_D

    def __init__(self, C=0.16, Cb=1.0, Pr=1.0):
        self.C, self.Cb, self.Pr = C, Cb, Pr
        self.required_halo = 1

    def prandtl(self, name):
        return self.Pr[name] if isinstance(self.Pr, dict) else self.Pr


def _Δ_q(d, lx, ly, lz):
    """Δx_qᶠᶜᶜ etc.: spacing * q."""
    loc = (lx, ly, lz)

    def f(i, j, k, grid, q):
        return grid.spacing(d, loc[d], (i, j, k)[d]) * val(q, i, j, k, grid)
    return f


def div_xy_ccc(i, j, k, grid, u, v):
    return 1 / grid.Az(i, j, k, C, C, C) * (dxc(i, j, k, grid, _Δ_q(1, F, C, C), u) +
                                            dyc(i, j, k, grid, _Δ_q(0, C, F, C), v))


def zeta3_ffc(i, j, k, grid, u, v):
    Γ = dxf(i, j, k, grid, _Δ_q(1, C, F, C), v) - dyf(i, j, k, grid, _Δ_q(0, F, C, C), u)
    return Γ / grid.Az(i, j, k, F, F, C)


# strain rates, velocity_tracer_gradients.jl:25-43
def Σ11(i, j, k, grid, u, v, w):
    return deriv(0, C, C, C)(i, j, k, grid, u)


def Σ22(i, j, k, grid, u, v, w):
    return deriv(1, C, C, C)(i, j, k, grid, v)


def Σ33(i, j, k, grid, u, v, w):
    return deriv(2, C, C, C)(i, j, k, grid, w)


def Σ12(i, j, k, grid, u, v, w):
    return grid.FT(0.5) * (deriv(1, F, F, C)(i, j, k, grid, u) + deriv(0, F, F, C)(i, j, k, grid, v))


def Σ13(i, j, k, grid, u, v, w):
    return grid.FT(0.5) * (deriv(2, F, C, F)(i, j, k, grid, u) + deriv(0, F, C, C)(i, j, k, grid, w))


def Σ23(i, j, k, grid, u, v, w):
    return grid.FT(0.5) * (deriv(2, C, F, F)(i, j, k, grid, v) + deriv(1, C, F, C)(i, j, k, grid, w))


def _zero(i, j, k, grid, *a):
    return grid.FT(0)


# ---- eddy viscosity of SmagorinskyLilly (smagorinsky_lilly.jl:85-170) ----------------------------------------------
def _sq(S):
    return lambda i, j, k, grid, u, v, w: S(i, j, k, grid, u, v, w) ** 2


def ΣijΣij_ccc(i, j, k, grid, u, v, w):
    """ΣᵢⱼΣᵢⱼᶜᶜᶜ (smagorinsky_lilly.jl:146-153): tr_Σ² + 2 ℑxyᶜᶜᵃ(Σ₁₂²) + 2 ℑxzᶜᵃᶜ(Σ₁₃²) + 2 ℑyzᵃᶜᶜ(Σ₂₃²); the double
    interpolations are outer(inner) as in interpolation_operators.jl:60-71"""
    tr = Σ11(i, j, k, grid, u, v, w) ** 2 + Σ22(i, j, k, grid, u, v, w) ** 2 + Σ33(i, j, k, grid, u, v, w) ** 2
    return (tr + 2 * Iyc(i, j, k, grid, Ixc, _sq(Σ12), u, v, w)
            + 2 * Izc(i, j, k, grid, Ixc, _sq(Σ13), u, v, w)
            + 2 * Izc(i, j, k, grid, Iyc, _sq(Σ23), u, v, w))


def smagorinsky_viscosity(i, j, k, grid, clo, dz_b, u, v, w):
    """calc_νᶜᶜᶜ (smagorinsky_lilly.jl:99-107); dz_b(i, j, k, grid) is ∂z_b at ccf, or None without buoyancy"""
    FT = grid.FT
    S2 = ΣijΣij_ccc(i, j, k, grid, u, v, w)
    if dz_b is None:
        N2 = FT(0) * S2
    else:
        N2 = np.maximum(FT(0), Izc(i, j, k, grid, dz_b))
    Δf = np.cbrt(grid.Δx(C, i) * grid.Δy(C, j) * grid.Δz(C, k))            # geo_mean_Δᶠ (turbulence_closure_utils.jl:29-30)
    with np.errstate(divide="ignore", invalid="ignore"):
        ς = np.where(S2 == 0, FT(0), np.sqrt(FT(1) - np.minimum(FT(1), FT(clo.Cb) * N2 / S2)))      # stability :83-87
    return ς * (FT(clo.C) * Δf) ** 2 * np.sqrt(2 * S2)                     # νₑ_deardorff :97


def _ν_at(loc, νe):
    """νᶜᶜᶜ / νᶠᶠᶜ / νᶠᶜᶠ / νᶜᶠᶠ of a cell-centred viscosity array (closure_kernel_operators.jl:84-90)"""
    if loc == (C, C, C):
        return lambda i, j, k, grid: νe[i, j, k]
    if loc == (F, F, C):
        return lambda i, j, k, grid: Iyf(i, j, k, grid, Ixf, νe)
    if loc == (F, C, F):
        return lambda i, j, k, grid: Izf(i, j, k, grid, Ixf, νe)
    if loc == (C, F, F):
        return lambda i, j, k, grid: Izf(i, j, k, grid, Iyf, νe)
    raise ValueError(loc)


def viscous_flux(comp, d, clo, νe=None):
    """viscous_flux_{u,v,w}{x,y,z} for closure formulation (abstract_scalar_diffusivity_closure.jl:172-193)."""
    form = clo.formulation
    Σ = {(0, 0): Σ11, (0, 1): Σ12, (0, 2): Σ13, (1, 0): Σ12, (1, 1): Σ22, (1, 2): Σ23,
         (2, 0): Σ13, (2, 1): Σ23, (2, 2): Σ33}[(comp, d)]
    if isinstance(clo, SmagorinskyLilly):          # viscosity(::SmagorinskyLilly, K) = K.νₑ, interpolated to the flux location
        νloc = _ν_at(_FLUXLOC[(comp, d)], νe)
        return lambda i, j, k, grid, u, v, w: -2 * (νloc(i, j, k, grid) * Σ(i, j, k, grid, u, v, w))

    def ν(grid):
        return grid.FT(clo.ν)
    if form == THREE_D:
        return lambda i, j, k, grid, u, v, w: -2 * (ν(grid) * Σ(i, j, k, grid, u, v, w))
    if form == HORIZONTAL:
        if (comp, d) in ((0, 0), (1, 1)):
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * div_xy_ccc(i, j, k, grid, u, v))
        if (comp, d) == (1, 0):
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * zeta3_ffc(i, j, k, grid, u, v))
        if (comp, d) == (0, 1):
            return lambda i, j, k, grid, u, v, w: +(ν(grid) * zeta3_ffc(i, j, k, grid, u, v))
        if (comp, d) == (2, 0):
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(0, F, C, F)(i, j, k, grid, w))
        if (comp, d) == (2, 1):
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(1, C, F, F)(i, j, k, grid, w))
        return _zero
    if form == VERTICAL:
        if d != 2:
            return _zero
        if comp == 0:
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(2, F, C, F)(i, j, k, grid, u))
        if comp == 1:
            return lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(2, C, F, F)(i, j, k, grid, v))
        return lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(2, C, C, C)(i, j, k, grid, w))
    raise ValueError(form)


# flux locations: (component, direction) -> location of the flux
_FLUXLOC = {(0, 0): (C, C, C), (0, 1): (F, F, C), (0, 2): (F, C, F),
            (1, 0): (F, F, C), (1, 1): (C, C, C), (1, 2): (C, F, F),
            (2, 0): (F, C, F), (2, 1): (C, F, F), (2, 2): (C, C, C)}


def div_τ(comp, i, j, k, grid, clo, u, v, w, νe=None):
    """∂ⱼ_τ₁ⱼ, ∂ⱼ_τ₂ⱼ, ∂ⱼ_τ₃ⱼ (closure_kernel_operators.jl:22-41)."""
    if clo is None:
        return grid.FT(0)
    loc = [C, C, C]
    loc[comp] = F
    terms = []
    for d in range(3):
        fl = viscous_flux(comp, d, clo, νe)
        A = area(d, *_FLUXLOC[(comp, d)])
        delta = DELTA[F if d == comp else C][d]
        terms.append(delta(i, j, k, grid,
                           lambda i, j, k, grid, A=A, fl=fl: A(i, j, k, grid) * fl(i, j, k, grid, u, v, w)))
    return 1 / grid.V(i, j, k, *loc) * (terms[0] + terms[1] + terms[2])


def div_q(i, j, k, grid, clo, κ, c, νe=None):
    """∇_dot_qᶜ (closure_kernel_operators.jl:43-48) with diffusive_flux_{x,y,z} =
    -κ ∂c (abstract_scalar_diffusivity_closure.jl:205-207).  SmagorinskyLilly: κ is the Prandtl number and the
    diffusivity κₑ = νₑ / Pr (an operation evaluated at cell centres, smagorinsky_lilly.jl:205-221) is interpolated
    to the flux location (closure_kernel_operators.jl:88-90)."""
    if clo is None:
        return grid.FT(0)
    form = clo.formulation
    κ = grid.FT(κ)
    locs = ((F, C, C), (C, F, C), (C, C, F))
    terms = []
    for d in range(3):
        active = (form == THREE_D) or (form == HORIZONTAL and d < 2) or (form == VERTICAL and d == 2)
        A = area(d, *locs[d])
        if isinstance(clo, SmagorinskyLilly):
            D = deriv(d, *locs[d])
            κe = lambda i, j, k, grid: νe[i, j, k] / κ
            κloc = (Ixf, Iyf, Izf)[d]
            fl = lambda i, j, k, grid, A=A, D=D, κloc=κloc: A(i, j, k, grid) * (-κloc(i, j, k, grid, κe) * D(i, j, k, grid, c))
        elif active:
            D = deriv(d, *locs[d])
            fl = lambda i, j, k, grid, A=A, D=D: A(i, j, k, grid) * (-κ * D(i, j, k, grid, c))
        else:
            fl = lambda i, j, k, grid, A=A: A(i, j, k, grid) * grid.FT(0)
        terms.append(DELTA[C][d](i, j, k, grid, fl))
    return 1 / grid.V(i, j, k, C, C, C) * (terms[0] + terms[1] + terms[2])
