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

    def __init__(self, formulation=THREE_D, ν=0.0, κ=0.0, time_discretization="Explicit"):
        self.formulation = formulation
        self.ν = ν
        self.κ = κ
        self.required_halo = 1
        # ExplicitTimeDiscretization / VerticallyImplicitTimeDiscretization (implicit_explicit_time_discretization.jl)
        assert time_discretization in ("Explicit", "VerticallyImplicit")
        assert not (time_discretization == "VerticallyImplicit" and formulation == HORIZONTAL)
        self.time_discretization = time_discretization

    @property
    def vitd(self):
        return self.time_discretization == "VerticallyImplicit"

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
    if hasattr(clo, "Cν") or isinstance(clo, SmagorinskyLilly):      # viscosity(::SmagorinskyLilly / ::AMD, K) = K.νₑ, interpolated to the flux location
        νloc = _ν_at(_FLUXLOC[(comp, d)], νe)
        return lambda i, j, k, grid, u, v, w: -2 * (νloc(i, j, k, grid) * Σ(i, j, k, grid, u, v, w))

    def ν(grid):
        return grid.FT(clo.ν)
    if getattr(clo, "vitd", False) and d == 2:
        # VerticallyImplicitTimeDiscretization on a vertically Bounded grid (abstract_scalar_diffusivity_closure.jl:232-255):
        # the explicit flux only at k == 1 and k == Nz + 1, elsewhere what stays explicit: -ν ∂x w / -ν ∂y w (u, v), 0 (w)
        import copy
        ex = copy.copy(clo)
        ex.time_discretization = "Explicit"
        explicit = viscous_flux(comp, d, ex, νe)
        if comp == 0:
            ivd = lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(0, F, C, F)(i, j, k, grid, w))
        elif comp == 1:
            ivd = lambda i, j, k, grid, u, v, w: -(ν(grid) * deriv(1, C, F, F)(i, j, k, grid, w))
        else:
            ivd = lambda i, j, k, grid, u, v, w: grid.FT(0)

        def fl(i, j, k, grid, u, v, w):
            kk = k.arr(2)
            edge = (kk == 1) | (kk == grid.Nz + 1)
            return np.where(edge, explicit(i, j, k, grid, u, v, w), ivd(i, j, k, grid, u, v, w))
        return fl
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


def div_q(i, j, k, grid, clo, κ, c, νe=None, κe_field=None):
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
        if κe_field is not None:        # AnisotropicMinimumDissipation: diffusivity(::AMD, K, id) = K.κₑ[id], a cell-centred field
            D = deriv(d, *locs[d])
            κloc = (Ixf, Iyf, Izf)[d]
            fl = lambda i, j, k, grid, A=A, D=D, κloc=κloc: A(i, j, k, grid) * (-κloc(i, j, k, grid, κe_field) * D(i, j, k, grid, c))
        elif isinstance(clo, SmagorinskyLilly):
            D = deriv(d, *locs[d])
            κe = lambda i, j, k, grid: νe[i, j, k] / κ
            κloc = (Ixf, Iyf, Izf)[d]
            fl = lambda i, j, k, grid, A=A, D=D, κloc=κloc: A(i, j, k, grid) * (-κloc(i, j, k, grid, κe) * D(i, j, k, grid, c))
        elif active and d == 2 and getattr(clo, "vitd", False):      # diffusive_flux_z with VITD (:251-255)
            D = deriv(d, *locs[d])

            def fl(i, j, k, grid, A=A, D=D):
                kk = k.arr(2)
                edge = (kk == 1) | (kk == grid.Nz + 1)
                return A(i, j, k, grid) * np.where(edge, -κ * D(i, j, k, grid, c), grid.FT(0))
        elif active:
            D = deriv(d, *locs[d])
            fl = lambda i, j, k, grid, A=A, D=D: A(i, j, k, grid) * (-κ * D(i, j, k, grid, c))
        else:
            fl = lambda i, j, k, grid, A=A: A(i, j, k, grid) * grid.FT(0)
        terms.append(DELTA[C][d](i, j, k, grid, fl))
    return 1 / grid.V(i, j, k, C, C, C) * (terms[0] + terms[1] + terms[2])


# ---- AnisotropicMinimumDissipation (anisotropic_minimum_dissipation.jl:180-359, velocity_tracer_gradients.jl:120-242) ------
class AnisotropicMinimumDissipation:
    """AnisotropicMinimumDissipation(FT; C=1/12, Cν=nothing, Cκ=nothing, Cb=nothing) (anisotropic_minimum_dissipation.jl:96-105):
    νₑ = max(0, -Cν δ² (r - Cb ζ) / q), κₑ = max(0, -Cκ δ² ϑ / σ) per tracer, recomputed in update_state!.  Constant
    Poincaré constants only (numbers, or a dict per tracer for Cκ)."""
    formulation = THREE_D

    def __init__(self, C=1 / 12, Cν=None, Cκ=None, Cb=None):
        self.Cν = C if Cν is None else Cν
        self.Cκ = C if Cκ is None else Cκ
        self.Cb = Cb
        self.required_halo = 1

    def Ck(self, name):
        return self.Cκ[name] if isinstance(self.Cκ, dict) else self.Cκ


def _Δf(d):
    """Δᶠx / Δᶠy / Δᶠz at ANY location are the ccc ones evaluated at the index passed (:255-267): 2 Δᶜ"""
    return lambda i, j, k, grid: 2 * grid.spacing(d, C, (i, j, k)[d])


_Δfx, _Δfy, _Δfz = _Δf(0), _Δf(1), _Δf(2)
# plain gradients at their natural locations (velocity_tracer_gradients.jl:6-22)
_dx_u = lambda i, j, k, g, u, v, w: deriv(0, C, C, C)(i, j, k, g, u)
_dy_v = lambda i, j, k, g, u, v, w: deriv(1, C, C, C)(i, j, k, g, v)
_dz_w = lambda i, j, k, g, u, v, w: deriv(2, C, C, C)(i, j, k, g, w)
_dx_v = lambda i, j, k, g, u, v, w: deriv(0, F, F, C)(i, j, k, g, v)
_dy_u = lambda i, j, k, g, u, v, w: deriv(1, F, F, C)(i, j, k, g, u)
_dx_w = lambda i, j, k, g, u, v, w: deriv(0, F, C, C)(i, j, k, g, w)          # ∂x_w = ∂xᶠᶜᶜ (velocity_tracer_gradients.jl:16)
_dz_u = lambda i, j, k, g, u, v, w: deriv(2, F, C, F)(i, j, k, g, u)
_dy_w = lambda i, j, k, g, u, v, w: deriv(1, C, F, C)(i, j, k, g, w)
_dz_v = lambda i, j, k, g, u, v, w: deriv(2, C, F, F)(i, j, k, g, v)
# normalised gradients (:125-149)
n_dx_u, n_dy_v, n_dz_w = _dx_u, _dy_v, _dz_w
n_dx_v = lambda i, j, k, g, u, v, w: _Δfx(i, j, k, g) / _Δfy(i, j, k, g) * _dx_v(i, j, k, g, u, v, w)
n_dy_u = lambda i, j, k, g, u, v, w: _Δfy(i, j, k, g) / _Δfx(i, j, k, g) * _dy_u(i, j, k, g, u, v, w)
n_dx_w = lambda i, j, k, g, u, v, w: _Δfx(i, j, k, g) / _Δfz(i, j, k, g) * _dx_w(i, j, k, g, u, v, w)
n_dz_u = lambda i, j, k, g, u, v, w: _Δfz(i, j, k, g) / _Δfx(i, j, k, g) * _dz_u(i, j, k, g, u, v, w)
n_dy_w = lambda i, j, k, g, u, v, w: _Δfy(i, j, k, g) / _Δfz(i, j, k, g) * _dy_w(i, j, k, g, u, v, w)
n_dz_v = lambda i, j, k, g, u, v, w: _Δfz(i, j, k, g) / _Δfy(i, j, k, g) * _dz_v(i, j, k, g, u, v, w)
n_S11, n_S22, n_S33 = n_dx_u, n_dy_v, n_dz_w
n_S12 = lambda i, j, k, g, u, v, w: g.FT(0.5) * (n_dy_u(i, j, k, g, u, v, w) + n_dx_v(i, j, k, g, u, v, w))
n_S13 = lambda i, j, k, g, u, v, w: g.FT(0.5) * (n_dz_u(i, j, k, g, u, v, w) + n_dx_w(i, j, k, g, u, v, w))
n_S23 = lambda i, j, k, g, u, v, w: g.FT(0.5) * (n_dz_v(i, j, k, g, u, v, w) + n_dy_w(i, j, k, g, u, v, w))


def _prod(a, b):
    return lambda i, j, k, g, u, v, w: a(i, j, k, g, u, v, w) * b(i, j, k, g, u, v, w)


def _Ixy(i, j, k, g, f, *a):      # ℑxyᶜᶜᵃ = ℑyᵃᶜᵃ(ℑxᶜᵃᵃ)
    return Iyc(i, j, k, g, Ixc, f, *a)


def _Ixz(i, j, k, g, f, *a):      # ℑxzᶜᵃᶜ = ℑzᵃᵃᶜ(ℑxᶜᵃᵃ)
    return Izc(i, j, k, g, Ixc, f, *a)


def _Iyz(i, j, k, g, f, *a):      # ℑyzᵃᶜᶜ = ℑzᵃᵃᶜ(ℑyᵃᶜᵃ)
    return Izc(i, j, k, g, Iyc, f, *a)


def amd_norm_tr_grad_u(i, j, k, g, u, v, w):
    """norm_tr_∇uᶜᶜᶜ (:315-328)"""
    a = (u, v, w)
    return (n_dx_u(i, j, k, g, *a) ** 2 + n_dy_v(i, j, k, g, *a) ** 2 + n_dz_w(i, j, k, g, *a) ** 2
            + _Ixy(i, j, k, g, _sq(n_dx_v), *a) + _Ixy(i, j, k, g, _sq(n_dy_u), *a)
            + _Ixz(i, j, k, g, _sq(n_dx_w), *a) + _Ixz(i, j, k, g, _sq(n_dz_u), *a)
            + _Iyz(i, j, k, g, _sq(n_dy_w), *a) + _Iyz(i, j, k, g, _sq(n_dz_v), *a))


def amd_r(i, j, k, g, u, v, w):
    """norm_uᵢₐ_uⱼₐ_Σᵢⱼᶜᶜᶜ (:269-313), term order kept"""
    a = (u, v, w)
    q = (i, j, k, g)
    t1 = (n_S11(*q, *a) * n_dx_u(*q, *a) ** 2
          + n_S22(*q, *a) * _Ixy(*q, _sq(n_dx_v), *a)
          + n_S33(*q, *a) * _Ixz(*q, _sq(n_dx_w), *a)
          + 2 * n_dx_u(*q, *a) * _Ixy(*q, _prod(n_dx_v, n_S12), *a)
          + 2 * n_dx_u(*q, *a) * _Ixz(*q, _prod(n_dx_w, n_S13), *a)
          + 2 * _Ixy(*q, n_dx_v, *a) * _Ixz(*q, n_dx_w, *a) * _Iyz(*q, n_S23, *a))
    t2 = (+ n_S11(*q, *a) * _Ixy(*q, _sq(n_dy_u), *a)
          + n_S22(*q, *a) * n_dy_v(*q, *a) ** 2
          + n_S33(*q, *a) * _Iyz(*q, _sq(n_dy_w), *a)
          + 2 * n_dy_v(*q, *a) * _Ixy(*q, _prod(n_dy_u, n_S12), *a)
          + 2 * _Ixy(*q, n_dy_u, *a) * _Iyz(*q, n_dy_w, *a) * _Ixz(*q, n_S13, *a)
          + 2 * n_dy_v(*q, *a) * _Iyz(*q, _prod(n_dy_w, n_S23), *a))
    t3 = (+ n_S11(*q, *a) * _Ixz(*q, _sq(n_dz_u), *a)
          + n_S22(*q, *a) * _Iyz(*q, _sq(n_dz_v), *a)
          + n_S33(*q, *a) * n_dz_w(*q, *a) ** 2
          + 2 * _Ixz(*q, n_dz_u, *a) * _Iyz(*q, n_dz_v, *a) * _Ixy(*q, n_S12, *a)
          + 2 * n_dz_w(*q, *a) * _Ixz(*q, _prod(n_dz_u, n_S13), *a)
          + 2 * n_dz_w(*q, *a) * _Iyz(*q, _prod(n_dz_v, n_S23), *a))
    return t1 + t2 + t3


def _amd_δ2(i, j, k, g):
    return 3 / (1 / _Δfx(i, j, k, g) ** 2 + 1 / _Δfy(i, j, k, g) ** 2 + 1 / _Δfz(i, j, k, g) ** 2)


def amd_viscosity(i, j, k, g, clo, bp, u, v, w):
    """calc_νᶜᶜᶜ (:180-199); bp(i, j, k, grid) is buoyancy_perturbation, or None (then the Cb term is zero, :330)"""
    FT = g.FT
    a = (u, v, w)
    q = amd_norm_tr_grad_u(i, j, k, g, *a)
    r = amd_r(i, j, k, g, *a)
    if clo.Cb is None or bp is None:
        Cbζ = FT(0) * r
    else:                      # Cb_norm_wᵢ_bᵢᶜᶜᶜ (:332-345) / Δᶠz
        dxb = lambda i, j, k, g: deriv(0, F, C, C)(i, j, k, g, bp)
        dyb = lambda i, j, k, g: deriv(1, C, F, C)(i, j, k, g, bp)
        dzb = lambda i, j, k, g: deriv(2, C, C, F)(i, j, k, g, bp)
        wx = _Ixz(i, j, k, g, n_dx_w, *a) * _Δfx(i, j, k, g) * Ixc(i, j, k, g, dxb)
        wy = _Iyz(i, j, k, g, n_dy_w, *a) * _Δfy(i, j, k, g) * Iyc(i, j, k, g, dyb)
        wz = n_dz_w(i, j, k, g, *a) * _Δfz(i, j, k, g) * Izc(i, j, k, g, dzb)
        Cbζ = FT(clo.Cb) * (wx + wy + wz) / _Δfz(i, j, k, g)
    δ2 = _amd_δ2(i, j, k, g)
    with np.errstate(divide="ignore", invalid="ignore"):
        ν = np.where(q == 0, FT(0), -FT(clo.Cν) * δ2 * (r - Cbζ) / q)
    return np.maximum(FT(0), ν)


def amd_diffusivity(i, j, k, g, Cκ, u, v, w, c):
    """calc_κᶜᶜᶜ (:201-220) with norm_θᵢ²ᶜᶜᶜ (:374-376) and norm_uᵢⱼ_cⱼ_cᵢᶜᶜᶜ (:347-372)"""
    FT = g.FT
    a = (u, v, w)
    ncx = lambda i, j, k, g: _Δfx(i, j, k, g) * deriv(0, F, C, C)(i, j, k, g, c)
    ncy = lambda i, j, k, g: _Δfy(i, j, k, g) * deriv(1, C, F, C)(i, j, k, g, c)
    ncz = lambda i, j, k, g: _Δfz(i, j, k, g) * deriv(2, C, C, F)(i, j, k, g, c)
    sq1 = lambda f: (lambda i, j, k, g: f(i, j, k, g) ** 2)
    q = (i, j, k, g)
    σ = Ixc(*q, sq1(ncx)) + Iyc(*q, sq1(ncy)) + Izc(*q, sq1(ncz))
    cx = (n_dx_u(*q, *a) * Ixc(*q, sq1(ncx))
          + _Ixy(*q, n_dx_v, *a) * Ixc(*q, ncx) * Iyc(*q, ncy)
          + _Ixz(*q, n_dx_w, *a) * Ixc(*q, ncx) * Izc(*q, ncz))
    cy = (_Ixy(*q, n_dy_u, *a) * Iyc(*q, ncy) * Ixc(*q, ncx)
          + n_dy_v(*q, *a) * Iyc(*q, sq1(ncy))
          + _Ixz(*q, n_dy_w, *a) * Iyc(*q, ncy) * Izc(*q, ncz))
    cz = (_Ixz(*q, n_dz_u, *a) * Izc(*q, ncz) * Ixc(*q, ncx)
          + _Iyz(*q, n_dz_v, *a) * Izc(*q, ncz) * Iyc(*q, ncy)
          + n_dz_w(*q, *a) * Izc(*q, sq1(ncz)))
    ϑ = cx + cy + cz
    δ2 = _amd_δ2(i, j, k, g)
    with np.errstate(divide="ignore", invalid="ignore"):
        κ = np.where(σ == 0, FT(0), -FT(Cκ) * δ2 * ϑ / σ)
    return np.maximum(FT(0), κ)


# ---- vertically implicit diffusion (vertically_implicit_diffusion_solver.jl:19-195) --------------------------------------------
def ivd_coefficients(grid, Δt, κ, z_face_field):
    """lower / main / upper diagonals (functions of k only for constant κ and regular x, y) of
    (1 - Δt ∂z κ ∂z) cⁿ⁺¹ = c★ as the reference builds them: ivd_lower_diagonal / ivd_diagonal / ivd_upper_diagonal for a
    field whose z location is Center (u, v, tracers: :27-46) or Face (w: :48-66).  Returned as the a (k = 1..Nz-1), b (1..Nz),
    c (1..Nz-1) vectors of BatchedTridiagonalSolver."""
    from .fields import R
    FT, Nz = grid.FT, grid.Nz
    κ, Δt = FT(κ), FT(Δt)
    dzc = lambda k: np.asarray(grid.Δz(C, R(k))).ravel()[0]
    dzf = lambda k: np.asarray(grid.Δz(F, R(k))).ravel()[0]
    κΔz2 = lambda kc, kf: κ / dzc(kc) / dzf(kf)                   # κ_Δz² :25
    if not z_face_field:
        upper = lambda k: FT(0) if k > Nz - 1 else -Δt * κΔz2(k, k + 1)
        lower = lambda k: FT(0) if k < 1 else -Δt * κΔz2(k + 1, k + 1)
    else:
        upper = lambda k: FT(0) if k < 1 else -Δt * κΔz2(k, k)
        lower = lambda k: FT(0) if k < 1 else -Δt * κΔz2(k + 1, k)
    diag = lambda k: FT(1) - Δt * FT(0) - upper(k) - lower(k - 1)
    a = np.array([lower(k) for k in range(1, Nz)], dtype=FT)
    b = np.array([diag(k) for k in range(1, Nz + 1)], dtype=FT)
    c = np.array([upper(k) for k in range(1, Nz)], dtype=FT)
    return a, b, c
