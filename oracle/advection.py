"""
Advection schemes and flux-form advection operators (test infrastructure -- see
oracle/__init__.py).

Follows Advection/weno_fifth_order.jl (constants :12-19, stencils :266-272, p-sums :299-305,
smoothness :311-317, weights :380-403, interpolation :489-498, coefficients :518-532,
stretched tables :562-584,740-772), centered_fourth_order.jl:17-33, centered_second_order.jl:16-32,
upwind_biased_{first,third,fifth}_order.jl, centered_advective_fluxes.jl:15-33,
upwind_biased_advective_fluxes.jl:10-128, topologically_conditional_interpolation.jl:19-80,
momentum_advection_operators.jl:52-86 and tracer_advection_operators.jl:31-35.

Generic naming: dimension d in {0,1,2}; location 'c'/'f' of the RESULT along d.  The
reference's e.g. `_left_biased_interpolate_yᵃᶠᵃ(i,j,k,grid,scheme,ψ)` is
`biased(LEFT, 1, Face, i, j, k, grid, scheme, ψ)` here.
"""
import numpy as np

from .grids import Bounded, Flat, Center, Face, flip
from .operators import DELTA, INTERP, A_q, area, val, _sh

LEFT, RIGHT = "left", "right"


class _Scheme:
    buffer = 0            # Nᴮ, the AbstractAdvectionScheme{Buffer} type parameter
    upwind = False

    @property
    def required_halo(self):
        return self.buffer + 1     # Advection.jl:40


class CenteredSecondOrder(_Scheme):
    buffer = 0


class CenteredFourthOrder(_Scheme):
    buffer = 1


class UpwindBiasedFirstOrder(_Scheme):
    buffer = 1
    upwind = True


class UpwindBiasedThirdOrder(_Scheme):
    buffer = 1
    upwind = True


class UpwindBiasedFifthOrder(_Scheme):
    buffer = 2
    upwind = True


def interp_weights(r, coord, i, bias, op):
    """ENO coefficients, weno_fifth_order.jl:740-772 (literal)."""
    coeff = []
    for j in range(0, 3):
        c = 0
        for m in range(j + 1, 4):
            num = 0
            for l in range(0, 4):
                if l != m:
                    prod = 1
                    for q in range(0, 4):
                        if q != m and q != l:
                            prod *= (coord[i + bias] - coord[op(i, r - q + 1)])
                    num += prod
            den = 1
            for l in range(0, 4):
                if l != m:
                    den *= (coord[op(i, r - m + 1)] - coord[op(i, r - l + 1)])
            c += num / den
        coeff.append(c * (coord[op(i, r - j)] - coord[op(i, r - j + 1)]))
    return tuple(coeff)


def calc_interpolating_coefficients(FT, coord, N):
    """weno_fifth_order.jl:562-584: four tables (r = -1, 0, 1, 2), each indexed 0..N+1,
    of 3-tuples converted to FT.  `coord` is an OffsetVector of nodes with halo >= 4."""
    tables = []
    for r in (-1, 0, 1, 2):
        t = np.zeros((N + 2, 3), dtype=FT)
        for i in range(0, N + 2):
            w = interp_weights(r, coord, i, 0, lambda a, b: a - b)
            t[i] = [FT(x) for x in w]
        tables.append(t)
    return tables


class WENO5(_Scheme):
    """WENO5(FT=Float64; grid=nothing, zweno=true) (weno_fifth_order.jl:164-180).  NOTE the
    reference's default is Z-WENO (`zweno = true`, :167) whatever its docstring says.
    With `grid`, stretched dimensions get per-index ENO coefficient tables
    (compute_stretched_weno_coefficients :182-209); smoothness stays uniform
    (stretched_smoothness=false)."""
    buffer = 2
    upwind = True

    def __init__(self, FT=np.float64, grid=None, zweno=True):
        if grid is not None:
            FT = grid.FT
        self.FT = np.dtype(FT).type
        self.zweno = zweno
        # coeff[d][loc] -> list of 4 tables or None
        self.coeff = [{Face: None, Center: None} for _ in range(3)]
        if grid is not None:
            g4 = grid.with_halo((4, 4, 4))
            for d in range(3):
                if not g4.regular[d]:
                    self.coeff[d][Face] = calc_interpolating_coefficients(self.FT, g4.nodesF[d], g4.N[d])
                    self.coeff[d][Center] = calc_interpolating_coefficients(self.FT, g4.nodesC[d], g4.N[d])


centered_fourth_order = CenteredFourthOrder()


# ---------------------------------------------------------------------------------------
# raw interpolants
# ---------------------------------------------------------------------------------------
def I3(d, loc, i, j, k, grid, c):
    """ℑ³xᶠᵃᵃ / ℑ³xᶜᵃᵃ, centered_fourth_order.jl:17-24: c[i] - δ(δ(c)) / 6."""
    return val(c, i, j, k, grid) - DELTA[loc][d](i, j, k, grid, DELTA[flip(loc)][d], c) / 6


def symmetric_interpolate(d, loc, i, j, k, grid, scheme, c):
    if isinstance(scheme, (CenteredFourthOrder, UpwindBiasedFifthOrder, WENO5)):
        # centered_fourth_order.jl:26-33; WENO5/U5 delegate to it (weno_fifth_order.jl:240-246)
        return INTERP[loc][d](i, j, k, grid, lambda i, j, k, grid, c: I3(d, flip(loc), i, j, k, grid, c), c)
    # C2 / U1 / U3: second order
    return INTERP[loc][d](i, j, k, grid, c)


def _weno(side, d, i, j, k, grid, scheme, ψ, idx, loc):
    """weno_{left,right}_biased_interpolate_xᶠᵃᵃ(i, j, k, grid, scheme, ψ, idx, loc)
    (weno_fifth_order.jl:489-498) with the weights of :380-403."""
    FT = scheme.FT

    def s(n):
        return val(ψ, *_sh((i, j, k), d, n), grid)
    if side == LEFT:      # left_stencil_x :266
        ψt = ((s(-3), s(-2), s(-1)), (s(-2), s(-1), s(0)), (s(-1), s(0), s(1)))
    else:                 # right_stencil_x :270
        ψt = ((s(-2), s(-1), s(0)), (s(-1), s(0), s(1)), (s(0), s(1), s(2)))
    ψ2, ψ1, ψ0 = ψt      # :381, :493

    c1312, c14 = FT(13 / 12), FT(1 / 4)

    def curv(p):
        return c1312 * (p[0] - 2 * p[1] + p[2]) ** 2

    def b_3m4p1(p):      # FT(1/4) * (3ψ1 - 4ψ2 + ψ3)^2
        return curv(p) + c14 * (3 * p[0] - 4 * p[1] + p[2]) ** 2

    def b_1m1(p):        # FT(1/4) * (ψ1 - ψ3)^2
        return curv(p) + c14 * (p[0] - p[2]) ** 2

    def b_1m4p3(p):      # FT(1/4) * (ψ1 - 4ψ2 + 3ψ3)^2
        return curv(p) + c14 * (p[0] - 4 * p[1] + 3 * p[2]) ** 2

    if side == LEFT:      # :311-313
        β0, β1, β2 = b_3m4p1(ψ0), b_1m1(ψ1), b_1m4p3(ψ2)
        C = (FT(3 / 10), FT(3 / 5), FT(1 / 10))         # (C3₀, C3₁, C3₂) :368
    else:                 # :315-317  (NOT the mirror image of the left formulas)
        β0, β1, β2 = b_1m4p3(ψ0), b_1m1(ψ1), b_3m4p1(ψ2)
        C = (FT(1 / 10), FT(3 / 5), FT(3 / 10))         # (C3₂, C3₁, C3₀)
    ε = FT(1e-6)
    if scheme.zweno:      # :386-390
        τ5 = abs(β2 - β0)
        α0 = C[0] * (1 + (τ5 / (β0 + ε)) ** 2)
        α1 = C[1] * (1 + (τ5 / (β1 + ε)) ** 2)
        α2 = C[2] * (1 + (τ5 / (β2 + ε)) ** 2)
    else:                 # :392-394
        α0 = C[0] / (β0 + ε) ** 2
        α1 = C[1] / (β1 + ε) ** 2
        α2 = C[2] / (β2 + ε) ** 2
    Σα = α0 + α1 + α2
    w0, w1, w2 = α0 / Σα, α1 / Σα, α2 / Σα

    tab = scheme.coeff[d][loc]
    if tab is None:       # uniform coefficients :518-524
        lp0 = (FT(1 / 3), FT(5 / 6), -FT(1 / 6))
        lp1 = (-FT(1 / 6), FT(5 / 6), FT(1 / 3))
        lp2 = (FT(1 / 3), -FT(7 / 6), FT(11 / 6))
        if side == LEFT:
            c0, c1, c2 = lp0, lp1, lp2
        else:
            c0, c1, c2 = lp2[::-1], lp1[::-1], lp0[::-1]
    else:                 # retrieve_coeff :526-539: table[r+2][idx]
        def get(r):
            t = tab[r + 1][idx.lo: idx.hi + 1]            # tables are indexed 0..N+1
            shape = [1, 1, 1]
            shape[d] = t.shape[0]
            return tuple(t[:, m].reshape(shape) for m in range(3))
        if side == LEFT:
            c0, c1, c2 = get(0), get(1), get(2)
        else:
            c0, c1, c2 = get(-1), get(0), get(1)

    def p(c, q):          # sum(coeff .* ψ) :299-305, left to right
        return c[0] * q[0] + c[1] * q[1] + c[2] * q[2]
    return w0 * p(c0, ψ0) + w1 * p(c1, ψ1) + w2 * p(c2, ψ2)


def biased_interpolate(side, d, loc, i, j, k, grid, scheme, ψ):
    """left/right_biased_interpolate_{x,y,z}{ᶠ,ᶜ}."""
    if loc == Center:
        # *_xᶜᵃᵃ(i) = *_xᶠᵃᵃ(i+1)  (weno :257-263 keeps idx = i, loc = Center for the tables)
        ijk = _sh((i, j, k), d, 1)
    else:
        ijk = [i, j, k]
    idx = (i, j, k)[d]
    if isinstance(scheme, WENO5):
        return _weno(side, d, *ijk, grid, scheme, ψ, idx, loc)

    def s(n):
        return val(ψ, *_sh(ijk, d, n), grid)
    if isinstance(scheme, UpwindBiasedFifthOrder):   # upwind_biased_fifth_order.jl:24-46
        if side == LEFT:
            return (-3 * s(1) + 27 * s(0) + 47 * s(-1) - 13 * s(-2) + 2 * s(-3)) / 60
        return (2 * s(2) - 13 * s(1) + 47 * s(0) + 27 * s(-1) - 3 * s(-2)) / 60
    if isinstance(scheme, UpwindBiasedThirdOrder):
        if side == LEFT:
            return (2 * s(0) + 5 * s(-1) - s(-2)) / 6
        return (-s(1) + 5 * s(0) + 2 * s(-1)) / 6
    if isinstance(scheme, UpwindBiasedFirstOrder):
        return s(-1) if side == LEFT else s(0)
    raise TypeError(scheme)


# ---------------------------------------------------------------------------------------
# topologically conditional interpolation (the underscore-prefixed functions)
# ---------------------------------------------------------------------------------------
def _outside(kind, idx, N, NB):
    """topologically_conditional_interpolation.jl:19-21 (literal predicates)."""
    if kind == "symmetric":
        return (idx > NB) & (idx < N + 1 - NB)
    if kind == LEFT:
        return (idx > NB) & (idx < N + 1 - (NB - 1))
    return (idx > NB - 1) & (idx < N + 1 - NB)


def _conditional(kind, d, loc, i, j, k, grid, scheme, ψ, high):
    if grid.topology[d] != Bounded:
        return high()
    idx = (i, j, k)[d].arr(d)
    mask = _outside(kind, idx, grid.N[d], scheme.buffer)
    low = INTERP[loc][d](i, j, k, grid, ψ)
    return np.where(mask, high(), low)     # ifelse evaluates both branches


def _symmetric(d, loc, i, j, k, grid, scheme, ψ):
    return _conditional("symmetric", d, loc, i, j, k, grid, scheme, ψ,
                        lambda: symmetric_interpolate(d, loc, i, j, k, grid, scheme, ψ))


def _biased(side, d, loc, i, j, k, grid, scheme, ψ):
    return _conditional(side, d, loc, i, j, k, grid, scheme, ψ,
                        lambda: biased_interpolate(side, d, loc, i, j, k, grid, scheme, ψ))


# ---------------------------------------------------------------------------------------
# fluxes
# ---------------------------------------------------------------------------------------
def upwind_biased_product(u, ψL, ψR):
    """upwind_biased_advective_fluxes.jl:10."""
    return ((u + abs(u)) * ψL + (u - abs(u)) * ψR) / 2


def _nat(a):
    loc = [Center, Center, Center]
    loc[a] = Face
    return loc


def advective_momentum_flux(a, b, i, j, k, grid, scheme, Ua, ψ):
    """advective_momentum_flux_{U,V,W}{u,v,w}: advection OF component b BY component a."""
    if a == b:
        floc = [Center, Center, Center]
        ul, ψl = Center, Center
        ud = a
    else:
        floc = [Center, Center, Center]
        floc[a] = Face
        floc[b] = Face
        ul, ψl = Face, Face
        ud = b
    if isinstance(scheme, CenteredSecondOrder):       # centered_second_order.jl:16-26
        Aq = A_q(a, *_nat(a))
        return INTERP[ul][ud](i, j, k, grid, Aq, Ua) * INTERP[ψl][a](i, j, k, grid, ψ)
    A = area(a, *floc)(i, j, k, grid)
    ũ = _symmetric(ud, ul, i, j, k, grid, scheme, Ua)
    if scheme.upwind:                                   # upwind_biased_advective_fluxes.jl:18-97
        ψL = _biased(LEFT, a, ψl, i, j, k, grid, scheme, ψ)
        ψR = _biased(RIGHT, a, ψl, i, j, k, grid, scheme, ψ)
        return A * upwind_biased_product(ũ, ψL, ψR)
    # centered_advective_fluxes.jl:15-27
    return A * ũ * _symmetric(a, ψl, i, j, k, grid, scheme, ψ)


def advective_tracer_flux(a, i, j, k, grid, scheme, Ua, c):
    """advective_tracer_flux_{x,y,z}."""
    floc = _nat(a)
    if isinstance(scheme, CenteredSecondOrder):       # centered_second_order.jl:28-32
        return A_q(a, *floc)(i, j, k, grid, Ua) * INTERP[Face][a](i, j, k, grid, c)
    if scheme.upwind:                                   # upwind_biased_advective_fluxes.jl:103-128
        ũ = Ua[i, j, k]
        cL = _biased(LEFT, a, Face, i, j, k, grid, scheme, c)
        cR = _biased(RIGHT, a, Face, i, j, k, grid, scheme, c)
        return area(a, *floc)(i, j, k, grid) * upwind_biased_product(ũ, cL, cR)
    # centered_advective_fluxes.jl:29-33
    return A_q(a, *floc)(i, j, k, grid, Ua) * _symmetric(a, Face, i, j, k, grid, scheme, c)


def div_Uu(b, i, j, k, grid, scheme, U, ψ):
    """div_𝐯u / div_𝐯v / div_𝐯w (b = 0, 1, 2), momentum_advection_operators.jl:52-86."""
    if scheme is None:
        return grid.FT(0)
    loc = _nat(b)
    terms = []
    for a in range(3):
        delta = DELTA[Face if a == b else Center][a]
        terms.append(delta(i, j, k, grid,
                           lambda i, j, k, grid, a=a: advective_momentum_flux(a, b, i, j, k, grid, scheme, U[a], ψ)))
    return 1 / grid.V(i, j, k, *loc) * (terms[0] + terms[1] + terms[2])


def div_Uc(i, j, k, grid, scheme, U, c):
    """tracer_advection_operators.jl:31-35."""
    if scheme is None:
        return grid.FT(0)
    terms = []
    for a in range(3):
        terms.append(DELTA[Center][a](i, j, k, grid,
                                      lambda i, j, k, grid, a=a: advective_tracer_flux(a, i, j, k, grid, scheme, U[a], c)))
    return 1 / grid.V(i, j, k, Center, Center, Center) * (terms[0] + terms[1] + terms[2])
