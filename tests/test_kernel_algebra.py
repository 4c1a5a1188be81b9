"""The specialised tendency kernels (csrc/weno_fast.cuh) evaluate the WENO5 reconstruction as a
re-arranged rational function (shared second differences, 4*beta, one folded reciprocal).  This
CPU test restates THAT algebra in numpy, operation for operation, and checks it against the oracle's
literal restatement of weno_fifth_order.jl -- so a wrong identity is caught here, without a GPU.
The GPU parity tests then check the compiled kernels themselves."""
import numpy as np
import pytest

from oracle import RectilinearGrid, Field, fill_halo_regions, WENO5, Periodic
from oracle.advection import biased_interpolate, LEFT, RIGHT
from oracle.fields import R


def weno_face_kernel_algebra(a, b, c, d, e, x2, x0, zweno):
    """csrc/weno_fast.cuh: wf::weno_face (Float64 branch), with an exact reciprocal."""
    t2, t1, t0 = (-2 * b + a) + c, (-2 * c + b) + d, (-2 * d + c) + e
    s2, s1, s0 = 2 * (x2 - b) + t2, b - d, 2 * (x0 - d) + t0
    k133, eps4 = 13.0 / 3.0, 4.0e-6
    B2, B1, B0 = s2 * s2 + k133 * (t2 * t2), s1 * s1 + k133 * (t1 * t1), s0 * s0 + k133 * (t0 * t0)
    r2 = 3 * t1 - 2 * t2
    m = 0.5 * (c + d)
    E0, E1, E2 = (B0 + eps4) ** 2, (B1 + eps4) ** 2, (B2 + eps4) ** 2
    P12, P02, P01 = E1 * E2, E0 * E2, E0 * E1
    if zweno:
        tt, PI = (B2 - B0) ** 2, E0 * P12
        g0, g1, g2 = 3 * (tt * P12 + PI), 6 * (tt * P02 + PI), tt * P01 + PI
    else:
        g0, g1, g2 = 3 * P12, 6 * P02, P01
    den = (g0 + g1) + g2
    S = g0 * t0 + (g1 * t1 + g2 * r2)
    return (S / den) * (-1.0 / 6.0) + m


def weno_upwind_kernel_algebra(pos, w, zweno):
    """csrc/weno_fast.cuh: wf::weno_upwind -- w[n] = psi[f-3+n], n = 0..5."""
    sel = lambda x, y: np.where(pos, x, y)
    c = sel(w[2], w[3])
    return weno_face_kernel_algebra(sel(w[0], w[5]), sel(w[1], w[4]), c, sel(w[3], w[2]), sel(w[4], w[1]),
                                    sel(c, w[5]), sel(c, w[1]), zweno)


@pytest.mark.parametrize("zweno", [True, False])
@pytest.mark.parametrize("scale", [1.0, 1e-4, 1e3])
@pytest.mark.parametrize("kind", ["random", "smooth"])
def test_kernel_weno_algebra_matches_oracle(zweno, scale, kind):
    N = 256
    g = RectilinearGrid(size=(N, 1, 1), x=(0, 1), y=(0, 1), z=(0, 1), topology=(Periodic, Periodic, Periodic))
    rng = np.random.default_rng(11)
    if kind == "random":
        a = scale * rng.uniform(-1, 1, N)
    else:
        x = (np.arange(N) + 0.5) / N
        a = scale * (np.sin(2 * np.pi * x) + 0.3 * np.cos(6 * np.pi * x + 0.2) + 1e-3 * rng.uniform(-1, 1, N))
    f = Field(g)
    f.set(a.reshape(N, 1, 1))
    fill_halo_regions(f)
    sch = WENO5(zweno=zweno)
    left = biased_interpolate(LEFT, 0, "f", R(1, N), R(1), R(1), g, sch, f)[:, 0, 0]
    right = biased_interpolate(RIGHT, 0, "f", R(1, N), R(1), R(1), g, sch, f)[:, 0, 0]
    # face i (1-based) sees psi[i-3 .. i+2]; periodic wrap
    w = [np.roll(a, -n) for n in (-3, -2, -1, 0, 1, 2)]
    mine_l = weno_upwind_kernel_algebra(np.ones(N, bool), w, zweno)
    mine_r = weno_upwind_kernel_algebra(np.zeros(N, bool), w, zweno)
    ref = np.max(np.abs(a))
    assert np.max(np.abs(mine_l - left)) <= 2e-14 * ref
    assert np.max(np.abs(mine_r - right)) <= 2e-14 * ref


def test_halley_reciprocal_reaches_double_precision():
    """wf::weno_face refines a 2^-23 reciprocal seed with ONE cubic step: q = S r0 (1 + e + e^2), e = 1 - den r0."""
    rng = np.random.default_rng(3)
    den = np.exp(rng.uniform(-70, 120, 100000))
    S = rng.uniform(-1, 1, den.size) * den
    r0 = (1.0 / den) * (1 + rng.uniform(-1, 1, den.size) * 2.0 ** -23)     # seed with the MUFU error bound
    e = 1.0 - den * r0
    q0 = S * r0
    q = q0 * (e * e + e) + q0
    assert np.max(np.abs(q - S / den) / np.abs(S / den)) < 4e-16


def test_interp4_is_the_average_of_two_fourth_order_interpolants():
    """wf::interp4: (I(c0) + I(c1))/2 with I(c) = c - (c+ - 2c + c-)/6 equals 7/12 (c0 + c1) - 1/12 (c- + c2)."""
    rng = np.random.default_rng(5)
    cm, c0, c1, c2 = rng.uniform(-1, 1, (4, 10000))
    I = lambda m, c, p: c - ((p - c) - (c - m)) / 6
    ref = 0.5 * (I(cm, c0, c1) + I(c0, c1, c2))
    mine = (-1.0 / 12.0) * (cm + c2) + (7.0 / 12.0) * (c0 + c1)
    assert np.max(np.abs(mine - ref)) < 5e-16
