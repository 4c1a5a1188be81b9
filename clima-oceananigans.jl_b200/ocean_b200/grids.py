"""Host-side mirror of `RectilinearGrid(arch, FT; size, x, y, z, extent, topology, halo)`
(reference src/Grids/rectilinear_grid.jl:249-279) for the `B200()` architecture.

In the Julia shim the reference's own constructor runs unchanged and the shim only marshals
its fields into an `ob200_grid_desc`; here, where no Julia exists, the same host logic is
written out: coordinate generation as in src/Grids/grid_generation.jl:28-112 (regular spacing
computed exactly and rounded once; stretched vectors with the reference's halo-extension
rules), default halo 3 and Flat -> N = 1, H = 0 (src/Grids/input_validation.jl:55-64)."""
import ctypes as C
from fractions import Fraction

import numpy as np

from . import _lib as L
from ._lib import lib, check

Periodic, Bounded, Flat, FullyConnected = "Periodic", "Bounded", "Flat", "FullyConnected"
Center, Face = "Center", "Face"
_TOPO = {Periodic: L.PERIODIC, Bounded: L.BOUNDED, Flat: L.FLAT, FullyConnected: L.FULLY_CONNECTED}


class B200:
    """The new architecture singleton next to CPU()/GPU() (src/Architectures.jl:68-75)."""

    def __init__(self, device=0):
        self.device = device
        check(lib.ob200_init(device))

    def __repr__(self):
        return f"B200(device={self.device})"


class _OV:
    """OffsetVector: numpy parent + Julia index of its first element."""

    def __init__(self, a, first):
        self.a, self.first = np.asarray(a), first

    def __getitem__(self, q):
        return self.a[q - self.first]

    def span(self, lo, hi):
        return self.a[lo - self.first: hi + 1 - self.first]


def _tlen(loc, topo, N, H):
    if topo == Flat:
        return N
    return N + 2 * H + (1 if (loc == Face and topo == Bounded) else 0)


def _regular(FT, topo, N, H, c):
    c1, c2 = Fraction(float(c[0])), Fraction(float(c[1]))
    if not c1 < c2:
        raise ValueError("coordinate interval must be increasing")
    Lr = c2 - c1
    d = Lr / N
    f0 = c1 - H * d
    f1 = f0 + (Lr + 2 * H * d if topo == Bounded else Lr + (2 * H - 1) * d)
    c0 = f0 + d / 2
    cN = c0 + Lr + d * (2 * H - 1)

    def rng(a, b, n):
        a, b = Fraction(float(FT(float(a)))), Fraction(float(FT(float(b))))
        if n == 1:
            return np.array([float(a)], dtype=FT)
        return np.array([float(a + (b - a) * i / (n - 1)) for i in range(n)], dtype=FT)
    F = _OV(rng(f0, f1, _tlen(Face, topo, N, H)), 1 - H)
    Cn = _OV(rng(c0, cN, _tlen(Center, topo, N, H)), 1 - H)
    return FT(float(Lr)), F, Cn, FT(float(d)), FT(float(d))


def _stretched(FT, topo, N, H, coord):
    face = (lambda i: coord(i)) if callable(coord) else (lambda i: coord[i - 1])
    Fi = np.array([face(i) for i in range(1, N + 2)], dtype=FT)
    Lr = Fi[-1] - Fi[0]
    if topo == Bounded:
        lo = [Fi[1] - Fi[0]] * H
        up = [Fi[-1] - Fi[-2]] * H
    else:
        lo = [Fi[N + 1 - H + i - 1] - Fi[N + 1 - H + i - 2] for i in range(1, H + 1)]
        up = [Fi[i] - Fi[i - 1] for i in range(1, H + 1)]
    up = up[::-1]
    Fm = [Fi[0] - np.sum(np.array(lo[i:], dtype=FT)) for i in range(H)]
    Fp = [Fi[-1] + np.sum(np.array(up[i:], dtype=FT)) for i in range(H)][::-1]
    Fall = np.concatenate([np.array(Fm, dtype=FT), Fi, np.array(Fp, dtype=FT)]).astype(FT)
    TC, TF = _tlen(Center, topo, N, H), _tlen(Face, topo, N, H)
    Call = np.array([(Fall[i + 1] + Fall[i]) / 2 for i in range(TC)], dtype=FT)
    dF = np.array([Call[i] - Call[i - 1] for i in range(1, TC)], dtype=FT)
    Fall = Fall[:TF]
    dC = np.array([Fall[i + 1] - Fall[i] for i in range(TF - 1)], dtype=FT)
    dF = np.concatenate([[dF[0]], dF, [dF[-1]]]).astype(FT)
    dF[1:] = dF[:-1].copy()
    return FT(Lr), _OV(Fall, 1 - H), _OV(Call, 1 - H), _OV(dF, -H), _OV(dC, 1 - H)


class RectilinearGrid:
    def __init__(self, architecture=None, FT=np.float64, size=None, x=None, y=None, z=None, extent=None,
                 topology=(Periodic, Periodic, Bounded), halo=None):
        self.global_size, self.multi = None, None
        if hasattr(architecture, "ranks") and hasattr(architecture, "child"):
            # RectilinearGrid(arch::MultiArch, ...) takes GLOBAL size / extent and builds the local slab
            # (reference src/Distributed/distributed_grids.jl:16-60)
            from . import distributed as D
            multi = architecture
            if multi.R > 1:
                if topology[1] != Periodic:
                    raise ValueError("slab decomposition needs a Periodic y topology")
                if extent is not None:
                    ext = (extent,) if np.isscalar(extent) else tuple(extent)
                    it = iter(ext)
                    cs = [x, y, z]
                    for d, t in enumerate(topology):
                        if t != Flat:
                            Ld = float(next(it))
                            cs[d] = (0.0, Ld) if d < 2 else (-Ld, 0.0)
                    x, y, z = cs
                    extent = None
                if not (isinstance(y, tuple) and len(y) == 2):
                    raise ValueError("the decomposed dimension must be regular")
                gsize = (size,) if np.isscalar(size) else tuple(size)
                self.global_size = gsize
                size = D.local_size(gsize, multi.ranks, 1)
                y = D.local_interval(y, multi.R, multi.local_rank)
                topology = (topology[0], FullyConnected, topology[2])
                self.multi = multi
            architecture = multi.child
        if not isinstance(architecture, B200):
            raise TypeError("this package only provides the B200() architecture")
        self.architecture = architecture
        self.FT = FT = np.dtype(FT).type
        if FT not in (np.float32, np.float64):
            raise TypeError("eltype must be Float32 or Float64")
        self.topology = tuple(topology)
        nflat = sum(t == Flat for t in self.topology)
        size = (size,) if np.isscalar(size) else tuple(size)
        if len(size) != 3 - nflat:
            raise ValueError("length(size) must equal the number of non-Flat dimensions")
        halo = (3,) * (3 - nflat) if halo is None else ((halo,) if np.isscalar(halo) else tuple(halo))
        coords = [x, y, z]
        if extent is not None:
            extent = (extent,) if np.isscalar(extent) else tuple(extent)
            it = iter(extent)
        si, hi = iter(size), iter(halo)
        N, H = [], []
        for d, t in enumerate(self.topology):
            if t == Flat:
                N.append(1), H.append(0)
            else:
                N.append(int(next(si))), H.append(int(next(hi)))
                if extent is not None:        # the "oceanic" default domain (Grids/input_validation.jl:92-95): z = (-Lz, 0)
                    Ld = float(next(it))
                    coords[d] = (0.0, Ld) if d < 2 else (-Ld, 0.0)
                if coords[d] is None:
                    raise ValueError("missing coordinate specification")
        self.N, self.H = tuple(N), tuple(H)
        self.Nx, self.Ny, self.Nz = self.N
        self.Hx, self.Hy, self.Hz = self.H
        self._coords = coords
        self.L, self.F, self.C, self.dF, self.dC, self.regular = [], [], [], [], [], []
        for d, t in enumerate(self.topology):
            c = coords[d]
            if t == Flat:
                one = np.ones(1, dtype=FT)
                r = (FT(1), _OV(one, 1), _OV(one, 1), FT(1), FT(1))
                reg = True
            elif isinstance(c, tuple) and len(c) == 2:
                r, reg = _regular(FT, t, N[d], H[d], c), True
            else:
                r, reg = _stretched(FT, t, N[d], H[d], c), False
            for lst, v in zip((self.L, self.F, self.C, self.dF, self.dC), r):
                lst.append(v)
            self.regular.append(reg)
        self.Lx, self.Ly, self.Lz = self.L
        self._make_handle()

    def _make_handle(self):
        d = L.GridDesc()
        d.ftype = L.F32 if self.FT == np.float32 else L.F64
        self._keep = []
        for k in range(3):
            d.N[k], d.H[k] = self.N[k], self.H[k]
            d.topology[k] = _TOPO[self.topology[k]]
            d.L[k] = float(self.L[k])
            d.regular[k] = int(self.regular[k])
            if self.regular[k]:
                d.delta[k] = float(self.dC[k])
            else:
                for name, ov in (("dC", self.dC[k]), ("dF", self.dF[k])):
                    arr = np.ascontiguousarray(ov.a, dtype=np.float64)
                    self._keep.append(arr)
                    getattr(d, name)[k] = arr.ctypes.data_as(C.POINTER(C.c_double))
                    getattr(d, name + "_first")[k] = ov.first
                    getattr(d, name + "_len")[k] = len(arr)
        h = C.c_void_p()
        check(lib.ob200_grid_create(C.byref(d), C.byref(h)))
        self.handle = h

    def with_halo(self, halo):
        """with_halo(new_halo, grid) (src/Grids/rectilinear_grid.jl)."""
        size = tuple(n for n, t in zip(self.N, self.topology) if t != Flat)
        hl = tuple(h for h, t in zip(halo, self.topology) if t != Flat)
        cs = []
        for d, t in enumerate(self.topology):
            if t == Flat:
                cs.append(None)
            elif self.regular[d]:
                cs.append(self._coords[d])
            else:
                cs.append(np.array(self.F[d].span(1, self.N[d] + 1)))
        g = RectilinearGrid(self.architecture, self.FT, size=size, x=cs[0], y=cs[1], z=cs[2],
                            topology=self.topology, halo=hl)
        g.multi, g.global_size = self.multi, self.global_size
        return g

    def local_slice(self):
        """index slices of this rank's slab inside GLOBAL interior arrays"""
        if self.multi is None:
            return (slice(None),) * 3
        r, n = self.multi.local_rank, self.N[1]
        return (slice(None), slice(r * n, (r + 1) * n), slice(None))

    def nodes(self, loc):
        """xnodes/ynodes/znodes of the interior points of a field at `loc`, broadcast-shaped."""
        out = []
        for d in range(3):
            n = self.N[d] + (1 if (loc[d] == Face and self.topology[d] == Bounded) else 0)
            src = self.F[d] if loc[d] == Face else self.C[d]
            shape = [1, 1, 1]
            shape[d] = n
            out.append(np.array(src.span(1, n)).reshape(shape))
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                lib.ob200_grid_destroy(self.handle)
        except Exception:
            pass
