"""
RectilinearGrid restatement (test infrastructure -- see oracle/__init__.py).

Follows Grids/rectilinear_grid.jl:249-279 (constructor), Grids/grid_generation.jl:28-112
(coordinate generation), Grids/grid_utils.jl:14-29 (total_length / total_extent),
Grids/input_validation.jl:55-64 (default halo 3, Flat -> 0) and
Operators/spacings_and_areas_and_volumes.jl:63-113,173-236 (spacings, areas, volumes).

Indexing convention used by the whole oracle: Julia's.  Interior cells are 1..N, halos
extend to 1-H .. N+H (Face fields on Bounded dims to N+1+H).  All vectors that the
reference stores as OffsetArrays are stored here as `OffsetVector`s with the same offsets.
"""
from fractions import Fraction

import numpy as np

Periodic, Bounded, Flat = "Periodic", "Bounded", "Flat"
Center, Face = "c", "f"


def flip(loc):
    """BoundaryConditions/apply_flux_bcs.jl:92-93."""
    return Face if loc == Center else Center


def total_length(loc, topo, N, H=0):
    """Grids/grid_utils.jl:24-29."""
    if topo == Flat:
        return N
    if loc == Face and topo == Bounded:
        return N + 1 + 2 * H
    return N + 2 * H


class OffsetVector:
    """1-D OffsetArray: element with Julia index q is `parent[q - first]`."""

    def __init__(self, parent, first):
        self.parent = np.asarray(parent)
        self.first = first

    def __getitem__(self, q):
        return self.parent[q - self.first]

    def slice(self, lo, hi):
        """values at Julia indices lo..hi inclusive"""
        return self.parent[lo - self.first: hi + 1 - self.first]

    @property
    def last(self):
        return self.first + len(self.parent) - 1


def _regular_nodes(FT, a, b, n):
    """range(FT(a), FT(b), length=n): Julia evaluates ranges in twice precision, i.e. to
    (nearly) the correctly rounded value of a + (i-1)(b-a)/(n-1)."""
    a = Fraction(float(FT(a)))
    b = Fraction(float(FT(b)))
    if n == 1:
        return np.array([FT(float(a))], dtype=FT)
    return np.array([FT(float(a + (b - a) * i / (n - 1))) for i in range(n)], dtype=FT)


def generate_regular_coordinate(FT, topo, N, H, coord):
    """Grids/grid_generation.jl:83-107.  BigFloat arithmetic is replaced by exact rationals."""
    c1, c2 = Fraction(float(coord[0])), Fraction(float(coord[1]))
    assert c1 < c2
    L = c2 - c1
    D = L / N
    Fm = c1 - H * D
    # total_extent: Grids/grid_utils.jl:14-15
    Fp = Fm + (L + 2 * H * D if topo == Bounded else L + (2 * H - 1) * D)
    Cm = Fm + D / 2
    Cp = Cm + L + D * (2 * H - 1)
    TF = total_length(Face, topo, N, H)
    TC = total_length(Center, topo, N, H)
    F = OffsetVector(_regular_nodes(FT, float(Fm), float(Fp), TF), 1 - H)
    C = OffsetVector(_regular_nodes(FT, float(Cm), float(Cp), TC), 1 - H)
    return FT(float(L)), F, C, FT(float(D)), FT(float(D))


def generate_stretched_coordinate(FT, topo, N, H, coord):
    """Grids/grid_generation.jl:28-80.  `coord` is a function of the face index (1-based)
    or a vector of N+1 faces."""
    get = (lambda i: coord(i)) if callable(coord) else (lambda i: coord[i - 1])
    interiorF = np.zeros(N + 1, dtype=FT)
    for i in range(1, N + 2):
        interiorF[i - 1] = get(i)
    L = interiorF[N] - interiorF[0]
    Fi = interiorF
    if topo == Bounded:
        dm = [Fi[1] - Fi[0] for _ in range(H)]          # lower_exterior_Δcoordᶠ(Bounded)  :17
        dp = [Fi[-1] - Fi[-2] for _ in range(H)]        # upper_exterior_Δcoordᶠ(Bounded)  :20
    else:
        n = len(Fi)
        # Fi[end - H + i] - Fi[end - H + i - 1], i = 1:H  (1-based)                        :16
        dm = [Fi[n - H + i - 1] - Fi[n - H + i - 2] for i in range(1, H + 1)]
        dp = [Fi[i] - Fi[i - 1] for i in range(1, H + 1)]                                 # :19
    dp = dp[::-1]
    c1, cN1 = interiorF[0], interiorF[N]
    Fm = [c1 - np.sum(np.array(dm[i - 1:H], dtype=FT)) for i in range(1, H + 1)]
    Fp = [cN1 + np.sum(np.array(dp[i - 1:H], dtype=FT)) for i in range(1, H + 1)][::-1]
    F = np.concatenate([np.array(Fm, dtype=FT), interiorF, np.array(Fp, dtype=FT)]).astype(FT)
    TC = total_length(Center, topo, N, H)
    C = np.array([(F[i + 1] + F[i]) / 2 for i in range(TC)], dtype=FT)
    dF = np.array([C[i] - C[i - 1] for i in range(1, TC)], dtype=FT)
    TF = total_length(Face, topo, N, H)
    F = F[:TF]
    dC = np.array([F[i + 1] - F[i] for i in range(TF - 1)], dtype=FT)
    dF = np.concatenate([[dF[0]], dF, [dF[-1]]]).astype(FT)
    for i in range(len(dF) - 1, 0, -1):
        dF[i] = dF[i - 1]
    return (FT(L), OffsetVector(F, 1 - H), OffsetVector(C, 1 - H),
            OffsetVector(dF, -H), OffsetVector(dC, 1 - H))


class RectilinearGrid:
    """RectilinearGrid(arch, FT; size, x, y, z | extent, topology, halo)
    (Grids/rectilinear_grid.jl:249-279).  `x`, `y`, `z` are 2-tuples (regular), callables of
    the face index or arrays of faces (stretched).  Flat dimensions are omitted from `size`
    and `halo` exactly as in the reference (Grids/input_validation.jl)."""

    def __init__(self, FT=np.float64, size=None, x=None, y=None, z=None, extent=None,
                 topology=(Periodic, Periodic, Bounded), halo=None):
        self.FT = FT = np.dtype(FT).type
        self.topology = tuple(topology)
        nflat = sum(t == Flat for t in topology)
        size = (size,) if np.isscalar(size) else tuple(size)
        assert len(size) == 3 - nflat, "size must have one entry per non-Flat dimension"
        if halo is None:
            halo = (3,) * (3 - nflat)                     # input_validation.jl:55
        halo = (halo,) if np.isscalar(halo) else tuple(halo)
        if extent is not None:
            extent = (extent,) if np.isscalar(extent) else tuple(extent)
            ext = iter(extent)
        coords = [x, y, z]
        N, H = [], []
        si, hi = iter(size), iter(halo)
        for d, t in enumerate(topology):
            if t == Flat:
                N.append(1)
                H.append(0)
                coords[d] = (0.0, 1.0)
            else:
                N.append(int(next(si)))
                H.append(int(next(hi)))
                if extent is not None:        # the "oceanic" default domain of input_validation.jl:92-95: z = (-Lz, 0)
                    Ld = float(next(ext))
                    coords[d] = (0.0, Ld) if d < 2 else (-Ld, 0.0)
        self.Nx, self.Ny, self.Nz = N
        self.Hx, self.Hy, self.Hz = H
        self.L, self.nodesF, self.nodesC, self.dF, self.dC, self.regular = [], [], [], [], [], []
        for d, t in enumerate(topology):
            c = coords[d]
            if t == Flat:
                # grid_generation.jl:110-112
                one = np.ones(N[d], dtype=FT)
                self.L.append(FT(1))
                self.nodesF.append(OffsetVector(one, 1))
                self.nodesC.append(OffsetVector(one, 1))
                self.dF.append(FT(1))
                self.dC.append(FT(1))
                self.regular.append(True)
            elif isinstance(c, tuple) and len(c) == 2 and not callable(c):
                L, F, C, dF, dC = generate_regular_coordinate(FT, t, N[d], H[d], c)
                self.L.append(L); self.nodesF.append(F); self.nodesC.append(C)
                self.dF.append(dF); self.dC.append(dC); self.regular.append(True)
            else:
                L, F, C, dF, dC = generate_stretched_coordinate(FT, t, N[d], H[d], c)
                self.L.append(L); self.nodesF.append(F); self.nodesC.append(C)
                self.dF.append(dF); self.dC.append(dC); self.regular.append(False)
        self.Lx, self.Ly, self.Lz = self.L

    # ---- sizes -------------------------------------------------------------------------
    @property
    def N(self):
        return (self.Nx, self.Ny, self.Nz)

    @property
    def H(self):
        return (self.Hx, self.Hy, self.Hz)

    def with_halo(self, halo):
        """Grids/rectilinear_grid.jl with_halo: same grid, new halo sizes."""
        g = object.__new__(RectilinearGrid)
        g.__dict__.update(self.__dict__)
        coords = []
        for d, t in enumerate(self.topology):
            if t == Flat:
                coords.append(None)
            elif self.regular[d]:
                coords.append((float(self.nodesF[d][1]), float(self.nodesF[d][1]) + float(self.L[d])))
            else:
                coords.append(np.array(self.nodesF[d].slice(1, self.N[d] + 1)))
        size = tuple(n for n, t in zip(self.N, self.topology) if t != Flat)
        hl = tuple(h for h, t in zip(halo, self.topology) if t != Flat)
        return RectilinearGrid(self.FT, size=size, x=coords[0], y=coords[1], z=coords[2],
                               topology=self.topology, halo=hl)

    def parent_size(self, loc):
        """Grids/new_data.jl:16-22,56-61."""
        return tuple(total_length(loc[d], self.topology[d], self.N[d], self.H[d]) for d in range(3))

    # ---- spacings (Operators/spacings_and_areas_and_volumes.jl:63-113) ------------------
    def spacing(self, d, loc, idx):
        """Δ along dimension d at location loc for index range `idx` (an R): scalar for a
        regular/Flat dimension, array broadcast along d for a stretched one."""
        v = self.dF[d] if loc == Face else self.dC[d]
        if self.regular[d]:
            return v
        a = v.slice(idx.lo, idx.hi)
        shape = [1, 1, 1]
        shape[d] = len(a)
        return a.reshape(shape)

    def Δx(self, loc, i):
        return self.spacing(0, loc, i)

    def Δy(self, loc, j):
        return self.spacing(1, loc, j)

    def Δz(self, loc, k):
        return self.spacing(2, loc, k)

    # areas and volumes (:173-236): Ax = Δy*Δz, Ay = Δx*Δz, Az = Δx*Δy, V = Az*Δz
    def Ax(self, i, j, k, lx, ly, lz):
        return self.Δy(ly, j) * self.Δz(lz, k)

    def Ay(self, i, j, k, lx, ly, lz):
        return self.Δx(lx, i) * self.Δz(lz, k)

    def Az(self, i, j, k, lx, ly, lz):
        return self.Δx(lx, i) * self.Δy(ly, j)

    def V(self, i, j, k, lx, ly, lz):
        return self.Az(i, j, k, lx, ly, lz) * self.Δz(lz, k)

    # ---- nodes --------------------------------------------------------------------------
    def nodes(self, loc, interior=True):
        """x, y, z node arrays of the interior points of a field at `loc`, shaped for
        broadcasting (Grids/grid_utils.jl xnodes/ynodes/znodes)."""
        out = []
        for d in range(3):
            n = self.N[d] + (1 if (loc[d] == Face and self.topology[d] == Bounded) else 0)
            src = self.nodesF[d] if loc[d] == Face else self.nodesC[d]
            a = np.array(src.slice(1, n))
            shape = [1, 1, 1]
            shape[d] = n
            out.append(a.reshape(shape))
        return out
