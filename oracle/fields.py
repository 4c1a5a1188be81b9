"""
Fields, boundary conditions and halo filling (test infrastructure -- see oracle/__init__.py).

Follows Fields/field.jl:16-31,165-194 (Field = grid + OffsetArray over a haloed parent),
Grids/new_data.jl:33-61 (parent sizes/offsets), BoundaryConditions/field_boundary_conditions.jl:13-30
(defaults), fill_halo_regions.jl:34-102 (ordering), fill_halo_regions_periodic.jl:15-105,
fill_halo_regions_flux.jl:16-35, fill_halo_regions_value_gradient.jl:7-99,
fill_halo_regions_open.jl:34-39 and apply_flux_bcs.jl:35-160.
"""
import numpy as np

from .grids import Periodic, Bounded, Flat, Center, Face, flip


class R:
    """Inclusive range of Julia indices along one axis; `r + n` shifts it.  The oracle
    evaluates the reference's pointwise functions f(i, j, k, ...) on whole index boxes at
    once by passing R objects where the reference passes integers."""
    __slots__ = ("lo", "hi")

    def __init__(self, lo, hi=None):
        self.lo = lo
        self.hi = lo if hi is None else hi

    def __add__(self, n):
        return R(self.lo + n, self.hi + n)

    def __sub__(self, n):
        return R(self.lo - n, self.hi - n)

    @property
    def n(self):
        return self.hi - self.lo + 1

    def arr(self, axis):
        shape = [1, 1, 1]
        shape[axis] = self.n
        return np.arange(self.lo, self.hi + 1).reshape(shape)

    def __repr__(self):
        return f"R({self.lo},{self.hi})"


def _r(x):
    return x if isinstance(x, R) else R(int(x))


class BoundaryCondition:
    """kind in {'Periodic','Flux','Value','Gradient','Open', None}; `condition` is a number
    or None (None + 'Flux' = NoFlux).  Function-valued conditions are out of scope."""

    def __init__(self, kind, condition=None):
        self.kind = kind
        self.condition = condition

    def __repr__(self):
        return f"BC({self.kind},{self.condition})"


def default_bc(topo, loc):
    """default_prognostic_bc / default_auxiliary_bc, field_boundary_conditions.jl:13-34."""
    if topo == Periodic:
        return BoundaryCondition("Periodic")
    if topo == Flat:
        return None
    if loc == Center:
        return BoundaryCondition("Flux", None)     # NoFluxBoundaryCondition
    return BoundaryCondition("Open", None)          # ImpenetrableBoundaryCondition


class FieldBoundaryConditions:
    SIDES = ("west", "east", "south", "north", "bottom", "top")

    def __init__(self, grid, loc, auxiliary=False, **kw):
        for d, (lo, hi) in enumerate((("west", "east"), ("south", "north"), ("bottom", "top"))):
            for side in (lo, hi):
                bc = kw.get(side)
                if bc is None:
                    bc = default_bc(grid.topology[d], loc[d])
                    if auxiliary and grid.topology[d] == Bounded and loc[d] == Face:
                        bc = None                     # default_auxiliary_bc(::Bounded, ::Face)
                setattr(self, side, bc)


class Field:
    def __init__(self, grid, loc=(Center, Center, Center), bcs=None, auxiliary=False):
        self.grid = grid
        self.loc = tuple(loc)
        self.parent = np.zeros(grid.parent_size(self.loc), dtype=grid.FT, order="F")
        self.H = grid.H
        self.bcs = bcs if bcs is not None else FieldBoundaryConditions(grid, self.loc, auxiliary)

    # Julia-style (offset) indexing with R ranges or ints
    def _sl(self, ijk):
        out = []
        for d, x in enumerate(ijk):
            x = _r(x)
            out.append(slice(x.lo - 1 + self.H[d], x.hi + self.H[d]))
        return tuple(out)

    def __getitem__(self, ijk):
        return self.parent[self._sl(ijk)]

    def __setitem__(self, ijk, v):
        self.parent[self._sl(ijk)] = v

    def size(self):
        """interior size (Fields/field.jl size(f)): N, +1 for Face on Bounded."""
        g = self.grid
        return tuple(g.N[d] + (1 if (self.loc[d] == Face and g.topology[d] == Bounded) else 0)
                     for d in range(3))

    @property
    def interior(self):
        n = self.size()
        return self[R(1, n[0]), R(1, n[1]), R(1, n[2])]

    def set(self, value):
        """Fields/set!.jl:20-65: array -> copied into the interior; function(x,y,z) ->
        evaluated at the field's nodes."""
        n = self.size()
        if callable(value):
            x, y, z = self.grid.nodes(self.loc)
            value = value(x, y, z) + np.zeros(n)
        self[R(1, n[0]), R(1, n[1]), R(1, n[2])] = np.asarray(value, dtype=self.grid.FT).reshape(n)


# ---------------------------------------------------------------------------------------
# fill_halo_regions!
# ---------------------------------------------------------------------------------------
def _fill_periodic(f, d):
    """fill_halo_regions_periodic.jl:37-105: H planes copied over the FULL parent extent of
    the other two dimensions (so successive x, y, z fills also fill edges and corners)."""
    H, N = f.H[d], f.grid.N[d]
    p = f.parent
    idx = [slice(None)] * 3

    def pl(a, b):
        s = list(idx)
        s[d] = slice(a, b)
        return tuple(s)
    p[pl(0, H)] = p[pl(N, N + H)]                  # c[i] = c[N+i]          (west)
    p[pl(N + H, N + 2 * H)] = p[pl(H, 2 * H)]      # c[N+H+i] = c[H+i]      (east)


def _getbc(bc, FT):
    return FT(0) if bc.condition is None else FT(bc.condition)


def _fill_bounded_side(f, d, side, bc):
    """One side of a non-periodic dimension.  The kernels run over the INTERIOR extent of the
    other two dimensions only (launch!(arch, grid, :yz, ...), fill_halo_regions.jl:163-170)."""
    if bc is None:
        return
    g = f.grid
    FT = g.FT
    N = g.N[d]
    n = g.N                                          # worksize = grid size, not field size
    rng = [R(1, n[0]), R(1, n[1]), R(1, n[2])]

    def at(q):
        r = list(rng)
        r[d] = R(q)
        return tuple(r)
    if bc.kind == "Flux":
        # fill_halo_regions_flux.jl:16-28: only the FIRST halo cell is mirrored
        if side == 0:
            f[at(0)] = f[at(1)]
        else:
            f[at(N + 1)] = f[at(N)]
    elif bc.kind == "Open":
        # fill_halo_regions_open.jl:34-39: boundary-normal component set on the wall face
        f[at(1 if side == 0 else N + 1)] = _getbc(bc, FT)
    elif bc.kind in ("Value", "Gradient"):
        # fill_halo_regions_value_gradient.jl:7-99
        iB = 1 if side == 0 else N + 1
        iI = 1 if side == 0 else N
        iH = 0 if side == 0 else N + 1
        Δ = g.spacing(d, flip(f.loc[d]), R(iB))
        if not np.isscalar(Δ):
            Δ = Δ.reshape(())[()]
        cI = f[at(iI)]
        if bc.kind == "Gradient":
            grad = _getbc(bc, FT)
        elif side == 0:
            grad = (cI - _getbc(bc, FT)) / (Δ / 2)
        else:
            grad = (_getbc(bc, FT) - cI) / (Δ / 2)
        f[at(iH)] = cI + grad * (-Δ if side == 0 else Δ)   # linearly_extrapolate
    else:
        raise ValueError(bc.kind)


def fill_halo_regions(fields):
    """fill_halo_regions.jl:34-102.  Non-periodic dimensions are filled before periodic ones
    (`fill_first`); within each class the order does not change the result because bounded
    fills touch only interior-extent halo cells and periodic fills span full parent extents."""
    if isinstance(fields, Field):
        fields = [fields]
    sides = (("west", "east"), ("south", "north"), ("bottom", "top"))
    for f in fields:
        topo = f.grid.topology
        for d in range(3):
            if topo[d] == Bounded:
                _fill_bounded_side(f, d, 0, getattr(f.bcs, sides[d][0]))
                _fill_bounded_side(f, d, 1, getattr(f.bcs, sides[d][1]))
        for d in range(3):
            if topo[d] == Periodic and f.H[d] > 0:
                _fill_periodic(f, d)


def apply_flux_bcs(G, f):
    """apply_x/y/z_bcs!, apply_flux_bcs.jl:35-160: G[1] += Q*A/V, G[N] -= Q*A/V for
    Flux boundary conditions with a (constant) non-trivial condition."""
    g = f.grid
    FT = g.FT
    n = g.N                                          # launch!(arch, grid, :xy, ...) etc.
    sides = (("west", "east"), ("south", "north"), ("bottom", "top"))
    for d in range(3):
        if g.topology[d] != Bounded:
            continue
        for s, name in enumerate(sides[d]):
            bc = getattr(f.bcs, name)
            if bc is None or bc.kind != "Flux" or bc.condition is None:
                continue
            N = g.N[d]
            rng = [R(1, n[0]), R(1, n[1]), R(1, n[2])]
            cell = 1 if s == 0 else N
            face = 1 if s == 0 else N + 1
            rc = list(rng); rc[d] = R(cell)
            rf = list(rng); rf[d] = R(face)
            loc_f = list(f.loc); loc_f[d] = flip(f.loc[d])
            area = (g.Ax, g.Ay, g.Az)[d](*rf, *loc_f)
            vol = g.V(*rc, *f.loc)
            q = FT(bc.condition)
            if s == 0:
                G[tuple(rc)] = G[tuple(rc)] + q * area / vol
            else:
                G[tuple(rc)] = G[tuple(rc)] - q * area / vol
