"""Output path either side of the time step: what is written is selected / reduced on the device and
only that travels to the host.

Mirrors OutputWriters/field_slicer.jl (`FieldSlicer`), fetch_output.jl:24-36 (`fetch_output`), the `AveragedField` outputs of
the examples (horizontal averages), and checkpointer.jl:64-95,201-262 (`Checkpointer`, `set!(model, filepath)`): a checkpoint
holds the parents (halos included) of every prognostic field, of the pressures, of G^n and G^- and the clock, so that a
restored model continues BIT FOR BIT."""
import ctypes as C
import os

import numpy as np

from ._lib import lib, check


class FieldSlicer:
    """FieldSlicer(i=:, j=:, k=:, with_halos=false): an index, a (first, last) pair (1-based, inclusive) or None per dimension"""

    def __init__(self, i=None, j=None, k=None, with_halos=False):
        self.ranges, self.with_halos = (i, j, k), bool(with_halos)

    def box(self, field):
        n, H = field.size(), field.grid.H
        lo, hi = [], []
        for d, r in enumerate(self.ranges):
            h = H[d] if self.with_halos else 0
            if r is None:
                a, b = 1 - h, n[d] + h
            elif isinstance(r, (int, np.integer)):
                a = b = int(r)
            else:
                a, b = int(r[0]), int(r[1])
            lo.append(a)
            hi.append(b)
        return lo, hi


def fetch_output(field, slicer=None, sync=True):
    """fetch_output(field, model, field_slicer): the sliced data as a host array"""
    lo, hi = (slicer or FieldSlicer()).box(field)
    return field.slice(lo, hi, sync=sync)


def horizontal_average(field, sync=True):
    """the profile mean(field, dims=(1, 2)) as a vector over z"""
    return field.average((1, 2), sync=sync).reshape(-1)


class Checkpointer:
    """Checkpointer(model; dir, prefix): write() stores <dir>/<prefix>_iteration<N>.npz; restore(model, path) is the reference's
    set!(model, filepath).  Restoring into a model built with the same arguments continues bit for bit."""

    def __init__(self, model, dir=".", prefix="checkpoint"):
        self.model, self.dir, self.prefix = model, dir, prefix

    @staticmethod
    def _fields(model):
        out = {n: model.fields[n] for n in model.names}
        for n in model.names:
            out["Gn_" + n] = model.Gn[n]
            out["Gm_" + n] = model.Gm[n]
        for n, f in model.pressures.items():
            if f is not None:
                out["pressure_" + n] = f
        for n, f in (getattr(model, "diffusivity_fields", None) or {}).items():     # LES closures: nu_e, kappa_e of every tracer
            if isinstance(f, dict):
                for t, g in f.items():
                    out[f"diffusivity_{n}_{t}"] = g
            elif f is not None:
                out["diffusivity_" + n] = f
        return out

    def write(self):
        m = self.model
        pdt = C.c_double()
        check(lib.ob200_model_previous_time_step(m.handle, C.byref(pdt)))
        data = {k: f.parent() for k, f in self._fields(m).items()}
        data["clock"] = np.array([m.clock.time, float(m.clock.iteration), pdt.value])
        os.makedirs(self.dir, exist_ok=True)
        path = os.path.join(self.dir, f"{self.prefix}_iteration{m.clock.iteration}.npz")
        np.savez(path, **data)
        return path

    @staticmethod
    def restore(model, path):
        with np.load(path) as z:
            for k, f in Checkpointer._fields(model).items():
                f.set_parent(z[k])
            t, it, pdt = z["clock"]
        check(lib.ob200_model_set_clock(model.handle, float(t), int(it), float(pdt)))
        return model
