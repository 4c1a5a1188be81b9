"""ctypes wrapper of oracle_cpu.c, the compiled OpenMP twin of the NumPy oracle (TEST
INFRASTRUCTURE: only tests/ and bench.py's cpu_baseline / --impl reference legs use it)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle_cpu.so")


def load():
    if not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", _HERE])
    lib = C.CDLL(_LIB)
    lib.oc_rk3_run.restype = C.c_int
    lib.oc_rk3_run.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int] + [C.c_void_p] * 4 + \
                              [C.c_int, C.c_double, C.c_int, C.c_int]
    lib.oc_max_threads.restype = C.c_int
    return lib


def rk3_run(N, L, u, v, w, b, nsteps, dt, zweno=True, project=True, nthreads=0):
    """N = (Nx, Ny, Nz) (Nz = 1 means Flat z); arrays are (Nx, Ny, Nz), returned updated (copies)."""
    lib = load()
    arrs = [np.asfortranarray(a, dtype=np.float64).copy(order="F") for a in (u, v, w, b)]
    Nc = (C.c_int * 3)(*N)
    Lc = (C.c_double * 3)(*[float(x) for x in L])
    lib.oc_rk3_run(Nc, Lc, int(zweno), *[a.ctypes.data_as(C.c_void_p) for a in arrs], int(nsteps), float(dt),
                   int(project), int(nthreads))
    return arrs


def max_threads():
    return load().oc_max_threads()
