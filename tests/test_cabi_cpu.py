"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/ocean_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ocean_b200.h")
LIB = os.path.join(ROOT, "clima-oceananigans.jl_b200", "ocean_b200", "libocean_b200.so")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "clima-oceananigans.jl_b200", "csrc"), "-j8"])
    return ctypes.CDLL(LIB)


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ob200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(built, n), f"{n} declared in include/ocean_b200.h but not exported"


def test_python_binding_covers_header():
    import ocean_b200._lib as L
    assert set(declared_symbols()) == set(L.SYMBOLS)


def test_version_and_no_cpu_fallback(built):
    import torch
    assert built.ob200_version() == 100
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ocean_b200 as ob
    with pytest.raises(ob.B200Error, match="no usable CUDA device"):
        ob.B200()
    # the C ABI reports errors by status + message, never by exception
    built.ob200_init.restype = ctypes.c_int32
    assert built.ob200_init(0) != 0
    buf = ctypes.create_string_buffer(256)
    built.ob200_last_error.restype = ctypes.c_size_t
    built.ob200_last_error(buf, 256)
    assert b"no CPU fallback" in buf.value


def test_sass_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_grid_matches_oracle_grid():
    """host mirror's coordinate generation (grids.py) against the oracle's restatement"""
    import numpy as np
    import oracle as O
    from ocean_b200 import grids as G
    zF = -np.linspace(1, 0, 10) ** 1.7
    for FT in (np.float64, np.float32):
        for topo in ("Periodic", "Bounded"):
            a = O.grids.generate_stretched_coordinate(FT, topo, 9, 3, zF)
            b = G._stretched(FT, topo, 9, 3, zF)
            for x, y in zip(a[1:], b[1:]):
                assert x.first == y.first and np.array_equal(x.parent, y.a)
            a = O.grids.generate_regular_coordinate(FT, topo, 12, 3, (0.0, 2 * np.pi))
            b = G._regular(FT, topo, 12, 3, (0.0, 2 * np.pi))
            assert a[0] == b[0] and a[3] == b[3]
            assert np.array_equal(a[1].parent, b[1].a) and np.array_equal(a[2].parent, b[2].a)


def test_host_weno_tables_match_oracle():
    import numpy as np
    import oracle as O
    from ocean_b200.model import _eno_weights
    from ocean_b200.grids import _stretched
    zF = -np.linspace(1, 0, 13) ** 1.4
    g = O.RectilinearGrid(size=(4, 4, 12), x=(0, 1), y=(0, 1), z=zF)
    s = O.WENO5(grid=g)
    _, F, Cn, _, _ = _stretched(np.float64, "Bounded", 12, 4, zF)
    for li, nodes in (("f", F), ("c", Cn)):
        for ri, r in enumerate((-1, 0, 1, 2)):
            for i in range(14):
                assert np.allclose(s.coeff[2][li][ri][i], _eno_weights(r, nodes, i), rtol=0, atol=0)


def test_field_slicer_index_boxes():
    """FieldSlicer (OutputWriters/field_slicer.jl): host logic only -- index ranges with and without halos"""
    import types
    from ocean_b200.output_writers import FieldSlicer
    f = types.SimpleNamespace(size=lambda: (8, 6, 5), grid=types.SimpleNamespace(H=(3, 2, 1)))
    assert FieldSlicer().box(f) == ([1, 1, 1], [8, 6, 5])
    assert FieldSlicer(with_halos=True).box(f) == ([-2, -1, 0], [11, 8, 6])
    assert FieldSlicer(k=3).box(f) == ([1, 1, 3], [8, 6, 3])
    assert FieldSlicer(i=(2, 4), j=5, with_halos=True).box(f) == ([2, 5, 0], [4, 5, 6])
