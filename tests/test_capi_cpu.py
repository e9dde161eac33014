"""The C-ABI library on a machine without a GPU: it loads, exports every symbol include/sgp_lattice.h declares,
its host-side constant functions agree with the oracle bit for bit, and argument errors come back as status codes
(no compute entry point is called with real work here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import MAT15_2, MAT15_3, RBF1, RBF2, ROOT

HEADER = os.path.join(ROOT, "include", "sgp_lattice.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(sgp_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(sg):
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_capi.SYMBOLS) == declared, "python binding list out of sync with the header"
    assert lib.sgp_abi_version() == 5


@pytest.mark.parametrize("coeffs", [RBF1, RBF2, MAT15_2, MAT15_3, [1.0], [0.5, 1.0, 0.5]])
def test_host_constants_match_oracle(sg, oracle, coeffs):
    v = sg.stencil_variance(coeffs)
    assert np.float32(v).view(np.int32) == np.float32(oracle.variance(coeffs)).view(np.int32)
    for d in (1, 2, 8, 11, 18, 24, 60, 126):
        a, b = sg.scale_factors(d, v), oracle.scale_factors(d, v)
        assert np.array_equal(a.view(np.int32), b.view(np.int32))
        assert np.float32(sg.slice_divisor(d)).view(np.int32) == np.float32(oracle.slice_divisor(d)).view(np.int32)


def test_status_codes_not_exceptions(sg):
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    out = C.c_float(0)
    arr = (C.c_float * 2)(1.0, 2.0)
    assert lib.sgp_stencil_variance(arr, 2, C.byref(out)) == -1   # even length
    assert b"odd" in lib.sgp_last_error()
    buf = (C.c_float * 4)()
    assert lib.sgp_scale_factors(0, C.c_float(0.4), buf) == -1
    assert lib.sgp_scale_factors(127, C.c_float(0.4), buf) == -1
    # dimension / size validation happens before any CUDA call
    assert lib.sgp_build_points(None, -1, 3, 3, buf, None, None, None, None, None) == -1
    assert lib.sgp_build_points(None, 10, 500, 500, buf, None, None, None, None, None) == -5
    assert lib.sgp_build_points(None, 1 << 40, 8, 8, buf, None, None, None, None, None) == -3
    assert lib.sgp_hash_insert(None, None, 10, 3, None, 1000, None, None, None) == -1   # null + not a power of two
    with pytest.raises(_capi.SgpError):
        _capi.check(lib.sgp_build_points(None, 5, 3, 3, buf, None, None, None, None, None))
    assert lib.sgp_hash_capacity(1000) == 2048 and lib.sgp_hash_capacity(1) == 1024
    assert lib.sgp_number_workspace_bytes(1000, 8) >= 9000 * 4


def test_one_call_filter_argument_handling(sg):
    """sgp_filter / sgp_filter_host validate their arguments and size their workspace on the host, before any CUDA call."""
    from simplex_gp_b200 import _capi
    lib = _capi.lib()
    c = (C.c_float * 3)(0.5, 1.0, 0.5)
    M = C.c_int64(-1)
    for fn in (lib.sgp_filter, lib.sgp_filter_host):
        assert fn(None, 4, None, 3, c, 3, 0, 4, 3, None, 4, None, 0, 0, C.byref(M), None) == 0    # N = 0: nothing to do
        assert M.value == 0
        assert fn(None, 4, None, 3, c, 2, 10, 4, 3, None, 4, None, 0, 0, None, None) == -1       # even stencil
        assert fn(None, 4, None, 3, c, 3, 10, 4, 3, None, 4, None, 0, 0, None, None) == -1       # null pointers
        assert fn(None, 4, None, 3, c, 3, 10, 4, 500, None, 4, None, 0, 0, None, None) == -1     # d out of range
    worst = lib.sgp_filter_workspace_bytes(1000, 8, 16, 1, 0)
    assert worst == lib.sgp_filter_workspace_bytes(1000, 8, 16, 1, 9000) == lib.sgp_filter_workspace_bytes(1000, 8, 16, 1, 10**9)
    assert lib.sgp_filter_workspace_bytes(1000, 8, 16, 1, 500) < worst
    assert lib.sgp_filter_host_workspace_bytes(1000, 8, 16, 1, 500) >= lib.sgp_filter_workspace_bytes(1000, 8, 16, 1, 500) + 4 * 1000 * (8 + 32)
    assert lib.sgp_filter_workspace_bytes(1000, 0, 16, 1, 0) == 0 and lib.sgp_filter_workspace_bytes(-1, 8, 16, 1, 0) == 0


def test_no_cpu_fallback(sg):
    """The product path must fail loudly without a CUDA device / library, never fall back to the oracle."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x = torch.randn(10, 3)
    with pytest.raises(RuntimeError):
        sg.Lattice(x, RBF1)
    with pytest.raises(RuntimeError):
        sg.filter(torch.randn(10, 2), x, torch.tensor(RBF1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "simplex-gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} mentions the oracle"
