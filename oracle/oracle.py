"""ctypes front-end of ``oracle/lattice_oracle.c`` (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product package never does.  All arrays are NumPy, fp32 /
int16 / int8 / int32, C-contiguous.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.C_ORACLE_SO
        if not os.path.exists(path) or build_oracle._stale(path, [os.path.join(build_oracle.HERE, "lattice_oracle.c")]):
            build_oracle.build_c_oracle()
        L = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        L.sgpo_variance.restype = C.c_float
        L.sgpo_variance.argtypes = [fp, C.c_int]
        L.sgpo_scale_factors.restype = None
        L.sgpo_scale_factors.argtypes = [C.c_int, C.c_float, fp]
        L.sgpo_slice_divisor.restype = C.c_float
        L.sgpo_slice_divisor.argtypes = [C.c_int]
        L.sgpo_lattice_build.restype = C.c_void_p
        L.sgpo_lattice_build.argtypes = [fp, C.c_int64, C.c_int, C.c_int64, fp, C.c_int, C.POINTER(C.c_int)]
        L.sgpo_lattice_build_reftable.restype = C.c_void_p
        L.sgpo_lattice_build_reftable.argtypes = [fp, C.c_int64, C.c_int, C.c_int64, fp, C.c_int, C.POINTER(C.c_int)]
        L.sgpo_lattice_free.restype = None
        L.sgpo_lattice_free.argtypes = [C.c_void_p]
        L.sgpo_build_neighbours.restype = C.c_int
        L.sgpo_build_neighbours.argtypes = [C.c_void_p]
        L.sgpo_lattice_M.restype = C.c_int64
        L.sgpo_lattice_M.argtypes = [C.c_void_p]
        L.sgpo_lattice_var.restype = C.c_float
        L.sgpo_lattice_var.argtypes = [C.c_void_p]
        for name, ty in [("scale", C.c_float), ("greedy", C.c_int16), ("rank", C.c_int8), ("weights", C.c_float),
                         ("offsets", C.c_int32), ("keys", C.c_int16), ("nbr", C.c_int32)]:
            f = getattr(L, "sgpo_lattice_" + name)
            f.restype = C.POINTER(ty)
            f.argtypes = [C.c_void_p]
        L.sgpo_mvm.restype = C.c_int
        L.sgpo_mvm.argtypes = [C.c_void_p, fp, C.c_int64, C.c_int64, fp, C.c_int, fp, C.c_int64, fp, fp]
        L.sgpo_filter.restype = C.c_int
        L.sgpo_filter.argtypes = [fp, fp, fp, C.c_int64, C.c_int64, C.c_int, C.c_int, fp, C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def variance(coeffs) -> np.float32:
    c = _f32(coeffs)
    return np.float32(lib().sgpo_variance(_fp(c), c.shape[0]))


def scale_factors(d: int, var) -> np.ndarray:
    out = np.empty(d, dtype=np.float32)
    lib().sgpo_scale_factors(d, C.c_float(float(var)), _fp(out))
    return out


def slice_divisor(d: int) -> np.float32:
    return np.float32(lib().sgpo_slice_divisor(d))


class OracleLattice:
    """Lattice structure of ``x[N,d]`` under stencil ``coeffs[2r+1]`` in the reference's numbering."""

    def __init__(self, x, coeffs, reference_table: bool = False):
        """``reference_table=True`` restates the reference's own hash table, growth defect included, and so reproduces
        the UNMODIFIED reference at any size (orphaned duplicate lattice points and all); the default is a correct
        table, which is what the product implements (see the header of lattice_oracle.c)."""
        x = _f32(x)
        assert x.ndim == 2
        self.coeffs = _f32(coeffs)
        self.N, self.d = x.shape
        self.order = self.coeffs.shape[0] // 2
        st = C.c_int(0)
        build = lib().sgpo_lattice_build_reftable if reference_table else lib().sgpo_lattice_build
        self._h = build(_fp(x), self.N, self.d, self.d, _fp(self.coeffs), self.coeffs.shape[0], C.byref(st))
        self.status = st.value
        if not self._h:
            raise RuntimeError(f"oracle build failed with status {st.value}")
        self.M = int(lib().sgpo_lattice_M(self._h))
        self.var = np.float32(lib().sgpo_lattice_var(self._h))
        self._nbr_ready = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None:
            _lib.sgpo_lattice_free(h)
            self._h = None

    def _arr(self, name, shape, dtype):
        n = int(np.prod(shape))
        if n == 0:
            return np.empty(shape, dtype=dtype)
        p = getattr(lib(), "sgpo_lattice_" + name)(self._h)
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(shape).astype(dtype, copy=True)

    @property
    def scale(self):
        return self._arr("scale", (self.d,), np.float32)

    @property
    def greedy(self):
        return self._arr("greedy", (self.N, self.d + 1), np.int16)

    @property
    def rank(self):
        return self._arr("rank", (self.N, self.d + 1), np.int8)

    @property
    def weights(self):
        return self._arr("weights", (self.N, self.d + 1), np.float32)

    @property
    def offsets(self):
        return self._arr("offsets", (self.N, self.d + 1), np.int32)

    @property
    def keys(self):
        return self._arr("keys", (self.M, self.d), np.int16)

    @property
    def nbr(self):
        """nbr[j, i, t]: lattice index of the neighbour of point i along axis j, t over o=-r..-1,1..r; -1 absent."""
        if not self._nbr_ready:
            st = lib().sgpo_build_neighbours(self._h)
            if st != 0:
                raise RuntimeError(f"oracle neighbour build failed: {st}")
            self._nbr_ready = True
        return self._arr("nbr", (self.d + 1, self.M, 2 * self.order), np.int32)

    def mvm(self, src, return_intermediates: bool = False):
        src = _f32(src)
        assert src.ndim == 2 and src.shape[0] == self.N
        Cc = src.shape[1]
        out = np.zeros((self.N, Cc), dtype=np.float32)
        sp = np.zeros((self.M, Cc), dtype=np.float32) if return_intermediates else None
        bl = np.zeros((self.M, Cc), dtype=np.float32) if return_intermediates else None
        st = lib().sgpo_mvm(self._h, _fp(src), Cc, Cc, _fp(self.coeffs), self.coeffs.shape[0], _fp(out), Cc,
                            _fp(sp) if sp is not None else None, _fp(bl) if bl is not None else None)
        if st != 0:
            raise RuntimeError(f"oracle mvm failed: {st}")
        self._nbr_ready = True
        return (out, sp, bl) if return_intermediates else out


def filter(src, ref, coeffs):
    """CPU restatement of the reference operator ``filter(src[N,C], ref[N,d], coeffs[2r+1]) -> out[N,C]``."""
    src, ref, coeffs = _f32(src), _f32(ref), _f32(coeffs)
    N, Cc = src.shape
    out = np.zeros((N, Cc), dtype=np.float32)
    M = C.c_int64(0)
    st = lib().sgpo_filter(_fp(src), _fp(ref), _fp(coeffs), N, Cc, ref.shape[1], coeffs.shape[0], _fp(out), C.byref(M))
    if st != 0:
        raise RuntimeError(f"oracle filter failed: {st}")
    return out
