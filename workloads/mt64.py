"""std::mt19937_64 (every synthetic input comes from it, seed 20261019) with the numpy-Generator calls
workloads/synth.py uses.  The C++ side (mt64.cpp, libstdc++ distributions) is compiled on first use with g++."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        so, src = os.path.join(HERE, "libmt64.so"), os.path.join(HERE, "mt64.cpp")
        if not os.path.exists(so) or os.path.getmtime(src) > os.path.getmtime(so):
            subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
        lib = C.CDLL(so)
        lib.mt64_new.restype = C.c_void_p
        lib.mt64_new.argtypes = [C.c_uint64]
        lib.mt64_free.argtypes = [C.c_void_p]
        lib.mt64_uniform.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t]
        lib.mt64_normal.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_size_t]
        lib.mt64_raw.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _LIB = lib
    return _LIB


class Mt64:
    def __init__(self, seed: int):
        self.lib = _lib()
        self.g = self.lib.mt64_new(int(seed) & (2**64 - 1))

    def __del__(self):
        try:
            self.lib.mt64_free(self.g)
        except Exception:
            pass

    @staticmethod
    def _shape(size):
        if size is None:
            return (), 1
        shape = (size,) if np.isscalar(size) else tuple(size)
        return shape, int(np.prod(shape)) if shape else 1

    def uniform(self, low=0.0, high=1.0, size=None):
        shape, n = self._shape(size)
        out = np.empty(n)
        self.lib.mt64_uniform(self.g, float(low), float(high), out.ctypes.data, n)
        return out.reshape(shape) if shape else float(out[0])

    def random(self, size=None):
        return self.uniform(0.0, 1.0, size)

    def standard_normal(self, size=None):
        shape, n = self._shape(size)
        out = np.empty(n)
        self.lib.mt64_normal(self.g, 0.0, 1.0, out.ctypes.data, n)
        return out.reshape(shape) if shape else float(out[0])

    def normal(self, loc=0.0, scale=1.0, size=None):
        return loc + scale * self.standard_normal(size)

    def choice(self, a, size=None):
        a = np.asarray(a)
        idx = np.minimum((self.uniform(0.0, 1.0, size) * a.size).astype(np.int64), a.size - 1)
        return a[idx]

    def permutation(self, n):
        return np.argsort(self.uniform(0.0, 1.0, int(n)), kind="stable")

    def dirichlet(self, alpha, size=None):
        """Dirichlet(alpha) through normalised Gamma draws; alpha = 1 (the only use) makes them -log U"""
        alpha = np.asarray(alpha, dtype=np.float64)
        assert np.all(alpha == 1.0), "only the flat Dirichlet is needed by the workloads"
        shape, n = self._shape(size)
        g = -np.log(1.0 - self.uniform(0.0, 1.0, (n, alpha.size)))
        out = g / g.sum(axis=1, keepdims=True)
        return out.reshape(shape + (alpha.size,)) if shape else out[0]


def default_rng(seed: int) -> Mt64:
    return Mt64(seed)
