"""ctypes loader of the C restatement of glibc's cos (oracle/glibc_cos.c).  TEST INFRASTRUCTURE:
imported by tests/ only; it is the checker of the device cosine ``pt_cos_libm`` / ``pt_gen_cheb_libm``."""
import ctypes

import numpy as np

from .build import build

_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        lib.oracle_glibc_cos_array.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.oracle_glibc_cos_array.restype = None
        lib.oracle_cheb_nodes.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.oracle_cheb_nodes.restype = None
        _lib = lib
    return _lib


def glibc_cos(x, use_fma):
    """cos(x) as the chosen glibc build computes it (x: float64 array)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _load().oracle_glibc_cos_array(x.ctypes.data, out.ctypes.data, x.size, int(bool(use_fma)))
    return out


def cheb_nodes(n, use_fma):
    """CGL nodes cos(fl(fl(j*pi)/(n-1))) (reference grid.py:75) with that cosine."""
    out = np.empty(int(n), dtype=np.float64)
    _load().oracle_cheb_nodes(out.ctypes.data, int(n), int(bool(use_fma)))
    return out
