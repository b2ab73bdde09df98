"""The C-ABI library: it loads without a GPU, exports every symbol include/*.h declares, and its
argument validation (which runs before any CUDA call) reports errors through pt_last_error()."""
import ctypes
import os
import re

import pytest

from pytristan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "pytristan_b200.h")) as fh:
        text = re.sub(r"/\*.*?\*/", "", fh.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m pytristan_b200.build`"
    assert os.path.dirname(_lib.LIB_PATH).endswith(os.path.join("pytristan_b200", "csrc"))


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 19
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_sizes():
    lib = _lib.load()
    assert lib.pt_version() == 100
    assert lib.pt_packed_size(257, 257) == 65 * 264 * 4
    assert lib.pt_packed_size(64, 64) == 16 * 64 * 4
    assert lib.pt_packed_size(0, 5) == 0


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    rc = lib.pt_dense_apply(None, 4, 4, None, None, 1, 1, 1.0, 0.0, None, 0, 0, None)
    assert rc == -1 and b"null pointer" in lib.pt_last_error()
    rc = lib.pt_gen_fourmat(8, 1.0, 3, ctypes.c_void_p(16), None)
    assert rc == -1 and b"order=3" in lib.pt_last_error()
    rc = lib.pt_gen_fornberg(None, 10, 1, 3, 0, 1, 0.1, ctypes.c_void_p(16), ctypes.c_void_p(16), 4, None)
    assert rc == -1 and b"even" in lib.pt_last_error()
    rc = lib.pt_stencil_apply(ctypes.c_void_p(16), ctypes.c_void_p(16), 7, 5, ctypes.c_void_p(16), ctypes.c_void_p(32),
                              1, 1, 0, None, 1.0, 0.0, None)
    assert rc == -1 and b"smaller than the stencil" in lib.pt_last_error()
    rc = lib.pt_grid_expand(None, (ctypes.c_int64 * 1)(4), 1, 3, None, None)
    assert rc in (-1, -3)
    rc = lib.pt_permute3d(ctypes.c_void_p(16), ctypes.c_void_p(32), 2, 2, 2, 0, None)   # perm 0 = (0,0,0)
    assert rc == -1 and b"not a permutation" in lib.pt_last_error()
    with pytest.raises(_lib.PytristanCudaError):
        _lib.check(rc, "pt_permute3d")
