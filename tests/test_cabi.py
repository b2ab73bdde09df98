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


def test_lu_plan_layout_and_validation_need_no_gpu():
    """Host logic of the factorisation-based solve: work and plan sizes, even block split, argument checks."""
    lib = _lib.load()
    assert lib.pt_lu_work_size(257) == 259 and lib.pt_lu_work_size(0) == 0
    # n <= nb: one block, no off-diagonal block rows, two padded triangles
    assert lib.pt_lu_plan_size(100, 128) == 2 * 112 * 112
    assert lib.pt_lu_plan_size(16, 16) == 2 * 16 * 16
    # n = 257, nb = 128 splits evenly into 88 + 88 + 81 (not 128 + 128 + 1): triangles of 96 x 96 each and the four
    # off-diagonal block rows L[1, 0:88], L[2, 0:176], U[0, 88:257], U[1, 176:257] in pt_operator_pack's layout
    tri = 3 * 2 * 96 * 96
    rect = sum(lib.pt_packed_size(m, k) for m, k in ((88, 88), (81, 176), (88, 169), (88, 81)))
    assert lib.pt_lu_plan_size(257, 128) == tri + rect
    assert lib.pt_lu_plan_size(1024, 128) == 8 * 2 * 128 * 128 + sum(2 * lib.pt_packed_size(128, 128 * b) for b in range(1, 8))
    assert lib.pt_lu_plan_size(0, 128) == 0
    p = ctypes.c_void_p(16)
    rc = lib.pt_lu_factor(None, 8, p, p, p, p, None)
    assert rc == -1 and b"null pointer" in lib.pt_last_error()
    rc = lib.pt_lu_factor(p, 10000, p, p, p, p, None)
    assert rc == -1 and b"outside" in lib.pt_last_error()
    rc = lib.pt_lu_plan(p, 64, 24, p, None)
    assert rc == -1 and b"block size" in lib.pt_last_error()
    rc = lib.pt_lu_solve(p, p, 64, 64, p, p, ctypes.c_void_p(32), 1, 1, None)
    assert rc == -1 and b"distinct" in lib.pt_last_error()
    rc = lib.pt_dense_apply_ld(p, 4, 4, p, 3, ctypes.c_void_p(32), 4, 1, 1, 1.0, 0.0, None)
    assert rc == -1 and b"slab strides" in lib.pt_last_error()
    assert lib.pt_debug_set_tri_lines(0) == 0
