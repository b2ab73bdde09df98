"""ctypes binding of the C-ABI library (include/pytristan_b200.h) and the device plumbing.

PyTorch is used only for device memory, streams and host<->device copies; every computation
goes through ``libpytristan_b200.so``.  There is no CPU fallback: without the library or
without a CUDA device the calls below raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpytristan_b200.so")

_c_i32 = ctypes.c_int
_c_i64 = ctypes.c_int64
_c_f64 = ctypes.c_double
_c_ptr = ctypes.c_void_p



class UniformRows(ctypes.Structure):
    """struct pt_uniform_rows (include/pytristan_b200.h)."""

    _fields_ = [("w_uni", _c_i32), ("lo", _c_i32), ("hi", _c_i32), ("n_bnd", _c_i32),
                ("wuni", ctypes.POINTER(_c_f64)), ("w_bnd", ctypes.POINTER(_c_f64)),
                ("start_bnd", ctypes.POINTER(_c_i32))]


class ScatterMap(ctypes.Structure):
    """struct pt_scatter_map (include/pytristan_b200.h)."""

    _fields_ = [("nranks", _c_i32), ("chunk", _c_i32), ("a_inner", _c_i64), ("s_outer", _c_i64),
                ("s_inner", _c_i64), ("s_i", _c_i64), ("offset", _c_i64), ("base", _c_ptr)]


def uniform_rows(band, n_out):
    """Build the pt_uniform_rows argument from a band dict (keys w, lo, hi, wuni, wmax, start,
    weights); returns None for non-uniform bands.  The arrays are kept alive on the struct."""
    if band.get("wuni") is None:
        return None
    lo, hi, wmax = int(band["lo"]), int(band["hi"]), int(band["wmax"])
    rows = list(range(0, lo)) + list(range(hi, n_out))
    u = UniformRows()
    u.w_uni, u.lo, u.hi, u.n_bnd = int(band["w"]), lo, hi, len(rows)
    u._wuni = (ctypes.c_double * int(band["w"]))(*np.asarray(band["wuni"], dtype=np.float64).tolist())
    u.wuni = ctypes.cast(u._wuni, ctypes.POINTER(_c_f64))
    if rows and len(rows) <= 8:
        wb = np.ascontiguousarray(np.asarray(band["weights"])[rows], dtype=np.float64).ravel()
        sb = np.ascontiguousarray(np.asarray(band["start"])[rows], dtype=np.int32)
        u._wb = (ctypes.c_double * wb.size)(*wb.tolist())
        u._sb = (ctypes.c_int32 * sb.size)(*sb.tolist())
        u.w_bnd = ctypes.cast(u._wb, ctypes.POINTER(_c_f64))
        u.start_bnd = ctypes.cast(u._sb, ctypes.POINTER(_c_i32))
    return u


# name -> (restype, argtypes); mirrors include/pytristan_b200.h one to one
SIGNATURES = {
    "pt_last_error": (ctypes.c_char_p, []),
    "pt_version": (_c_i32, []),
    "pt_device_info": (_c_i32, [ctypes.POINTER(_c_i32)] * 3),
    "pt_gen_linspace": (_c_i32, [_c_ptr, _c_f64, _c_f64, _c_i64, _c_ptr]),
    "pt_gen_cheb": (_c_i32, [_c_ptr, _c_i64, _c_i32, _c_f64, _c_f64, _c_ptr]),
    "pt_cheb_map": (_c_i32, [_c_ptr, _c_i64, _c_f64, _c_f64, _c_ptr]),
    "pt_gen_cheb_libm": (_c_i32, [_c_ptr, _c_i64, _c_i32, _c_f64, _c_f64, _c_i32, _c_ptr]),
    "pt_cos_libm": (_c_i32, [_c_ptr, _c_ptr, _c_i64, _c_i32, _c_ptr]),
    "pt_grid_expand": (_c_i32, [_c_ptr, ctypes.POINTER(_c_i64), _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "pt_grid_axpoints": (_c_i32, [_c_ptr, _c_i64, _c_i64, _c_ptr, _c_ptr]),
    "pt_gen_chebmat": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "pt_gen_barymat": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr]),
    "pt_gen_fourmat": (_c_i32, [_c_i32, _c_f64, _c_i32, _c_ptr, _c_ptr]),
    "pt_gen_fourmat_general": (_c_i32, [_c_i32, _c_f64, _c_i32, _c_ptr, _c_ptr, _c_ptr]),
    "pt_gen_fornberg": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_i32, _c_f64, _c_ptr,
                                 _c_ptr, _c_i32, _c_ptr]),
    "pt_band_to_dense": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "pt_kron_lift": (_c_i32, [_c_ptr, _c_i32, _c_i64, _c_i64, _c_ptr, _c_ptr]),
    "pt_packed_size": (_c_i64, [_c_i32, _c_i32]),
    "pt_operator_pack": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_i64, _c_ptr, _c_ptr]),
    "pt_dense_apply": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64,
                                _c_f64, _c_ptr, _c_i32, _c_i64, _c_ptr]),
    "pt_flip_symmetrize": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr]),
    "pt_operator_symmetry": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "pt_packed_eo_size": (_c_i64, [_c_i32]),
    "pt_operator_pack_eo": (_c_i32, [_c_ptr, _c_i32, _c_ptr, _c_ptr, _c_ptr]),
    "pt_dense_apply_eo": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64, _c_f64,
                                   _c_ptr, _c_i32, _c_i64, _c_ptr]),
    "pt_dense_apply_eo_scatter": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64,
                                           _c_f64, ctypes.POINTER(ScatterMap), _c_ptr]),
    "pt_packed_eo_ws_size": (_c_i64, [_c_i32]),
    "pt_operator_pack_eo_ws": (_c_i32, [_c_ptr, _c_i32, _c_ptr, _c_ptr]),
    "pt_dense_apply_eo_ws": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64, _c_f64,
                                      _c_ptr, _c_i32, _c_i64, _c_ptr]),
    "pt_dense_apply_eo_ws_scatter": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64,
                                              _c_f64, ctypes.POINTER(ScatterMap), _c_ptr]),
    "pt_dense_apply_eo_ws_scatter_copy": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64,
                                                   _c_f64, ctypes.POINTER(ScatterMap), ctypes.POINTER(ScatterMap), _c_ptr]),
    "pt_fourier_twiddles": (_c_i32, [_c_i32, _c_ptr, _c_ptr]),
    "pt_fourier_apply": (_c_i32, [_c_ptr, _c_i32, _c_f64, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64, _c_f64,
                                  _c_ptr, _c_i64, _c_ptr]),
    "pt_stencil_apply": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64,
                                  _c_i32, ctypes.POINTER(UniformRows), _c_f64, _c_f64, _c_ptr]),
    "pt_stencil_apply_ex": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64,
                                     _c_i64, _c_i32, ctypes.POINTER(UniformRows), _c_f64, _c_f64, _c_ptr]),
    "pt_stencil_grad3": (_c_i32, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i64,
                                  ctypes.POINTER(_c_ptr), ctypes.POINTER(_c_ptr), ctypes.POINTER(_c_i32),
                                  ctypes.POINTER(ctypes.POINTER(UniformRows)), _c_ptr]),
    "pt_stencil_grad3_halo": (_c_i32, [_c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32,
                                       ctypes.POINTER(_c_ptr), ctypes.POINTER(_c_ptr), ctypes.POINTER(_c_i32),
                                       ctypes.POINTER(ctypes.POINTER(UniformRows)), _c_ptr]),
    "pt_permute3d": (_c_i32, [_c_ptr, _c_ptr, _c_i64, _c_i64, _c_i64, _c_i32, _c_ptr]),
    "pt_replace_rows": (_c_i32, [_c_ptr, _c_i32, _c_ptr, _c_i32, _c_ptr, _c_ptr]),
    "pt_invert_work_size": (_c_i64, [_c_i32]),
    "pt_invert": (_c_i32, [_c_ptr, _c_i32, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "pt_dense_apply_ld": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_i64, _c_ptr, _c_i64, _c_i64, _c_i64, _c_f64, _c_f64, _c_ptr]),
    "pt_lu_work_size": (_c_i64, [_c_i32]),
    "pt_lu_factor": (_c_i32, [_c_ptr, _c_i32, _c_ptr, _c_ptr, _c_ptr, _c_ptr, _c_ptr]),
    "pt_lu_plan_size": (_c_i64, [_c_i32, _c_i32]),
    "pt_lu_plan": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr]),
    "pt_lu_solve": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_ptr]),
    "pt_debug_set_tri_lines": (_c_i32, [_c_i32]),
    "pt_measure_dmma_peak": (_c_i32, [ctypes.POINTER(_c_f64), _c_i32]),
    "pt_peer_alloc": (_c_i32, [_c_i64, ctypes.POINTER(_c_ptr)]),
    "pt_peer_free": (_c_i32, [_c_ptr]),
    "pt_ipc_export": (_c_i32, [_c_ptr, _c_ptr]),
    "pt_ipc_open": (_c_i32, [_c_ptr, ctypes.POINTER(_c_ptr)]),
    "pt_ipc_close": (_c_i32, [_c_ptr]),
    "pt_peer_barrier": (_c_i32, [_c_ptr, _c_i32, _c_i32, ctypes.c_uint64, _c_i64, _c_ptr, _c_ptr]),
    "pt_dense_apply_scatter": (_c_i32, [_c_ptr, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64, _c_i64, _c_f64,
                                        _c_f64, ctypes.POINTER(ScatterMap), _c_ptr]),
    "pt_peer_scatter": (_c_i32, [_c_ptr, _c_i32, _c_i64, _c_i64, ctypes.POINTER(ScatterMap), _c_ptr]),
    "pt_stencil_apply_halo": (_c_i32, [_c_ptr, _c_ptr, _c_i32, _c_i32, _c_i32, _c_i32, _c_ptr, _c_ptr, _c_i64,
                                       _c_ptr, _c_i64, _c_ptr, _c_i64, _c_i64, ctypes.POINTER(UniformRows),
                                       _c_f64, _c_f64, _c_ptr]),
}

SCALE_NONE, SCALE_ROW, SCALE_AFTER, SCALE_BEFORE = 0, 1, 2, 3
LIBM_GLIBC_FMA, LIBM_GLIBC_NOFMA = 1, 2

# arguments on which the two x86-64 builds of glibc's cos disagree: (x, cos_fma(x), cos_nofma(x)) bits
_LIBM_PROBES = (
    (0x4006994ac2a0c6a4, 0xbfee687cff59a44f, 0xbfee687cff59a450),
    (0x3ffc4fd49259333f, 0xbfc94408859fa3c5, 0xbfc94408859fa3c4),
    (0x40001a0458d397a5, 0xbfdb5eda16e37d6d, 0xbfdb5eda16e37d6e),
    (0x3fe49acc91af868a, 0x3fe997a81c219b77, 0x3fe997a81c219b76),
    (0x3fe877bd71546efa, 0x3fe717bd07934ca8, 0x3fe717bd07934ca7),
    (0x3ffdb39d631e5038, 0xbfd20735577cc5de, 0xbfd20735577cc5df),
)
_libm_model = None


def host_libm_model():
    """Which build of glibc's cos ``numpy.cos`` runs on this host (pt_libm_model), or 0 when it
    is neither: decided by evaluating numpy.cos on arguments where the builds differ."""
    global _libm_model
    if _libm_model is None:
        x = np.array([p[0] for p in _LIBM_PROBES], dtype=np.uint64).view(np.float64)
        got = np.cos(x).view(np.uint64)
        if all(int(g) == p[1] for g, p in zip(got, _LIBM_PROBES)):
            _libm_model = LIBM_GLIBC_FMA
        elif all(int(g) == p[2] for g, p in zip(got, _LIBM_PROBES)):
            _libm_model = LIBM_GLIBC_NOFMA
        else:
            _libm_model = 0
    return _libm_model

_lib = None


class PytristanCudaError(RuntimeError):
    """A C-ABI call failed; the message is the library's pt_last_error()."""


def load():
    """Load (once) and return the ctypes handle of the C-ABI library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        # the library is built in-tree by __graft_entry__.build(); as a safety net build it here
        # when a toolkit is present (the GPU image has nvcc) — never fall back to a CPU path
        try:
            from .build import build as _build

            _build()
        except Exception as exc:
            raise ImportError(
                f"{LIB_PATH} is missing and could not be built ({exc}): run "
                "`python -m pytristan_b200.build` (pytristan_b200 has no CPU fallback)"
            ) from exc
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().pt_last_error()
        raise PytristanCudaError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def torch_mod():
    import torch

    return torch


def require_cuda():
    """Return ``torch`` after making sure a CUDA device is usable; raise otherwise."""
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise RuntimeError(
            "pytristan_b200 needs a CUDA device (B200, sm_100a): there is no CPU fallback"
        )
    load()
    return torch


def stream_ptr():
    torch = torch_mod()
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a torch tensor as c_void_p (None for None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def to_device(arr, dtype=None):
    """Host ndarray -> contiguous CUDA tensor (reinterpreting unsupported dtypes as bytes is the
    caller's business)."""
    torch = require_cuda()
    a = np.ascontiguousarray(arr, dtype=dtype)
    if not a.flags.writeable:
        a = np.array(a)           # torch.from_numpy refuses to alias read-only memory silently
    return torch.from_numpy(a).cuda(non_blocking=False)


def empty(shape, dtype=None):
    torch = require_cuda()
    return torch.empty(shape, dtype=dtype or torch.float64, device="cuda")
