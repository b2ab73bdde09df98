"""Fused multi-operator passes: callers want gradients and Laplacians, not
single derivatives; fusing them cuts the algorithmic traffic.

``gradient(ops, u)`` returns the three first derivatives of 3-D fields.  When the three operators
are uniform non-periodic ``FinDiffMat``s of the same width it runs ``pt_stencil_grad3`` — one read
of ``u``, three writes, 32 B/point instead of 48 — otherwise it applies the operators one by one.
"""
import ctypes

import numpy as np

from . import _lib
from .operators import FinDiffMat, _field_to_device

__all__ = ["gradient"]


def gradient(ops, u, outs=None):
    """``[D0 u, D1 u, D2 u]`` for operators bound to axes 0, 1, 2 of the same 3-D grid.

    ``u``: one or more flat fields (CUDA tensor or ndarray); ``outs``: optional list of three
    CUDA tensors of the same size.  Returns arrays of the same kind as ``u``.
    """
    torch = _lib.require_cuda()
    ops = list(ops)
    if len(ops) != 3 or any(op.npts is None or len(op.npts) != 3 for op in ops):
        raise ValueError("gradient needs three operators bound to a 3-D grid")
    if [op.axis for op in ops] != [0, 1, 2] or any(tuple(op.npts) != tuple(ops[0].npts) for op in ops):
        raise ValueError("gradient: operators must be bound to axes 0, 1, 2 of the same grid")
    d_u, kind = _field_to_device(u)
    nx, ny, nz = (int(v) for v in ops[0].npts)
    total = nx * ny * nz
    if d_u.numel() % total:
        raise ValueError("gradient: field size is not a multiple of the grid size")
    nfields = d_u.numel() // total
    if outs is None:
        d_out = [torch.empty_like(d_u) for _ in range(3)]
    else:
        d_out = list(outs)
        for o in d_out:
            if not (isinstance(o, torch.Tensor) and o.is_cuda and o.dtype == torch.float64 and o.is_contiguous()
                    and o.numel() == d_u.numel()):
                raise TypeError("outs must be contiguous float64 CUDA tensors of the field's size")

    fused = all(isinstance(op, FinDiffMat) and op._band is not None and op._band["wuni"] is not None
                and not op._band["periodic"] for op in ops)
    rc = -3
    if fused:
        lib = _lib.load()
        bands = [op._band for op in ops]
        for band, n in zip(bands, (nx, ny, nz)):
            if "uni" not in band:
                band["uni"] = _lib.uniform_rows(band, n)
        w_arr = (ctypes.c_void_p * 3)(*[b["d_w"].data_ptr() for b in bands])
        s_arr = (ctypes.c_void_p * 3)(*[b["d_start"].data_ptr() for b in bands])
        m_arr = (ctypes.c_int * 3)(*[int(b["wmax"]) for b in bands])
        u_arr = (ctypes.POINTER(_lib.UniformRows) * 3)(*[ctypes.pointer(b["uni"]) for b in bands])
        rc = lib.pt_stencil_grad3(_lib.dptr(d_u), _lib.dptr(d_out[0]), _lib.dptr(d_out[1]), _lib.dptr(d_out[2]),
                                  nx, ny, nz, nfields, w_arr, s_arr, m_arr, u_arr, _lib.stream_ptr())
        if rc not in (0, -3):
            _lib.check(rc, "pt_stencil_grad3")
    if rc != 0:      # not fusable: one pass per operator
        for op, o in zip(ops, d_out):
            op.apply(d_u, out=o)
    if kind == "numpy":
        return [o.cpu().numpy().reshape(np.shape(u)) for o in d_out]
    return [o.view(d_u.shape) for o in d_out]
