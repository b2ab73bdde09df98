"""Boundary-condition rows and solves along a grid axis (the step after the
hot path in a collocation solver built on these matrices).

``with_rows(op, rows, new_rows)``  — copy of an operator's matrix with some rows replaced
                                      (``pt_replace_rows``): Dirichlet rows are identity rows,
                                      Neumann/Robin rows come from another operator.
``dirichlet(op)``                  — first and last rows replaced by identity rows.
``AxisSolver(A, npts, axis)``      — solves ``A y = f`` along ``axis`` for every grid line of a
                                      batch of fields.  ``A`` (n x n, n <= 4096) is inverted once
                                      on the device (``pt_invert``, Gauss-Jordan with partial
                                      pivoting); each solve is then the hot path itself —
                                      ``pt_dense_apply`` with the inverse, all lines at once, no
                                      transposes — optionally followed by steps of iterative refinement
                                      (residual with ``A``, correction with the inverse: two more passes
                                      each) that bring the residual to the backward-stable level.
"""
import ctypes

import numpy as np

from . import _lib
from .operators import _DiffMat

__all__ = ["with_rows", "dirichlet", "AxisSolver"]


def _device_matrix(a):
    if isinstance(a, _DiffMat):
        return a.device().clone(), int(a.shape[0])
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if arr.ndim != 2 or arr.shape[0] != arr.shape[1]:
        raise ValueError("expected a square matrix")
    return _lib.to_device(arr), int(arr.shape[0])


def with_rows(op, rows, new_rows):
    """Matrix of ``op`` with ``rows`` replaced by ``new_rows`` (``len(rows) x n``); returns an
    operator bound like ``op`` (a plain ``_DiffMat``: it is no longer a differentiation matrix)."""
    torch = _lib.require_cuda()
    d_mat, n = _device_matrix(op)
    rows = np.asarray(rows, dtype=np.int64).ravel()
    rows = np.where(rows < 0, rows + n, rows)
    if rows.size and (rows.min() < 0 or rows.max() >= n):
        raise IndexError("row index out of range")
    if len(set(rows.tolist())) != rows.size:
        raise ValueError("rows must be unique")
    new = np.ascontiguousarray(new_rows, dtype=np.float64).reshape(rows.size, n)
    d_rows = torch.from_numpy(rows.astype(np.int32)).cuda()
    d_new = _lib.to_device(new)
    _lib.check(_lib.load().pt_replace_rows(_lib.dptr(d_mat), n, _lib.dptr(d_rows), int(rows.size), _lib.dptr(d_new),
                                           _lib.stream_ptr()), "pt_replace_rows")
    npts = getattr(op, "npts", None) or (n,)
    axis = getattr(op, "axis", None)
    return _DiffMat._wrap(d_mat, npts, 0 if axis is None else axis, getattr(op, "order", None), getattr(op, "grid_num", None))


def dirichlet(op):
    """``op`` with its first and last rows replaced by identity rows (Dirichlet conditions at both
    ends of the axis: the right-hand side carries the boundary values in those rows)."""
    n = int(np.shape(op)[0])
    eye = np.zeros((2, n))
    eye[0, 0] = 1.0
    eye[1, n - 1] = 1.0
    return with_rows(op, [0, n - 1], eye)


class AxisSolver:
    """``A y = f`` along one grid axis, for every grid line of a batch of fields.

    ``AxisSolver(A, npts=None, axis=None, refine=0)``: ``A`` is an ``n x n`` matrix (ndarray or operator;
    an operator supplies ``npts``/``axis``).  ``solve(f)`` accepts what ``apply`` accepts.  Raises
    ``numpy.linalg.LinAlgError`` when ``A`` is singular.

    ``refine`` steps of iterative refinement follow the product with the inverse, each two more passes of
    the hot path: ``r = f - A y`` (``pt_dense_apply`` with alpha = -1, beta = 1) and ``y += A^-1 r``.  One
    step brings the residual of an ill-conditioned collocation system (cond ~ n^4) from
    ``cond * eps * |f|`` down to the backward-stable level ``eps * |A| |y|`` — what an LU solve with
    refinement delivers — without leaving the tensor cores or transposing anything.
    """

    def __init__(self, A, npts=None, axis=None, refine=0):
        torch = _lib.require_cuda()
        lib = _lib.load()
        d_a, n = _device_matrix(A)
        if n > 4096:
            raise ValueError("AxisSolver: n must be <= 4096")
        if npts is None:
            npts = getattr(A, "npts", None) or (n,)
        if axis is None:
            axis = getattr(A, "axis", None)
            axis = 0 if axis is None else axis
        if int(npts[axis]) != n:
            raise ValueError("AxisSolver: the matrix size does not match the grid axis")
        d_inv = _lib.empty((n, n))
        work = _lib.empty((int(lib.pt_invert_work_size(n)),))
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(lib.pt_invert(_lib.dptr(d_a), n, _lib.dptr(d_inv), _lib.dptr(work), _lib.dptr(info),
                                 _lib.stream_ptr()), "pt_invert")
        step = int(info.item())
        if step:
            raise np.linalg.LinAlgError(f"AxisSolver: singular matrix (elimination step {step - 1})")
        self.n, self.npts, self.axis = n, tuple(int(v) for v in npts), int(axis)
        self.refine = int(refine)
        if self.refine < 0:
            raise ValueError("AxisSolver: refine must be >= 0")
        self.inverse = _DiffMat._wrap(d_inv, self.npts, self.axis, None, None)
        self.matrix = _DiffMat._wrap(d_a, self.npts, self.axis, None, None)      # for the residuals

    def solve(self, f, out=None, refine=None):
        """``y`` with ``A y[.., :, ..] = f[.., :, ..]`` along the axis (same kind and shape as ``f``);
        ``refine`` overrides the number of refinement steps given to the constructor."""
        steps = self.refine if refine is None else int(refine)
        if steps <= 0:
            return self.inverse.apply(f, out=out)
        from .operators import _field_to_device

        d_f, kind = _field_to_device(f)
        after, n, before = self.inverse._split(d_f.numel())
        d_y = self.inverse._apply_device(d_f, after, n, before, out, 1.0, 0.0, None, None)
        d_r = _lib.empty((d_f.numel(),))
        for _ in range(steps):
            d_r.copy_(d_f.reshape(-1))
            self.matrix._apply_device(d_y.reshape(-1), after, n, before, d_r, -1.0, 1.0, None, None)     # r = f - A y
            self.inverse._apply_device(d_r, after, n, before, d_y.reshape(-1), 1.0, 1.0, None, None)      # y += A^-1 r
        if kind == "numpy":
            return d_y.cpu().numpy().reshape(np.shape(f))
        return d_y.view(d_f.shape)
