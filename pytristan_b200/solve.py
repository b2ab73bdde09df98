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

    ``AxisSolver(A, npts=None, axis=None, refine=0, method="lu", block=128)``: ``A`` is an ``n x n`` matrix
    (ndarray or operator; an operator supplies ``npts``/``axis``).  ``solve(f)`` accepts what ``apply`` accepts.
    Raises ``numpy.linalg.LinAlgError`` when ``A`` is singular.

    ``method="lu"`` (default): ``P A = L U`` with partial pivoting once on the device (``pt_lu_factor``), the
    factors packed per block row of ``block`` rows (``pt_lu_plan``); ``solve`` is ``pt_lu_solve`` — a
    row-permuting copy, then blocked forward and backward substitution: the off-diagonal block rows (all of
    the O(n^2) work per line) are the hot path itself (``pt_dense_apply_ld`` on a row range of the fields,
    tensor cores), the ``block x block`` diagonal triangles a substitution kernel (one thread per grid line).
    Componentwise backward stable like any LU solve; neither ``A^-1`` nor an inverted block is ever formed.
    ``factors()`` returns ``(perm, L, U)`` as ndarrays (``A[perm] = L @ U``).

    ``method="inverse"``: ``A`` (n <= 4096) is inverted once (``pt_invert``, Gauss-Jordan with partial
    pivoting) and a solve is ONE pass of the hot path with the inverse — the fastest form, with an error of
    ``cond(A) * eps``; ``inverse`` holds the operator.

    ``refine`` steps of iterative refinement follow either form, each one residual ``r = f - A y``
    (``pt_dense_apply`` with alpha = -1, beta = 1) and one more solve ``y += solve(r)``.  One step brings the
    residual of an ill-conditioned collocation system (cond ~ n^4) solved through the inverse from
    ``cond * eps * |f|`` down to the backward-stable level ``eps * |A| |y|``.
    """

    def __init__(self, A, npts=None, axis=None, refine=0, method="lu", block=128):
        torch = _lib.require_cuda()
        lib = _lib.load()
        if method not in ("lu", "inverse"):
            raise ValueError("AxisSolver: method must be 'lu' or 'inverse'")
        d_a, n = _device_matrix(A)
        if n > (8192 if method == "lu" else 4096):
            raise ValueError("AxisSolver: n must be <= 8192 (method='lu') / 4096 (method='inverse')")
        if npts is None:
            npts = getattr(A, "npts", None) or (n,)
        if axis is None:
            axis = getattr(A, "axis", None)
            axis = 0 if axis is None else axis
        if int(npts[axis]) != n:
            raise ValueError("AxisSolver: the matrix size does not match the grid axis")
        self.refine = int(refine)
        if self.refine < 0:
            raise ValueError("AxisSolver: refine must be >= 0")
        block = int(block)
        if block < 16 or block > 128 or block % 16:
            raise ValueError("AxisSolver: block must be a multiple of 16 in [16, 128]")
        self.n, self.npts, self.axis, self.method, self.block = n, tuple(int(v) for v in npts), int(axis), method, block
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.inverse = None
        if method == "inverse":
            d_inv = _lib.empty((n, n))
            work = _lib.empty((int(lib.pt_invert_work_size(n)),))
            _lib.check(lib.pt_invert(_lib.dptr(d_a), n, _lib.dptr(d_inv), _lib.dptr(work), _lib.dptr(info),
                                     _lib.stream_ptr()), "pt_invert")
            self.inverse = _DiffMat._wrap(d_inv, self.npts, self.axis, None, None)
        else:
            self._lu = _lib.empty((n, n))
            self._perm = torch.empty(n, dtype=torch.int32, device="cuda")
            work = _lib.empty((int(lib.pt_lu_work_size(n)),))
            _lib.check(lib.pt_lu_factor(_lib.dptr(d_a), n, _lib.dptr(self._lu), _lib.dptr(self._perm), _lib.dptr(work),
                                        _lib.dptr(info), _lib.stream_ptr()), "pt_lu_factor")
        step = int(info.item())
        if step:
            raise np.linalg.LinAlgError(f"AxisSolver: singular matrix (elimination step {step - 1})")
        if method == "lu":
            self._plan = _lib.empty((max(1, int(lib.pt_lu_plan_size(n, block))),))
            _lib.check(lib.pt_lu_plan(_lib.dptr(self._lu), n, block, _lib.dptr(self._plan), _lib.stream_ptr()),
                       "pt_lu_plan")
        self.matrix = _DiffMat._wrap(d_a, self.npts, self.axis, None, None)      # for the residuals
        self._scratch = None

    def factors(self):
        """``(perm, L, U)`` of ``method="lu"`` as ndarrays: ``A[perm] = L @ U``, ``L`` unit lower triangular."""
        if self.method != "lu":
            raise ValueError("AxisSolver.factors: only for method='lu'")
        lu = self._lu.cpu().numpy()
        return self._perm.cpu().numpy().astype(np.int64), np.tril(lu, -1) + np.eye(self.n), np.triu(lu)

    def _solve_device(self, d_f, after, n, before, d_out):
        """``d_out = A^-1 d_f`` (flat device tensors, distinct)."""
        if self.method == "inverse":
            return self.inverse._apply_device(d_f, after, n, before, d_out, 1.0, 0.0, None, None)
        if self._scratch is None or self._scratch.numel() != d_f.numel():
            self._scratch = _lib.empty((d_f.numel(),))
        _lib.check(_lib.load().pt_lu_solve(_lib.dptr(self._plan), _lib.dptr(self._perm), n, self.block, _lib.dptr(d_f),
                                           _lib.dptr(d_out), _lib.dptr(self._scratch), after, before,
                                           _lib.stream_ptr()), "pt_lu_solve")
        return d_out

    def solve(self, f, out=None, refine=None):
        """``y`` with ``A y[.., :, ..] = f[.., :, ..]`` along the axis (same kind and shape as ``f``);
        ``refine`` overrides the number of refinement steps given to the constructor."""
        from .operators import _field_to_device

        steps = self.refine if refine is None else int(refine)
        if steps <= 0 and self.method == "inverse":
            return self.inverse.apply(f, out=out)
        d_f, kind = _field_to_device(f)
        after, n, before = self.matrix._split(d_f.numel())
        d_ff = d_f.reshape(-1)
        d_y = self.matrix._prepare_out(d_ff, out, 0.0).reshape(-1)
        self._solve_device(d_ff, after, n, before, d_y)
        if steps > 0:
            d_r = _lib.empty((d_f.numel(),))
            d_c = _lib.empty((d_f.numel(),)) if self.method == "lu" else None
        for _ in range(steps):
            d_r.copy_(d_ff)
            self.matrix._apply_device(d_y, after, n, before, d_r, -1.0, 1.0, None, None)          # r = f - A y
            if self.method == "inverse":
                self.inverse._apply_device(d_r, after, n, before, d_y, 1.0, 1.0, None, None)      # y += A^-1 r
            else:
                self._solve_device(d_r, after, n, before, d_c)
                d_y.add_(d_c)
        if kind == "numpy":
            return d_y.cpu().numpy().reshape(np.shape(f))
        return d_y.view(d_f.shape)
