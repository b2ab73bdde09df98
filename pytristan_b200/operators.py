"""Differential operators — ``ChebMat``, ``FourMat``, ``FinDiffMat`` — and their getters.

The reference announces these as ``numpy.ndarray`` subclasses with registry-backed getters
(reference README.md:7-10,16,24-26) but ships no code for them; names come from README.md:16,
the getter/drop family mirrors the grid one (reference grid.py:328-421, :514-610) on a second
``ObjectManager`` whose composite keys ``"grid-axis-order[-accuracy]"`` are what
``ObjectManager.nums`` already parses (reference _manager.py:7-12).

Every operator is an ``n x n`` host ndarray (so it *is* a NumPy array, as the README promises)
whose entries were generated on the GPU (``pt_gen_chebmat`` / ``pt_gen_fourmat`` /
``pt_gen_fornberg``) and which keeps device-resident forms for application:

* dense operators keep the fragment-major packed copy read by the FP64 tensor-core kernels
  (``pt_dense_apply``: K1 for a strided axis, K2 for axis 0); operators that are
  centro(anti)symmetric — every ``FourMat``, every ``ChebMat`` on a CGL axis — keep the two packed
  half-size operators of the even-odd decomposition instead and are applied with half the
  tensor-core work (``pt_dense_apply_eo``; decided once per operator by ``pt_operator_symmetry``);
* ``FinDiffMat`` keeps the banded row-start form read by the stencil kernels
  (``pt_stencil_apply``: K3/K4).

``op.apply(u)`` / ``op @ u`` differentiate a field (flat, in the Grid's linearisation, any
number of stacked fields) along the operator's axis — the implicit form of the Kronecker
lifting ``kron(I_after, kron(D, I_before))``, which ``op.kron()`` materialises for small grids.
Fields may be CUDA ``torch`` tensors (zero-copy, result stays on the device; DLPack-compatible)
or host ndarrays (copied in and out).  There is no CPU fallback.
"""
import ctypes
import operator
import warnings

import numpy as np

from . import _lib
from ._manager import ObjectManager
from .grid import DeviceGrid, Grid, _get_grid_manager

__all__ = [
    "ChebMat",
    "FourMat",
    "FinDiffMat",
    "PolarLaplacian",
    "get_cheb_mat",
    "get_four_mat",
    "get_findiff_mat",
    "drop_cheb_mat",
    "drop_four_mat",
    "drop_findiff_mat",
    "drop_all_mats",
    "_get_mat_manager",
]


def _prod(seq):
    out = 1
    for v in seq:
        out *= int(v)
    return out


def _resolve_axis(grid, axis):
    """(nodes, npts, axis, grid_num) from a Grid, a registered grid number or a 1-D node array."""
    if isinstance(grid, (int, np.integer)) and not isinstance(grid, bool):
        found = getattr(_get_grid_manager(), str(int(grid)), None)
        if found is None:
            raise ValueError(f"No grid with identifier {grid} is registered.")
        grid = found
    if isinstance(grid, DeviceGrid) or (isinstance(grid, Grid) and grid.npts is not None):
        try:
            operator.index(axis)
        except TypeError as exc:
            raise TypeError("Axis' index axis must be an integer.") from exc
        nodes = grid.axpoints(axis)
        axis = axis + grid.num_dim if axis < 0 else axis
        return np.asarray(nodes, dtype=np.float64), tuple(grid.npts), int(axis), grid.num
    nodes = np.asarray(grid, dtype=np.float64)
    if nodes.ndim != 1:
        raise ValueError("Expected a Grid, a registered grid identifier or a 1-D array of nodes.")
    if axis not in (0, -1):
        raise IndexError("A bare node array has a single axis.")
    return nodes, (nodes.size,), 0, None


def _nodes_symmetric(x):
    """Nodes symmetric about their midpoint to within 8 ulp of their span (oracle: nodes_symmetric)."""
    x = np.asarray(x, dtype=np.float64)
    if x.size < 2:
        return False
    c2 = x[0] + x[-1]
    scale = max(float(np.max(np.abs(x))), abs(float(x[-1] - x[0])))
    return bool(np.max(np.abs(x + x[::-1] - c2)) <= 8.0 * np.finfo(np.float64).eps * scale)


def _is_uniform(x):
    n = x.size
    if n < 3:
        return True
    h = (x[-1] - x[0]) / (n - 1)
    tol = 8.0 * np.finfo(np.float64).eps * max(float(np.max(np.abs(x))), abs(float(h)))
    return bool(np.max(np.abs(np.diff(x) - h)) <= tol)


def _is_affine_cgl(x, tol=1e-13):
    """True when the nodes are an affine image (either orientation) of the CGL set."""
    n = x.size
    if n < 2 or x[-1] == x[0]:
        return False
    t = (x - x[0]) / (x[-1] - x[0])
    ref = (1.0 - np.cos(np.arange(n) * np.pi / (n - 1))) / 2.0
    return bool(np.max(np.abs(t - ref)) <= tol)


def _field_to_device(u):
    """(device tensor float64 contiguous, kind) where kind tells how to hand the result back."""
    torch = _lib.require_cuda()
    if isinstance(u, torch.Tensor):
        if not u.is_cuda:
            raise TypeError("torch fields must live on a CUDA device (no CPU path)")
        if u.dtype != torch.float64:
            raise TypeError("fields must be float64")
        return u.contiguous(), "torch"
    if hasattr(u, "__dlpack__") and not isinstance(u, np.ndarray):
        t = torch.from_dlpack(u)
        if not t.is_cuda or t.dtype != torch.float64:
            raise TypeError("DLPack fields must be float64 CUDA arrays")
        return t.contiguous(), "torch"
    a = np.ascontiguousarray(u, dtype=np.float64)
    return torch.from_numpy(a).cuda(), "numpy"


# Even-odd application of centrosymmetric operators (csrc/dense_apply_eo.cu): operators whose
# entries satisfy J D J = +-D to within EO_TOLERANCE (entrywise, relative) are applied as two
# half-size tensor-core products.  Set EVEN_ODD = False to force the plain kernel.
EVEN_ODD = True
EVEN_ODD_WS = True          # large fields: the warp-specialised even-odd kernel (csrc/dense_apply_eo_ws.cu)
FUSED_FIELD_COPY = True     # pencil redistributions: the field rides in the DMMA kernel (pt_dense_apply_eo_ws_scatter_copy)
STATS = {"fused_field_copies": 0}       # how often that happened (tests and bench.py report it)
EO_TOLERANCE = 4.0e-15


class _DiffMat(np.ndarray):
    """Common machinery: attributes, device copies, application, lifting."""

    _KIND = "mat"

    def __array_finalize__(self, obj):
        if obj is None:
            return
        # views/slices/ufunc results are plain matrices: they lose the grid binding and the
        # device-resident forms (same policy as Grid.__array_finalize__, reference grid.py:99-109)
        self.axis = None
        self.order = None
        self.npts = None
        self.grid_num = None
        self._key = None
        self._dev = None
        self._packed = None
        self._band = None
        self._eo = None

    # -- construction helpers ------------------------------------------------------------------
    @classmethod
    def _wrap(cls, d_mat, npts, axis, order, grid_num):
        host = d_mat.cpu().numpy()
        # The device-resident forms below are what apply()/@ read.  The host entries are therefore
        # frozen: an in-place edit (D[0, :] = 0 ...) would silently leave the GPU applying the old
        # operator.  Boundary rows go through solve.with_rows / solve.dirichlet; a writable
        # ``D.copy()`` is a plain matrix whose device forms are rebuilt on every use (_cacheable).
        host.setflags(write=False)
        obj = host.view(cls)
        obj.axis, obj.order, obj.npts, obj.grid_num = axis, order, tuple(npts), grid_num
        obj._key = None
        obj._dev = d_mat
        obj._packed = None
        obj._band = None
        obj._eo = None
        if EVEN_ODD:
            obj._eo_forms()          # eligibility + packed halves now, not inside the first apply()
        return obj

    @property
    def key(self):
        """Registry key ``[grid, axis, order(, accuracy)]`` or None when unregistered."""
        return self._key

    @property
    def _cacheable(self):
        """Device forms are cached only while the host entries cannot change (read-only array)."""
        return not self.flags.writeable

    def device(self):
        """Row-major ``n_out x n_in`` CUDA tensor holding the same entries as the ndarray."""
        if self._dev is not None and self._cacheable:
            return self._dev
        dev = _lib.to_device(np.ascontiguousarray(self), dtype=np.float64)
        self._dev = dev if self._cacheable else None
        return dev

    def _packed_device(self):
        if self._packed is not None and self._cacheable:
            return self._packed
        lib = _lib.load()
        d = self.device()
        n_out, n_in = int(self.shape[0]), int(self.shape[1])
        packed = _lib.empty((lib.pt_packed_size(n_out, n_in),))
        _lib.check(
            lib.pt_operator_pack(_lib.dptr(d), n_out, n_in, n_in, _lib.dptr(packed),
                                 _lib.stream_ptr()),
            "pt_operator_pack",
        )
        self._packed = packed if self._cacheable else None
        return packed

    def _eo_forms(self):
        """``(sign, packed_even, packed_odd, packed_ws)`` when the operator is centro(anti)symmetric to within
        EO_TOLERANCE, else None.  Decided on the device when the operator is built (``_wrap``), so
        that ``apply`` never allocates or synchronises; the verdict of ``pt_operator_symmetry`` lands
        in a caller-owned device scalar and is read back here, on the Python side."""
        if not self._cacheable:
            return None               # writable copy: entries may change, plain kernel every time
        if getattr(self, "_eo", None) is None:
            self._eo = False
            n = int(self.shape[0])
            if self.ndim == 2 and self.shape[0] == self.shape[1] and n >= 16:
                lib = _lib.load()
                torch = _lib.torch_mod()
                d = self.device()
                order = getattr(self, "order", None)
                signs = (1, -1) if order is None else ((1,) if order % 2 == 0 else (-1,))
                d_worst = torch.zeros(len(signs), dtype=torch.float64, device="cuda")
                for k, sign in enumerate(signs):
                    _lib.check(lib.pt_operator_symmetry(_lib.dptr(d), n, sign, _lib.dptr(d_worst[k:]),
                                                        _lib.stream_ptr()), "pt_operator_symmetry")
                worst = d_worst.cpu().tolist()
                for sign, w in zip(signs, worst):
                    if w <= EO_TOLERANCE:
                        size = lib.pt_packed_eo_size(n)
                        pe, po = _lib.empty((size,)), _lib.empty((size,))
                        _lib.check(lib.pt_operator_pack_eo(_lib.dptr(d), n, _lib.dptr(pe), _lib.dptr(po),
                                                           _lib.stream_ptr()), "pt_operator_pack_eo")
                        pw = None
                        if n >= 128:      # large operators: also the layout of the warp-specialised kernel
                            pw = _lib.empty((lib.pt_packed_eo_ws_size(n),))
                            _lib.check(lib.pt_operator_pack_eo_ws(_lib.dptr(d), n, _lib.dptr(pw), _lib.stream_ptr()),
                                       "pt_operator_pack_eo_ws")
                        self._eo = (sign, pe, po, pw)
                        break
        return self._eo or None

    @property
    def even_odd(self):
        """True when ``apply`` runs the even-odd kernel (centrosymmetric operator, EVEN_ODD on)."""
        return bool(EVEN_ODD and self._band is None and self._eo_forms() is not None)

    # -- geometry -----------------------------------------------------------------------------------
    def _split(self, size, npts=None, axis=None):
        npts = self.npts if npts is None else tuple(npts)
        axis = self.axis if axis is None else axis
        if npts is None or axis is None:
            raise ValueError("operator is not bound to a grid axis (views lose the binding)")
        total = _prod(npts)
        if total == 0 or size % total:
            raise ValueError(
                f"field of {size} elements is not a whole number of fields on a grid of {total} points"
            )
        before = _prod(npts[:axis])
        after = _prod(npts[axis + 1:]) * (size // total)
        return after, int(npts[axis]), before

    # -- application ------------------------------------------------------------------------------
    def apply(self, u, out=None, alpha=1.0, beta=0.0, scale=None, scale_mode=None):
        """Differentiate ``u`` along this operator's grid axis.

        ``u``: one or more fields (flat in the Grid's linearisation, stacked along leading
        dimensions), CUDA ``torch`` tensor or host ndarray, float64.  Returns an array of the
        same kind and shape.  ``out`` (CUDA tensor) receives ``alpha*scale*(D u) + beta*out``;
        ``scale``/``scale_mode`` ('row' | 'after' | 'before') is the fused diagonal scaling of
        ``pt_dense_apply`` (dense operators only).
        """
        d_u, kind = _field_to_device(u)
        after, n, before = self._split(d_u.numel())
        d_y = self._apply_device(d_u, after, n, before, out, alpha, beta, scale, scale_mode)
        if kind == "numpy":
            return d_y.cpu().numpy().reshape(np.shape(u))
        return d_y.view(d_u.shape)

    def _prepare_out(self, d_u, out, beta):
        torch = _lib.torch_mod()
        if out is None:
            if beta != 0.0:
                raise ValueError("beta != 0 needs an `out` tensor to accumulate into")
            return torch.empty_like(d_u)
        if not (isinstance(out, torch.Tensor) and out.is_cuda and out.dtype == torch.float64
                and out.is_contiguous() and out.numel() == d_u.numel()):
            raise TypeError("out must be a contiguous float64 CUDA tensor of the field's size")
        if out.data_ptr() == d_u.data_ptr():
            raise ValueError("in-place application is not supported")
        return out

    def _apply_device(self, d_u, after, n, before, out, alpha, beta, scale, scale_mode):
        lib = _lib.load()
        d_y = self._prepare_out(d_u, out, beta)
        mode, d_scale, slen = _lib.SCALE_NONE, None, 0
        if scale is not None:
            mode = {"row": _lib.SCALE_ROW, "after": _lib.SCALE_AFTER,
                    "before": _lib.SCALE_BEFORE}[scale_mode]
            d_scale, _ = _field_to_device(scale)
            slen = d_scale.numel()
            need = {"row": n, "after": 1, "before": before}[scale_mode]
            if slen < need:
                raise ValueError("scale vector is too short")
        eo = self._eo_forms() if EVEN_ODD else None
        if eo is not None:
            if eo[3] is not None and EVEN_ODD_WS:
                # warp-specialised kernel (bulk-copied operator panels, mbarrier ring); -3 = shape not covered
                rc = lib.pt_dense_apply_eo_ws(_lib.dptr(eo[3]), n, eo[0], _lib.dptr(d_u), _lib.dptr(d_y), after, before,
                                              float(alpha), float(beta), _lib.dptr(d_scale), mode, slen,
                                              _lib.stream_ptr())
                if rc == 0:
                    return d_y
                if rc != -3:
                    _lib.check(rc, "pt_dense_apply_eo_ws")
            _lib.check(
                lib.pt_dense_apply_eo(_lib.dptr(eo[1]), _lib.dptr(eo[2]), n, eo[0], _lib.dptr(d_u), _lib.dptr(d_y),
                                      after, before, float(alpha), float(beta), _lib.dptr(d_scale), mode, slen,
                                      _lib.stream_ptr()),
                "pt_dense_apply_eo",
            )
            return d_y
        _lib.check(
            lib.pt_dense_apply(_lib.dptr(self._packed_device()), n, n, _lib.dptr(d_u),
                               _lib.dptr(d_y), after, before, float(alpha), float(beta),
                               _lib.dptr(d_scale), mode, slen, _lib.stream_ptr()),
            "pt_dense_apply",
        )
        return d_y

    def _apply_scatter(self, d_u, after, n, before, c, alpha, beta, smap, field_map=None):
        """``alpha * D u + beta * c`` stored through the pencil-redistribution map ``smap``
        (``pt_dense_apply_scatter``: the epilogue writes into the peers' buffers).  With ``field_map`` the
        field ``u`` itself is redistributed through that second map BY THE SAME KERNEL when the
        warp-specialised kernel covers the shape (``pt_dense_apply_eo_ws_scatter_copy``); returns True
        in that case, False when the caller still has to move ``u`` (``pt_peer_scatter``)."""
        lib = _lib.load()
        eo = self._eo_forms() if EVEN_ODD else None
        if eo is not None:
            if eo[3] is not None and EVEN_ODD_WS:
                if field_map is not None and FUSED_FIELD_COPY:
                    rc = lib.pt_dense_apply_eo_ws_scatter_copy(_lib.dptr(eo[3]), n, eo[0], _lib.dptr(d_u), _lib.dptr(c),
                                                               after, before, float(alpha), float(beta),
                                                               ctypes.byref(smap), ctypes.byref(field_map),
                                                               _lib.stream_ptr())
                    if rc == 0:
                        STATS["fused_field_copies"] += 1
                        return True
                    if rc != -3:
                        _lib.check(rc, "pt_dense_apply_eo_ws_scatter_copy")
                rc = lib.pt_dense_apply_eo_ws_scatter(_lib.dptr(eo[3]), n, eo[0], _lib.dptr(d_u), _lib.dptr(c), after,
                                                      before, float(alpha), float(beta), ctypes.byref(smap),
                                                      _lib.stream_ptr())
                if rc == 0:
                    return False
                if rc != -3:
                    _lib.check(rc, "pt_dense_apply_eo_ws_scatter")
            _lib.check(
                lib.pt_dense_apply_eo_scatter(_lib.dptr(eo[1]), _lib.dptr(eo[2]), n, eo[0], _lib.dptr(d_u),
                                              _lib.dptr(c), after, before, float(alpha), float(beta),
                                              ctypes.byref(smap), _lib.stream_ptr()),
                "pt_dense_apply_eo_scatter",
            )
            return False
        _lib.check(
            lib.pt_dense_apply_scatter(_lib.dptr(self._packed_device()), n, n, _lib.dptr(d_u), _lib.dptr(c),
                                       after, before, float(alpha), float(beta), ctypes.byref(smap),
                                       _lib.stream_ptr()),
            "pt_dense_apply_scatter",
        )
        return False

    def matmul(self, other):
        """``D @ other``.

        Whole fields on the bound grid are differentiated along the bound axis (the
        Kronecker-lifted product ``L_k @ u`` without forming ``L_k``): an operand whose last
        extent is ``prod(npts)``, or a flat vector holding a whole number of fields.  Anything else
        with leading extent ``n_in`` is multiplied as a matrix (``D @ M``), also for row/column
        slices of an operator (``D[1:-1, :] @ u``: a non-square ``n_out x n_in`` product).  On a 1-D
        grid (``prod(npts) == n``) a 2-D operand keeps NumPy's meaning, ``D @ M``; ``apply()`` is the
        unambiguous entry point for stacks of fields."""
        torch = _lib.torch_mod()
        if not (isinstance(other, (np.ndarray, torch.Tensor)) or hasattr(other, "__dlpack__")):
            other = np.asarray(other, dtype=np.float64)
        shape = tuple(other.shape)
        n_out, n_in = int(self.shape[0]), int(self.shape[1])
        if self.npts is not None and self.axis is not None and len(shape) >= 1:
            total = _prod(self.npts)
            size = _prod(shape)
            whole = total > 0 and size % total == 0
            if whole and (len(shape) == 1 or (shape[-1] == total and total != n_in)):
                return self.apply(other)
        if len(shape) >= 1 and shape[0] == n_in:
            d_u, kind = _field_to_device(other)
            cols = d_u.numel() // n_in
            d_y = self._matrix_product(d_u, cols)
            out_shape = (n_out,) + shape[1:]
            if kind == "numpy":
                return d_y.cpu().numpy().reshape(out_shape)
            return d_y.view(out_shape)
        if self.npts is None:
            raise ValueError(f"matmul: operand of shape {shape} does not match the {n_out} x {n_in} matrix "
                             "(views of an operator are plain matrices, not bound to a grid)")
        return self.apply(other)

    def _matrix_product(self, d_u, cols):
        """``D @ M`` for the ``n_out x n_in`` matrix of this array (plain DMMA kernel)."""
        lib = _lib.load()
        n_out, n_in = int(self.shape[0]), int(self.shape[1])
        d_y = _lib.empty((n_out * cols,))
        _lib.check(
            lib.pt_dense_apply(_lib.dptr(self._packed_device()), n_out, n_in, _lib.dptr(d_u),
                               _lib.dptr(d_y), 1, cols, 1.0, 0.0, None, _lib.SCALE_NONE, 0,
                               _lib.stream_ptr()),
            "pt_dense_apply",
        )
        return d_y

    def __matmul__(self, other):
        return self.matmul(other)

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        if ufunc is np.matmul and method == "__call__" and out is None and inputs[0] is self \
                and self.ndim == 2:
            return self.matmul(inputs[1])
        args = [np.asarray(i) if isinstance(i, _DiffMat) else i for i in inputs]
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, _DiffMat) else o for o in out)
        res = getattr(ufunc, method)(*args, **kwargs)
        return res

    def kron(self):
        """The explicit lifted matrix ``kron(I_after, kron(D, I_before))`` for one field on the
        bound grid (``P x P`` host ndarray; small grids only)."""
        lib = _lib.load()
        after, n, before = self._split(_prod(self.npts))
        total = after * n * before
        lifted = _lib.empty((total, total))
        _lib.check(
            lib.pt_kron_lift(_lib.dptr(self.device()), n, before, after, _lib.dptr(lifted),
                             _lib.stream_ptr()),
            "pt_kron_lift",
        )
        return lifted.cpu().numpy()

    def __repr__(self):
        tag = "unregistered" if self._key is None else "key " + "-".join(map(str, self._key))
        return (f"{type(self).__name__}(n={self.shape[0]}, axis={self.axis}, order={self.order}, "
                f"{tag})")


class ChebMat(_DiffMat):
    """Chebyshev collocation differentiation matrix on a Chebyshev-Gauss-Lobatto axis (any affine
    map, either orientation — it is built from the axis' actual nodes).

    ``ChebMat(grid, axis=0, order=1, weights="auto")``; ``grid`` may be a ``Grid``, a registered
    grid number or a 1-D node array.  Entries: D1_ij = (c_i/c_j)/(x_i-x_j); higher orders by
    Welfert's recursion; diagonals by the negative-sum trick.  Nodes that are not an affine image
    of the CGL set (a custom stretching mapper) get the polynomial collocation matrix on the
    actual nodes instead (barycentric weights, ``weights="bary"``).
    """

    _KIND = "cheb"

    def __new__(cls, grid, axis=0, order=1, weights="auto"):
        nodes, npts, axis, gnum = _resolve_axis(grid, axis)
        order = operator.index(order)
        if order < 1:
            raise ValueError("order must be a positive integer")
        if nodes.size < 2:
            raise ValueError("a Chebyshev axis needs at least 2 nodes")
        if weights not in ("auto", "cgl", "bary"):
            raise ValueError("weights must be 'auto', 'cgl' or 'bary'")
        if weights == "auto":
            weights = "cgl" if _is_affine_cgl(nodes) else "bary"
        lib = _lib.load()
        d_x = _lib.to_device(nodes)
        n = nodes.size
        d_mat = _lib.empty((n, n))
        if weights == "cgl":
            _lib.check(lib.pt_gen_chebmat(_lib.dptr(d_x), n, order, _lib.dptr(d_mat),
                                          _lib.stream_ptr()), "pt_gen_chebmat")
        else:
            # nodes from a non-affine mapper: polynomial collocation with barycentric weights
            work = _lib.empty((2 * n,))
            _lib.check(lib.pt_gen_barymat(_lib.dptr(d_x), n, order, _lib.dptr(d_mat), _lib.dptr(work),
                                          _lib.stream_ptr()), "pt_gen_barymat")
        if _nodes_symmetric(nodes):
            # flipping trick on node sets that are symmetric about their midpoint (every affine CGL
            # set; stretched meshes with an odd mapper): exactly centro(anti)symmetric matrix, which
            # the even-odd kernel applies with half the tensor-core work
            _lib.check(lib.pt_flip_symmetrize(_lib.dptr(d_mat), n, 1 if order % 2 == 0 else -1,
                                              _lib.stream_ptr()), "pt_flip_symmetrize")
        obj = cls._wrap(d_mat, npts, axis, order, gnum)
        obj.weights_kind = weights
        return obj


class FourMat(_DiffMat):
    """Fourier spectral differentiation matrix (orders 1-8) on an equispaced periodic axis
    whose endpoint is excluded, e.g. the theta axis of ``get_polar_grid`` (reference
    grid.py:505).  ``period`` defaults to ``n * spacing``.

    ``method`` selects how ``apply`` evaluates the product: ``"dense"`` (default; the matrix on the
    FP64 tensor cores, ``pt_dense_apply``) or ``"fft"`` (``pt_fourier_apply``: the same linear map
    through shared-memory FFTs, O(n log n) and HBM-bound; power-of-two ``n`` up to 4096, otherwise
    the dense path runs).  The ndarray itself is the dense matrix either way."""

    _KIND = "four"

    def __new__(cls, grid, axis=0, order=1, period=None, method="dense"):
        if method not in ("dense", "fft"):
            raise ValueError("FourMat: method must be 'dense' or 'fft'")
        nodes, npts, axis, gnum = _resolve_axis(grid, axis)
        order = operator.index(order)
        if not 1 <= order <= 8:
            raise ValueError("FourMat generates orders 1 to 8")
        n = nodes.size
        if n < 2:
            raise ValueError("a Fourier axis needs at least 2 nodes")
        if not _is_uniform(nodes):
            raise ValueError("FourMat needs an equispaced axis")
        if period is None:
            period = n * ((nodes[-1] - nodes[0]) / (n - 1))
        lib = _lib.load()
        d_mat = _lib.empty((n, n))
        if order <= 2:       # closed forms (Trefethen ch. 3)
            _lib.check(lib.pt_gen_fourmat(n, float(period), order, _lib.dptr(d_mat),
                                          _lib.stream_ptr()), "pt_gen_fourmat")
        else:                # direct summation of the symbol (i k s)^p
            work = _lib.empty((n,))
            _lib.check(lib.pt_gen_fourmat_general(n, float(period), order, _lib.dptr(d_mat),
                                                  _lib.dptr(work), _lib.stream_ptr()),
                       "pt_gen_fourmat_general")
        obj = cls._wrap(d_mat, npts, axis, order, gnum)
        obj.period = float(period)
        obj.method = method
        obj._tw = None
        return obj

    def __array_finalize__(self, obj):
        super().__array_finalize__(obj)
        self.period = getattr(obj, "period", None)
        self.method = "dense"
        self._tw = None

    def _apply_device(self, d_u, after, n, before, out, alpha, beta, scale, scale_mode):
        if (getattr(self, "method", "dense") == "fft" and (scale is None or scale_mode == "after")
                and self.period is not None):
            lib = _lib.load()
            if n >= 4 and n & (n - 1) == 0 and n <= 4096:
                d_scale, slen = None, 0
                if scale is not None:
                    d_scale, _ = _field_to_device(scale)
                    slen = d_scale.numel()
                if self._tw is None:
                    self._tw = _lib.empty((2 * n,))
                    _lib.check(lib.pt_fourier_twiddles(n, _lib.dptr(self._tw), _lib.stream_ptr()),
                               "pt_fourier_twiddles")
                d_y = self._prepare_out(d_u, out, beta)
                rc = lib.pt_fourier_apply(_lib.dptr(self._tw), n, float(self.period), int(self.order),
                                          _lib.dptr(d_u), _lib.dptr(d_y), after, before, float(alpha),
                                          float(beta), _lib.dptr(d_scale), slen, _lib.stream_ptr())
                if rc == 0:
                    return d_y
                if rc != -3:
                    _lib.check(rc, "pt_fourier_apply")
                out = d_y      # unsupported layout (odd stride / alignment): dense kernel below
        return super()._apply_device(d_u, after, n, before, out, alpha, beta, scale, scale_mode)


class FinDiffMat(_DiffMat):
    """Finite-difference differentiation matrix from Fornberg weights on the axis' actual nodes
    (mapped, non-uniform axes included).

    ``FinDiffMat(grid, axis=0, order=1, accuracy=2, periodic=False)``.  Interior rows are centred
    (``2*floor((order+1)/2) - 1 + accuracy`` nodes), rows near the ends one-sided
    (``order + accuracy`` nodes) unless ``periodic``.  The ndarray is the dense banded matrix;
    ``start``/``weights`` expose the compact row-start form the kernels read.
    """

    _KIND = "findiff"

    def __new__(cls, grid, axis=0, order=1, accuracy=2, periodic=False, uniform=None):
        nodes, npts, axis, gnum = _resolve_axis(grid, axis)
        order = operator.index(order)
        accuracy = operator.index(accuracy)
        if order < 1 or accuracy < 1 or accuracy % 2:
            raise ValueError("order must be >= 1 and accuracy a positive even integer")
        n = nodes.size
        w = 2 * ((order + 1) // 2) - 1 + accuracy
        wb = order + accuracy
        wmax = w if periodic else max(w, wb)
        if n < wmax:
            raise ValueError(f"axis of {n} nodes is shorter than the {wmax}-point stencil")
        if uniform is None:
            uniform = _is_uniform(nodes)
        if periodic and not _is_uniform(nodes):
            raise ValueError("periodic finite differences need an equispaced axis")
        h = float((nodes[-1] - nodes[0]) / (n - 1))
        lib = _lib.load()
        torch = _lib.require_cuda()
        d_x = _lib.to_device(nodes)
        d_start = torch.empty(n, dtype=torch.int32, device="cuda")
        d_w = _lib.empty((n, wmax))
        _lib.check(
            lib.pt_gen_fornberg(_lib.dptr(d_x), n, order, accuracy, int(bool(periodic)),
                                int(bool(uniform)), h, _lib.dptr(d_start), _lib.dptr(d_w), wmax,
                                _lib.stream_ptr()),
            "pt_gen_fornberg",
        )
        d_mat = _lib.empty((n, n))
        _lib.check(
            lib.pt_band_to_dense(_lib.dptr(d_w), _lib.dptr(d_start), wmax, n,
                                 int(bool(periodic)), _lib.dptr(d_mat), _lib.stream_ptr()),
            "pt_band_to_dense",
        )
        obj = cls._wrap(d_mat, npts, axis, order, gnum)
        hw = w // 2
        weights = d_w.cpu().numpy()
        if uniform:
            lo, hi = (0, n) if periodic else (hw, n - hw)
            wuni = np.ascontiguousarray(weights[lo if not periodic else 0, :w])
        else:
            lo, hi, wuni = 0, 0, None
        obj._band = dict(d_start=d_start, d_w=d_w, wmax=wmax, w=w, lo=lo, hi=hi, wuni=wuni,
                         periodic=bool(periodic), uniform=bool(uniform), h=h,
                         start=d_start.cpu().numpy(), weights=weights)
        obj.accuracy = accuracy
        obj.periodic = bool(periodic)
        return obj

    @property
    def start(self):
        """int32[n]: first column of every row's stencil (the bit-exact index layout)."""
        return self._band["start"]

    @property
    def weights(self):
        """float64[n, wmax]: stencil weights, zero padded."""
        return self._band["weights"]

    @property
    def uniform(self):
        return self._band["uniform"]

    def _apply_device(self, d_u, after, n, before, out, alpha, beta, scale, scale_mode):
        if scale is not None:
            raise ValueError("FinDiffMat.apply has no fused scaling")
        if self._band is None:
            raise ValueError("operator lost its banded form (views do); rebuild it")
        lib = _lib.load()
        band = self._band
        d_y = self._prepare_out(d_u, out, beta)
        if "uni" not in band:
            band["uni"] = _lib.uniform_rows(band, n)
        uni = band["uni"]
        _lib.check(
            lib.pt_stencil_apply(_lib.dptr(band["d_w"]), _lib.dptr(band["d_start"]), band["wmax"],
                                 n, _lib.dptr(d_u), _lib.dptr(d_y), after, before,
                                 int(band["periodic"]), ctypes.byref(uni) if uni is not None else None,
                                 float(alpha), float(beta), _lib.stream_ptr()),
            "pt_stencil_apply",
        )
        return d_y

    def _matrix_product(self, d_u, cols):
        if self._band is None or self.shape[0] != self.shape[1]:
            return super()._matrix_product(d_u, cols)     # row/column slice: a plain dense matrix
        return self._apply_device(d_u, 1, int(self.shape[1]), cols, None, 1.0, 0.0, None, None)


class PolarLaplacian:
    """Laplacian on a polar grid (axis 0 = theta, Fourier; axis 1 = r, Chebyshev or any mapped
    radial mesh without a node at r = 0):

        lap u = (D2_r + R^-1 D_r) u + R^-2 D2_theta u

    applied as two dense tensor-core passes; the second one accumulates into the first with the
    1/r^2 scaling fused in its epilogue (``pt_dense_apply`` with PT_SCALE_AFTER).
    """

    def __init__(self, grid, radial="cheb", accuracy=6, azimuthal="dense"):
        if isinstance(grid, (int, np.integer)):
            grid = getattr(_get_grid_manager(), str(int(grid)))
        if grid.num_dim != 2:
            raise ValueError("PolarLaplacian needs a 2-D polar grid")
        if radial not in ("cheb", "findiff"):
            raise ValueError("radial must be 'cheb' (CGL radial nodes) or 'findiff' (any mapped mesh)")
        self.npts = tuple(grid.npts)
        r = grid.axpoints(1)
        if np.any(r == 0.0):
            raise ValueError("the radial axis has a node at r = 0 (use fornberg=True)")
        self.r = r
        if radial == "cheb":
            d1 = ChebMat(grid, axis=1, order=1)
            d2 = ChebMat(grid, axis=1, order=2)
        else:   # arbitrary (mapped, non-uniform) radial mesh: Fornberg weights on the actual nodes
            d1 = FinDiffMat(grid, axis=1, order=1, accuracy=accuracy)
            d2 = FinDiffMat(grid, axis=1, order=2, accuracy=accuracy)
        _lib.require_cuda()
        d_r = _lib.to_device(r)
        radial_mat = d2.device() + d1.device() / d_r[:, None]
        if radial == "cheb" and np.max(np.abs(r + r[::-1])) <= 8.0 * np.finfo(np.float64).eps * np.max(np.abs(r)):
            # radial nodes symmetric about r = 0 (fornberg=True): D2 + D1/r is centrosymmetric; the
            # flipping trick makes it exactly so (the 1/r rows differ by an ulp of r otherwise)
            _lib.check(_lib.load().pt_flip_symmetrize(_lib.dptr(radial_mat), int(r.size), 1, _lib.stream_ptr()),
                       "pt_flip_symmetrize")
        self.radial = _DiffMat._wrap(radial_mat, self.npts, 1, 2, grid.num)
        # azimuthal="fft": D2_theta through pt_fourier_apply (1/r^2 still fused in its epilogue)
        self.azimuthal = FourMat(grid, axis=0, order=2, method=azimuthal)
        self.inv_r2 = 1.0 / (d_r * d_r)
        self.launches = 2      # kernels per application (radial + azimuthal)

    def apply(self, u, out=None):
        d_u, kind = _field_to_device(u)
        after, n, before = self._split(d_u.numel())
        d_y = self._apply_device(d_u, after, n, before, out, 1.0, 0.0, None, None)
        if kind == "numpy":
            return d_y.cpu().numpy().reshape(np.shape(u))
        return d_y.view(d_u.shape)

    @property
    def even_odd(self):
        """True when every dense pass of the Laplacian runs the even-odd kernel."""
        az = self.azimuthal
        return bool(self.radial.even_odd and (getattr(az, "method", "dense") == "fft" or az.even_odd))

    def executed_flop_ratio(self):
        """Executed / dense-form tensor-core flops of one application (0.5 per pass that runs the
        even-odd kernel; an FFT azimuthal pass counts on neither side)."""
        nt, nr = int(self.npts[0]), int(self.npts[1])
        dense = nr + (0 if getattr(self.azimuthal, "method", "dense") == "fft" else nt)
        done = nr * (0.5 if self.radial.even_odd else 1.0)
        if getattr(self.azimuthal, "method", "dense") != "fft":
            done += nt * (0.5 if self.azimuthal.even_odd else 1.0)
        return done / dense

    # the two methods the host pipeline (apply_host) and the benchmarks drive operators through
    def _split(self, size):
        return self.radial._split(size)

    def _apply_device(self, d_u, after, n, before, out, alpha, beta, scale, scale_mode):
        if scale is not None or alpha != 1.0 or beta != 0.0:
            raise ValueError("PolarLaplacian has no fused scaling or accumulation")
        d_y = self.radial._apply_device(d_u, after, n, before, out, 1.0, 0.0, None, None)
        a0, n0, b0 = self.azimuthal._split(d_u.numel())
        self.azimuthal._apply_device(d_u, a0, n0, b0, d_y, 1.0, 1.0, self.inv_r2, "after")
        return d_y

    __matmul__ = apply


# ---- registry ------------------------------------------------------------------------------------


def _get_mat_manager(kind, _managers={}):
    """One registry per operator class ('cheb' | 'four' | 'findiff')."""
    if kind not in ("cheb", "four", "findiff"):
        raise ValueError(f"unknown operator kind {kind!r}")
    return _managers.setdefault(kind, ObjectManager())


def _grid_number(grid):
    if isinstance(grid, (Grid, DeviceGrid)):
        if grid.num is None:
            raise ValueError("the grid is not registered: create it with get_grid")
        return int(grid.num)
    try:
        return operator.index(grid)
    except TypeError as exc:
        raise TypeError("grid must be a registered Grid or its integer identifier") from exc


def _get_mat(cls, grid, axis, order, extra, overwrite, **kwargs):
    gnum = _grid_number(grid)
    found = getattr(_get_grid_manager(), str(gnum), None)
    if found is None:
        raise ValueError(f"No grid with identifier {gnum} is registered.")
    try:
        axis = operator.index(axis)
        order = operator.index(order)
    except TypeError as exc:
        raise TypeError("axis and order must be integers") from exc
    if axis >= found.num_dim or axis < -found.num_dim:
        raise IndexError(f"axis {axis} out of bounds for a grid with {found.num_dim} dimensions")
    axis = axis + found.num_dim if axis < 0 else axis
    parts = [gnum, axis, order, *extra]
    name = "-".join(str(p) for p in parts)
    manager = _get_mat_manager(cls._KIND)
    mat = getattr(manager, name, None)
    if mat is None or overwrite:
        mat = cls(found, axis=axis, order=order, **kwargs)
        mat._key = parts
        setattr(manager, name, mat)
    return mat


def get_cheb_mat(grid, axis=0, order=1, overwrite=False):
    """Registered ``ChebMat`` for (grid, axis, order): built on first use, then returned as the
    same object (README.md:24-26 "re-usage of once allocated objects")."""
    return _get_mat(ChebMat, grid, axis, order, (), overwrite)


def get_four_mat(grid, axis=0, order=1, overwrite=False):
    """Registered ``FourMat`` for (grid, axis, order)."""
    return _get_mat(FourMat, grid, axis, order, (), overwrite)


def get_findiff_mat(grid, axis=0, order=1, accuracy=2, periodic=False, overwrite=False):
    """Registered ``FinDiffMat`` for (grid, axis, order, accuracy); periodic and non-periodic
    operators of the same key do not coexist (use ``overwrite=True`` to switch)."""
    return _get_mat(FinDiffMat, grid, axis, order, (operator.index(accuracy),), overwrite,
                    accuracy=accuracy, periodic=periodic)


def _drop_mat(kind, key, nitem):
    try:
        nitem = operator.index(nitem)
    except TypeError as exc:
        raise TypeError("Number of drop items nitem must be an integer.") from exc
    if nitem < 0:
        raise ValueError("Number of drop items nitem must be a positive integer.")
    manager = _get_mat_manager(kind)
    if key is None:
        if nitem == 0:
            warnings.warn("No operators were dropped because key is None and nitem=0.",
                          RuntimeWarning)
            return
        names = list(vars(manager))[-nitem:][::-1]
    else:
        if nitem:
            raise ValueError("key and nitem != 0 cannot be combined")
        keys = np.asarray(key)
        if not np.issubdtype(keys.dtype, np.integer):
            raise TypeError("key must be a sequence of integers (grid, axis, order[, accuracy])")
        if keys.ndim == 1:
            keys = keys[np.newaxis]
        if keys.ndim != 2:
            raise ValueError("key must be one composite key or a list of composite keys")
        names = ["-".join(str(v) for v in row) for row in keys.tolist()]
    for name in names:
        try:
            delattr(manager, name)
        except AttributeError:
            warnings.warn(f"Operator with key {name} could not be dropped as it's not registered "
                          f"in the manager.", RuntimeWarning)


def drop_cheb_mat(key=None, nitem=0):
    """Drop ``ChebMat``s by composite key(s) ``(grid, axis, order)`` or the last ``nitem``."""
    _drop_mat("cheb", key, nitem)


def drop_four_mat(key=None, nitem=0):
    """Drop ``FourMat``s by composite key(s) or the last ``nitem``."""
    _drop_mat("four", key, nitem)


def drop_findiff_mat(key=None, nitem=0):
    """Drop ``FinDiffMat``s by composite key(s) ``(grid, axis, order, accuracy)`` or the last
    ``nitem``."""
    _drop_mat("findiff", key, nitem)


def drop_all_mats():
    """Empty the three operator registries."""
    for kind in ("cheb", "four", "findiff"):
        manager = _get_mat_manager(kind)
        for name in list(vars(manager)):
            delattr(manager, name)
