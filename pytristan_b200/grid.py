"""Multi-dimensional meshes — drop-in for the reference's grid module.

Public surface and semantics follow the reference (src/pytristan/grid.py:23-32): ``cheb``,
``Grid`` (a ``numpy.ndarray`` subclass holding the linearised coordinate arrays, shape
``(ndim, prod(npts))``, axis 0 fastest), the registry getters ``get_grid`` /
``get_polar_grid`` and the ``drop_*`` family.

What is different underneath: the O(ndim * prod(npts)) expansion of the 1-D axis arrays into
the Grid layout — where the reference spends all of its time (three full-size NumPy copies,
grid.py:83-88) — runs on the GPU (``pt_grid_expand``), and the result is copied into the host
ndarray the API promises.  The 1-D node arrays themselves (O(sum(npts)) doubles) are produced
on the host with the reference's NumPy formulas, which is what makes the Grid bit-identical to
the reference's on the same host.  There is no CPU fallback for the
expansion: without a CUDA device constructing a Grid raises.
"""
import ctypes
import operator
import sys
import warnings

import numpy as np

from . import _lib
from ._manager import ObjectManager

__all__ = [
    "cheb",
    "cheb_device",
    "Grid",
    "DeviceGrid",
    "_get_grid_manager",
    "get_grid",
    "get_polar_grid",
    "drop_grid",
    "drop_last_grid",
    "drop_all_grids",
]

_PINNED_THRESHOLD = 8 << 20  # bytes: larger grids are read back through pinned memory


def cheb(npts, xmin=None, xmax=None):
    """(Mapped) Chebyshev-Gauss-Lobatto points.

    ``x_j = cos(pi*j/(N-1))``, j = 0..N-1 (descending from 1 to -1); when either bound is given
    the points are mapped to ``(1-x)/2*(xmax-xmin)+xmin`` (ascending from xmin to xmax).
    Same arithmetic, operation for operation, as the reference (grid.py:75-77), evaluated by
    NumPy on the host so the nodes are bit-identical to the reference's.

    Raises TypeError if ``npts`` is not an integer, ValueError if it is negative.
    """
    try:
        n = operator.index(npts)
    except TypeError as exc:
        raise TypeError("npts must be an integer.") from exc
    if n < 0:
        raise ValueError(f"Number of grid points, {n}, must be a positive integer.")
    nodes = np.cos(np.arange(n) * np.pi / (n - 1))
    if xmin is None and xmax is None:
        return nodes
    return (1.0 - nodes) / 2.0 * (xmax - xmin) + xmin


def cheb_device(npts, xmin=None, xmax=None):
    """``cheb`` evaluated on the GPU: a CUDA float64 tensor holding the same nodes **bit for bit**.

    The device kernel (``pt_gen_cheb_libm``) reproduces the cosine of the host's libm, which is
    what ``numpy.cos`` — and therefore the reference (grid.py:75) — evaluates; glibc's cosine is not
    correctly rounded, so an accurate device cosine differs from it at a few nodes for N >= 512.
    When the host libm is not one of the two glibc builds the kernel models, the nodes are
    computed by ``cheb`` and uploaded (still bit-identical; only O(N) values)."""
    try:
        n = operator.index(npts)
    except TypeError as exc:
        raise TypeError("npts must be an integer.") from exc
    if n < 0:
        raise ValueError(f"Number of grid points, {n}, must be a positive integer.")
    torch = _lib.require_cuda()
    model = _lib.host_libm_model()
    mapped = not (xmin is None and xmax is None)
    if model == 0 or n < 2:
        return torch.from_numpy(np.ascontiguousarray(cheb(n, xmin, xmax), dtype=np.float64)).cuda()
    lo, hi = (float(xmin), float(xmax)) if mapped else (0.0, 0.0)   # one bound None -> TypeError like the reference
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().pt_gen_cheb_libm(_lib.dptr(x), n, int(mapped), lo, hi, model, _lib.stream_ptr()),
               "pt_gen_cheb_libm")
    return x


def _expand_axes(axes, dtype):
    """Axis arrays -> ``(ndim, prod(npts))`` host ndarray in the Grid layout, via the GPU.

    out[d, l] = axes[d][(l // prod(npts[:d])) % npts[d]]  (reference grid.py:83-88).
    """
    ndim = len(axes)
    npts = [int(a.shape[0]) for a in axes]
    total = 1
    for n in npts:
        total *= n
    out = np.empty((ndim, total), dtype=dtype)
    if total == 0:
        return out
    if dtype.itemsize not in (1, 2, 4, 8, 16) or dtype.kind not in "biufc":
        raise TypeError(f"Grid: unsupported coordinate dtype {dtype}")
    torch = _lib.require_cuda()
    lib = _lib.load()
    flat = np.concatenate([np.ascontiguousarray(a, dtype=dtype) for a in axes])
    d_axes = torch.from_numpy(flat.view(np.uint8)).cuda()
    nbytes = ndim * total * dtype.itemsize
    d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    c_npts = (ctypes.c_int64 * ndim)(*npts)
    _lib.check(
        lib.pt_grid_expand(_lib.dptr(d_axes), c_npts, ndim, dtype.itemsize, _lib.dptr(d_out),
                           _lib.stream_ptr()),
        "pt_grid_expand",
    )
    if nbytes >= _PINNED_THRESHOLD:
        pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        pinned.copy_(d_out)
        out = pinned.numpy().view(dtype).reshape(ndim, total)
    else:
        torch.from_numpy(out.view(np.uint8).reshape(-1)).copy_(d_out)
    return out


def _expand_device_axes(d_axes, npts, dtype):
    """Concatenated axis arrays ALREADY on the device -> ``(ndim, prod(npts))`` CUDA tensor."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    ndim = len(npts)
    total = 1
    for n in npts:
        total *= int(n)
    out = torch.empty((ndim, total), dtype=d_axes.dtype, device="cuda")
    if total:
        c_npts = (ctypes.c_int64 * ndim)(*[int(n) for n in npts])
        _lib.check(
            lib.pt_grid_expand(_lib.dptr(d_axes), c_npts, ndim, np.dtype(dtype).itemsize, _lib.dptr(out),
                               _lib.stream_ptr()),
            "pt_grid_expand",
        )
    return out


def _expand_axes_device(axes, dtype):
    """Host axis arrays -> ``(ndim, prod(npts))`` CUDA tensor in the Grid layout (no host copy)."""
    torch = _lib.require_cuda()
    tdtype = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
              np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32}.get(np.dtype(dtype))
    if tdtype is None:
        raise TypeError(f"DeviceGrid: unsupported coordinate dtype {dtype}")
    npts = [int(a.shape[0]) for a in axes]
    flat = np.concatenate([np.ascontiguousarray(a, dtype=dtype) for a in axes])
    if flat.size == 0:
        return torch.empty((len(axes), 0), dtype=tdtype, device="cuda")
    return _expand_device_axes(torch.from_numpy(flat).cuda(), npts, dtype)


def _linspace_device(start, stop, num):
    """``numpy.linspace(start, stop, num)`` evaluated on the GPU (``pt_gen_linspace``: the same
    ``fl(fl(i*step)+start)`` arithmetic with explicit rounding, last node overwritten with ``stop``),
    bit-identical to NumPy's for float64 bounds."""
    torch = _lib.require_cuda()
    try:
        n = operator.index(num)
    except TypeError as exc:
        raise TypeError(f"object of type {type(num)} cannot be safely interpreted as an integer.") from exc
    if n < 0:
        raise ValueError(f"Number of samples, {n}, must be non-negative.")
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(_lib.load().pt_gen_linspace(_lib.dptr(x), float(start), float(stop), n, _lib.stream_ptr()),
               "pt_gen_linspace")
    return x


class DeviceGrid:
    """Device-resident mesh with the Grid layout: ``tensor`` is a CUDA
    ``(ndim, prod(npts))`` array with the same values, bit for bit, as ``Grid`` would hold, but it
    is never copied to the host — this is how meshes the reference cannot materialise on a host
    (1024^3: 24 GiB, ~72 GiB peak in reference grid.py:83-88) are built in milliseconds.

    ``from_bounds`` generates the 1-D nodes on the device as well: ``pt_gen_linspace`` for unmapped
    axes and ``cheb_device`` (``pt_gen_cheb_libm``, the host libm's cosine bit for bit) for axes
    mapped with ``cheb``; other mappers are called on the host as the reference does
    (grid.py:201-207).  ``axpoints`` is the strided read of a grid row (``pt_grid_axpoints``,
    reference grid.py:277-280) brought to the host.

    Same constructors and accessors as ``Grid`` (``from_bounds``, ``from_arrs``, ``axpoints``,
    ``npts``, ``num_dim``, ``num``); operators accept it wherever they accept a ``Grid``.
    """

    def __init__(self, arrs):
        torch = _lib.require_cuda()
        arrs = list(arrs)
        if not arrs:
            raise ValueError("DeviceGrid needs at least one axis")
        if any(isinstance(a, torch.Tensor) for a in arrs):
            # axes generated on the device (from_bounds): float64 nodes, never copied to the host
            d_axes = [a.to(device="cuda", dtype=torch.float64).reshape(-1) if isinstance(a, torch.Tensor)
                      else torch.from_numpy(np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.float64)).cuda()
                      for a in arrs]
            self.npts = tuple(int(a.numel()) for a in d_axes)
            self.num_dim = len(d_axes)
            self._num = None
            self.tensor = _expand_device_axes(torch.cat(d_axes), self.npts, np.dtype(np.float64))
            return
        axes = [np.asarray(a).ravel() for a in arrs]
        self.npts = tuple(len(a) for a in axes)
        self.num_dim = len(axes)
        self._num = None
        self.tensor = _expand_axes_device(axes, np.result_type(*axes))

    @classmethod
    def from_bounds(cls, *bounds, axes=[], mappers=[]):
        return cls(Grid._axes_from_bounds(bounds, axes, mappers, device=True))

    @classmethod
    def from_arrs(cls, *arrs):
        for arr in arrs:
            if np.asarray(arr).ndim != 1:
                raise ValueError("Coordinate arrays must be one-dimensional array-like.")
        return cls(arrs)

    @property
    def num(self):
        return self._num

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def axpoints_device(self, axis):
        """1-D nodes of ``axis`` as a CUDA tensor: strided gather of row ``axis`` of the grid
        (``pt_grid_axpoints``; float64 grids)."""
        torch = _lib.require_cuda()
        try:
            operator.index(axis)
        except TypeError as exc:
            raise TypeError("Axis' index axis must be an integer.") from exc
        if axis >= self.num_dim or axis < -self.num_dim:
            raise IndexError(f"Axis' index {axis} out of bounds for the grid with {self.num_dim} dimensions.")
        axis = axis + self.num_dim if axis < 0 else axis
        stride = 1
        for n in self.npts[:axis]:
            stride *= n
        n = self.npts[axis]
        if self.tensor.dtype != torch.float64:
            return self.tensor[axis, :stride * n:stride].contiguous()
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().pt_grid_axpoints(_lib.dptr(self.tensor[axis]), stride, n, _lib.dptr(x),
                                                _lib.stream_ptr()), "pt_grid_axpoints")
        return x

    def axpoints(self, axis):
        return self.axpoints_device(axis).cpu().numpy()

    def to_host(self):
        """The equivalent host ``Grid`` (copies ``ndim * prod(npts)`` elements)."""
        return Grid([self.axpoints(k) for k in range(self.num_dim)])

    def __repr__(self):
        tag = "no unique identifier" if self._num is None else f"unique identifier: {self._num}"
        return f"Instance of {type(self).__name__} with {tag}."


class Grid(np.ndarray):
    """Linearised N-D mesh: row ``i`` holds the coordinate along axis ``i`` of every grid point,
    points ordered with axis 0 fastest (column-major linearisation, reference grid.py:85-88).

    Attributes: ``num`` (registry identifier or None), ``num_dim``, ``npts``.
    """

    def __new__(cls, arrs):
        axes = [np.asarray(a).ravel() for a in arrs]
        if not axes:
            obj = np.empty((0,), dtype=np.float64).view(cls)
        else:
            obj = _expand_axes(axes, np.result_type(*axes)).view(cls)
        obj._num = None
        obj.num_dim = obj.shape[0]
        obj.npts = tuple(len(a) for a in arrs)
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        # Views, slices and ufunc results do not inherit the layout-dependent attributes: the
        # user is responsible for them (reference grid.py:99-109).
        self._num = None
        self.num_dim = None
        self.npts = None

    @classmethod
    def from_bounds(cls, *bounds, axes=[], mappers=[]):
        """Grid from ``(lower, upper, npts)`` triples, one per dimension (see
        ``_axes_from_bounds`` for the validation and error matrix)."""
        return cls(cls._axes_from_bounds(bounds, axes, mappers))

    @staticmethod
    def _axes_from_bounds(bounds, axes, mappers, device=False):
        """1-D axis arrays from ``(lower, upper, npts)`` triples, one per dimension.

        ``axes``/``mappers``: indices of the axes to map and the mapper for each, called as
        ``mapper(npts, lower, upper)``; unmapped axes use ``numpy.linspace``.

        Raises ValueError (malformed bound, len(axes) != len(mappers), repeated axes),
        TypeError (axes/mappers not sized, non-integer axes), IndexError (axis out of range) —
        the matrix of reference grid.py:163-199.
        """
        for bound in bounds:
            if np.asarray(bound).shape != (3,):
                raise ValueError(
                    "Wrong bound format: each bound must be one-dimensional and hold three "
                    "elements: lower bound, upper bound and number of grid points."
                )
        try:
            lengths_match = len(axes) == len(mappers)
        except TypeError as exc:
            raise TypeError("Axes and mappers must be one-dimensional array-like.") from exc
        if not lengths_match:
            raise ValueError("The number of axes must be equal to the number of mappers.")

        ndim = len(bounds)
        mapper_of = {}
        if len(axes):
            idx = np.asarray(axes)
            if not np.issubdtype(idx.dtype, np.integer):
                raise TypeError("Axes' indices in `axes` must be integer numbers.")
            if ((idx >= ndim) | (idx < -ndim)).any():
                raise IndexError(
                    f"Axis' index out of bounds for the grid with {ndim} dimensions in axes. "
                    f"Allowed interval is from {-ndim} to {ndim - 1}."
                )
            idx = np.where(idx < 0, idx + ndim, idx)
            if np.unique(idx).size != idx.size:
                raise ValueError(
                    "All indices in `axes` must be unique; check for repeated values and for "
                    "negative indices that refer to the same axes as positive ones."
                )
            mapper_of = dict(zip(idx.tolist(), mappers))

        # A mapped axis is built as the reference builds it (grid.py:203): the call
        # ``mapper(*((bound[2],) + bound[:2]))`` — which is also what makes a *list* bound of a
        # mapped axis a TypeError there (tuple + list), kept here.
        if device:       # DeviceGrid: nodes generated on the GPU, bit-identical to the host formulas
            return tuple(
                (cheb_device(*((bound[2],) + bound[:2])) if mapper_of[k] is cheb
                 else mapper_of[k](*((bound[2],) + bound[:2]))) if k in mapper_of
                else _linspace_device(*bound)
                for k, bound in enumerate(bounds)
            )
        return tuple(
            mapper_of[k](*((bound[2],) + bound[:2])) if k in mapper_of else np.linspace(*bound)
            for k, bound in enumerate(bounds)
        )

    @classmethod
    def from_arrs(cls, *arrs):
        """Grid from 1-D coordinate arrays.  ValueError if any of them is not 1-D."""
        for arr in arrs:
            if np.asarray(arr).ndim != 1:
                raise ValueError("Coordinate arrays must be one-dimensional array-like.")
        return cls(arrs)

    def __repr__(self):
        tag = "no unique identifier" if self._num is None else f"unique identifier: {self._num}"
        return f"Instance of {type(self).__name__} with {tag}."

    def axpoints(self, axis):
        """1-D coordinates along ``axis`` (negative indices allowed).

        TypeError for a non-integer axis, IndexError when out of range (reference
        grid.py:264-273).  Implemented, as in the reference, as a strided read of row ``axis``.
        """
        try:
            operator.index(axis)
        except TypeError as exc:
            raise TypeError("Axis' index axis must be an integer.") from exc
        ndim = self.num_dim
        if axis >= ndim or axis < -ndim:
            raise IndexError(
                f"Axis' index {axis} out of bounds for the grid with {ndim} dimensions in axes. "
                f"Allowed interval is from {-ndim} to {ndim - 1}."
            )
        axis = axis + ndim if axis < 0 else axis
        stride = 1
        for n in self.npts[:axis]:
            stride *= n
        stop = stride * self.npts[axis]
        return np.array(np.asarray(self)[axis, :stop:stride])

    def mgrids(self):
        """Coordinate matrices of shape ``npts``, equal to ``numpy.meshgrid(..., indexing='ij')``."""
        return [row.reshape(self.npts, order="F") for row in self]

    @property
    def num(self):
        return self._num

    @num.setter
    def num(self, number):
        # Only get_grid may name a grid: it also registers it (reference grid.py:300-320).
        caller = sys._getframe(1)
        fname, func = caller.f_code.co_filename, caller.f_code.co_name
        if not fname.endswith("get_grid.py") and "get_grid" not in func:
            raise RuntimeError(
                f"{func} in {fname} attempted to assign a name to an instance of "
                f"{type(self).__name__}. Only pytristan.get_grid may do that: it registers the "
                f"instance in the grid manager so that it can be re-used."
            )
        self._num = number


def _get_grid_manager(_grid_manager_instance=ObjectManager()):
    """The process-wide grid registry (single instance bound at definition time)."""
    return _grid_manager_instance


def get_grid(*args, from_bounds=False, axes=[], mappers=[], num=None, overwrite=False, device=False):
    """Return the registered grid ``num`` or build, register and return a new one.

    ``args`` are 1-D coordinate arrays, or ``(lower, upper, npts)`` bounds when
    ``from_bounds=True``.  Without ``num`` the new grid gets ``max(existing) + 1`` (0 for an
    empty registry).  ``overwrite=True`` rebuilds an existing identifier.  ``device=True`` (an
    extension) registers a ``DeviceGrid`` that lives on the GPU only.

    Raises TypeError for a non-integer ``num`` and ValueError when the grid does not exist and
    no data is given (reference grid.py:392-421).
    """
    manager = _get_grid_manager()
    if num is None:
        known = manager.nums()
        num = max(known) + 1 if known else 0
    else:
        try:
            operator.index(num)
        except TypeError as exc:
            raise TypeError("Unique identifier num must be an integer") from exc

    grid = getattr(manager, str(num), None)
    if grid is None or overwrite:
        if not args:
            raise ValueError("Could not create a new grid - no grid data supplied.")
        kind = DeviceGrid if device else Grid
        if from_bounds:
            grid = kind.from_bounds(*args, axes=axes, mappers=mappers)
        else:
            grid = kind.from_arrs(*args)
        if device:
            grid._num = num
        else:
            grid.num = num
        setattr(manager, str(num), grid)
    return grid


def get_polar_grid(nt, nr, axes=[], mappers=[], fornberg=False, num=None):
    """2-D polar grid: axis 0 is theta on ``[-pi, pi)`` (``nt`` equispaced points, endpoint
    excluded), axis 1 is r on ``[0, 1]`` with ``nr`` points — or, with ``fornberg=True``, on
    ``[-1, 1]`` with ``2*nr`` points (Fornberg's redundant grid that avoids treating r = 0 as a
    boundary; reference grid.py:498-511)."""
    if fornberg:
        nr, rmin = 2 * nr, -1.0
    else:
        rmin = 0.0
    return get_grid(
        (-np.pi, np.pi - 2.0 * np.pi / nt, nt),
        (rmin, 1.0, nr),
        from_bounds=True,
        num=num,
        axes=axes,
        mappers=mappers,
    )


def _drop_from(manager, what, num, nitem):
    """Shared drop logic of the grid and operator registries (reference grid.py:559-600)."""
    try:
        nitem = operator.index(nitem)
    except TypeError as exc:
        raise TypeError("Number of drop items nitem must be an integer.") from exc
    if nitem < 0:
        raise ValueError("Number of drop items nitem must be a positive integer.")

    if num is None:
        if nitem == 0:
            warnings.warn(
                f"No {what}s were dropped because num is None and nitem=0.", RuntimeWarning
            )
            return
        names = list(vars(manager))[-nitem:][::-1]
    else:
        if nitem:
            raise ValueError(
                "Providing num different from None and nitem different from 0 at the same time "
                f"is ambiguous: drop the last N {what}s and the {what}s with given identifiers in "
                "two consecutive calls."
            )
        ids = np.asarray(num)
        if not np.issubdtype(ids.dtype, np.integer):
            raise TypeError("num must be an integer or an array-like of integers.")
        if ids.ndim == 0:
            ids = ids[np.newaxis]
        names = _names_from_ids(ids)

    for name in names:
        try:
            delattr(manager, name)
        except AttributeError:
            warnings.warn(
                f"{what.capitalize()} with identifier {name} could not be dropped as it's not "
                f"registered in the manager.",
                RuntimeWarning,
            )


def _names_from_ids(ids):
    if ids.ndim != 1:
        raise ValueError("num cannot have more that one dimension.")
    return [str(i) for i in ids.tolist()]


def drop_grid(num=None, nitem=0):
    """Remove grids from the registry: those with identifier(s) ``num``, or the last ``nitem``.

    TypeError / ValueError / RuntimeWarning matrix of reference grid.py:559-600.
    """
    _drop_from(_get_grid_manager(), "grid", num, nitem)


def drop_last_grid():
    """Drop the most recently registered grid."""
    drop_grid(nitem=1)


def drop_all_grids():
    """Drop every registered grid.  Like the reference (grid.py:608-610) this is
    ``drop_grid(num=manager.nums())``, so on an EMPTY registry it raises TypeError (an empty list is
    not an array of integers) — callers that may find the registry empty check ``nums()`` first."""
    drop_grid(num=_get_grid_manager().nums())
