"""GPU parity: the CUDA path (through the C-ABI, via the Python façade) against the oracle on
the same seeded inputs.  Tolerances follow BASELINE.json / SURVEY.md §8(d): index layout and
grid construction bit-exact; derivatives within 1e-12 componentwise relative to (|D||u|)_i and
1e-12 normwise on the N(0,1) synthetic fields."""
import numpy as np
import pytest
from conftest import clear_grids

from oracle import operators as oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _torch():
    import torch

    return torch


def rel_componentwise(y, yref, bound):
    return float(np.max(np.abs(y - yref) / np.maximum(bound, 1e-300)))


def rel_normwise(y, yref):
    return float(np.linalg.norm((y - yref).ravel()) / max(np.linalg.norm(yref.ravel()), 1e-300))


def check_apply(op, D, U, npts, axis):
    y = op.apply(U)
    yref = oracle.apply_along_axis(D, U, npts, axis)
    bound = oracle.abs_bound(D, U, npts, axis)
    assert y.shape == U.shape
    assert rel_componentwise(y, yref, bound) <= RTOL
    assert rel_normwise(y, yref) <= RTOL


# ---------------------------------------------------------------------------------------------
# grids
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(7,), (2, 3), (5, 1, 4), (33, 17, 9), (1024, 257), (3, 4, 5, 6)])
def test_grid_expand_bit_exact(shape):
    import pytristan_b200 as pt

    rng = np.random.default_rng(1)
    axes = [np.sort(rng.standard_normal(n)) for n in shape]
    g = pt.Grid.from_arrs(*axes)
    ref = oracle.grid_expand(axes)
    assert g.shape == ref.shape and g.dtype == ref.dtype
    assert g.tobytes() == ref.tobytes()
    assert g.npts == shape and g.num_dim == len(shape)
    for k in range(len(shape)):
        assert np.array_equal(g.axpoints(k), axes[k])


def test_grid_dtypes_and_empty():
    import pytristan_b200 as pt

    g = pt.Grid.from_arrs(np.arange(3), np.arange(4))
    assert g.dtype == np.int64 and np.array_equal(np.asarray(g), oracle.grid_expand([np.arange(3), np.arange(4)]))
    g = pt.Grid.from_arrs(np.arange(3, dtype=np.float32), np.arange(2, dtype=np.float32))
    assert g.dtype == np.float32 and g.shape == (2, 6)
    g = pt.Grid.from_arrs(np.arange(3) + 1j, np.arange(2.0))
    assert g.dtype == np.complex128 and np.array_equal(np.asarray(g), oracle.grid_expand([np.arange(3) + 1j, np.arange(2.0)]))
    g = pt.Grid.from_arrs(np.arange(0.0), np.arange(3.0))
    assert g.shape == (2, 0)
    g = pt.Grid.from_arrs(np.array([True, False]), np.array([False, True, True]))
    assert g.dtype == np.bool_ and g.shape == (2, 6)


def test_grid_golden_vector():
    # reference test_grid.py:69-74 golden layout
    import pytristan_b200 as pt

    arr = np.array([[0.0, 1.0, 0.0, 1.0, 0.0, 1.0], [-1.0, -1.0, 0.0, 0.0, 1.0, 1.0]])
    assert np.array_equal(arr, pt.Grid.from_bounds((0.0, 1.0, 2), (-1.0, 1.0, 3)))
    assert np.array_equal(arr, pt.Grid.from_arrs(np.arange(2.0), np.linspace(-1.0, 1.0, 3)))


def test_device_generators():
    from pytristan_b200 import _lib

    lib = _lib.load()
    torch = _torch()
    for (a, b, n) in [(0.0, 1.0, 2), (-1.0, 1.0, 257), (0.0, 2 * np.pi - 2 * np.pi / 1024, 1024),
                      (-np.pi, np.pi - 2.0 * np.pi / 1000, 1000), (3.0, 3.0, 5), (0.5, -2.5, 1), (0.1, 0.7, 4097)]:
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        _lib.check(lib.pt_gen_linspace(_lib.dptr(x), a, b, n, _lib.stream_ptr()), "linspace")
        assert x.cpu().numpy().tobytes() == np.linspace(a, b, n).tobytes(), (a, b, n)
    worst = 0
    for n in (2, 10, 64, 257, 1024, 2048):
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        _lib.check(lib.pt_gen_cheb(_lib.dptr(x), n, 0, 0.0, 0.0, _lib.stream_ptr()), "cheb")
        ref = oracle.cheb_nodes(n)
        got = x.cpu().numpy()
        ulp = np.abs(got - ref) / np.spacing(np.maximum(np.abs(ref), 1e-300))
        worst = max(worst, float(ulp[np.abs(ref) > 1e-3].max()))
        assert np.max(np.abs(got - ref)) <= 2.3e-16
        # the affine map itself is bit exact given the nodes
        xm = torch.from_numpy(ref.copy()).cuda()
        _lib.check(lib.pt_cheb_map(_lib.dptr(xm), n, -2.0, 5.0, _lib.stream_ptr()), "map")
        assert xm.cpu().numpy().tobytes() == oracle.cheb_nodes(n, -2.0, 5.0).tobytes()
    assert worst <= 2.0


# ---------------------------------------------------------------------------------------------
# operator generation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [2, 5, 64, 257])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_chebmat_entries(n, order):
    import pytristan_b200 as pt

    if order > n - 1:
        pytest.skip("order too high for the node count")
    for x in (oracle.cheb_nodes(n), oracle.cheb_nodes(n, 0.0, 3.0)):
        D = pt.ChebMat(x, order=order)
        ref = oracle.cheb_mat(x, order)
        assert isinstance(D, np.ndarray) and D.shape == (n, n)
        scale = np.abs(ref).sum(axis=1, keepdims=True)
        assert np.max(np.abs(np.asarray(D) - ref) / scale) <= 1e-14
        # off-diagonal entries are produced by the same IEEE operations: bit identical
        off = ~np.eye(n, dtype=bool)
        assert np.array_equal(np.asarray(D)[off], ref[off])


@pytest.mark.parametrize("n", [2, 3, 16, 17, 1024])
@pytest.mark.parametrize("order", [1, 2])
def test_fourmat_entries(n, order):
    import pytristan_b200 as pt

    period = 2 * np.pi
    x = -np.pi + period / n * np.arange(n)
    D = pt.FourMat(x, order=order, period=period)
    ref = oracle.four_mat(n, period, order)
    assert np.max(np.abs(np.asarray(D) - ref)) <= 1e-13 * np.max(np.abs(ref))


@pytest.mark.parametrize("order,accuracy", [(1, 2), (1, 6), (2, 6), (3, 4), (4, 2)])
@pytest.mark.parametrize("kind", ["uniform", "cheb", "periodic"])
def test_findiff_band(order, accuracy, kind):
    import pytristan_b200 as pt

    n = 40
    if kind == "uniform":
        x, periodic = np.linspace(-1.0, 2.0, n), False
    elif kind == "cheb":
        x, periodic = oracle.cheb_nodes(n, 0.0, 1.0), False
    else:
        x, periodic = np.linspace(0.0, 2 * np.pi - 2 * np.pi / n, n), True
    F = pt.FinDiffMat(x, order=order, accuracy=accuracy, periodic=periodic)
    band = oracle.findiff_band(x, order, accuracy, periodic)
    assert F.start.dtype == np.int32 and np.array_equal(F.start, band["start"])   # index layout: bit exact
    assert F.weights.shape == band["W"].shape
    assert np.array_equal(F.weights != 0.0, band["W"] != 0.0)
    scale = np.abs(band["W"]).sum(axis=1, keepdims=True)
    assert np.max(np.abs(F.weights - band["W"]) / scale) <= 1e-13
    ref = oracle.band_to_dense(band["start"], band["W"], periodic)
    assert np.max(np.abs(np.asarray(F) - ref)) <= 1e-13 * np.max(np.abs(ref))
    assert F.uniform == band["uniform"]


# ---------------------------------------------------------------------------------------------
# dense application (K1 / K2)
# ---------------------------------------------------------------------------------------------
def test_config1_cheb64_sin():
    """BASELINE config 1: 1D ChebMat N=64 first derivative of sin(x)."""
    import pytristan_b200 as pt

    g = pt.Grid.from_bounds((-1.0, 1.0, 64), axes=[0], mappers=[pt.cheb])
    x = g.axpoints(0)
    D = pt.ChebMat(g, axis=0, order=1)
    u = np.sin(x)
    y = D @ u
    yref = oracle.cheb_mat(x, 1) @ u
    assert np.max(np.abs(y - np.cos(x))) < 2e-12          # spectral accuracy (known answer)
    assert np.max(np.abs(y - yref)) / np.max(np.abs(yref)) < 1e-12
    assert np.max(np.abs(np.matmul(D, u) - y)) == 0.0


DENSE_SHAPES = [
    # npts, axis, batch
    ((64,), 0, 1),
    ((64,), 0, 5),
    ((9, 7), 0, 3),
    ((9, 7), 1, 3),
    ((8, 6, 10), 1, 2),
    ((8, 6, 10), 2, 1),
    ((33, 257), 1, 2),      # odd `before`: 8-byte staging path
    ((257, 12), 0, 3),      # odd n on the contiguous axis: 8-byte staging path
    ((130, 40), 0, 7),
    ((48, 136), 1, 3),
    ((256, 200), 0, 1),
    ((1024, 64), 0, 2),
    ((128, 520), 1, 1),
]


@pytest.mark.parametrize("npts,axis,batch", DENSE_SHAPES)
def test_dense_apply_random(npts, axis, batch):
    import pytristan_b200 as pt

    rng = np.random.default_rng(1234)
    n = npts[axis]
    D = rng.standard_normal((n, n))
    # a random matrix bound to the axis: exercises the kernels on dense data
    op2 = pt.operators._DiffMat._wrap(pt.operators._lib.to_device(D), npts, axis, 1, None)
    U = rng.standard_normal((batch, int(np.prod(npts))))
    check_apply(op2, D, U, npts, axis)
    # torch in, torch out, same numbers
    torch = _torch()
    yt = op2.apply(torch.from_numpy(U).cuda())
    assert yt.is_cuda and np.array_equal(yt.cpu().numpy(), op2.apply(U))


def test_dense_alpha_beta_scale():
    import pytristan_b200 as pt

    torch = _torch()
    rng = np.random.default_rng(7)
    npts, batch = (24, 40), 3
    for axis in (0, 1):
        n = npts[axis]
        D = rng.standard_normal((n, n))
        op = pt.operators._DiffMat._wrap(pt.operators._lib.to_device(D), npts, axis, 1, None)
        U = rng.standard_normal((batch, 24 * 40))
        Y0 = rng.standard_normal((batch, 24 * 40))
        base = oracle.apply_along_axis(D, U, npts, axis)
        after, _, before = oracle.axis_split(npts, axis, U.size)
        for mode, length in (("row", n), ("after", 5), ("before", before)):
            s = rng.standard_normal(length)
            out = torch.from_numpy(Y0.copy()).cuda()
            op.apply(torch.from_numpy(U).cuda(), out=out, alpha=0.5, beta=-2.0, scale=s, scale_mode=mode)
            V = base.reshape(after, n, before)
            if mode == "row":
                S = s[None, :, None]
            elif mode == "after":
                S = s[np.arange(after) % length][:, None, None]
            else:
                S = s[None, None, :]
            ref = (0.5 * S * V).reshape(U.shape) - 2.0 * Y0
            assert np.max(np.abs(out.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_config2_channel_grid():
    """BASELINE config 2 at a reduced batch: Fourier-x (1024) x Chebyshev-y (257), d/dx and
    d2/dy2 of stacked fields (full batch of 64 is exercised by bench.py and the property test)."""
    import pytristan_b200 as pt

    g = pt.Grid.from_bounds((0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), (-1.0, 1.0, 257), axes=[1], mappers=[pt.cheb])
    Dx = pt.FourMat(g, axis=0, order=1)
    Dyy = pt.ChebMat(g, axis=1, order=2)
    rng = np.random.default_rng(1234)
    U = rng.standard_normal((4, 1024 * 257))
    x, y = g.axpoints(0), g.axpoints(1)
    check_apply(Dx, oracle.four_mat(1024, 1024 * ((x[-1] - x[0]) / 1023), 1), U, g.npts, 0)
    check_apply(Dyy, oracle.cheb_mat(y, 2), U, g.npts, 1)
    # smooth field: d/dx of sin(3x)cos(y) and d2/dy2 — known answers
    X, Y = np.asarray(g)[0], np.asarray(g)[1]
    f = np.sin(3 * X) * np.cos(Y)
    assert np.max(np.abs(Dx @ f - 3 * np.cos(3 * X) * np.cos(Y))) < 1e-10
    assert np.max(np.abs(Dyy @ f + f)) < 1e-6


def test_linearity_and_kron():
    import pytristan_b200 as pt

    g = pt.Grid.from_bounds((0.0, 1.0, 6), (-1.0, 1.0, 9), (0.0, 2.0, 5), axes=[1], mappers=[pt.cheb])
    rng = np.random.default_rng(3)
    u, v = rng.standard_normal(6 * 9 * 5), rng.standard_normal(6 * 9 * 5)
    for axis in (0, 1, 2):
        op = pt.ChebMat(g, axis=axis, order=1) if axis == 1 else pt.FinDiffMat(g, axis=axis, order=1, accuracy=2)
        L = op.kron()
        ref = oracle.kron_lift(np.asarray(op), g.npts, axis)
        assert np.array_equal(L, ref)
        y = op @ u
        assert np.max(np.abs(L @ u - y)) <= 1e-12 * np.max(np.abs(y))
        z = op @ (2.0 * u - 3.0 * v)
        assert np.max(np.abs(z - (2.0 * (op @ u) - 3.0 * (op @ v)))) <= 1e-11 * np.max(np.abs(z))


# ---------------------------------------------------------------------------------------------
# stencils (K3 / K4 / table)
# ---------------------------------------------------------------------------------------------
STENCIL_SHAPES = [((40,), 0, 1), ((64, 9), 0, 2), ((9, 64), 1, 2), ((33, 31, 12), 1, 1), ((16, 8, 50), 2, 3),
                  ((1000, 7), 0, 1), ((7, 1000), 1, 1), ((257, 33), 0, 2), ((33, 257), 1, 2), ((4098, 3), 0, 1)]


@pytest.mark.parametrize("npts,axis,batch", STENCIL_SHAPES)
@pytest.mark.parametrize("order,accuracy", [(1, 6), (2, 6), (1, 2)])
@pytest.mark.parametrize("kind", ["uniform", "mapped", "periodic"])
def test_stencil_apply(npts, axis, batch, order, accuracy, kind):
    import pytristan_b200 as pt

    n = npts[axis]
    axes = [np.linspace(0.0, 1.0, m) for m in npts]
    periodic = kind == "periodic"
    if kind == "mapped":
        axes[axis] = oracle.cheb_nodes(n, 0.0, 1.0)
    elif periodic:
        axes[axis] = np.linspace(0.0, 1.0 - 1.0 / n, n)
    g = pt.Grid.from_arrs(*axes)
    F = pt.FinDiffMat(g, axis=axis, order=order, accuracy=accuracy, periodic=periodic)
    band = oracle.findiff_band(axes[axis], order, accuracy, periodic)
    rng = np.random.default_rng(1234)
    U = rng.standard_normal((batch, int(np.prod(npts))))
    y = F.apply(U)
    yref = oracle.stencil_apply(band, U, npts, axis)
    bound = oracle.abs_bound(oracle.band_to_dense(band["start"], band["W"], periodic), U, npts, axis)
    assert rel_componentwise(y, yref, bound) <= RTOL
    assert rel_normwise(y, yref) <= RTOL


def test_stencil_polynomial_exactness():
    import pytristan_b200 as pt

    x = np.linspace(-1.0, 1.0, 64)
    g = pt.Grid.from_arrs(x, np.linspace(0, 1, 5))
    F = pt.FinDiffMat(g, axis=0, order=1, accuracy=6)
    X = np.asarray(g)[0]
    y = F @ (X ** 5)
    assert np.max(np.abs(y - 5 * X ** 4)) < 1e-11


# ---------------------------------------------------------------------------------------------
# polar Laplacian (config 3 at reduced size) and permutation copy
# ---------------------------------------------------------------------------------------------
def test_polar_laplacian():
    import pytristan_b200 as pt

    clear_grids()
    g = pt.get_polar_grid(64, 24, axes=[1], mappers=[pt.cheb], fornberg=True)
    lap = pt.PolarLaplacian(g)
    th, r = g.axpoints(0), g.axpoints(1)
    T, R = np.asarray(g)[0], np.asarray(g)[1]
    rng = np.random.default_rng(1234)
    U = rng.standard_normal((3, T.size))
    y = lap.apply(U)
    yref = oracle.polar_laplacian(th, r, U)
    assert rel_normwise(y, yref) <= 1e-12
    # harmonic function r^3 cos(3 theta): Laplacian vanishes; r^2: Laplacian = 4
    f = R ** 3 * np.cos(3 * T)
    assert np.max(np.abs(lap.apply(f))) < 1e-8
    assert np.max(np.abs(lap.apply(R ** 2) - 4.0)) < 1e-8
    clear_grids()


@pytest.mark.parametrize("dims", [(5, 7, 9), (33, 65, 34), (1, 40, 70)])
def test_permute3d(dims):
    import itertools

    from pytristan_b200 import _lib

    torch = _torch()
    lib = _lib.load()
    d2, d1, d0 = dims
    a = np.random.default_rng(0).standard_normal(dims)
    t = torch.from_numpy(a).cuda()
    for p in itertools.permutations(range(3)):
        out = torch.empty(a.size, dtype=torch.float64, device="cuda")
        code = p[0] + 3 * p[1] + 9 * p[2]
        _lib.check(lib.pt_permute3d(_lib.dptr(t), _lib.dptr(out), d2, d1, d0, code, _lib.stream_ptr()), "permute")
        ref = np.ascontiguousarray(np.transpose(a, p))
        assert np.array_equal(out.cpu().numpy().reshape(ref.shape), ref), p


# ---------------------------------------------------------------------------------------------
# fused three-axis gradient
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("npts", [(64, 16, 14), (70, 33, 21), (130, 18, 40), (14, 14, 14)])
@pytest.mark.parametrize("accuracy", [2, 6])
def test_fused_gradient(npts, accuracy):
    import pytristan_b200 as pt

    torch = _torch()
    axes = [np.linspace(0.0, 1.0 + k, n) for k, n in enumerate(npts)]
    ops = []
    for k in range(3):
        F = pt.FinDiffMat(axes[k], order=1, accuracy=accuracy)
        F.npts, F.axis = tuple(npts), k
        ops.append(F)
    rng = np.random.default_rng(1234)
    U = rng.standard_normal((2, int(np.prod(npts))))
    got = pt.gradient(ops, U)
    for k in range(3):
        band = oracle.findiff_band(axes[k], 1, accuracy)
        ref = oracle.stencil_apply(band, U, npts, k)
        assert np.max(np.abs(got[k] - ref)) <= 1e-12 * np.max(np.abs(ref)), k
        assert np.array_equal(got[k], ops[k].apply(U)) or np.max(np.abs(got[k] - ops[k].apply(U))) <= 1e-13 * np.max(np.abs(ref))
    # non-fusable combination (mapped axis) falls back to three passes, same answers
    ops[1] = pt.FinDiffMat(oracle.cheb_nodes(npts[1], 0.0, 1.0), order=1, accuracy=accuracy)
    ops[1].npts, ops[1].axis = tuple(npts), 1
    got = pt.gradient(ops, torch.from_numpy(U).cuda())
    band = oracle.findiff_band(oracle.cheb_nodes(npts[1], 0.0, 1.0), 1, accuracy)
    ref = oracle.stencil_apply(band, U, npts, 1)
    assert np.max(np.abs(got[1].cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("tile", [16, 32, 108, 116])
def test_fused_gradient_tile_shapes_agree_bitwise(tile):
    """Every tile shape of pt_stencil_grad3 (64x16, 64x32, 128x8, 128x16 points) performs the same
    per-point arithmetic: results must be bit-identical to the default shape's."""
    import pytristan_b200 as pt
    from pytristan_b200 import _lib

    torch = _torch()
    npts = (150, 37, 29)
    axes = [np.linspace(0.0, 1.0 + k, n) for k, n in enumerate(npts)]
    ops = []
    for k in range(3):
        F = pt.FinDiffMat(axes[k], order=1, accuracy=6)
        F.npts, F.axis = npts, k
        ops.append(F)
    U = torch.from_numpy(np.random.default_rng(7).standard_normal(int(np.prod(npts)))).cuda()
    lib = _lib.load()
    ref = [g.clone() for g in pt.gradient(ops, U)]
    try:
        lib.pt_debug_set_grad_ty(tile)
        got = pt.gradient(ops, U)
        for k in range(3):
            assert torch.equal(got[k], ref[k]), (tile, k)
    finally:
        lib.pt_debug_set_grad_ty(108)         # default: 128 x 8


# ---------------------------------------------------------------------------------------------
# collocation on non-affinely mapped nodes; Fourier orders > 2
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [9, 64, 257, 1024])
@pytest.mark.parametrize("order", [1, 2])
def test_barymat_entries(n, order):
    import pytristan_b200 as pt

    c = oracle.cheb_nodes(n)[::-1]
    x = c + 0.15 * (1.0 - c * c)             # mildly stretched CGL nodes: monotone, not an affine image
    assert not oracle.is_affine_cgl(x)
    D = pt.ChebMat(x, order=order)
    assert D.weights_kind == "bary"
    ref = oracle.bary_mat(x, order)
    scale = np.abs(ref).sum(axis=1, keepdims=True)
    assert np.max(np.abs(np.asarray(D) - ref) / scale) <= 1e-14
    # on true CGL nodes the barycentric matrix agrees with the c-pattern one
    xc = oracle.cheb_nodes(n, 0.0, 2.0)
    Db = pt.ChebMat(xc, order=order, weights="bary")
    Dc = pt.ChebMat(xc, order=order)
    assert Dc.weights_kind == "cgl"
    sc = np.abs(np.asarray(Dc)).sum(axis=1, keepdims=True)
    assert np.max(np.abs(np.asarray(Db) - np.asarray(Dc)) / sc) <= 2e-11     # two formulas, O(n eps) apart
    # exact on polynomials of moderate degree (interpolation on stretched nodes is only as stable
    # as the node distribution allows, so no claim is made for general smooth functions)
    if order == 1:
        p = np.polynomial.Polynomial([1.0, -2.0, 0.5, 3.0, -1.0])
        assert np.max(np.abs(D @ p(x) - p.deriv(1)(x))) <= 1e-9 * np.abs(ref).sum(axis=1).max()


@pytest.mark.parametrize("n", [8, 17, 64, 255])
@pytest.mark.parametrize("order", [3, 4, 5])
def test_fourmat_high_order(n, order):
    import pytristan_b200 as pt

    period = 2 * np.pi
    x = -np.pi + period / n * np.arange(n)
    D = pt.FourMat(x, order=order, period=period)
    ref = oracle.four_mat(n, period, order)
    assert np.max(np.abs(np.asarray(D) - ref)) <= 1e-12 * np.max(np.abs(ref))
    if n >= 64:
        f = np.exp(np.sin(x))
        k = np.fft.fftfreq(n, 1.0 / n)
        if n % 2 == 0 and order % 2 == 1:
            k[n // 2] = 0.0
        spec = np.real(np.fft.ifft((1j * k) ** order * np.fft.fft(f)))
        assert np.max(np.abs(D @ f - spec)) <= 1e-12 * np.max(np.abs(ref) @ np.abs(f))     # relative to |D||f|


# ---------------------------------------------------------------------------------------------
# host-buffer pipeline (the e2e path of bench.py)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nfields,chunk", [(1, None), (7, 2), (16, None), (5, 8)])
def test_apply_host_pipeline(nfields, chunk):
    import pytristan_b200 as pt

    torch = _torch()
    g = pt.Grid.from_bounds((0.0, 2 * np.pi - 2 * np.pi / 48, 48), (-1.0, 1.0, 33), axes=[1], mappers=[pt.cheb])
    ops = [pt.FourMat(g, axis=0, order=1), pt.ChebMat(g, axis=1, order=2), pt.FinDiffMat(g, axis=0, order=1, accuracy=4, periodic=True)]
    rng = np.random.default_rng(1234)
    U = rng.standard_normal((nfields, 48 * 33))
    ref = [op.apply(U) for op in ops]
    got = pt.apply_host(ops, U, chunk_fields=chunk)                      # ndarray in, ndarrays out
    assert all(isinstance(a, np.ndarray) and np.array_equal(a, b) for a, b in zip(got, ref))
    hin = torch.from_numpy(U).pin_memory()
    houts = [torch.empty_like(hin).pin_memory() for _ in ops]
    res = pt.apply_host(ops, hin, houts, chunk_fields=chunk)             # pinned tensors in and out
    torch.cuda.synchronize()
    assert all(r.data_ptr() == o.data_ptr() for r, o in zip(res, houts))
    assert all(np.array_equal(o.numpy(), b) for o, b in zip(houts, ref))
    assert pt.launches_per_apply(ops) == 3
    with pytest.raises(ValueError):
        pt.apply_host(ops, U.ravel()[:-1])


# ---------------------------------------------------------------------------------------------
# alternative staging paths of the dense kernel (bulk-copy engine / mbarrier pipeline)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [2])
def test_dense_bulk_copy_variants_bitwise(mode):
    """mode 2: the plain (non-symmetric operator) kernel with bulk-copied operator panels.  Same tiles and
    summation order as the default cp.async kernel, so results must be bit-identical.  (The bulk-copy
    engine is on the DEFAULT path of the even-odd operators: csrc/dense_apply_eo_ws.cu, tests/test_dense_eo_ws.py.)"""
    import ctypes

    from pytristan_b200 import _lib
    from pytristan_b200.operators import _DiffMat

    torch = _torch()
    lib = _lib.load()
    lib.pt_debug_set_dense_tma.argtypes = [ctypes.c_int]
    try:
        for npts, axis, batch in [((130, 64), 0, 3), ((64, 136), 1, 3), ((258, 40), 0, 1), ((40, 258), 1, 2), ((1024, 66), 0, 1), ((64, 2048), 1, 1)]:
            n = npts[axis]
            gen = torch.Generator(device="cuda").manual_seed(n)
            op = _DiffMat._wrap(torch.randn(n, n, dtype=torch.float64, device="cuda", generator=gen), npts, axis, 1, None)
            U = torch.randn(batch, int(np.prod(npts)), dtype=torch.float64, device="cuda", generator=gen)
            lib.pt_debug_set_dense_tma(0)
            y0 = op.apply(U)
            lib.pt_debug_set_dense_tma(mode)
            y1 = op.apply(U)
            assert torch.equal(y0, y1), (mode, npts, axis)
    finally:
        lib.pt_debug_set_dense_tma(0)
