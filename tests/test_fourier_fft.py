"""FourMat applied through the FFT (pt_fourier_apply) against the NumPy oracle's dense matrix and
against the dense tensor-core path: same linear map, componentwise within 1e-12 of |D||u|."""
import numpy as np
import pytest
from conftest import clear_grids

from oracle import operators as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-12    # componentwise relative to (|D||u|)_i, SURVEY.md 8d


def _check(n, order, npts, axis, nfields, period=2 * np.pi, alpha=1.0, beta=0.0, seed=0):
    import torch

    import pytristan_b200 as pt

    rng = np.random.default_rng(seed + n + 7 * order)
    x = np.arange(n) * (period / n)
    F = pt.FourMat(x, order=order, method="fft")
    F.npts, F.axis = tuple(npts), axis
    U = rng.standard_normal((nfields, int(np.prod(npts))))
    D = oracle.four_mat(n, period, order)
    ref = oracle.apply_along_axis(D, U, npts, axis)
    bound = oracle.abs_bound(D, U, npts, axis)
    d_u = torch.from_numpy(U).cuda()
    if beta != 0.0:
        Y0 = rng.standard_normal(U.shape)
        out = torch.from_numpy(Y0).cuda()
        got = F.apply(d_u, out=out, alpha=alpha, beta=beta).cpu().numpy()
        ref = alpha * ref + beta * Y0
        bound = abs(alpha) * bound + abs(beta) * np.abs(Y0)
    else:
        got = F.apply(d_u, alpha=alpha).cpu().numpy()
        ref = alpha * ref
        bound = abs(alpha) * bound
    err = np.max(np.abs(got - ref) / bound)
    assert err <= TOL, (n, order, npts, axis, err)
    # and against the dense DMMA path of the same operator
    F.method = "dense"
    dense = F.apply(d_u).cpu().numpy() * alpha
    if beta == 0.0:
        assert np.max(np.abs(got - dense) / bound) <= TOL


@pytest.mark.parametrize("n", [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("order", [1, 2])
def test_contiguous_axis(n, order):
    _check(n, order, (n, 5), 0, 3)          # 15 rows: odd count, the last row is paired with zeros


@pytest.mark.parametrize("order", [3, 4, 5, 8])
def test_higher_orders(order):
    _check(64, order, (64, 4), 0, 2, period=3.0)
    _check(32, order, (6, 32, 3), 1, 2, period=3.0)


@pytest.mark.parametrize("n,before,after", [(8, 2, 3), (16, 6, 2), (64, 10, 3), (256, 64, 2), (1024, 24, 1), (2048, 6, 1),
                                            (4096, 2, 1)])
def test_strided_axis(n, before, after):
    _check(n, 1, (before, n, after), 1, 2)
    _check(n, 2, (before, n, after), 1, 1, alpha=-0.5, beta=2.0)


def test_fused_scaling_and_accumulation_contiguous():
    _check(128, 1, (128, 9), 0, 2, alpha=1.5, beta=-0.25)


def test_unsupported_shapes_take_the_dense_path():
    import torch

    import pytristan_b200 as pt

    for n, npts, axis in [(24, (24, 4), 0), (16, (5, 16), 1)]:     # not a power of two; odd stride
        x = np.arange(n) * (2 * np.pi / n)
        F = pt.FourMat(x, order=1, method="fft")
        F.npts, F.axis = npts, axis
        U = np.random.default_rng(3).standard_normal((2, int(np.prod(npts))))
        D = oracle.four_mat(n, 2 * np.pi, 1)
        ref = oracle.apply_along_axis(D, U, npts, axis)
        got = F.apply(torch.from_numpy(U).cuda()).cpu().numpy()
        assert np.max(np.abs(got - ref) / oracle.abs_bound(D, U, npts, axis)) <= TOL


def test_known_answer_config2_axis():
    """d/dx of exp(sin x) on the 1024-point periodic axis of BASELINE config 2."""
    import pytristan_b200 as pt

    n = 1024
    x = np.arange(n) * (2 * np.pi / n)
    F = pt.FourMat(x, order=1, method="fft")
    u = np.exp(np.sin(x))
    assert np.max(np.abs(F.apply(u) - np.cos(x) * u)) < 1e-11
    F2 = pt.FourMat(x, order=2, method="fft")
    assert np.max(np.abs(F2.apply(u) - (np.cos(x) ** 2 - np.sin(x)) * u)) < 1e-9


def test_polar_laplacian_with_fft_azimuthal_operator():
    """Config 3 at reduced size: lap u = (D2_r + D_r/r) u + (1/r^2) D2_theta u with D2_theta through
    the FFT and the 1/r^2 scaling fused in its store phase (PT_SCALE_AFTER)."""
    import pytristan_b200 as pt

    clear_grids()
    g = pt.get_polar_grid(64, 24, axes=[1], mappers=[pt.cheb], fornberg=True)
    lap = pt.PolarLaplacian(g, azimuthal="fft")
    dense = pt.PolarLaplacian(g)
    th, r = g.axpoints(0), g.axpoints(1)
    T, R = np.asarray(g)[0], np.asarray(g)[1]
    U = np.random.default_rng(1234).standard_normal((3, T.size))
    y = lap.apply(U)
    yref = oracle.polar_laplacian(th, r, U)
    assert np.linalg.norm(y - yref) / np.linalg.norm(yref) <= 1e-12
    assert np.linalg.norm(y - dense.apply(U)) / np.linalg.norm(yref) <= 1e-12
    assert np.max(np.abs(lap.apply(R ** 3 * np.cos(3 * T)))) < 1e-8
    assert np.max(np.abs(lap.apply(R ** 2) - 4.0)) < 1e-8
    clear_grids()


def test_fft_after_scale_strided_axis():
    import torch

    import pytristan_b200 as pt

    n, before, after = 64, 6, 5
    x = np.arange(n) * (2 * np.pi / n)
    F = pt.FourMat(x, order=2, method="fft")
    F.npts, F.axis = (before, n, after), 1
    rng = np.random.default_rng(5)
    U = rng.standard_normal((1, before * n * after))
    sc = rng.uniform(0.5, 2.0, after)
    D = oracle.four_mat(n, 2 * np.pi, 2)
    ref = oracle.apply_along_axis(D, U, F.npts, 1).reshape(after, n, before) * sc[:, None, None]
    bound = oracle.abs_bound(D, U, F.npts, 1).reshape(after, n, before) * sc[:, None, None]
    got = F.apply(torch.from_numpy(U).cuda(), scale=sc, scale_mode="after").cpu().numpy().reshape(after, n, before)
    assert np.max(np.abs(got - ref) / bound) <= TOL


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("rows", [1, 2, 7, 300])
def test_n1024_contiguous_warp_kernel(order, rows):
    """n = 1024 on the contiguous axis runs the warp-per-sequence kernel: orders with a compile-time
    symbol (1, 2) and the general one, odd row counts (the last row is paired with zeros), one row."""
    if order > 2 and rows not in (7,):
        pytest.skip("general-order symbol: one shape is enough")
    _check(1024, order, (1024, rows), 0, 1, period=2.5)


def test_n1024_contiguous_fused_scaling_accumulation_and_row_scale():
    import torch

    import pytristan_b200 as pt

    _check(1024, 1, (1024, 9), 0, 2, alpha=1.5, beta=-0.25)
    _check(1024, 2, (1024, 4), 0, 1, alpha=-2.0, beta=0.5)
    n, rows = 1024, 13
    x = np.arange(n) * (2 * np.pi / n)
    F = pt.FourMat(x, order=2, method="fft")
    F.npts, F.axis = (n, rows), 0
    rng = np.random.default_rng(11)
    U = rng.standard_normal((2, n * rows))
    sc = rng.uniform(0.5, 2.0, rows)
    D = oracle.four_mat(n, 2 * np.pi, 2)
    ref = oracle.apply_along_axis(D, U, F.npts, 0).reshape(2, rows, n) * sc[None, :, None]
    bound = oracle.abs_bound(D, U, F.npts, 0).reshape(2, rows, n) * sc[None, :, None]
    got = F.apply(torch.from_numpy(U).cuda(), scale=sc, scale_mode="after").cpu().numpy().reshape(2, rows, n)
    assert np.max(np.abs(got - ref) / bound) <= TOL


@pytest.mark.parametrize("cfg", [0, 1, 3, 4, 5, 8, 11])
def test_every_tuning_config_is_the_same_map(cfg):
    """pt_debug_set_fft_config picks the kernel shape / twiddle source; every id must give the operator."""
    from pytristan_b200 import _lib

    lib = _lib.load()
    try:
        lib.pt_debug_set_fft_config(cfg)
        _check(1024, 1, (1024, 11), 0, 1)
        _check(1024, 2, (6, 1024, 2), 1, 1)
        _check(256, 1, (256, 9), 0, 2)
        _check(2048, 1, (2048, 3), 0, 1)
        _check(64, 3, (10, 64, 3), 1, 1)
    finally:
        lib.pt_debug_set_fft_config(-1)
