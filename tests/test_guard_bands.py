"""Out-of-bounds checks of our own (compute-sanitizer is closed on this pool: profiles/r02_sanitizer_unavailable.txt).

Every apply kernel runs on fields embedded in a larger allocation:
* the words around the OUTPUT hold a canary that must be untouched afterwards (out-of-bounds WRITES),
* the words around the INPUT hold NaN, so that a value read from outside the field (out-of-bounds READS
  that reach a result) poisons the output, which is compared with the oracle.
Shapes are chosen to hit the edge paths: ragged column tiles, odd n, odd strides, partial k chunks,
the small-tile / warp-specialised / plain dense kernels, all three stencil kernels, the FFT path.
"""
import numpy as np
import pytest

from oracle import operators as oracle

pytestmark = pytest.mark.gpu
GUARD = 1024          # doubles on each side
CANARY = -7.25e77


def _embedded(values):
    import torch

    n = values.size
    buf = torch.full((n + 2 * GUARD + 2,), float("nan"), dtype=torch.float64, device="cuda")
    # keep 16-byte alignment of the payload (GUARD is even)
    view = buf[GUARD:GUARD + n]
    view.copy_(torch.from_numpy(np.ascontiguousarray(values).ravel()).cuda())
    return buf, view


def _guarded_out(n):
    import torch

    buf = torch.full((n + 2 * GUARD + 2,), CANARY, dtype=torch.float64, device="cuda")
    return buf, buf[GUARD:GUARD + n]


def _check(buf, view, ref, bound, what, tol=1e-12):
    got = view.cpu().numpy()
    assert not np.isnan(got).any(), (what, "NaN in the output: something outside the field was read")
    assert np.max(np.abs(got - ref) / bound) <= tol, what
    lo, hi = buf[:GUARD].cpu().numpy(), buf[GUARD + view.numel():].cpu().numpy()
    assert np.all(lo == CANARY) and np.all(hi == CANARY), (what, "canary overwritten: out-of-bounds write")


DENSE_SHAPES = [  # (n, before, after): small-tile, ragged, odd n / odd stride (8-byte paths), ws-eligible ones
    (17, 1, 5), (64, 1, 33), (130, 1, 67), (257, 1, 70), (64, 3, 5), (66, 10, 7), (129, 64, 3), (131, 7, 9),
    (256, 1, 9473), (260, 1, 4801), (257, 150, 64), (131, 66, 149), (272, 1000, 5), (512, 2400, 2),
]


@pytest.mark.parametrize("n,before,after", DENSE_SHAPES)
@pytest.mark.parametrize("kind", ["symmetric", "plain"])
def test_dense_kernels_stay_inside_their_fields(n, before, after, kind):
    from pytristan_b200.operators import _DiffMat
    import torch

    rng = np.random.default_rng(n * 31 + before)
    D = rng.standard_normal((n, n))
    if kind == "symmetric":
        D = oracle.flip_symmetrize(D, 1)
    op = _DiffMat._wrap(torch.from_numpy(D).cuda(), (before, n, after), 1, 2, None)
    assert (op._eo_forms() is not None) == (kind == "symmetric" and n >= 16)
    U = rng.standard_normal(before * n * after)
    _, u = _embedded(U)
    ybuf, y = _guarded_out(U.size)
    op.apply(u, out=y)
    ref = oracle.apply_along_axis(D, U[None], (before, n, after), 1)[0]
    bound = oracle.abs_bound(D, U[None], (before, n, after), 1)[0]
    _check(ybuf, y, ref, bound, (kind, n, before, after))


@pytest.mark.parametrize("npts,axis,periodic,mapped", [
    ((70,), 0, False, False), ((33, 5), 0, False, False), ((6, 41), 1, False, False), ((5, 7, 23), 2, False, False),
    ((64,), 0, True, False), ((10, 32), 1, True, False), ((9, 30), 1, False, True), ((31, 3), 0, False, True),
    ((258, 130), 0, False, False), ((130, 258), 1, False, False),
])
def test_stencil_kernels_stay_inside_their_fields(npts, axis, periodic, mapped):
    import pytristan_b200 as pt

    n = npts[axis]
    if mapped:
        x = oracle.cheb_nodes(n, 0.0, 1.0)
    elif periodic:
        x = np.linspace(0.0, 1.0 - 1.0 / n, n)
    else:
        x = np.linspace(0.0, 1.0, n)
    F = pt.FinDiffMat(x, order=1, accuracy=6, periodic=periodic)
    F.npts, F.axis = tuple(npts), axis
    rng = np.random.default_rng(n)
    U = rng.standard_normal((3, int(np.prod(npts))))
    _, u = _embedded(U)
    ybuf, y = _guarded_out(U.size)
    F.apply(u, out=y)
    band = oracle.findiff_band(x, 1, 6, periodic)
    ref = oracle.stencil_apply(band, U, npts, axis).ravel()
    bound = oracle.stencil_apply(dict(band, W=np.abs(band["W"])), np.abs(U), npts, axis).ravel()
    _check(ybuf, y, ref, bound, (npts, axis, periodic, mapped))


@pytest.mark.parametrize("n,before,after", [(8, 1, 3), (64, 1, 5), (1024, 1, 7), (1024, 1, 1), (1024, 1, 1186), (16, 6, 2),
                                            (256, 64, 2), (512, 2, 3), (1024, 4, 3), (2048, 1, 3), (2048, 2, 2)])
def test_fft_kernel_stays_inside_its_fields(n, before, after):
    import pytristan_b200 as pt

    x = np.arange(n) * (2 * np.pi / n)
    F = pt.FourMat(x, order=1, method="fft")
    F.npts, F.axis = (before, n, after), 1
    rng = np.random.default_rng(n + before)
    U = rng.standard_normal(before * n * after)
    _, u = _embedded(U)
    ybuf, y = _guarded_out(U.size)
    F.apply(u, out=y)
    D = oracle.four_mat(n, 2 * np.pi, 1)
    ref = oracle.apply_along_axis(D, U[None], (before, n, after), 1)[0]
    bound = oracle.abs_bound(D, U[None], (before, n, after), 1)[0]
    _check(ybuf, y, ref, bound, ("fft", n, before, after))


def test_fused_gradient_stays_inside_its_fields():
    import pytristan_b200 as pt

    npts = (37, 21, 19)
    axes = [np.linspace(0.0, 1.0, m) for m in npts]
    ops = []
    for k in range(3):
        F = pt.FinDiffMat(axes[k], order=1, accuracy=6)
        F.npts, F.axis = npts, k
        ops.append(F)
    U = np.random.default_rng(5).standard_normal(int(np.prod(npts)))
    _, u = _embedded(U)
    outs = [_guarded_out(U.size) for _ in range(3)]
    pt.gradient(ops, u, outs=[o[1] for o in outs])
    for k in range(3):
        band = oracle.findiff_band(axes[k], 1, 6)
        ref = oracle.stencil_apply(band, U[None], npts, k)[0]
        bound = oracle.stencil_apply(dict(band, W=np.abs(band["W"])), np.abs(U)[None], npts, k)[0]
        _check(outs[k][0], outs[k][1], ref, bound, ("grad3", k))
