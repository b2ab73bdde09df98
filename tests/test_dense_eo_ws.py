"""Warp-specialised even-odd kernel (csrc/dense_apply_eo_ws.cu): bit-identical to the cp.async
even-odd kernel for even n, within 1e-12 of |D||u| of the NumPy oracle for odd n (where the middle
row and column are rank-1 terms beside an exactly (n-1)/2-sized half problem)."""
import ctypes

import numpy as np
import pytest

from oracle import operators as oracle

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _sym(rng, n, sign):
    return oracle.flip_symmetrize(rng.standard_normal((n, n)), sign)


def _forms(D, sign):
    """(device matrix, packed even, packed odd, packed ws) through the C-ABI."""
    import torch

    from pytristan_b200 import _lib

    lib = _lib.load()
    n = D.shape[0]
    d = torch.from_numpy(np.ascontiguousarray(D)).cuda()
    size = lib.pt_packed_eo_size(n)
    pe, po = _lib.empty((size,)), _lib.empty((size,))
    _lib.check(lib.pt_operator_pack_eo(_lib.dptr(d), n, _lib.dptr(pe), _lib.dptr(po), _lib.stream_ptr()), "pack_eo")
    pw = _lib.empty((lib.pt_packed_eo_ws_size(n),))
    _lib.check(lib.pt_operator_pack_eo_ws(_lib.dptr(d), n, _lib.dptr(pw), _lib.stream_ptr()), "pack_eo_ws")
    return d, pe, po, pw


def _apply(kind, forms, n, sign, u, y, after, before, alpha=1.0, beta=0.0, scale=None, mode=0):
    from pytristan_b200 import _lib

    lib = _lib.load()
    _, pe, po, pw = forms
    slen = 0 if scale is None else scale.numel()
    if kind == "ws":
        rc = lib.pt_dense_apply_eo_ws(_lib.dptr(pw), n, sign, _lib.dptr(u), _lib.dptr(y), after, before, alpha, beta,
                                      _lib.dptr(scale), mode, slen, _lib.stream_ptr())
    else:
        rc = lib.pt_dense_apply_eo(_lib.dptr(pe), _lib.dptr(po), n, sign, _lib.dptr(u), _lib.dptr(y), after, before,
                                   alpha, beta, _lib.dptr(scale), mode, slen, _lib.stream_ptr())
    return rc


# (n, after, before): enough column tiles for one tile per SM (the kernel declines smaller fields)
EVEN_CASES = [(128, 9600, 1), (256, 9473, 1), (1024, 2500, 1), (260, 4800, 1), (272, 4801, 1),
              (128, 75, 128), (256, 3, 3200), (512, 2, 2400), (260, 37, 130), (272, 5, 1000), (130, 160, 62)]


@pytest.mark.parametrize("n,after,before", EVEN_CASES)
@pytest.mark.parametrize("sign", [1, -1])
def test_ws_is_bit_identical_to_the_cp_async_kernel(n, after, before, sign):
    import torch

    rng = np.random.default_rng(n + after + before + sign)
    D = _sym(rng, n, sign)
    forms = _forms(D, sign)
    gen = torch.Generator(device="cuda").manual_seed(n * 7 + before)
    u = torch.randn(after * n * before, dtype=torch.float64, device="cuda", generator=gen)
    y0 = torch.randn(after * n * before, dtype=torch.float64, device="cuda", generator=gen)
    for (alpha, beta) in [(1.0, 0.0), (-0.75, 0.5)]:
        ya, yb = y0.clone(), y0.clone()
        assert _apply("ws", forms, n, sign, u, ya, after, before, alpha, beta) == 0, "shape declined by the ws kernel"
        assert _apply("eo", forms, n, sign, u, yb, after, before, alpha, beta) == 0
        assert torch.equal(ya, yb), (n, after, before, sign, alpha)
    # fused scalings (row / after / before)
    for mode, length in ((1, n), (2, 7), (3, before)):
        sc = torch.rand(length, dtype=torch.float64, device="cuda", generator=gen) + 0.5
        ya, yb = y0.clone(), y0.clone()
        assert _apply("ws", forms, n, sign, u, ya, after, before, 1.25, 0.0, sc, mode) == 0
        assert _apply("eo", forms, n, sign, u, yb, after, before, 1.25, 0.0, sc, mode) == 0
        assert torch.equal(ya, yb), (n, after, before, sign, mode)
    # and against the oracle on a slice of the field (first two slabs)
    a_chk = min(after, 2)
    ya = torch.empty_like(u)
    assert _apply("ws", forms, n, sign, u, ya, after, before) == 0
    U = u[:a_chk * n * before].cpu().numpy()
    ref = oracle.apply_along_axis(D, U[None], (before, n, a_chk), 1)[0]
    bound = oracle.abs_bound(D, U[None], (before, n, a_chk), 1)[0]
    assert np.max(np.abs(ya[:a_chk * n * before].cpu().numpy() - ref) / bound) <= TOL


# odd n on a strided axis: middle row and column split off
ODD_CASES = [(129, 150, 64), (257, 10, 1024), (257, 64, 150), (385, 3, 1600), (131, 149, 66), (273, 4, 1200)]


@pytest.mark.parametrize("n,after,before", ODD_CASES)
@pytest.mark.parametrize("sign", [1, -1])
def test_ws_odd_n_matches_the_oracle(n, after, before, sign):
    import torch

    rng = np.random.default_rng(n + after + before + 2 * sign)
    D = _sym(rng, n, sign)
    forms = _forms(D, sign)
    U = rng.standard_normal(after * n * before)
    Y0 = rng.standard_normal(U.size)
    u = torch.from_numpy(U).cuda()
    npts = (before, n, after)
    for (alpha, beta) in [(1.0, 0.0), (0.5, -1.5)]:
        y = torch.from_numpy(Y0.copy()).cuda()
        assert _apply("ws", forms, n, sign, u, y, after, before, alpha, beta) == 0, "shape declined by the ws kernel"
        ref = alpha * oracle.apply_along_axis(D, U[None], npts, 1)[0] + beta * Y0
        bound = abs(alpha) * oracle.abs_bound(D, U[None], npts, 1)[0] + abs(beta) * np.abs(Y0)
        err = np.abs(y.cpu().numpy() - ref) / bound
        assert np.max(err) <= TOL, (n, after, before, sign, alpha, int(np.argmax(err)))
        # the cp.async kernel computes the same thing in another order: both within tolerance of each other
        y2 = torch.from_numpy(Y0.copy()).cuda()
        assert _apply("eo", forms, n, sign, u, y2, after, before, alpha, beta) == 0
        assert np.max(np.abs((y - y2).cpu().numpy()) / bound) <= TOL
    # scalings on the middle row too
    for mode, length in ((1, n), (2, 5), (3, before)):
        sc = rng.random(length) + 0.5
        y = torch.empty_like(u)
        assert _apply("ws", forms, n, sign, u, y, after, before, 2.0, 0.0, torch.from_numpy(sc).cuda(), mode) == 0
        V = oracle.apply_along_axis(D, U[None], npts, 1)[0].reshape(after, n, before)
        B = oracle.abs_bound(D, U[None], npts, 1)[0].reshape(after, n, before)
        if mode == 1:
            f = sc[None, :, None]
        elif mode == 2:
            f = sc[np.arange(after) % length][:, None, None]
        else:
            f = sc[None, None, :]
        ref = 2.0 * f * V
        assert np.max(np.abs(y.cpu().numpy().reshape(after, n, before) - ref) / (2.0 * f * B)) <= TOL, (n, mode)


def test_ws_declines_what_it_does_not_cover():
    import torch

    rng = np.random.default_rng(0)
    n = 256
    forms = _forms(_sym(rng, n, 1), 1)
    u = torch.zeros(200 * n * 64 + 1, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(u)
    assert _apply("ws", forms, n, 1, u[:4 * n * 64], y[:4 * n * 64], 4, 64) == -3            # fewer tiles than SMs
    assert _apply("ws", forms, n, 1, u[1:1 + 200 * n * 64], y[:200 * n * 64], 200, 64) == -3  # unaligned field
    assert _apply("ws", forms, n, 1, u[:200 * n * 63], y[:200 * n * 63], 200, 63) == -3      # odd stride
    f2 = _forms(_sym(rng, 257, 1), 1)
    assert _apply("ws", f2, 257, 1, u[:9600 * 257], y[:9600 * 257], 9600, 1) == -3           # odd n, contiguous axis
    f3 = _forms(_sym(rng, 64, 1), 1)
    assert _apply("ws", f3, 64, 1, u[:800 * 64 * 64], y[:800 * 64 * 64], 800, 64) == -3      # n < 128


def test_operators_take_the_ws_kernel_by_default_and_config2_is_covered():
    """BASELINE config 2 through the public classes: d/dx (n = 1024, contiguous axis) and d2/dy2
    (n = 257, strided axis) run the warp-specialised kernel and agree with the cp.async one."""
    import torch

    import pytristan_b200 as pt
    import pytristan_b200.operators as ops_mod
    from pytristan_b200 import _lib

    x = np.linspace(0.0, 2 * np.pi - 2 * np.pi / 1024, 1024)
    yv = pt.cheb(257, -1.0, 1.0)
    Dx, Dy = pt.FourMat(x, order=1), pt.ChebMat(yv, order=2)
    for op, axis in ((Dx, 0), (Dy, 1)):
        op.npts, op.axis = (1024, 257), axis
        assert op._eo_forms()[3] is not None
    gen = torch.Generator(device="cuda").manual_seed(1)
    u = torch.randn(16, 1024 * 257, dtype=torch.float64, device="cuda", generator=gen)
    lib = _lib.load()
    for op, axis in ((Dx, 0), (Dy, 1)):
        y_ws = op.apply(u)
        try:
            ops_mod.EVEN_ODD_WS = False
            y_eo = op.apply(u)
        finally:
            ops_mod.EVEN_ODD_WS = True
        Dn = np.asarray(op)
        U = u[:2].cpu().numpy()
        ref = oracle.apply_along_axis(Dn, U, (1024, 257), axis)
        bound = oracle.abs_bound(Dn, U, (1024, 257), axis)
        assert np.max(np.abs(y_ws[:2].cpu().numpy() - ref) / bound) <= TOL
        if axis == 0:
            assert torch.equal(y_ws, y_eo)                      # even n: same bits
        else:
            assert float((y_ws - y_eo).abs().max()) <= 1e-12 * float(bound.max())
        # the ws kernel really ran (it would return -3 otherwise)
        after, nn, before = op._split(u.numel())
        eo = op._eo_forms()
        out = torch.empty_like(u)
        rc = lib.pt_dense_apply_eo_ws(_lib.dptr(eo[3]), nn, eo[0], _lib.dptr(u), _lib.dptr(out), after, before, 1.0, 0.0,
                                      None, 0, 0, _lib.stream_ptr())
        assert rc == 0 and torch.equal(out.view_as(y_ws), y_ws)


def test_ws_scatter_epilogue_matches_the_cp_async_scatter():
    """Single-destination scatter maps (the pencil redistribution with one rank): same bits from both kernels."""
    import torch

    from pytristan_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    for (n, after, before, sign) in [(256, 9500, 1, 1), (256, 4, 2400, -1), (257, 4, 2400, 1)]:
        D = _sym(rng, n, sign)
        forms = _forms(D, sign)
        _, pe, po, pw = forms
        gen = torch.Generator(device="cuda").manual_seed(n + before)
        u = torch.randn(after * n * before, dtype=torch.float64, device="cuda", generator=gen)
        c = torch.randn(after * n * before, dtype=torch.float64, device="cuda", generator=gen)
        # destination = the same layout shifted by an even offset inside a larger buffer
        outs = []
        for kind in ("ws", "eo"):
            dest = torch.zeros(after * n * before + 64, dtype=torch.float64, device="cuda")
            table = torch.tensor([dest.data_ptr()], dtype=torch.int64, device="cuda")
            m = _lib.ScatterMap()
            m.nranks, m.chunk = 1, n
            m.a_inner, m.s_outer, m.s_inner, m.s_i, m.offset = 1, n * before, 0, before, 32
            m.base = table.data_ptr()
            if kind == "ws":
                rc = lib.pt_dense_apply_eo_ws_scatter(_lib.dptr(pw), n, sign, _lib.dptr(u), _lib.dptr(c), after, before,
                                                      1.5, 0.25, ctypes.byref(m), _lib.stream_ptr())
            else:
                rc = lib.pt_dense_apply_eo_scatter(_lib.dptr(pe), _lib.dptr(po), n, sign, _lib.dptr(u), _lib.dptr(c),
                                                   after, before, 1.5, 0.25, ctypes.byref(m), _lib.stream_ptr())
            assert rc == 0, (kind, n, after, before)
            outs.append(dest)
        if n % 2 == 0:
            assert torch.equal(outs[0], outs[1])
        else:
            U = u.cpu().numpy()
            bound = 1.5 * oracle.abs_bound(D, U[None], (before, n, after), 1)[0] + 0.25 * np.abs(c.cpu().numpy())
            assert np.max(np.abs((outs[0] - outs[1])[32:32 + U.size].cpu().numpy()) / bound) <= TOL
        assert float(outs[0][:32].abs().max()) == 0.0 and float(outs[0][32 + u.numel():].abs().max()) == 0.0


def test_ws_fused_field_copy_redistributes_the_field_bit_for_bit():
    """pt_dense_apply_eo_ws_scatter_copy: besides the result (scatter epilogue) the producers store the field
    itself through a second map.  Single-destination maps here; the pencil maps over several ranks are
    covered by tests/test_peer.py (emulated ranks) and, with >= 2 GPUs, by the real fabric."""
    import torch

    from pytristan_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(11)
    for (n, after, before, sign) in [(256, 9500, 1, 1), (260, 4801, 1, -1), (256, 4, 2400, -1), (257, 4, 2400, 1),
                                     (130, 160, 62, 1), (385, 3, 1600, -1)]:
        D = _sym(rng, n, sign)
        _, pe, po, pw = _forms(D, sign)
        gen = torch.Generator(device="cuda").manual_seed(n + before)
        u = torch.randn(after * n * before, dtype=torch.float64, device="cuda", generator=gen)
        total = u.numel()

        def maps(dest_y, dest_u):
            out = []
            for dest, off in ((dest_y, 32), (dest_u, 6)):
                table = torch.tensor([dest.data_ptr()], dtype=torch.int64, device="cuda")
                m = _lib.ScatterMap()
                m.nranks, m.chunk = 1, n + (n & 1)            # an even chunk that covers n
                m.a_inner, m.s_outer, m.s_inner, m.s_i, m.offset = 1, n * before, 0, before, off
                m.base = table.data_ptr()
                m._keep = table
                out.append(m)
            return out

        y1 = torch.zeros(total + 64, dtype=torch.float64, device="cuda")
        c1 = torch.full((total + 64,), float("nan"), dtype=torch.float64, device="cuda")
        my, mu = maps(y1, c1)
        rc = lib.pt_dense_apply_eo_ws_scatter_copy(_lib.dptr(pw), n, sign, _lib.dptr(u), None, after, before, 1.0, 0.0,
                                                   ctypes.byref(my), ctypes.byref(mu), _lib.stream_ptr())
        assert rc == 0, (n, after, before)
        assert torch.equal(c1[6:6 + total], u), ("field copy", n, after, before)      # every element, exactly once or same value
        assert torch.isnan(c1[:6]).all() and torch.isnan(c1[6 + total:]).all()
        y2 = torch.zeros(total + 64, dtype=torch.float64, device="cuda")
        my2, _ = maps(y2, c1)
        assert lib.pt_dense_apply_eo_ws_scatter(_lib.dptr(pw), n, sign, _lib.dptr(u), None, after, before, 1.0, 0.0,
                                                ctypes.byref(my2), _lib.stream_ptr()) == 0
        assert torch.equal(y1, y2), ("result unchanged by the fused copy", n, after, before)
