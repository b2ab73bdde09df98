"""Seeded random-shape sweeps of the apply kernels against the NumPy oracle: ragged sizes, odd
strides (the 8-byte paths), every dense tile shape, scatter maps with arbitrary strides (scattered
into this rank's own buffer, so one GPU suffices), FFT sizes and layouts."""
import ctypes

import numpy as np
import pytest

from oracle import operators as oracle

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _torch():
    import torch

    return torch


def _shapes(rng, count, nmax):
    out = []
    for _ in range(count):
        n = int(rng.integers(2, nmax))
        before = int(rng.choice([1, 1, 2, 3, 5, 8, 17, 64, 130]))
        after = int(rng.integers(1, 6)) if before > 1 else int(rng.integers(1, 300))
        out.append((before, n, after))
    return out


@pytest.mark.parametrize("config", [-1, 2, 5, 9])
def test_dense_apply_random_shapes(config):
    import pytristan_b200 as pt
    from pytristan_b200 import _lib
    from pytristan_b200.operators import _DiffMat

    torch = _torch()
    lib = _lib.load()
    rng = np.random.default_rng(100 + config)
    try:
        lib.pt_debug_set_dense_config(config)
        for (before, n, after) in _shapes(rng, 14, 300):
            D = rng.standard_normal((n, n))
            U = rng.standard_normal(after * n * before)
            op = _DiffMat._wrap(torch.from_numpy(D).cuda(), (before, n, after), 1, 1, None)
            alpha, beta = (1.0, 0.0) if rng.random() < 0.5 else (-0.75, 0.5)
            Y0 = rng.standard_normal(U.size)
            out = torch.from_numpy(Y0.copy()).cuda()
            got = op.apply(torch.from_numpy(U).cuda(), out=out, alpha=alpha, beta=beta).cpu().numpy()
            ref = alpha * oracle.apply_along_axis(D, U[None], (before, n, after), 1)[0] + beta * Y0
            bound = abs(alpha) * oracle.abs_bound(D, U[None], (before, n, after), 1)[0] + abs(beta) * np.abs(Y0)
            assert np.max(np.abs(got - ref) / bound) <= TOL, (config, before, n, after)
    finally:
        lib.pt_debug_set_dense_config(-1)


def test_dense_tile_shapes_agree_bitwise():
    """Every tile shape accumulates over k in the same order and rounds the epilogue explicitly:
    the shape-based choice never changes a result bit (what makes sharded runs reproduce 1-GPU runs)."""
    from pytristan_b200 import _lib
    from pytristan_b200.operators import _DiffMat

    torch = _torch()
    lib = _lib.load()
    rng = np.random.default_rng(9)
    try:
        for (before, n, after) in [(1, 257, 90), (64, 130, 3), (3, 77, 5), (1, 64, 1000), (256, 512, 2)]:
            D = torch.from_numpy(rng.standard_normal((n, n))).cuda()
            U = torch.from_numpy(rng.standard_normal(after * n * before)).cuda()
            Y0 = torch.from_numpy(rng.standard_normal(after * n * before)).cuda()
            op = _DiffMat._wrap(D, (before, n, after), 1, 1, None)
            res = []
            for config in (2, 3, 5, 9):
                lib.pt_debug_set_dense_config(config)
                res.append(op.apply(U, out=Y0.clone(), alpha=1.25, beta=-0.3))
            for r in res[1:]:
                assert torch.equal(r, res[0]), (before, n, after)
    finally:
        lib.pt_debug_set_dense_config(-1)


def test_stencil_apply_random_shapes():
    import pytristan_b200 as pt

    torch = _torch()
    rng = np.random.default_rng(77)
    for (before, n, after) in _shapes(rng, 24, 120):
        n = max(n, 12)
        order = int(rng.integers(1, 3))
        acc = int(rng.choice([2, 4, 6]))
        kind = rng.choice(["uniform", "periodic", "mapped"])
        if kind == "mapped":
            x = oracle.cheb_nodes(n, 0.0, 1.0)
        elif kind == "periodic":
            x = np.linspace(0.0, 1.0 - 1.0 / n, n)
        else:
            x = np.linspace(0.0, 1.0, n)
        F = pt.FinDiffMat(x, order=order, accuracy=acc, periodic=(kind == "periodic"))
        F.npts, F.axis = (before, n, after), 1
        U = rng.standard_normal((2, before * n * after))
        band = oracle.findiff_band(x, order, acc, periodic=(kind == "periodic"))
        assert np.array_equal(F.start, band["start"])
        ref = oracle.stencil_apply(band, U, (before, n, after), 1)
        got = F.apply(torch.from_numpy(U).cuda()).cpu().numpy()
        assert np.max(np.abs(got - ref)) <= TOL * np.max(np.abs(ref)), (before, n, after, order, acc, kind)


def test_fft_random_shapes():
    import pytristan_b200 as pt

    torch = _torch()
    rng = np.random.default_rng(55)
    for _ in range(20):
        n = int(2 ** rng.integers(2, 12))
        before = int(rng.choice([1, 1, 2, 4, 6, 10, 34]))
        after = int(rng.integers(1, 4)) if before > 1 else int(rng.integers(1, 40))
        if before * n * after > 2_000_000:
            after = 1
        order = int(rng.integers(1, 5))
        period = float(rng.uniform(0.5, 7.0))
        x = np.arange(n) * (period / n)
        F = pt.FourMat(x, order=order, method="fft")
        F.npts, F.axis = (before, n, after), 1
        U = rng.standard_normal((1, before * n * after))
        D = oracle.four_mat(n, period, order)
        ref = oracle.apply_along_axis(D, U, F.npts, 1)
        got = F.apply(torch.from_numpy(U).cuda()).cpu().numpy()
        assert np.max(np.abs(got - ref) / oracle.abs_bound(D, U, F.npts, 1)) <= TOL, (before, n, after, order)


def test_scatter_maps_random_single_rank():
    """pt_dense_apply_scatter / pt_peer_scatter with random destination maps whose every destination
    is a region of ONE local buffer: exercises the address arithmetic (even and odd strides, the
    vector and scalar store paths, K1 and K2) against the NumPy restatement `scatter_dest`."""
    from pytristan_b200 import _lib
    from pytristan_b200.operators import _DiffMat
    from pytristan_b200.peer import scatter_dest

    torch = _torch()
    lib = _lib.load()
    rng = np.random.default_rng(31)
    for trial in range(16):
        nd = int(rng.integers(1, 4))                       # destinations
        chunk = int(rng.integers(1, 9)) * (2 if trial % 2 == 0 else 1) * (3 if trial % 3 == 0 else 1)
        n = nd * chunk - int(rng.integers(0, min(chunk, 3)))   # last destination may be ragged
        before = int(rng.choice([1, 2, 3, 4, 6]))
        a_inner = int(rng.integers(1, 4))
        a_outer = int(rng.integers(1, 4))
        after = a_inner * a_outer
        # destination layout (per destination): [a_outer][i % chunk][a_inner][b], padded strides
        pad = int(rng.integers(0, 3)) * (2 if trial % 2 == 0 else 1)
        s_inner = before + pad
        s_i = a_inner * s_inner if before > 1 else 1
        if before == 1:
            s_inner = chunk + pad                          # b == 0: i is the fastest destination index
            s_outer = a_inner * s_inner
        else:
            s_outer = chunk * s_i
        region = a_outer * s_outer + 8
        offset = int(rng.integers(0, 3)) * 2
        total = nd * region + offset + 8
        dest = torch.full((total,), float("nan"), dtype=torch.float64, device="cuda")
        table = torch.tensor([dest.data_ptr() + 8 * q * region for q in range(nd)], dtype=torch.int64, device="cuda")
        m = _lib.ScatterMap()
        m.nranks, m.chunk, m.a_inner, m.s_outer, m.s_inner, m.s_i, m.offset = nd, chunk, a_inner, s_outer, s_inner, s_i, offset
        m.base = table.data_ptr()

        D = rng.standard_normal((n, n))
        if trial % 3 == 0 and n >= 16:
            D = oracle.flip_symmetrize(D, -1)               # goes through pt_dense_apply_eo_scatter
        U = rng.standard_normal(after * n * before)
        C = rng.standard_normal(after * n * before)
        op = _DiffMat._wrap(torch.from_numpy(D).cuda(), (before, n, after), 1, 1, None)
        d_u, d_c = torch.from_numpy(U).cuda(), torch.from_numpy(C).cuda()
        a, i, b = np.meshgrid(np.arange(after), np.arange(n), np.arange(before), indexing="ij")
        q, off = scatter_dest(a, i, b, chunk, a_inner, s_outer, s_inner, s_i, offset)
        flat = (q * region + off).ravel()
        assert np.unique(flat).size == flat.size, "test map is not injective"

        # (1) pure redistribution
        _lib.check(lib.pt_peer_scatter(_lib.dptr(d_u), n, after, before, ctypes.byref(m), _lib.stream_ptr()), "scatter")
        got = dest.cpu().numpy()
        assert np.array_equal(got[flat], U), (trial, "peer_scatter")
        mask = np.ones(total, bool); mask[flat] = False
        assert np.all(np.isnan(got[mask])), (trial, "peer_scatter wrote outside its map")

        # (2) operator with scatter epilogue, alpha/beta and the local C term
        dest.fill_(float("nan"))
        op._apply_scatter(d_u, after, n, before, d_c, 1.25, -0.5, m)
        got = dest.cpu().numpy()
        ref = 1.25 * oracle.apply_along_axis(D, U[None], (before, n, after), 1)[0] - 0.5 * C
        bound = 1.25 * oracle.abs_bound(D, U[None], (before, n, after), 1)[0] + 0.5 * np.abs(C)
        assert np.max(np.abs(got[flat] - ref) / bound) <= TOL, (trial, "dense scatter")
        assert np.all(np.isnan(got[mask])), (trial, "dense scatter wrote outside its map")
        # bit-identical to the un-scattered kernel
        plain = op.apply(d_u, out=d_c.clone(), alpha=1.25, beta=-0.5).cpu().numpy()
        assert np.array_equal(got[flat], plain), (trial, "scatter != plain bits")


def test_field_sharding_is_bitwise_invisible():
    """fields shard across GPUs with no communication and the results must equal the
    un-sharded run bit for bit — every kernel computes an output element with the same arithmetic
    whatever the number of fields in the launch (checked here by splitting one batch in two)."""
    import pytristan_b200 as pt

    torch = _torch()
    rng = np.random.default_rng(2)
    nx, ny = 64, 33
    x = np.arange(nx) * (2 * np.pi / nx)
    y = pt.cheb(ny, -1.0, 1.0)
    ops = []
    for op, axis in ((pt.FourMat(x, order=1), 0), (pt.ChebMat(y, order=2), 1), (pt.FourMat(x, order=2, method="fft"), 0),
                     (pt.FinDiffMat(x, order=1, accuracy=6, periodic=True), 0), (pt.FinDiffMat(y, order=2, accuracy=4), 1)):
        op.npts, op.axis = (nx, ny), axis
        ops.append(op)
    U = torch.from_numpy(rng.standard_normal((10, nx * ny))).cuda()
    for k, op in enumerate(ops):
        whole = op.apply(U)
        parts = torch.cat([op.apply(U[:3].contiguous()), op.apply(U[3:].contiguous())])
        if getattr(op, "method", "dense") == "fft":
            # the FFT path transforms two rows at once (z = u1 + i u2): which row a row is paired
            # with changes with the split (3 fields x 33 rows is odd), and with it the last bits
            bound = oracle.abs_bound(np.asarray(op), U.cpu().numpy(), (nx, ny), op.axis)
            assert np.max(np.abs((whole - parts).cpu().numpy()) / bound) <= 1e-13, k
        else:
            assert torch.equal(whole, parts), (k, type(op).__name__)
