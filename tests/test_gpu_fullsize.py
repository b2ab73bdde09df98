"""GPU parity at the FULL sizes of BASELINE.json's configs, through size-independent properties
(known answers, polynomial exactness, linearity, D*const = 0) and oracle checks on sampled
fields/lines — the oracle itself cannot hold these sizes comfortably."""
import numpy as np
import pytest

from oracle import operators as oracle

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("clean_registries")]


def _torch():
    import torch

    return torch


def _free_gb():
    torch = _torch()
    free, _ = torch.cuda.mem_get_info()
    return free / 2 ** 30


def test_config2_full_batch():
    """1024 x 257 channel grid, d/dx (FourMat) and d2/dy2 (ChebMat) on a batch of 64 fields."""
    import pytristan_b200 as pt

    torch = _torch()
    g = pt.get_grid((0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), (-1.0, 1.0, 257), from_bounds=True, axes=[1], mappers=[pt.cheb])
    Dx = pt.get_four_mat(g, axis=0, order=1)
    Dyy = pt.get_cheb_mat(g, axis=1, order=2)
    P = 1024 * 257
    gen = torch.Generator(device="cuda").manual_seed(1234)
    U = torch.randn(64, P, dtype=torch.float64, device="cuda", generator=gen)
    V = torch.randn(64, P, dtype=torch.float64, device="cuda", generator=gen)
    x, y = g.axpoints(0), g.axpoints(1)
    refs = {0: oracle.four_mat(1024, 1024 * ((x[-1] - x[0]) / 1023), 1), 1: oracle.cheb_mat(y, 2)}
    for op, axis in ((Dx, 0), (Dyy, 1)):
        Y = op.apply(U)
        # oracle on sampled fields of the batch
        for f in (0, 31, 63):
            u = U[f].cpu().numpy()
            ref = oracle.apply_along_axis(refs[axis], u, g.npts, axis)
            bound = oracle.abs_bound(refs[axis], u, g.npts, axis)
            assert np.max(np.abs(Y[f].cpu().numpy() - ref) / bound) <= 1e-12
        # linearity over the whole batch
        Z = op.apply(2.0 * U - 3.0 * V)
        W = 2.0 * Y - 3.0 * op.apply(V)
        assert float((Z - W).abs().max() / W.abs().max()) <= 1e-12
        # constants are annihilated (relative to |D| 1)
        ones = torch.ones(1, P, dtype=torch.float64, device="cuda")
        scale = np.abs(refs[axis]).sum(axis=1).max()
        assert float(op.apply(ones).abs().max()) <= 1e-12 * scale
    # known answer on a smooth field
    X, Yc = np.asarray(g)[0], np.asarray(g)[1]
    f = np.sin(3 * X) * np.exp(Yc)
    assert np.max(np.abs(Dx @ f - 3 * np.cos(3 * X) * np.exp(Yc))) < 1e-10
    assert np.max(np.abs(Dyy @ f - f)) < 5e-7


@pytest.mark.parametrize("radial", ["cheb", "findiff"])
def test_config3_polar_laplacian_full(radial):
    """Polar grid 1024 (theta, Fourier) x 512 (r), mapped non-uniform radial mesh, batch of 16."""
    import pytristan_b200 as pt

    torch = _torch()
    if radial == "cheb":
        g = pt.get_polar_grid(1024, 256, axes=[1], mappers=[pt.cheb], fornberg=True)      # 512 CGL radial nodes on [-1, 1]
    else:
        def stretch(npts, lo, hi):                                                    # custom non-affine mapper
            s = np.linspace(-1.0, 1.0, npts)
            return lo + (hi - lo) * 0.5 * (1.0 + np.tanh(1.2 * s) / np.tanh(1.2))

        g = pt.get_polar_grid(1024, 256, axes=[1], mappers=[stretch], fornberg=True)
    assert g.npts == (1024, 512)
    lap = pt.PolarLaplacian(g, radial=radial)
    T, R = np.asarray(g)[0], np.asarray(g)[1]
    th, r = g.axpoints(0), g.axpoints(1)
    # analytic: r^2 -> 4 ; r^3 cos(3 theta) is harmonic
    # Chebyshev D2 at N = 512 carries O(N^4 eps) ~ 1e-5 rounding at the end nodes (inherent to the
    # collocation matrix, the oracle shows the same); the FD variant is truncation limited
    tol = 1e-4 if radial == "cheb" else 2e-3
    assert np.max(np.abs(lap.apply(R ** 2) - 4.0)) < tol
    assert np.max(np.abs(lap.apply(R ** 3 * np.cos(3 * T)))) < tol * 10
    # oracle on 2 sampled fields of a batch of 16 random fields
    gen = torch.Generator(device="cuda").manual_seed(1234)
    U = torch.randn(16, T.size, dtype=torch.float64, device="cuda", generator=gen)
    Y = lap.apply(U)
    u = U[[0, 15]].cpu().numpy()
    if radial == "cheb":
        ref = oracle.polar_laplacian(th, r, u)
    else:
        b1, b2 = oracle.findiff_band(r, 1, 6), oracle.findiff_band(r, 2, 6)
        Ar = oracle.band_to_dense(b2["start"], b2["W"]) + oracle.band_to_dense(b1["start"], b1["W"]) / r[:, None]
        D2t = oracle.four_mat(1024, 1024 * ((th[-1] - th[0]) / 1023), 2)
        inv = np.tile(np.repeat(1.0 / r ** 2, 1024), 2).reshape(u.shape)
        ref = oracle.apply_along_axis(Ar, u, g.npts, 1) + inv * oracle.apply_along_axis(D2t, u, g.npts, 0)
    got = Y[[0, 15]].cpu().numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) <= 1e-12


@pytest.mark.parametrize("n", [512, 1024])
def test_config4_fd6_cube(n):
    """6th-order FinDiffMat d/dx, d/dy, d/dz on an n^3 field: exact for polynomials of degree <= 6."""
    import pytristan_b200 as pt

    torch = _torch()
    if _free_gb() < 5.0 * n ** 3 * 8 / 2 ** 30:
        pytest.skip("not enough device memory")
    x = np.linspace(0.0, 1.0, n)
    ops = []
    for k in range(3):
        F = pt.FinDiffMat(x, order=1, accuracy=6)
        F.npts, F.axis = (n, n, n), k
        ops.append(F)
    t = torch.from_numpy(x).cuda()
    X, Y, Z = t.view(1, 1, n), t.view(1, n, 1), t.view(n, 1, 1)          # axis 0 (x) fastest
    u = (X ** 6 - 2.0 * X ** 3 * Y ** 2 + Y ** 5 * Z + 3.0 * Z ** 4 + X * Y * Z).contiguous().view(-1)
    exact = [6.0 * X ** 5 - 6.0 * X ** 2 * Y ** 2 + Y * Z, -4.0 * X ** 3 * Y + 5.0 * Y ** 4 * Z + X * Z,
             Y ** 5 + 12.0 * Z ** 3 + X * Y]
    out = torch.empty_like(u)
    for k in range(3):
        ops[k].apply(u, out=out)
        err = float((out.view(n, n, n) - exact[k]).abs().max())
        assert err <= 2e-9, (k, err)          # rounding only: weights are O(n), values O(1)
    assert np.array_equal(ops[0].start, oracle.findiff_band(x, 1, 6)["start"])
    del u, out, X, Y, Z, exact
    torch.cuda.empty_cache()

    # ---- oracle parity on a RANDOM field: sampled lines and planes, componentwise <= 1e-12 of |W||u| ------
    # (hard part 6: the NumPy oracle cannot hold the whole cube; lines along the
    # differentiated axis and planes containing it are independent problems it can)
    gen = torch.Generator(device="cuda").manual_seed(1234 + n)
    u = torch.randn(n ** 3, dtype=torch.float64, device="cuda", generator=gen)
    out = torch.empty_like(u)
    band = oracle.findiff_band(x, 1, 6)
    absband = dict(band, W=np.abs(band["W"]))
    rng = np.random.default_rng(n)
    V, O = u.view(n, n, n), out.view(n, n, n)            # [z][y][x], x fastest
    worst = 0.0
    for k in range(3):
        ops[k].apply(u, out=out)
        # 48 random lines along axis k (+ the corner lines)
        picks = [(0, 0), (n - 1, n - 1), (0, n - 1)] + [tuple(rng.integers(0, n, 2)) for _ in range(48)]
        for (p, q) in picks:
            sl = {0: (p, q, slice(None)), 1: (p, slice(None), q), 2: (slice(None), p, q)}[k]
            line, got = V[sl].cpu().numpy(), O[sl].cpu().numpy()
            ref = oracle.stencil_apply(band, line[None], (n,), 0)[0]
            bound = oracle.stencil_apply(absband, np.abs(line)[None], (n,), 0)[0]
            worst = max(worst, float(np.max(np.abs(got - ref) / bound)))
        # two planes containing axis k: a 2-D grid problem for the oracle
        for fixed in (int(rng.integers(0, n)), n - 1):
            if k == 0:      # x-y plane at z = fixed: grid (nx, ny), axis 0
                plane, got, npts2, ax = V[fixed].cpu().numpy(), O[fixed].cpu().numpy(), (n, n), 0
            elif k == 1:    # x-y plane at z = fixed: axis 1
                plane, got, npts2, ax = V[fixed].cpu().numpy(), O[fixed].cpu().numpy(), (n, n), 1
            else:           # x-z plane at y = fixed: grid (nx, nz), axis 1
                plane, got, npts2, ax = V[:, fixed, :].cpu().numpy(), O[:, fixed, :].cpu().numpy(), (n, n), 1
            ref = oracle.stencil_apply(band, plane.reshape(1, -1), npts2, ax)
            bound = oracle.stencil_apply(absband, np.abs(plane).reshape(1, -1), npts2, ax)
            worst = max(worst, float(np.max(np.abs(got.reshape(1, -1) - ref) / bound)))
    assert worst <= 1e-12, worst
    # the fused one-read gradient against the three passes (another summation order: 1e-12 of max|y|)
    outs = [torch.empty_like(u) for _ in range(3)] if _free_gb() > 4.0 * n ** 3 * 8 / 2 ** 30 else None
    if outs is not None:
        pt.gradient(ops, u, outs=outs)
        for k in range(3):
            ops[k].apply(u, out=out)
            assert float((outs[k] - out).abs().max()) <= 1e-12 * float(out.abs().max()), k


def test_config5a_cheb2048_batched():
    """ChebMat N = 2048 applied to 65536 fields (1 GiB), batch contiguous (K1, before = 65536)."""
    import pytristan_b200 as pt

    torch = _torch()
    n, nb = 2048, 65536
    if _free_gb() < 3.0:
        pytest.skip("not enough device memory")
    x = pt.cheb(n)
    D = pt.ChebMat(x, order=1)
    D.npts, D.axis = (nb, n), 1
    gen = torch.Generator(device="cuda").manual_seed(1234)
    U = torch.randn(n, nb, dtype=torch.float64, device="cuda", generator=gen)
    Y = D.apply(U.view(-1)).view(n, nb)
    Dref = oracle.cheb_mat(x, 1)
    cols = np.array([0, 1, 777, 32768, 65534, 65535])
    u = U[:, cols].cpu().numpy()
    ref = Dref @ u
    bound = np.abs(Dref) @ np.abs(u)
    assert np.max(np.abs(Y[:, cols].cpu().numpy() - ref) / bound) <= 1e-12
    # D differentiates x exactly (up to |D||x| rounding) on every field of a small stack
    lin = torch.from_numpy(x).cuda().view(n, 1).expand(n, 4).contiguous()
    D4 = pt.ChebMat(x, order=1)
    D4.npts, D4.axis = (4, n), 1
    scale = np.abs(Dref).sum(axis=1).max()
    assert float((D4.apply(lin.view(-1)).view(n, 4) - 1.0).abs().max()) <= 1e-12 * scale


def test_config5b_cheb512_three_axes_single_gpu():
    """3-D Chebyshev 512^3: D2x + D2y + D2z with the fused accumulate epilogue, sampled lines vs oracle."""
    import pytristan_b200 as pt

    torch = _torch()
    n = 256 if _free_gb() < 6.0 else 512
    x = pt.cheb(n)
    ops = []
    for k in range(3):
        D = pt.ChebMat(x, order=2)
        D.npts, D.axis = (n, n, n), k
        ops.append(D)
    gen = torch.Generator(device="cuda").manual_seed(1234)
    u = torch.randn(n ** 3, dtype=torch.float64, device="cuda", generator=gen)
    lap = ops[0].apply(u)
    ops[1].apply(u, out=lap, beta=1.0)
    ops[2].apply(u, out=lap, beta=1.0)
    D2 = oracle.cheb_mat(x, 2)
    U3 = u.view(n, n, n)
    L3 = lap.view(n, n, n)
    rng = np.random.default_rng(5)
    worst = 0.0
    for _ in range(24):
        k, j, i = (int(v) for v in rng.integers(0, n, 3))
        lx, ly, lz = U3[k, j, :].cpu().numpy(), U3[k, :, i].cpu().numpy(), U3[:, j, i].cpu().numpy()
        ref = D2[i] @ lx + D2[j] @ ly + D2[k] @ lz
        bound = np.abs(D2[i]) @ np.abs(lx) + np.abs(D2[j]) @ np.abs(ly) + np.abs(D2[k]) @ np.abs(lz)
        worst = max(worst, abs(float(L3[k, j, i]) - ref) / bound)
    assert worst <= 1e-12
