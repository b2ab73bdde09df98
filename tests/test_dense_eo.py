"""Even-odd application of centrosymmetric operators (csrc/dense_apply_eo.cu) against the NumPy
oracle: the same D @ u, computed as two half-size tensor-core products."""
import numpy as np
import pytest

from oracle import operators as oracle

TOL = 1e-12


def test_oracle_flip_and_fourier_symmetry():
    rng = np.random.default_rng(0)
    for n in (5, 8):
        for sign in (1, -1):
            D = oracle.flip_symmetrize(rng.standard_normal((n, n)), sign)
            assert np.array_equal(D, sign * D[::-1, ::-1])
    for n, order in [(64, 1), (64, 2), (65, 1), (257, 2)]:
        x = oracle.cheb_nodes(n, -1.0, 1.0)
        D = oracle.cheb_mat(x, order)
        s = 1 if order % 2 == 0 else -1
        assert np.array_equal(D, s * D[::-1, ::-1])
        # the flipped matrix is the unflipped one to rounding (the nodes are symmetric to an ulp)
        D0 = oracle.cheb_mat(x, order, flip=False)
        assert np.max(np.abs(D - D0) / np.abs(D0).sum(axis=1, keepdims=True)) < 1e-11
    for n, order in [(64, 1), (64, 2), (63, 1), (1024, 1)]:
        D = oracle.four_mat(n, 2 * np.pi, order)
        s = 1 if order % 2 == 0 else -1
        assert np.array_equal(D, s * D[::-1, ::-1])


def _sym_matrix(rng, n, sign):
    return oracle.flip_symmetrize(rng.standard_normal((n, n)), sign)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [16, 17, 33, 64, 130, 257, 258, 512])
@pytest.mark.parametrize("sign", [1, -1])
def test_eo_matches_oracle_on_symmetrised_random_operators(n, sign):
    import torch

    from pytristan_b200 import _lib
    from pytristan_b200.operators import _DiffMat

    rng = np.random.default_rng(n * 3 + sign)
    D = _sym_matrix(rng, n, sign)
    d_D = torch.from_numpy(D).cuda()
    for (before, after) in [(1, 37), (1, 200), (2, 9), (5, 4), (64, 3), (130, 2)]:
        npts = (before, n, after)
        op = _DiffMat._wrap(d_D, npts, 1, 2 if sign == 1 else 1, None)
        assert op._eo_forms() is not None and op._eo_forms()[0] == sign
        U = rng.standard_normal(before * n * after)
        Y0 = rng.standard_normal(U.size)
        for (alpha, beta) in [(1.0, 0.0), (-0.75, 0.5)]:
            out = torch.from_numpy(Y0.copy()).cuda()
            got = op.apply(torch.from_numpy(U).cuda(), out=out, alpha=alpha, beta=beta).cpu().numpy()
            ref = alpha * oracle.apply_along_axis(D, U[None], npts, 1)[0] + beta * Y0
            bound = abs(alpha) * oracle.abs_bound(D, U[None], npts, 1)[0] + abs(beta) * np.abs(Y0)
            assert np.max(np.abs(got - ref) / bound) <= TOL, (n, sign, before, after, alpha)
    # every tile shape gives the same bits
    lib = _lib.load()
    op = _DiffMat._wrap(d_D, (6, n, 5), 1, 2 if sign == 1 else 1, None)
    U = torch.from_numpy(rng.standard_normal(6 * n * 5)).cuda()
    res = []
    try:
        for cfg in (0, 1, 2, 3):
            lib.pt_debug_set_eo_config(cfg)
            res.append(op.apply(U))
    finally:
        lib.pt_debug_set_eo_config(-1)
    assert all(torch.equal(res[0], r) for r in res[1:])


@pytest.mark.gpu
def test_eo_fused_scalings():
    import torch

    from pytristan_b200.operators import _DiffMat

    rng = np.random.default_rng(5)
    n, before, after = 66, 10, 7
    D = _sym_matrix(rng, n, 1)
    op = _DiffMat._wrap(torch.from_numpy(D).cuda(), (before, n, after), 1, 2, None)
    U = rng.standard_normal((1, before * n * after))
    ref = oracle.apply_along_axis(D, U, (before, n, after), 1).reshape(after, n, before)
    bound = oracle.abs_bound(D, U, (before, n, after), 1).reshape(after, n, before)
    for mode, vec, shape in [("row", rng.uniform(0.5, 2, n), (1, n, 1)), ("after", rng.uniform(0.5, 2, after), (after, 1, 1)),
                             ("before", rng.uniform(0.5, 2, before), (1, 1, before))]:
        got = op.apply(torch.from_numpy(U).cuda(), scale=vec, scale_mode=mode).cpu().numpy().reshape(after, n, before)
        sc = vec.reshape(shape)
        assert np.max(np.abs(got - ref * sc) / (bound * sc)) <= TOL, mode


@pytest.mark.gpu
def test_real_operators_take_the_even_odd_path_and_agree_with_the_plain_kernel():
    import torch

    import pytristan_b200 as pt
    import pytristan_b200.operators as ops_mod

    rng = np.random.default_rng(11)
    cases = [(pt.ChebMat(pt.cheb(257, -1.0, 1.0), order=2), (32, 257, 3), 1),
             (pt.ChebMat(pt.cheb(64), order=1), (64, 50), 0),
             (pt.FourMat(np.arange(1024) * (2 * np.pi / 1024), order=1), (1024, 40), 0),
             (pt.FourMat(np.arange(96) * (3.0 / 96), order=2), (7, 96, 4), 1),
             (pt.FourMat(np.arange(63) * (2 * np.pi / 63), order=1), (63, 33), 0)]
    for op, npts, axis in cases:
        op.npts, op.axis = npts, axis
        assert op._eo_forms() is not None, (type(op).__name__, npts)
        U = rng.standard_normal((2, int(np.prod(npts))))
        D = np.asarray(op)
        ref = oracle.apply_along_axis(D, U, npts, axis)
        bound = oracle.abs_bound(D, U, npts, axis)
        d_u = torch.from_numpy(U).cuda()
        got = op.apply(d_u).cpu().numpy()
        assert np.max(np.abs(got - ref) / bound) <= TOL
        try:
            ops_mod.EVEN_ODD = False
            plain = op.apply(d_u).cpu().numpy()
        finally:
            ops_mod.EVEN_ODD = True
        assert np.max(np.abs(got - plain) / bound) <= TOL
    # operators without the symmetry keep the plain kernel
    x = oracle.cheb_nodes(40)[::-1]
    stretched = pt.ChebMat(x + 0.15 * (1.0 - x * x), order=1)
    assert stretched._eo_forms() is None
    # a stretching mapper that is odd about the centre keeps the node set symmetric: the barycentric
    # matrix is flipped like the CGL one and takes the even-odd path
    xs = np.sinh(2.0 * x) / np.sinh(2.0)
    assert oracle.nodes_symmetric(xs) and not oracle.is_affine_cgl(xs)
    for order in (1, 2):
        op = pt.ChebMat(xs, order=order)
        assert op.weights_kind == "bary" and op._eo_forms() is not None
        ref = oracle.bary_mat(xs, order)
        assert np.max(np.abs(np.asarray(op) - ref) / np.abs(ref).sum(axis=1, keepdims=True)) <= 1e-13
        op.npts, op.axis = (40, 9), 0
        U = rng.standard_normal((2, 360))
        got = op.apply(torch.from_numpy(U).cuda()).cpu().numpy()
        assert np.max(np.abs(got - oracle.apply_along_axis(ref, U, (40, 9), 0)) / oracle.abs_bound(ref, U, (40, 9), 0)) <= TOL


@pytest.mark.gpu
def test_config1_known_answer_through_the_even_odd_path():
    import pytristan_b200 as pt

    x = pt.cheb(64)
    D = pt.ChebMat(x, order=1)
    assert D._eo_forms() is not None
    assert np.max(np.abs(D @ np.sin(x) - np.cos(x))) < 2e-12
