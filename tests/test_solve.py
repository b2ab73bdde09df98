"""Boundary rows and solves along an axis (pytristan_b200.solve) against numpy.linalg."""
import numpy as np
import pytest

from oracle import operators as oracle

pytestmark = pytest.mark.gpu


def test_with_rows_and_dirichlet():
    import pytristan_b200 as pt

    x = pt.cheb(17)
    D2 = pt.ChebMat(x, order=2)
    D1 = pt.ChebMat(x, order=1)
    A = pt.dirichlet(D2)
    ref = np.asarray(D2).copy()
    ref[0] = 0.0; ref[0, 0] = 1.0
    ref[-1] = 0.0; ref[-1, -1] = 1.0
    assert np.array_equal(np.asarray(A), ref)
    B = pt.with_rows(D2, [-1], np.asarray(D1)[-1:])           # Neumann row at the last node
    ref = np.asarray(D2).copy(); ref[-1] = np.asarray(D1)[-1]
    assert np.array_equal(np.asarray(B), ref)
    with pytest.raises(IndexError):
        pt.with_rows(D2, [17], np.zeros((1, 17)))
    with pytest.raises(ValueError):
        pt.with_rows(D2, [0, 0], np.zeros((2, 17)))


@pytest.mark.parametrize("n", [2, 5, 33, 64, 257])
def test_inverse_matches_numpy(n):
    import pytristan_b200 as pt

    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n)) + np.diag(np.arange(n) % 3 == 0) * 1e-3      # needs pivoting
    A[0, 0] = 0.0
    s = pt.AxisSolver(A, method="inverse")
    inv = np.asarray(s.inverse)
    ref = np.linalg.inv(A)
    cond = np.linalg.cond(A)
    assert np.max(np.abs(inv - ref)) <= 1e-13 * cond * np.max(np.abs(ref))
    assert np.max(np.abs(inv @ A - np.eye(n))) <= 1e-13 * cond


def test_singular_matrix_raises():
    import pytristan_b200 as pt

    A = np.ones((6, 6))
    with pytest.raises(np.linalg.LinAlgError):
        pt.AxisSolver(A)
    with pytest.raises(np.linalg.LinAlgError):
        pt.AxisSolver(A, method="inverse")


def test_poisson_chebyshev_known_answer():
    """u'' = f on [-1, 1], u(+-1) given: u = sin(2x) + x."""
    import pytristan_b200 as pt

    n = 48
    x = pt.cheb(n)
    A = pt.dirichlet(pt.ChebMat(x, order=2))
    u = np.sin(2 * x) + x
    f = -4 * np.sin(2 * x)
    f[0], f[-1] = u[0], u[-1]
    y = pt.AxisSolver(A).solve(f)
    assert np.max(np.abs(y - u)) < 1e-11
    y = pt.AxisSolver(A, method="inverse").solve(f)
    assert np.max(np.abs(y - u)) < 1e-11


@pytest.mark.parametrize("method", ["lu", "inverse"])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_helmholtz_along_each_axis_of_a_3d_grid(axis, method):
    """(D2 - k^2) y = f with Dirichlet rows along one axis, for every line of a batch of 3-D fields."""
    import torch

    import pytristan_b200 as pt

    npts = (12, 20, 9)
    n = npts[axis]
    x = pt.cheb(n, 0.0, 2.0)
    D2 = pt.ChebMat(x, order=2)
    A = np.asarray(D2) - 2.5 ** 2 * np.eye(n)
    A[0] = 0.0; A[0, 0] = 1.0
    A[-1] = 0.0; A[-1, -1] = 1.0
    solver = pt.AxisSolver(A, npts=npts, axis=axis, method=method, block=16)
    rng = np.random.default_rng(11)
    F = rng.standard_normal((3, int(np.prod(npts))))
    y = solver.solve(torch.from_numpy(F).cuda()).cpu().numpy()
    # oracle: numpy.linalg.solve on every line
    ref = oracle.apply_along_axis(np.linalg.inv(A), F, npts, axis)
    back = oracle.apply_along_axis(A, y, npts, axis)
    cond = np.linalg.cond(A)
    assert np.max(np.abs(y - ref)) <= 1e-13 * cond * np.max(np.abs(ref))
    assert np.max(np.abs(back - F)) <= 1e-12 * cond * np.max(np.abs(F))


@pytest.mark.parametrize("n", [129, 257])
def test_iterative_refinement_reaches_the_backward_stable_residual(n):
    """Chebyshev Poisson operator with Dirichlet rows (cond ~ n^4): the product with the explicit inverse
    leaves a residual ~ cond * eps; one refinement step (r = f - A y, y += A^-1 r: two more DMMA passes)
    brings it to ~ eps * |A| |y|, like an LU solve with refinement."""
    import torch

    import pytristan_b200 as pt

    x = pt.cheb(n, -1.0, 1.0)
    A = pt.dirichlet(pt.ChebMat(x, order=2))
    An = np.asarray(A)
    npts = (24, n)
    rng = np.random.default_rng(n)
    F = rng.standard_normal((2, 24 * n))
    Fd = torch.from_numpy(F).cuda()
    plain = pt.AxisSolver(A, npts=npts, axis=1, method="inverse")
    refined = pt.AxisSolver(A, npts=npts, axis=1, refine=1, method="inverse")
    y0 = plain.solve(Fd).cpu().numpy()
    y1 = refined.solve(Fd).cpu().numpy()
    assert np.array_equal(plain.solve(Fd, refine=1).cpu().numpy(), y1)

    def scaled_residual(y):          # componentwise backward error: |f - A y| / (|A| |y| + |f|)
        r = oracle.apply_along_axis(An, y, npts, 1) - F
        return np.max(np.abs(r) / (oracle.abs_bound(An, y, npts, 1) + np.abs(F)))

    r0, r1 = scaled_residual(y0), scaled_residual(y1)
    assert r1 <= 50 * np.finfo(np.float64).eps, (r0, r1)
    assert r1 <= r0
    ref = np.linalg.solve(An, F.reshape(2, n, 24).transpose(1, 0, 2).reshape(n, -1))
    ref = ref.reshape(n, 2, 24).transpose(1, 0, 2).reshape(2, -1)
    cond = np.linalg.cond(An)
    assert np.max(np.abs(y1 - ref)) <= 1e-14 * cond * np.max(np.abs(ref))
    # host arrays go through the same path
    assert np.array_equal(refined.solve(F), y1)


# ---- factorisation-based solve (pt_lu_factor / pt_lu_plan / pt_lu_solve) ---------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 33, 64, 130, 257, 600])
def test_lu_factors_match_the_oracle(n):
    """P A = L U on the device against oracle.lu_factor (same pivot rule): identical row order, factors to
    rounding, reconstruction to eps * |L||U|."""
    import pytristan_b200 as pt

    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    A[0, 0] = 0.0 if n > 1 else 2.0
    s = pt.AxisSolver(A)
    perm, L, U = s.factors()
    perm_ref, LU_ref = oracle.lu_factor(A)
    assert np.array_equal(perm, perm_ref)
    eps = np.finfo(np.float64).eps
    bound = np.abs(L) @ np.abs(U)
    assert np.max(np.abs(L)) <= 1.0
    assert np.all(np.abs(A[perm] - L @ U) <= 2 * n * eps * bound + 1e-300)
    LU = np.tril(L, -1) + U
    growth = max(1.0, np.linalg.cond(A)) ** 0.5
    assert np.max(np.abs(LU - LU_ref)) <= 1e-12 * growth * np.max(np.abs(LU_ref))


@pytest.mark.parametrize("block", [16, 64, 128])
@pytest.mark.parametrize("npts,axis", [((37,), 0), ((130, 6), 0), ((6, 130), 1), ((5, 257, 4), 1), ((3, 7, 300), 2),
                                       ((7, 64, 3), 1)])
def test_lu_solve_matches_lapack(npts, axis, block):
    """Blocked substitution on row ranges of the fields (both kernels: contiguous and strided axis, odd strides,
    ragged last block) against numpy.linalg.solve on every grid line."""
    import torch

    import pytristan_b200 as pt

    n = npts[axis]
    rng = np.random.default_rng(n + block)
    A = rng.standard_normal((n, n)) + 0.5 * np.sqrt(n) * np.eye(n)
    A[[0, n // 2]] = A[[n // 2, 0]]                                 # pivoting moves rows
    solver = pt.AxisSolver(A, npts=npts, axis=axis, block=block)
    F = rng.standard_normal((3, int(np.prod(npts))))
    Fd = torch.from_numpy(F).cuda()
    y = solver.solve(Fd).cpu().numpy()
    assert np.array_equal(Fd.cpu().numpy(), F)                       # the right-hand side is left alone
    ref = oracle.solve_along_axis(A, F, npts, axis)
    eps = np.finfo(np.float64).eps
    cond = np.linalg.cond(A)
    assert np.max(np.abs(y - ref)) <= 1e-13 * cond * np.max(np.abs(ref))
    # componentwise backward error of an LU solve: |f - A y| <= c n eps |A||y| (growth-free matrices)
    r = oracle.apply_along_axis(A, y, npts, axis) - F
    assert np.max(np.abs(r) / (oracle.abs_bound(A, y, npts, axis) + np.abs(F))) <= 4 * n * eps
    assert np.array_equal(solver.solve(F), y)                        # host arrays: same path
    out = torch.empty_like(Fd)
    assert solver.solve(Fd, out=out).data_ptr() == out.data_ptr() and np.array_equal(out.cpu().numpy(), y)


@pytest.mark.parametrize("n", [129, 257, 1024])
def test_lu_solve_is_backward_stable_on_the_chebyshev_poisson_operator(n):
    """cond ~ n^4: the LU solve is at LAPACK's level (componentwise) and below the product with the explicit inverse,
    backward stable normwise, and componentwise after one refinement step."""
    import torch

    import pytristan_b200 as pt

    x = pt.cheb(n, -1.0, 1.0)
    A = pt.dirichlet(pt.ChebMat(x, order=2))
    An = np.asarray(A)
    npts = (16, n)
    rng = np.random.default_rng(n)
    F = rng.standard_normal((2, 16 * n))
    Fd = torch.from_numpy(F).cuda()
    y_lu = pt.AxisSolver(A, npts=npts, axis=1).solve(Fd).cpu().numpy()
    y_inv = pt.AxisSolver(A, npts=npts, axis=1, method="inverse").solve(Fd).cpu().numpy()

    def backward_error(y):
        r = oracle.apply_along_axis(An, y, npts, 1) - F
        return np.max(np.abs(r) / (oracle.abs_bound(An, y, npts, 1) + np.abs(F)))

    ref = oracle.solve_along_axis(An, F, npts, 1)
    e_lu, e_inv, e_lapack = backward_error(y_lu), backward_error(y_inv), backward_error(ref)
    # Componentwise, Gaussian elimination with partial pivoting is only as good as the row scaling allows (Dirichlet
    # rows of size 1 beside rows of size n^4): the bar is LAPACK's own figure on the same systems ...
    assert e_lu <= 8 * e_lapack, (e_lu, e_inv, e_lapack)
    assert e_lu <= e_inv
    # ... normwise it is backward stable per grid line
    r = (oracle.apply_along_axis(An, y_lu, npts, 1) - F).reshape(2, n, 16)
    yl = y_lu.reshape(2, n, 16)
    eta = np.max(np.abs(r), axis=1) / (np.max(np.sum(np.abs(An), axis=1)) * np.max(np.abs(yl), axis=1)
                                       + np.max(np.abs(F.reshape(2, n, 16)), axis=1))
    assert np.max(eta) <= n * np.finfo(np.float64).eps, np.max(eta)
    assert np.max(np.abs(y_lu - ref)) <= 1e-14 * np.linalg.cond(An) * np.max(np.abs(ref))
    # and one refinement step makes it componentwise backward stable as well
    y_ref1 = pt.AxisSolver(A, npts=npts, axis=1, refine=1).solve(Fd).cpu().numpy()
    assert backward_error(y_ref1) <= max(50, n) * np.finfo(np.float64).eps


@pytest.mark.parametrize("after,before", [(5, 1), (1, 6), (3, 8), (2, 7)])
def test_dense_apply_on_row_ranges(after, before):
    """pt_dense_apply_ld: rows [r0, r0 + n_in) of U -> rows [q0, q0 + n_out) of Y, the rest of Y untouched."""
    import torch

    from pytristan_b200 import _lib

    lib = _lib.load()
    n, n_in, n_out, r0, q0 = 50, 22, 12, 10, 32
    rng = np.random.default_rng(after * 10 + before)
    D = rng.standard_normal((n_out, n_in))
    U = rng.standard_normal((after, n, before))
    Y0 = rng.standard_normal((after, n, before))
    d_D = torch.from_numpy(D).cuda()
    packed = torch.zeros(int(lib.pt_packed_size(n_out, n_in)), dtype=torch.float64, device="cuda")
    _lib.check(lib.pt_operator_pack(_lib.dptr(d_D), n_out, n_in, n_in, _lib.dptr(packed), _lib.stream_ptr()), "pack")
    d_U = torch.from_numpy(U).cuda()
    d_Y = torch.from_numpy(Y0).cuda()
    ld = n * before
    rc = lib.pt_dense_apply_ld(_lib.dptr(packed), n_out, n_in, d_U.data_ptr() + 8 * r0 * before, ld,
                               d_Y.data_ptr() + 8 * q0 * before, ld, after, before, -1.0, 1.0, _lib.stream_ptr())
    _lib.check(rc, "pt_dense_apply_ld")
    ref = Y0.copy()
    ref[:, q0:q0 + n_out, :] -= np.einsum("ik,akb->aib", D, U[:, r0:r0 + n_in, :])
    got = d_Y.cpu().numpy()
    bound = np.einsum("ik,akb->aib", np.abs(D), np.abs(U[:, r0:r0 + n_in, :])) + np.abs(Y0[:, q0:q0 + n_out, :])
    assert np.array_equal(got[:, :q0], Y0[:, :q0]) and np.array_equal(got[:, q0 + n_out:], Y0[:, q0 + n_out:])
    assert np.max(np.abs(got - ref)[:, q0:q0 + n_out] / bound) <= 1e-14
    assert lib.pt_dense_apply_ld(_lib.dptr(packed), n_out, n_in, d_U.data_ptr(), n_in * before - 1, d_Y.data_ptr(), ld,
                                 after, before, 1.0, 0.0, _lib.stream_ptr()) != 0


def test_lu_solve_large_n_is_normwise_backward_stable():
    """n = 2000 (16 block rows of 125 -> 128 rows, multi-level off-diagonal products): normwise backward error at
    the n * eps level and agreement with LAPACK to the conditioning."""
    import torch

    import pytristan_b200 as pt

    n, lines = 2000, 24
    rng = np.random.default_rng(7)
    A = rng.standard_normal((n, n)) + 0.1 * np.sqrt(n) * np.eye(n)
    solver = pt.AxisSolver(A, npts=(lines, n), axis=1)
    F = rng.standard_normal((1, lines * n))
    y = solver.solve(torch.from_numpy(F).cuda()).cpu().numpy()
    r = oracle.apply_along_axis(A, y, (lines, n), 1) - F
    eta = np.max(np.abs(r)) / (np.max(np.sum(np.abs(A), axis=1)) * np.max(np.abs(y)) + np.max(np.abs(F)))
    assert eta <= n * np.finfo(np.float64).eps, eta
    ref = oracle.solve_along_axis(A, F, (lines, n), 1)
    assert np.max(np.abs(y - ref)) <= 1e-13 * np.linalg.cond(A) * np.max(np.abs(ref))
