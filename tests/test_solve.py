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
    s = pt.AxisSolver(A)
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


@pytest.mark.parametrize("axis", [0, 1, 2])
def test_helmholtz_along_each_axis_of_a_3d_grid(axis):
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
    solver = pt.AxisSolver(A, npts=npts, axis=axis)
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
    plain = pt.AxisSolver(A, npts=npts, axis=1)
    refined = pt.AxisSolver(A, npts=npts, axis=1, refine=1)
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
