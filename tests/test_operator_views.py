"""Operator objects as ndarrays: slices, copies, in-place edits and the ``@`` dispatch
(non-square views, stale device caches, ambiguous stacks of n fields, host sync of
``apply_host`` with CPU tensors)."""
import numpy as np
import pytest
from conftest import clear_grids

from oracle import operators as oracle

pytestmark = pytest.mark.gpu


def _cheb(n, order=1):
    import pytristan_b200 as pt

    x = pt.cheb(n, -1.0, 1.0)
    return pt.ChebMat(x, order=order), oracle.cheb_mat(x, order)


@pytest.mark.parametrize("n", [17, 64, 130])
def test_row_and_column_slices_multiply_as_plain_matrices(n):
    rng = np.random.default_rng(n)
    D, Dref = _cheb(n, 2)
    u = rng.standard_normal(n)
    M = rng.standard_normal((n, 5))
    for sl in (np.s_[1:-1, :], np.s_[:3, :], np.s_[2:, :]):
        V = D[sl]
        assert V.npts is None and V.shape == Dref[sl].shape
        got = V @ u
        ref = Dref[sl] @ u
        assert got.shape == ref.shape
        bound = np.abs(Dref[sl]) @ np.abs(u)
        assert np.max(np.abs(got - ref) / bound) <= 1e-12
        got = V @ M
        assert got.shape == (Dref[sl].shape[0], 5)
        assert np.max(np.abs(got - Dref[sl] @ M) / (np.abs(Dref[sl]) @ np.abs(M))) <= 1e-12
    # interior block: n-2 x n-2
    V = D[1:-1, 1:-1]
    w = rng.standard_normal(n - 2)
    ref = Dref[1:-1, 1:-1] @ w
    got = V @ w
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / (np.abs(Dref[1:-1, 1:-1]) @ np.abs(w))) <= 1e-12
    with pytest.raises(ValueError):
        V @ rng.standard_normal(n)        # wrong length for the sliced matrix


def test_findiff_slices_are_dense_matrices():
    import pytristan_b200 as pt

    x = np.linspace(0.0, 1.0, 40)
    F = pt.FinDiffMat(x, order=1, accuracy=4)
    Fref = np.asarray(F).copy()
    u = np.random.default_rng(0).standard_normal(40)
    got = F[1:-1, :] @ u
    assert got.shape == (38,)
    assert np.max(np.abs(got - Fref[1:-1] @ u)) <= 1e-12 * np.max(np.abs(Fref[1:-1] @ u))


def test_operator_entries_are_frozen_and_copies_track_their_edits():
    D, Dref = _cheb(33, 1)
    with pytest.raises(ValueError):
        D[0, :] = 0.0                      # the device forms would go stale: entries are read-only
    u = np.random.default_rng(1).standard_normal(33)
    A = D.copy()                           # writable plain matrix: device forms rebuilt on every use
    assert A.flags.writeable and A.npts is None
    y0 = A @ u
    assert np.max(np.abs(y0 - Dref @ u) / (np.abs(Dref) @ np.abs(u))) <= 1e-12
    A[0, :] = 0.0
    A[0, 0] = 1.0
    A[-1, :] = 0.0
    A[-1, -1] = 1.0
    ref = np.asarray(A) @ u
    got = A @ u
    assert got[0] == u[0] and got[-1] == u[-1]
    assert np.max(np.abs(got - ref) / (np.abs(np.asarray(A)) @ np.abs(u))) <= 1e-12
    # the supported way to impose boundary rows keeps the fast path
    import pytristan_b200 as pt

    Bc = pt.dirichlet(D)
    assert np.array_equal(np.asarray(Bc), np.asarray(A))
    assert np.array_equal(Bc @ u, got) or np.max(np.abs((Bc @ u) - got)) <= 1e-12 * np.max(np.abs(got))


def test_matmul_of_exactly_n_stacked_fields_differentiates_them():
    """64 fields on a 64 x 64 grid have shape (64, 4096): they are fields, not a matrix operand."""
    import pytristan_b200 as pt

    clear_grids()
    pt.drop_all_mats()
    try:
        g = pt.get_grid((-1.0, 1.0, 64), (0.0, 1.0, 64), from_bounds=True, axes=[0], mappers=[pt.cheb])
        for axis in (0, 1):
            D = pt.get_cheb_mat(g, axis=0, order=1) if axis == 0 else pt.get_findiff_mat(g, axis=1, order=1, accuracy=4)
            U = np.random.default_rng(axis).standard_normal((64, 64 * 64))
            Dn = np.asarray(D)
            ref = oracle.apply_along_axis(Dn, U, g.npts, axis)
            got = D @ U
            assert got.shape == U.shape
            assert np.max(np.abs(got - ref) / oracle.abs_bound(Dn, U, g.npts, axis)) <= 1e-12
            assert np.array_equal(got, D.apply(U))
            # a (n, k) operand that is not a stack of whole fields is a matrix product
            M = np.random.default_rng(5).standard_normal((64, 7))
            assert np.max(np.abs(D @ M - Dn @ M) / (np.abs(Dn) @ np.abs(M) + 1e-300)) <= 1e-12
    finally:
        pt.drop_all_mats()
        clear_grids()


def test_matmul_on_a_1d_grid_keeps_numpy_semantics():
    D, Dref = _cheb(48, 1)
    M = np.random.default_rng(2).standard_normal((48, 48))
    got = D @ M
    assert np.max(np.abs(got - Dref @ M) / (np.abs(Dref) @ np.abs(M))) <= 1e-12
    u = M[:, 0].copy()
    assert np.max(np.abs(D @ u - Dref @ u) / (np.abs(Dref) @ np.abs(u))) <= 1e-12


def test_apply_host_outputs_are_ready_on_return_for_cpu_tensors():
    """No torch.cuda.synchronize() between the call and the read of the pinned outputs."""
    import torch

    import pytristan_b200 as pt

    x = pt.cheb(65, -1.0, 1.0)
    D = pt.ChebMat(x, order=1)
    D.npts, D.axis = (32, 65), 1
    rng = np.random.default_rng(3)
    for _ in range(5):
        U = rng.standard_normal((48, 32 * 65))
        h_in = torch.from_numpy(U).pin_memory()
        h_out = [torch.full((48, 32 * 65), float("nan"), dtype=torch.float64).pin_memory()]
        out = pt.apply_host([D], h_in, h_out, chunk_fields=5)[0]
        got = out.numpy().copy()                         # read immediately
        ref = oracle.apply_along_axis(np.asarray(D), U, (32, 65), 1)
        assert not np.isnan(got).any()
        assert np.max(np.abs(got - ref) / oracle.abs_bound(np.asarray(D), U, (32, 65), 1)) <= 1e-12


def test_first_apply_neither_allocates_symmetry_scratch_nor_changes_the_choice():
    """Even-odd eligibility is decided when the operator is built (pt_operator_symmetry writes a
    caller-owned device scalar); apply() finds the packed halves ready."""
    D, _ = _cheb(64, 2)
    assert D._eo not in (None,) and D.even_odd
    import pytristan_b200 as pt

    x = np.linspace(0.0, 1.0, 50) ** 2            # one-sided stretching: not symmetric
    B = pt.ChebMat(x, order=1)
    assert B._eo is False and not B.even_odd


def test_captured_apply_replays_the_same_bits():
    """CUDA-graph replay of operator applications (launch-bound regime): same kernels, same bits, and the
    graph follows in-place refills of its input buffer."""
    import torch

    import pytristan_b200 as pt

    clear_grids()
    try:
        g = pt.get_polar_grid(64, 24, axes=[1], mappers=[pt.cheb], fornberg=True)
        lap = pt.PolarLaplacian(g)
        x = pt.cheb(65, -1.0, 1.0)
        D = pt.ChebMat(x, order=1)
        D.npts, D.axis = (32, 65), 1
        F = pt.FinDiffMat(np.linspace(0.0, 1.0, 32), order=1, accuracy=4)
        F.npts, F.axis = (32, 65), 0
        gen = torch.Generator(device="cuda").manual_seed(3)
        for ops, size in (([lap], 3 * 64 * 48), ([D, F], 5 * 32 * 65)):
            u = torch.randn(size, dtype=torch.float64, device="cuda", generator=gen)
            cap = pt.CapturedApply(ops, u)
            for _ in range(3):
                u.copy_(torch.randn(size, dtype=torch.float64, device="cuda", generator=gen))
                outs = cap.replay()
                torch.cuda.synchronize()
                for op, y in zip(ops, outs):
                    assert torch.equal(y, op.apply(u).reshape(-1))
        with pytest.raises(TypeError):
            pt.CapturedApply([D], np.zeros(32 * 65))
    finally:
        clear_grids()
