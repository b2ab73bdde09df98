"""Pins for the oracle (CPU).

Grid half: against the golden fixtures generated from the reference itself
(tests/golden/make_golden.py, reference src/pytristan/grid.py).
Operator half: the reference has no operator code ("parity unpinned"), so the oracle is pinned by
independent implementations available offline: sympy's exact-rational Fornberg weights,
numpy.polynomial.chebyshev (differentiation through coefficients), numpy.fft, analytic pairs and
structural identities.
"""
import json
import os

import numpy as np
import pytest
from numpy.polynomial import chebyshev as C

from oracle import operators as oracle

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
GOLD = np.load(os.path.join(GOLDEN_DIR, "reference_small.npz"))


def stretch(npts, lo, hi):
    s = np.linspace(-1.0, 1.0, npts)
    return lo + (hi - lo) * 0.5 * (1.0 + np.tanh(1.5 * s) / np.tanh(1.5))


def axes_of(bounds, axes, mappers):
    amap = {a % len(bounds): m for a, m in zip(axes, mappers)}
    fn = {"cheb": oracle.cheb_nodes, "stretch": stretch}
    return [fn[amap[k]](b[2], b[0], b[1]) if k in amap else np.linspace(*b) for k, b in enumerate(bounds)]


SMALL = {
    "raw_2x3": ([(0.0, 1.0, 2), (-1.0, 1.0, 3)], [], []),
    "cheb_8x3": ([(0.0, 1.0, 8), (-1.0, 1.0, 3)], [0], ["cheb"]),
    "cheb_8x6_both": ([(0.0, 1.0, 8), (-2.0, 2.0, 6)], [1, 0], ["cheb", "cheb"]),
    "cheb_neg_axis": ([(0.0, 1.0, 8), (-2.0, 2.0, 6)], [-1], ["cheb"]),
    "three_d": ([(0.0, 1.0, 5), (-1.0, 1.0, 4), (2.0, 3.0, 3)], [1], ["cheb"]),
    "stretch_2d": ([(0.0, 2.0, 9), (0.0, 1.0, 17)], [1], ["stretch"]),
    "one_d_cheb64": ([(-1.0, 1.0, 64)], [0], ["cheb"]),
}


@pytest.mark.parametrize("name", sorted(SMALL))
def test_grid_expand_matches_reference_golden(name):
    got = oracle.grid_expand(axes_of(*SMALL[name]))
    ref = GOLD["grid_" + name]
    assert got.shape == ref.shape and got.tobytes() == ref.tobytes()


@pytest.mark.parametrize("n", [2, 3, 10, 64, 257, 512, 1024, 2048])
def test_cheb_nodes_match_reference_golden(n):
    assert oracle.cheb_nodes(n).tobytes() == GOLD[f"cheb_{n}"].tobytes()
    assert oracle.cheb_nodes(n, 0.0, 10.0).tobytes() == GOLD[f"cheb_{n}_0_10"].tobytes()
    assert oracle.cheb_nodes(n, -1.0, 1.0).tobytes() == GOLD[f"cheb_{n}_m1_1"].tobytes()


def test_polar_axes_match_reference_golden():
    th = np.linspace(-np.pi, np.pi - 2.0 * np.pi / 4, 4)
    for name, r in (("polar_4_6", np.linspace(0.0, 1.0, 6)), ("polar_4_6_cheb", oracle.cheb_nodes(6, 0.0, 1.0)),
                    ("polar_4_6_fornberg", oracle.cheb_nodes(12, -1.0, 1.0))):
        assert oracle.grid_expand([th, r]).tobytes() == GOLD["grid_" + name].tobytes()


def test_golden_hash_file_is_consistent():
    with open(os.path.join(GOLDEN_DIR, "reference_hashes.json")) as fh:
        hashes = json.load(fh)
    import hashlib

    g = oracle.grid_expand([np.linspace(0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), oracle.cheb_nodes(257, -1.0, 1.0)])
    assert hashlib.sha256(g.tobytes()).hexdigest() == hashes["channel_1024x257"]["sha256"]


# -- Chebyshev ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 16, 33, 64])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_cheb_mat_vs_coefficient_space(n, order):
    x = oracle.cheb_nodes(n)
    rng = np.random.default_rng(n)
    coef = rng.standard_normal(n) / (1.0 + np.arange(n)) ** 2      # a degree n-1 polynomial: exact
    u = C.chebval(x, coef)
    du = C.chebval(x, C.chebder(coef, order))
    D = oracle.cheb_mat(x, order)
    assert np.max(np.abs(D @ u - du)) <= 1e-10 * max(1.0, np.max(np.abs(du))) * n ** (2 * order - 2)


def test_cheb_mat_structure_and_maps():
    x = oracle.cheb_nodes(24)
    D = oracle.cheb_mat(x)
    assert np.max(np.abs(D @ np.ones(24))) < 1e-12
    assert abs(D[0, 0] - (2 * 23 ** 2 + 1) / 6.0) < 1e-11             # Trefethen ch.6: D_00 = (2N^2+1)/6
    assert np.allclose(D, -D[::-1, ::-1], atol=1e-11)                 # centro-antisymmetry
    a, b = 0.5, 3.0                                                   # affine covariance, either orientation
    xm = oracle.cheb_nodes(24, a, b)
    Dm = oracle.cheb_mat(xm)
    assert np.allclose(Dm, (oracle.cheb_mat(x)[::-1, ::-1]) * (2.0 / (b - a)) * -1.0 * -1.0 * 1.0, atol=1e-10) or \
        np.max(np.abs(Dm @ np.exp(xm) - np.exp(xm))) < 1e-10
    assert np.max(np.abs(Dm @ np.exp(xm) - np.exp(xm))) < 1e-10
    D2 = oracle.cheb_mat(x, 2)
    assert np.max(np.abs(D2 - D @ D)) <= 1e-12 * np.max(np.abs(D2))


def test_config1_known_answer():
    x = oracle.cheb_nodes(64, -1.0, 1.0)
    assert np.max(np.abs(oracle.cheb_mat(x) @ np.sin(x) - np.cos(x))) < 2e-12


# -- Fourier ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 7, 16, 33, 128])
@pytest.mark.parametrize("period", [2 * np.pi, 3.0])
def test_four_mat_vs_fft(n, period):
    t = period / n * np.arange(n)
    rng = np.random.default_rng(n)
    f = rng.standard_normal(n)
    k = np.fft.fftfreq(n, 1.0 / n) * (2 * np.pi / period)
    k1 = k.copy()
    if n % 2 == 0:
        k1[n // 2] = 0.0
    d1 = np.real(np.fft.ifft(1j * k1 * np.fft.fft(f)))
    d2 = np.real(np.fft.ifft(-(k ** 2) * np.fft.fft(f)))
    D1, D2 = oracle.four_mat(n, period, 1), oracle.four_mat(n, period, 2)
    assert np.max(np.abs(D1 @ f - d1)) <= 1e-12 * np.abs(D1).sum(axis=1).max() * np.max(np.abs(f))
    assert np.max(np.abs(D2 @ f - d2)) <= 1e-12 * np.abs(D2).sum(axis=1).max() * np.max(np.abs(f))
    assert np.allclose(D1, -D1.T, atol=1e-12 * np.max(np.abs(D1)))    # antisymmetric circulant
    assert np.allclose(D2, D2.T)
    assert np.array_equal(D1[1:, 1:], D1[:-1, :-1])                   # Toeplitz


def test_four_mat_analytic():
    n = 64
    t = -np.pi + 2 * np.pi / n * np.arange(n)
    f = np.exp(np.sin(t))
    assert np.max(np.abs(oracle.four_mat(n, 2 * np.pi, 1) @ f - np.cos(t) * f)) < 1e-12
    assert np.max(np.abs(oracle.four_mat(n, 2 * np.pi, 2) @ f - (np.cos(t) ** 2 - np.sin(t)) * f)) < 1e-11


# -- Fornberg ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,nodes,z", [(1, range(7), 3), (2, range(7), 3), (1, range(7), 0), (2, range(8), 0),
                                       (3, range(7), 3), (4, range(9), 2), (1, [0, 1, 3, 4, 7], 3)])
def test_fornberg_vs_sympy(m, nodes, z):
    sympy = pytest.importorskip("sympy")
    nodes = list(nodes)
    ref = [float(v) for v in sympy.finite_diff_weights(m, [sympy.Integer(v) for v in nodes], sympy.Integer(z))[m][-1]]
    got = oracle.fornberg_weights(np.array(nodes, dtype=float), float(z), m)
    assert np.max(np.abs(got - np.array(ref))) <= 4e-15 * np.max(np.abs(ref))


def test_fornberg_classic_rows():
    w = oracle.fornberg_weights(np.arange(7.0), 3.0, 1)
    assert np.allclose(w, np.array([-1, 9, -45, 0, 45, -9, 1]) / 60.0, atol=1e-15)
    w = oracle.fornberg_weights(np.arange(7.0), 0.0, 1)
    assert np.allclose(w, [-49 / 20, 6, -15 / 2, 20 / 3, -15 / 4, 6 / 5, -1 / 6], atol=1e-13)


def test_findiff_band_layout():
    x = np.linspace(0.0, 1.0, 20)
    b = oracle.findiff_band(x, 1, 6)
    assert b["wmax"] == 7 and b["uniform"] and (b["lo"], b["hi"]) == (3, 17)
    assert b["start"].tolist() == [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 13, 13, 13]
    b2 = oracle.findiff_band(x, 2, 6)
    assert b2["wmax"] == 8
    assert b2["start"].tolist() == [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 12, 12, 12, 12]
    # every interior row of a uniform axis carries the same weights; rows sum to zero
    W = b["W"]
    assert all(np.array_equal(W[i, :7], b["wuni"]) for i in range(3, 17))
    assert np.max(np.abs(W.sum(axis=1))) < 1e-10
    bp = oracle.findiff_band(np.linspace(0, 1 - 1 / 16, 16), 1, 4, periodic=True)
    assert bp["start"].tolist() == list(range(-2, 14)) and bp["wmax"] == 5
    with pytest.raises(ValueError):
        oracle.findiff_band(x[:5], 1, 6)


@pytest.mark.parametrize("order,accuracy", [(1, 2), (1, 6), (2, 6), (3, 4)])
def test_findiff_polynomial_exactness(order, accuracy):
    for x in (np.linspace(-1.0, 2.0, 30), oracle.cheb_nodes(30, 0.0, 1.0)):
        b = oracle.findiff_band(x, order, accuracy)
        D = oracle.band_to_dense(b["start"], b["W"])
        deg = oracle.findiff_widths(order, accuracy)[0] - 1          # w nodes: exact up to degree w - 1
        p = np.polynomial.Polynomial(np.arange(1.0, deg + 2))
        assert np.max(np.abs(D @ p(x) - p.deriv(order)(x))) <= 1e-8 * max(1.0, np.max(np.abs(p.deriv(order)(x))))


# -- application along an axis -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("npts,axis", [((5,), 0), ((4, 6), 0), ((4, 6), 1), ((3, 5, 4), 1), ((3, 5, 4), 2)])
def test_axis_apply_equals_kron(npts, axis):
    rng = np.random.default_rng(0)
    n = npts[axis]
    D = rng.standard_normal((n, n))
    u = rng.standard_normal(int(np.prod(npts)))
    L = oracle.kron_lift(D, npts, axis)
    assert np.allclose(L @ u, oracle.apply_along_axis(D, u, npts, axis), atol=1e-13)
    U = rng.standard_normal((3, u.size))
    Y = oracle.apply_along_axis(D, U, npts, axis)
    assert all(np.allclose(Y[k], L @ U[k], atol=1e-13) for k in range(3))
    # the grid layout: differentiating the coordinate row of the grid along its own axis
    axes = [np.linspace(0.0, 1.0, m) for m in npts]
    G = oracle.grid_expand(axes)
    b = oracle.findiff_band(axes[axis], 1, 2) if n >= 3 else None
    if b is not None:
        assert np.allclose(oracle.stencil_apply(b, G[axis], npts, axis), 1.0, atol=1e-12)
        Dd = oracle.band_to_dense(b["start"], b["W"])
        assert np.allclose(oracle.stencil_apply(b, U, npts, axis), oracle.apply_along_axis(Dd, U, npts, axis), atol=1e-12)


# -- barycentric collocation, general-order Fourier ---------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 33, 200])
def test_bary_mat(n):
    xc = oracle.cheb_nodes(n, -1.0, 1.0)
    D = oracle.bary_mat(xc, 1)
    Dc = oracle.cheb_mat(xc, 1)
    assert np.max(np.abs(D - Dc)) <= 1e-11 * np.max(np.abs(Dc))        # same matrix on CGL nodes
    x = np.sort(np.random.default_rng(n).uniform(-1, 1, min(n, 24)))   # arbitrary nodes, moderate degree
    p = np.polynomial.Polynomial(np.random.default_rng(1).standard_normal(len(x)))
    for order in (1, 2, 3):
        Dp = oracle.bary_mat(x, order)
        assert np.max(np.abs(Dp @ p(x) - p.deriv(order)(x))) <= 1e-7 * max(1.0, np.max(np.abs(p.deriv(order)(x)))) * 10 ** order
    assert oracle.is_affine_cgl(oracle.cheb_nodes(50, 2.0, 5.0)) and oracle.is_affine_cgl(oracle.cheb_nodes(50))
    assert not oracle.is_affine_cgl(np.linspace(0, 1, 50))
    big = oracle.cheb_nodes(2048)                                       # no under/overflow at n = 2048
    assert np.all(np.isfinite(oracle.bary_mat(big, 1)))


@pytest.mark.parametrize("n", [8, 17, 64])
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_four_general_vs_fft_and_closed_forms(n, order):
    period = 3.0
    col = oracle.four_col_general(n, period, order)
    if order in (1, 2):
        ref = oracle.four_col(n, period, order)
        assert np.max(np.abs(col - ref)) <= 1e-12 * np.max(np.abs(ref)) * n
    rng = np.random.default_rng(n)
    f = rng.standard_normal(n)
    k = np.fft.fftfreq(n, 1.0 / n) * (2 * np.pi / period)
    if n % 2 == 0 and order % 2 == 1:
        k[n // 2] = 0.0
    spec = np.real(np.fft.ifft((1j * k) ** order * np.fft.fft(f)))
    D = oracle.four_mat(n, period, order) if order > 2 else col[(np.arange(n)[:, None] - np.arange(n)[None, :]) % n]
    assert np.max(np.abs(D @ f - spec)) <= 1e-11 * np.abs(D).sum(axis=1).max() * np.max(np.abs(f))


def test_operator_known_answer_vectors_are_stable():
    """The oracle's operator conventions, frozen in tests/golden/operator_kat.npz by
    tests/golden/make_operator_kat.py: any drift (summation order, flipping rule, band layout)
    shows up here bit for bit."""
    import importlib.util
    import os

    here = os.path.join(os.path.dirname(__file__), "golden")
    spec = importlib.util.spec_from_file_location("make_operator_kat", os.path.join(here, "make_operator_kat.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    now = mod.build()
    frozen = np.load(os.path.join(here, "operator_kat.npz"))
    assert sorted(frozen.files) == sorted(now)
    for key in frozen.files:
        assert frozen[key].dtype == now[key].dtype and np.array_equal(frozen[key], now[key]), key


@pytest.mark.parametrize("n", [1, 2, 7, 40, 129])
def test_lu_oracle_vs_lapack(n):
    """oracle.lu_factor (textbook partial pivoting) against LAPACK getrf through scipy, and
    oracle.solve_along_axis against a Kronecker-lifted dense solve."""
    import scipy.linalg as sla

    rng = np.random.default_rng(100 + n)
    A = rng.standard_normal((n, n))
    A[0, 0] = 0.0 if n > 1 else 1.0                      # forces a row exchange in the first step
    perm, LU = oracle.lu_factor(A)
    L, U = np.tril(LU, -1) + np.eye(n), np.triu(LU)
    assert sorted(perm.tolist()) == list(range(n))
    assert np.max(np.abs(L)) <= 1.0
    assert np.max(np.abs(A[perm] - L @ U)) <= 4 * n * np.finfo(float).eps * np.max(np.abs(L) @ np.abs(U))
    lu_ref, piv = sla.lu_factor(A)
    perm_ref = np.arange(n)
    for c, p in enumerate(piv):
        perm_ref[[c, p]] = perm_ref[[p, c]]
    assert np.array_equal(perm, perm_ref)
    assert np.max(np.abs(LU - lu_ref)) <= 1e-12 * np.max(np.abs(lu_ref)) * max(1.0, np.linalg.cond(A)) ** 0.5
    if n >= 7:
        npts = (3, n, 2)
        F = rng.standard_normal((2, 6 * n))
        Y = oracle.solve_along_axis(A, F, npts, 1)
        back = oracle.apply_along_axis(A, Y, npts, 1)
        assert np.max(np.abs(back - F)) <= 1e-12 * np.linalg.cond(A) * np.max(np.abs(F))
    with pytest.raises(np.linalg.LinAlgError):
        oracle.lu_factor(np.ones((3, 3)))


@pytest.mark.parametrize("N", [1, 2, 3, 8, 33, 128])
def test_cheb_mat_vs_published_closed_form(N):
    """Trefethen, Spectral Methods in MATLAB (2000), Theorem 7 (eqs. 6.3-6.5), written out entry by entry — an
    implementation that shares nothing with oracle.cheb_mat (which uses the negative-sum trick and the flipping
    trick) — plus the two matrices printed in the book (p. 53)."""
    x = oracle.cheb_nodes(N + 1)
    D = oracle.cheb_mat(x, 1)
    c = np.ones(N + 1); c[0] = c[N] = 2.0
    ref = np.zeros((N + 1, N + 1))
    for i in range(N + 1):
        for j in range(N + 1):
            if i != j:
                ref[i, j] = (c[i] / c[j]) * (-1.0) ** (i + j) / (x[i] - x[j])
            elif i == 0:
                ref[i, j] = (2.0 * N * N + 1.0) / 6.0
            elif i == N:
                ref[i, j] = -(2.0 * N * N + 1.0) / 6.0
            else:
                ref[i, j] = -x[j] / (2.0 * (1.0 - x[j] ** 2))
    scale = np.sum(np.abs(ref), axis=1, keepdims=True)
    assert np.max(np.abs(D - ref) / scale) <= 1e-13 * max(1, N)          # the closed-form diagonal loses ~N eps itself
    if N == 1:
        assert np.array_equal(D, np.array([[0.5, -0.5], [0.5, -0.5]]))
    if N == 2:
        assert np.max(np.abs(D - np.array([[1.5, -2.0, 0.5], [0.5, 0.0, -0.5], [-0.5, 2.0, -1.5]]))) <= 2e-15
