"""CPU ORACLE (test infrastructure, not product code) for the operator half of the hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this
module; nothing under ``pytristan_b200/`` does.

PARITY UNPINNED BY THE REFERENCE: the mounted reference snapshot (pytristan 0.2a) contains no
operator code — ``ChebMat``/``FourMat``/``FinDiffMat`` exist only as prose (reference
README.md:7-10,16).  This file is therefore a NumPy restatement of the textbook definitions the
reference names and cites (README.md:8-10; grid.py:462-464 cites Fornberg 1995 and Trefethen
2000), on the reference's own grid layout (grid.py:85-88) and node conventions (grid.py:75-77,
:505-506).  It is pinned instead by independent checks in tests/test_oracle.py:
``sympy.finite_diff_weights`` (exact rational Fornberg weights), ``numpy.polynomial.chebyshev``
(differentiation through coefficients), ``numpy.fft`` (ik multiplication) and analytic
derivative pairs.

Conventions: a field on a grid with ``npts = (n0, .., n_{d-1})`` is flat, axis 0 fastest.  For
axis k: ``before = prod(npts[:k])``, ``after = prod(npts[k+1:]) * batch``; element (a, i, b) is at
``(a*n + i)*before + b``.
"""
import numpy as np

# --------------------------------------------------------------------------------------------
# nodes (reference grid.py:75-77) and the Grid layout (grid.py:83-88)
# --------------------------------------------------------------------------------------------


def cheb_nodes(n, xmin=None, xmax=None):
    """CGL nodes with the reference's operation order (grid.py:75-77)."""
    x = np.cos(np.arange(n) * np.pi / (n - 1))
    if xmin is not None or xmax is not None:
        x = (1.0 - x) / 2.0 * (xmax - xmin) + xmin
    return x


def grid_expand(axes):
    """Axis arrays -> (ndim, P) array, axis 0 fastest: row d repeats each node of axis d
    ``prod(npts[:d])`` times and tiles that ``prod(npts[d+1:])`` times.  Equivalent to the
    reference's meshgrid(indexing='ij') + flatten(order='F') (grid.py:83-88); equality with
    the imported reference is checked in tests/test_reference_parity.py."""
    axes = [np.asarray(a).ravel() for a in axes]
    if not axes:
        return np.empty((0,))
    dtype = np.result_type(*axes)
    npts = [len(a) for a in axes]
    total = int(np.prod(npts, dtype=np.int64))
    out = np.empty((len(axes), total), dtype=dtype)
    before = 1
    for d, a in enumerate(axes):
        after = total // (before * npts[d]) if npts[d] else 0
        out[d] = np.tile(np.repeat(a.astype(dtype), before), after) if total else []
        before *= npts[d]
    return out


def axis_split(npts, axis, size):
    """(after, n, before) of a flat field of ``size`` elements (a whole number of fields)."""
    npts = tuple(int(n) for n in npts)
    total = int(np.prod(npts, dtype=np.int64))
    if total == 0 or size % total:
        raise ValueError("field size is not a multiple of the grid size")
    before = int(np.prod(npts[:axis], dtype=np.int64))
    after = int(np.prod(npts[axis + 1:], dtype=np.int64)) * (size // total)
    return after, npts[axis], before


# --------------------------------------------------------------------------------------------
# Chebyshev collocation (Trefethen 2000 ch. 6; Welfert 1997 recursion for higher orders)
# --------------------------------------------------------------------------------------------


def _seq_row_sum(M):
    """Left-to-right sequential sum of every row (the order the device kernel uses)."""
    return np.cumsum(M, axis=1)[:, -1]


def nodes_symmetric(x):
    """True when the nodes are symmetric about their midpoint to within 8 ulp of their span (every
    affine CGL set; stretched meshes whose mapper is odd about the centre)."""
    x = np.asarray(x, dtype=np.float64)
    if x.size < 2:
        return False
    c2 = x[0] + x[-1]
    scale = max(float(np.max(np.abs(x))), abs(float(x[-1] - x[0])))
    return bool(np.max(np.abs(x + x[::-1] - c2)) <= 8.0 * np.finfo(np.float64).eps * scale)


def flip_symmetrize(D, sign):
    """Flipping trick (Don & Solomonoff 1995): rows below the middle, and the right part of the
    middle row of an odd n, become the mirror image of the rows above, D[i,j] = sign*D[n-1-i,n-1-j]
    (same rule as pt_flip_symmetrize)."""
    n = D.shape[0]
    D = D.copy()
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    mi, mj = n - 1 - i, n - 1 - j
    sel = (i > mi) | ((i == mi) & (j > mj))
    D[sel] = sign * D[mi[sel], mj[sel]]
    if sign < 0 and n % 2 == 1:
        D[n // 2, n // 2] = 0.0        # the centre of a centro-antisymmetric matrix is its own negative
    return D


def cheb_mat(x, order=1, flip=None):
    """Collocation differentiation matrix of ``order`` on nodes ``x`` (any affine image of the
    CGL set, either orientation).  With ``flip`` (default, what ChebMat generates) the lower half
    is the mirror image of the upper half, so the matrix is exactly centro(anti)symmetric.

    D1_ij = (c_i/c_j)/(x_i-x_j), c_0 = c_{n-1} = 2 else 1, c_i <- c_i(-1)^i;
    D^(l)_ij = l/(x_i-x_j) * ((c_i/c_j) D^(l-1)_ii - D^(l-1)_ij);
    every diagonal by the negative-sum trick, summed left to right.
    One rounding per operation (no FMA), the same order as pt_gen_chebmat.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    c = np.ones(n)
    c[0] = c[-1] = 2.0
    c[1::2] *= -1.0
    off = ~np.eye(n, dtype=bool)
    cr = c[:, None] / c[None, :]
    dx = np.where(off, x[:, None] - x[None, :], 1.0)
    D = np.where(off, cr / dx, 0.0)
    D[~off] = -_seq_row_sum(D)
    for l in range(2, order + 1):
        diag = np.diag(D).copy()
        Dn = np.where(off, (l / dx) * (cr * diag[:, None] - D), 0.0)
        Dn[~off] = -_seq_row_sum(Dn)
        D = Dn
    if flip is None:
        flip = nodes_symmetric(x)       # every affine CGL set is symmetric about its midpoint
    return flip_symmetrize(D, 1 if order % 2 == 0 else -1) if flip else D


def bary_mat(x, order=1, flip=None):
    """Polynomial collocation differentiation matrix on ARBITRARY distinct nodes (what a
    non-affine mapper of the reference produces, grid.py:201-207): barycentric weights
    w_j = 1/prod_{k!=j}(x_j-x_k) (Berrut & Trefethen 2004), D1_ij = (w_j/w_i)/(x_i-x_j), higher
    orders by the recursion of ``cheb_mat`` with c_i/c_j replaced by w_j/w_i.  The products are
    kept as mantissa*2^exponent (frexp after every multiply) exactly as pt_gen_barymat does."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    mant = np.ones(n)
    expo = np.zeros(n, dtype=np.int64)
    for k in range(n):
        f = x - x[k]
        f[k] = 1.0
        mant, ex = np.frexp(mant * f)
        expo += ex
    off = ~np.eye(n, dtype=bool)
    cr = np.ldexp(mant[:, None] / mant[None, :], (expo[:, None] - expo[None, :]).astype(np.int32))
    dx = np.where(off, x[:, None] - x[None, :], 1.0)
    D = np.where(off, cr / dx, 0.0)
    D[~off] = -_seq_row_sum(D)
    for l in range(2, order + 1):
        diag = np.diag(D).copy()
        Dn = np.where(off, (l / dx) * (cr * diag[:, None] - D), 0.0)
        Dn[~off] = -_seq_row_sum(Dn)
        D = Dn
    if flip is None:
        flip = nodes_symmetric(x)       # same rule as ChebMat: symmetric nodes -> flipping trick
    return flip_symmetrize(D, 1 if order % 2 == 0 else -1) if flip else D


def is_affine_cgl(x, tol=1e-13):
    """True when ``x`` is an affine image (either orientation) of the CGL set."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 2 or x[-1] == x[0]:
        return False
    t = (x - x[0]) / (x[-1] - x[0])
    ref = (1.0 - np.cos(np.arange(n) * np.pi / (n - 1))) / 2.0
    return bool(np.max(np.abs(t - ref)) <= tol)


# --------------------------------------------------------------------------------------------
# Fourier spectral (Trefethen 2000 ch. 3; Weideman & Reddy 2000 `fourdif`)
# --------------------------------------------------------------------------------------------


def four_col(n, period, order):
    """First column of the circulant Fourier differentiation matrix: col[k] is the entry with
    (i - j) mod n == k.  Entries with k > n/2 come from the (anti)symmetry of the column so that
    the trigonometric argument never exceeds pi/2 (same folding as pt_gen_fourmat).

    With s = 2*pi/period, theta = k*pi/n:
      n even: D1 = s/2 (-1)^k cot(theta),  D2 = -s^2 (-1)^k / (2 sin^2 theta), D2_0 = -s^2 (n^2/12 + 1/6)
      n odd : D1 = s/2 (-1)^k / sin(theta), D2 = -s^2/2 (-1)^k cos(theta)/sin^2(theta),
              D2_0 = -s^2 (n^2 - 1)/12
    """
    if order not in (1, 2):
        raise ValueError("four_col: order must be 1 or 2")
    s = (2.0 * np.pi) / period
    col = np.zeros(n)
    k = np.arange(1, n)
    fold = 2 * k > n
    kk = np.where(fold, n - k, k)
    theta = kk * np.pi / n
    sign = np.where(kk % 2 == 1, -1.0, 1.0)
    sn, cs = np.sin(theta), np.cos(theta)
    even = n % 2 == 0
    if order == 1:
        v = ((s * 0.5) * sign) * (cs / sn) if even else ((s * 0.5) * sign) / sn
        col[k] = np.where(fold, -v, v)
        if even:
            col[n // 2] = 0.0          # cot(pi/2) = 0 exactly (the rounded pi/2 would leave 6e-17)
    else:
        s2 = s * s
        sn2 = sn * sn
        col[k] = -((s2 * sign) / (2.0 * sn2)) if even else -((((s2 * 0.5) * sign) * cs) / sn2)
        nn = float(n) * float(n)
        col[0] = -(s2 * (nn / 12.0 + 1.0 / 6.0)) if even else -(s2 * ((nn - 1.0) / 12.0))
    return col


def four_col_general(n, period, order):
    """First column for any order: (1/n) sum_k (i k s)^p exp(2 pi i k m / n) over the retained
    modes |k| <= (n-1)/2, plus the Nyquist mode k = n/2 for even n and even order (it is dropped
    for odd orders, as in the order-1 closed form).  Sequential summation in k."""
    s = (2.0 * np.pi) / period
    m = np.arange(n)
    K = (n - 1) // 2
    even_p = order % 2 == 0
    sgn = (-1.0 if (order // 2) % 2 else 1.0) if even_p else (1.0 if ((order - 1) // 2) % 2 else -1.0)
    acc = np.zeros(n)
    for k in range(1, K + 1):
        ks = k * s
        pw = 1.0
        for _ in range(order):
            pw = pw * ks
        frac = ((2 * k * m) % (2 * n)) / float(n)
        trig = np.cos(np.pi * frac) if even_p else np.sin(np.pi * frac)
        acc = acc + ((2.0 * sgn) * pw) * trig
    if n % 2 == 0 and even_p:
        ks = (n // 2) * s
        pw = 1.0
        for _ in range(order):
            pw = pw * ks
        acc = acc + (sgn * pw) * np.where(m % 2 == 1, -1.0, 1.0)
    return acc / float(n)


def four_mat(n, period, order=1):
    """Dense circulant matrix D[i, j] = col[(i - j) mod n] (closed forms for orders 1 and 2,
    direct summation of the symbol for higher orders)."""
    col = four_col(n, period, order) if order in (1, 2) else four_col_general(n, period, order)
    i = np.arange(n)
    return col[(i[:, None] - i[None, :]) % n]


# --------------------------------------------------------------------------------------------
# Finite differences (Fornberg 1988)
# --------------------------------------------------------------------------------------------


def fornberg_weights(xs, z, m):
    """Weights of the m-th derivative at ``z`` on nodes ``xs`` (Fornberg, Math. Comp. 51, 1988).
    Scalar loops, one rounding per operation — the order pt_gen_fornberg uses."""
    xs = [float(v) for v in xs]
    N = len(xs)
    C = np.zeros((N, m + 1))
    C[0, 0] = 1.0
    c1 = 1.0
    c4 = xs[0] - z
    for i in range(1, N):
        mn = min(i, m)
        c2 = 1.0
        c5 = c4
        c4 = xs[i] - z
        for j in range(i):
            c3 = xs[i] - xs[j]
            c2 = c2 * c3
            if j == i - 1:
                for k in range(mn, 0, -1):
                    C[i, k] = (c1 * (k * C[i - 1, k - 1] - c5 * C[i - 1, k])) / c2
                C[i, 0] = ((-c1 * c5) * C[i - 1, 0]) / c2
            for k in range(mn, 0, -1):
                C[j, k] = (c4 * C[j, k] - k * C[j, k - 1]) / c3
            C[j, 0] = (c4 * C[j, 0]) / c3
        c1 = c2
    return C[:, m].copy()


def is_uniform(x):
    """Equispaced test used by the operator layer: every spacing within 8 ulp of the mean."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 3:
        return True
    h = (x[-1] - x[0]) / (n - 1)
    tol = 8.0 * np.finfo(np.float64).eps * max(np.max(np.abs(x)), abs(h))
    return bool(np.max(np.abs(np.diff(x) - h)) <= tol)


def findiff_widths(order, accuracy):
    w = 2 * ((order + 1) // 2) - 1 + accuracy
    wb = order + accuracy
    return w, wb


def findiff_band(x, order=1, accuracy=2, periodic=False, uniform=None):
    """Banded row-start form of the finite-difference matrix on nodes ``x``.

    Returns dict(start int32[n], W float64[n, wmax], wmax, w, lo, hi, wuni, h, uniform):
    interior rows use ``w = 2*floor((order+1)/2) - 1 + accuracy`` centred nodes, rows within
    ``w//2`` of an end use ``order + accuracy`` one-sided nodes (non-periodic), and
    ``start[i] = clip(i - w//2, 0, n - wmax)`` (periodic: ``i - w//2``, wrapped on use).
    Uniform axes get integer-offset weights scaled by ``h**-order`` (identical interior rows);
    rows ``lo <= i < hi`` are those canonical rows and ``wuni`` their weights.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if accuracy % 2:
        raise ValueError("accuracy must be even")
    w, wb = findiff_widths(order, accuracy)
    hw = w // 2
    wmax = w if periodic else max(w, wb)
    if n < wmax:
        raise ValueError("axis shorter than the stencil")
    if uniform is None:
        uniform = is_uniform(x)
    h = (x[-1] - x[0]) / (n - 1)
    hm = h
    for _ in range(1, order):
        hm = hm * h
    period = n * h
    start = np.empty(n, dtype=np.int32)
    W = np.zeros((n, wmax))
    for i in range(n):
        if periodic:
            s0, first, cnt = i - hw, i - hw, w
        else:
            s0 = min(max(i - hw, 0), n - wmax)
            if i - hw < 0:
                first, cnt = 0, wb
            elif i + hw > n - 1:
                first, cnt = n - wb, wb
            else:
                first, cnt = i - hw, w
        if uniform:
            wt = fornberg_weights(np.arange(cnt, dtype=np.float64), float(i - first), order) / hm
        else:
            idx = first + np.arange(cnt)
            shift = np.where(idx < 0, -period, np.where(idx >= n, period, 0.0))
            wt = fornberg_weights(x[idx % n] + shift, x[i], order)
        start[i] = s0
        W[i, first - s0:first - s0 + cnt] = wt
    if uniform:
        lo, hi = (0, n) if periodic else (hw, n - hw)
        wuni = fornberg_weights(np.arange(w, dtype=np.float64), float(hw), order) / hm
    else:
        lo = hi = 0
        wuni = None
    return dict(start=start, W=W, wmax=wmax, w=w, lo=lo, hi=hi, wuni=wuni, h=h, uniform=uniform,
                periodic=periodic)


def band_to_dense(start, W, periodic=False):
    n, wmax = W.shape
    D = np.zeros((n, n))
    for i in range(n):
        for s in range(wmax):
            j = int(start[i]) + s
            if periodic:
                j %= n
            D[i, j] += W[i, s]
    return D


# --------------------------------------------------------------------------------------------
# application along one axis; Kronecker lifting
# --------------------------------------------------------------------------------------------


def apply_along_axis(D, U, npts, axis):
    """Y[a, :, b] = D @ U[a, :, b] on the (after, n, before) view of the flat field(s) U."""
    U = np.asarray(U, dtype=np.float64)
    after, n, before = axis_split(npts, axis, U.size)
    V = U.reshape(after, n, before)
    if before == 1:
        Y = V[:, :, 0] @ D.T
        return Y.reshape(U.shape[:-1] + (-1,)) if D.shape[0] != n else Y.reshape(U.shape)
    Y = np.matmul(D, V)
    return Y.reshape(U.shape) if D.shape[0] == n else Y


def stencil_apply(band, U, npts, axis):
    """Banded application, sum of products in slot order s = 0..wmax-1."""
    U = np.asarray(U, dtype=np.float64)
    after, n, before = axis_split(npts, axis, U.size)
    V = U.reshape(after, n, before)
    start, W = band["start"], band["W"]
    Y = np.zeros_like(V)
    rows = np.arange(n)
    for s in range(band["wmax"]):
        j = start.astype(np.int64) + s
        if band["periodic"]:
            j %= n
        Y += W[rows, s][None, :, None] * V[:, j, :]
    return Y.reshape(U.shape)


def kron_lift(D, npts, axis):
    """Explicit kron(I_after, kron(D, I_before)) for one field (small grids only)."""
    after, n, before = axis_split(npts, axis, int(np.prod(npts)))
    return np.kron(np.eye(after), np.kron(D, np.eye(before)))


def abs_bound(D, U, npts, axis):
    """(|D||u|) per output element: the denominator of the componentwise parity metric."""
    return apply_along_axis(np.abs(D), np.abs(U), npts, axis)


def polar_laplacian(theta, r, U):
    """(D2_r + R^-1 D_r) u + R^-2 D2_theta u on a polar grid (theta fastest), batch allowed."""
    nt, nr = len(theta), len(r)
    h = (theta[-1] - theta[0]) / (nt - 1)
    D2t = four_mat(nt, nt * h, 2)
    Ar = cheb_mat(r, 2) + cheb_mat(r, 1) / r[:, None]
    if np.max(np.abs(r + r[::-1])) <= 8.0 * np.finfo(np.float64).eps * np.max(np.abs(r)):
        Ar = flip_symmetrize(Ar, 1)       # radial nodes symmetric about 0: same rule as PolarLaplacian
    U = np.asarray(U, dtype=np.float64)
    Y = apply_along_axis(Ar, U, (nt, nr), 1)
    T = apply_along_axis(D2t, U, (nt, nr), 0)
    scale = np.tile(np.repeat(1.0 / (r * r), nt), U.size // (nt * nr)).reshape(U.shape)
    return Y + scale * T


# ---- solves along an axis (no reference code exists: textbook LU, Golub & Van Loan
# Alg. 3.4.1, pinned in tests/test_oracle.py against scipy.linalg.lu_factor / numpy.linalg.solve) ------------
def lu_factor(A):
    """P A = L U by Gaussian elimination with partial pivoting (first row of maximal modulus).  Returns
    (perm, LU): A[perm] = L @ U with L = tril(LU, -1) + I and U = triu(LU)."""
    LU = np.array(A, dtype=np.float64)
    n = LU.shape[0]
    perm = np.arange(n)
    for c in range(n):
        p = c + int(np.argmax(np.abs(LU[c:, c])))
        if LU[p, c] == 0.0:
            raise np.linalg.LinAlgError(f"singular matrix (elimination step {c})")
        if p != c:
            LU[[c, p]] = LU[[p, c]]
            perm[[c, p]] = perm[[p, c]]
        LU[c + 1:, c] /= LU[c, c]
        LU[c + 1:, c + 1:] -= np.outer(LU[c + 1:, c], LU[c, c + 1:])
    return perm, LU


def solve_along_axis(A, F, npts, axis):
    """Y with A @ Y[a, :, b] = F[a, :, b] on the (after, n, before) view of the flat field(s) F
    (numpy.linalg.solve = LAPACK gesv on all grid lines at once)."""
    F = np.asarray(F, dtype=np.float64)
    after, n, before = axis_split(npts, axis, F.size)
    V = F.reshape(after, n, before).transpose(1, 0, 2).reshape(n, -1)
    Y = np.linalg.solve(np.asarray(A, dtype=np.float64), V)
    return Y.reshape(n, after, before).transpose(1, 0, 2).reshape(F.shape)
