"""Freeze known-answer vectors of the OPERATOR half of the oracle (oracle/operators.py).

The reference ships no operator code, so these are not reference outputs: they pin the oracle's own
conventions (flipping trick, Fourier antipodal zero, Fornberg row-start layout, sequential sums)
against drift between rounds, and they are what the GPU generators are compared with bit for bit
(off-diagonal entries) in tests/test_gpu_golden.py.  Independent pins of the same functions
(sympy.finite_diff_weights, numpy.polynomial.chebyshev, numpy.fft, analytic derivatives) live in
tests/test_oracle.py.

    python tests/golden/make_operator_kat.py      # writes tests/golden/operator_kat.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import operators as o  # noqa: E402


def build():
    out = {}
    for n in (8, 9, 33):
        x = o.cheb_nodes(n, -1.0, 1.0)
        out[f"cheb_nodes_{n}"] = x
        for order in (1, 2, 3):
            out[f"cheb_mat_{n}_o{order}"] = o.cheb_mat(x, order)
    xd = o.cheb_nodes(12)                        # the reference's default orientation (+1 -> -1)
    out["cheb_mat_desc12_o1"] = o.cheb_mat(xd, 1)
    c = o.cheb_nodes(17)[::-1]
    out["bary_nodes_17"] = c + 0.15 * (1.0 - c * c)
    out["bary_mat_17_o1"] = o.bary_mat(out["bary_nodes_17"], 1)
    out["bary_mat_17_o2"] = o.bary_mat(out["bary_nodes_17"], 2)
    for n in (8, 9, 32):
        for order in (1, 2, 3, 4):
            out[f"four_mat_{n}_o{order}"] = o.four_mat(n, 2 * np.pi, order)
    out["four_mat_16_period3_o1"] = o.four_mat(16, 3.0, 1)
    xu = np.linspace(0.0, 1.0, 21)
    xm = o.cheb_nodes(21, 0.0, 1.0)
    for tag, x, periodic in (("uni", xu, False), ("map", xm, False), ("per", np.linspace(0.0, 1.0 - 1.0 / 20, 20), True)):
        for (order, acc) in ((1, 2), (1, 6), (2, 4)):
            band = o.findiff_band(x, order, acc, periodic)
            out[f"fd_{tag}_o{order}a{acc}_start"] = np.asarray(band["start"], dtype=np.int64)
            out[f"fd_{tag}_o{order}a{acc}_W"] = np.asarray(band["W"], dtype=np.float64)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "operator_kat.npz"), **build())
    print("wrote", len(build()), "arrays")
