"""GPU: the product's grids against the golden fixtures generated from the reference
(tests/golden/make_golden.py) — bit-exact — and the operator registry on real operators."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("clean_registries")]

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


def stretch(npts, lo, hi):
    s = np.linspace(-1.0, 1.0, npts)
    return lo + (hi - lo) * 0.5 * (1.0 + np.tanh(1.5 * s) / np.tanh(1.5))


def _build(pt, bounds, axes, mappers):
    fn = {"cheb": pt.cheb, "stretch": stretch}
    return pt.Grid.from_bounds(*bounds, axes=axes, mappers=[fn[m] for m in mappers])


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
def test_small_grids_bit_exact(name):
    import pytristan_b200 as pt

    gold = np.load(os.path.join(GOLDEN_DIR, "reference_small.npz"))
    g = _build(pt, *SMALL[name])
    ref = gold["grid_" + name]
    assert g.shape == ref.shape and g.dtype == ref.dtype and np.asarray(g).tobytes() == ref.tobytes()


def test_polar_grids_bit_exact():
    import pytristan_b200 as pt

    gold = np.load(os.path.join(GOLDEN_DIR, "reference_small.npz"))
    g1 = pt.get_polar_grid(4, 6)
    g2 = pt.get_polar_grid(4, 6, axes=[1], mappers=[pt.cheb])
    g3 = pt.get_polar_grid(4, 6, axes=[1], mappers=[pt.cheb], fornberg=True)
    for g, name in ((g1, "polar_4_6"), (g2, "polar_4_6_cheb"), (g3, "polar_4_6_fornberg")):
        assert np.asarray(g).tobytes() == gold["grid_" + name].tobytes()
    assert [g.num for g in (g1, g2, g3)] == [0, 1, 2]


def test_large_grids_hash():
    """BASELINE config 2 and 3 grids, compared through sha256 of the reference's bytes."""
    import pytristan_b200 as pt

    with open(os.path.join(GOLDEN_DIR, "reference_hashes.json")) as fh:
        hashes = json.load(fh)
    g = pt.Grid.from_bounds((0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), (-1.0, 1.0, 257), axes=[1], mappers=[pt.cheb])
    assert hashlib.sha256(np.asarray(g).tobytes()).hexdigest() == hashes["channel_1024x257"]["sha256"]
    g = pt.Grid.from_bounds((0.0, 1.0, 48), (0.0, 2.0, 48), (-1.0, 1.0, 48), axes=[2], mappers=[pt.cheb])
    assert hashlib.sha256(np.asarray(g).tobytes()).hexdigest() == hashes["cube_48"]["sha256"]
    for nt, nr, fb, key in ((1024, 512, False, "polar_1024_512_n"), (1024, 256, True, "polar_1024_256_f")):
        g = pt.get_polar_grid(nt, nr, axes=[1], mappers=[pt.cheb], fornberg=fb)
        assert list(g.shape) == hashes[key]["shape"]
        assert hashlib.sha256(np.asarray(g).tobytes()).hexdigest() == hashes[key]["sha256"]
        assert hashlib.sha256(g.axpoints(1).tobytes()).hexdigest() == hashes[key]["r_sha256"]


def test_pinned_readback_path():
    import pytristan_b200 as pt
    from oracle import operators as oracle

    axes = [np.linspace(0.0, 1.0, 160), np.linspace(0.0, 1.0, 150), np.linspace(-1.0, 0.0, 40)]   # 23 MB
    g = pt.Grid.from_arrs(*axes)
    assert np.asarray(g).tobytes() == oracle.grid_expand(axes).tobytes()


def test_operator_registry():
    import pytristan_b200 as pt

    g0 = pt.get_grid((0.0, 1.0, 12), (-1.0, 1.0, 9), from_bounds=True, axes=[1], mappers=[pt.cheb])
    g1 = pt.get_grid((0.0, 2 * np.pi - 2 * np.pi / 16, 16), (0.0, 1.0, 10), from_bounds=True)
    a = pt.get_cheb_mat(g0, axis=1, order=1)
    assert pt.get_cheb_mat(0, axis=1, order=1) is a and pt.get_cheb_mat(g0, axis=-1) is a
    assert a.key == [0, 1, 1] and isinstance(a, np.ndarray) and isinstance(a, pt.ChebMat)
    b = pt.get_cheb_mat(g0, axis=1, order=2)
    assert b is not a and b.order == 2
    f = pt.get_four_mat(g1, axis=0, order=1)
    fd = pt.get_findiff_mat(g1, axis=1, order=1, accuracy=4)
    assert pt._get_mat_manager("cheb").nums() == [[0, 1, 1], [0, 1, 2]]
    assert pt._get_mat_manager("four").nums() == [[1, 0, 1]]
    assert pt._get_mat_manager("findiff").nums() == [[1, 1, 1, 4]]
    a2 = pt.get_cheb_mat(g0, axis=1, order=1, overwrite=True)
    assert a2 is not a and np.array_equal(a2, a)
    pt.drop_cheb_mat(key=[0, 1, 2])
    assert pt._get_mat_manager("cheb").nums() == [[0, 1, 1]]
    pt.drop_cheb_mat(nitem=1)
    assert pt._get_mat_manager("cheb").nums() == []
    pt.drop_all_mats()
    assert pt._get_mat_manager("four").nums() == [] and pt._get_mat_manager("findiff").nums() == []
    # operators are NumPy arrays: ufuncs and slicing work, results are plain matrices
    s = a + 2.0 * b
    assert type(s) is np.ndarray and s.shape == (9, 9)
    v = a[:3]
    assert isinstance(v, pt.ChebMat) and v.npts is None
    u = np.linspace(0.0, 1.0, 12 * 9)
    assert np.array_equal(np.matmul(a, u), a.apply(u)) and np.array_equal(a @ u, a.apply(u))
    assert (a @ np.eye(9)).shape == (9, 9)
    assert np.max(np.abs(a @ np.eye(9) - np.asarray(a))) <= 1e-13 * np.max(np.abs(a))
    assert f.period == pytest.approx(2 * np.pi) and fd.accuracy == 4
