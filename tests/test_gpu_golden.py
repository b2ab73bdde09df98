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


def test_device_grid_matches_host_grid():
    import pytristan_b200 as pt

    bounds = [(0.0, 2 * np.pi - 2 * np.pi / 64, 64), (-1.0, 1.0, 33), (0.0, 1.0, 5)]
    host = pt.Grid.from_bounds(*bounds, axes=[1], mappers=[pt.cheb])
    dev = pt.DeviceGrid.from_bounds(*bounds, axes=[1], mappers=[pt.cheb])
    assert dev.npts == host.npts and dev.num_dim == host.num_dim and dev.shape == host.shape
    assert dev.tensor.is_cuda and dev.tensor.cpu().numpy().tobytes() == np.asarray(host).tobytes()
    for k in range(3):
        assert np.array_equal(dev.axpoints(k), host.axpoints(k))
    assert np.array_equal(np.asarray(dev.to_host()), np.asarray(host))
    # registry + operators accept it
    g = pt.get_grid(*bounds, from_bounds=True, axes=[1], mappers=[pt.cheb], device=True)
    assert isinstance(g, pt.DeviceGrid) and g.num == 0 and pt.get_grid(num=0) is g
    D = pt.get_cheb_mat(g, axis=1, order=1)
    u = np.random.default_rng(0).standard_normal(64 * 33 * 5)
    Dh = pt.ChebMat(host, axis=1, order=1)
    assert np.array_equal(D.apply(u), Dh.apply(u))
    with pytest.raises(IndexError):
        g.axpoints(3)


@pytest.mark.parametrize("case", ["channel", "polar_fornberg", "cheb3d"])
def test_device_grid_nodes_are_generated_on_the_device_bit_for_bit(case):
    """DeviceGrid.from_bounds builds its 1-D nodes on the GPU (pt_gen_linspace, pt_gen_cheb_libm)
    and reads them back through pt_grid_axpoints: same bits as the host Grid (= the reference's)."""
    import pytristan_b200 as pt

    if case == "channel":
        bounds, axes = [(0.0, 2 * np.pi - 2 * np.pi / 1024, 1024), (-1.0, 1.0, 257)], [1]
    elif case == "polar_fornberg":
        bounds, axes = [(-np.pi, np.pi - 2.0 * np.pi / 1024, 1024), (-1.0, 1.0, 1024)], [1]
    else:
        bounds, axes = [(-1.0, 1.0, 96), (0.0, 3.0, 65), (-2.0, 5.0, 33)], [0, 1, 2]
    mappers = [pt.cheb] * len(axes)
    host = pt.Grid.from_bounds(*bounds, axes=axes, mappers=mappers)
    dev = pt.DeviceGrid.from_bounds(*bounds, axes=axes, mappers=mappers)
    assert dev.npts == host.npts
    assert dev.tensor.cpu().numpy().tobytes() == np.asarray(host).tobytes()
    for k in range(len(bounds)):
        assert dev.axpoints(k).tobytes() == host.axpoints(k).tobytes()
        assert dev.axpoints_device(k).is_cuda
    # a user mapper that is not `cheb` is called on the host, as the reference does
    stretch = lambda n, lo, hi: lo + (hi - lo) * np.linspace(0.0, 1.0, n) ** 2   # noqa: E731
    h2 = pt.Grid.from_bounds((0.0, 1.0, 17), (0.0, 2.0, 9), axes=[1], mappers=[stretch])
    d2 = pt.DeviceGrid.from_bounds((0.0, 1.0, 17), (0.0, 2.0, 9), axes=[1], mappers=[stretch])
    assert d2.tensor.cpu().numpy().tobytes() == np.asarray(h2).tobytes()
    with pytest.raises(TypeError):
        pt.DeviceGrid.from_bounds([0.0, 1.0, 5], axes=[0], mappers=[pt.cheb])      # list bound of a mapped axis


def test_device_grid_config4_size():
    """The 1024^3 mesh of BASELINE config 4 (24 GiB, not materialisable by the reference on this
    host) built on the device; checked through slices and checksums."""
    import torch

    import pytristan_b200 as pt

    free, _ = torch.cuda.mem_get_info()
    n = 1024 if free > 40 * 2 ** 30 else 256
    x = np.linspace(0.0, 1.0, n)
    g = pt.DeviceGrid.from_arrs(x, 2.0 * x, -x)
    t = g.tensor
    assert tuple(t.shape) == (3, n ** 3)
    tx = torch.from_numpy(x).cuda()
    assert torch.equal(t[0].view(n, n, n)[7, 5], tx)                      # axis 0 fastest
    assert torch.equal(t[1].view(n, n, n)[3, :, 9], 2.0 * tx)
    assert torch.equal(t[2].view(n, n, n)[:, 11, 2], -tx)
    assert float(t[0].sum()) == pytest.approx(float(x.sum()) * n * n, rel=1e-12)


def test_operator_generators_against_the_frozen_known_answers():
    """Device generators vs tests/golden/operator_kat.npz (frozen oracle outputs): the stencil index
    layout and the off-diagonal Chebyshev entries bit for bit, every entry within 1e-13 of its
    row's absolute sum."""
    import pytristan_b200 as pt

    kat = np.load(os.path.join(GOLDEN_DIR, "operator_kat.npz"))

    def close(got, ref, tol=1e-13):
        scale = np.abs(ref).sum(axis=1, keepdims=True)
        return np.max(np.abs(np.asarray(got) - ref) / np.where(scale > 0, scale, 1.0)) <= tol

    for n in (8, 9, 33):
        x = kat[f"cheb_nodes_{n}"]
        for order in (1, 2, 3):
            ref = kat[f"cheb_mat_{n}_o{order}"]
            D = np.asarray(pt.ChebMat(x, order=order))
            assert close(D, ref)
            if order == 1:
                off = ~np.eye(n, dtype=bool)
                assert np.array_equal(D[off], ref[off])
    assert close(pt.ChebMat(kat["bary_nodes_17"], order=2), kat["bary_mat_17_o2"])
    for n in (8, 9, 32):
        x = np.arange(n) * (2 * np.pi / n)
        for order in (1, 2, 3, 4):
            assert close(pt.FourMat(x, order=order), kat[f"four_mat_{n}_o{order}"], 1e-12 if order > 2 else 1e-13)
    cases = {"uni": (np.linspace(0.0, 1.0, 21), False), "per": (np.linspace(0.0, 1.0 - 1.0 / 20, 20), True)}
    from oracle import operators as oracle
    cases["map"] = (oracle.cheb_nodes(21, 0.0, 1.0), False)
    for tag, (x, periodic) in cases.items():
        for (order, acc) in ((1, 2), (1, 6), (2, 4)):
            F = pt.FinDiffMat(x, order=order, accuracy=acc, periodic=periodic)
            assert np.array_equal(np.asarray(F.start, dtype=np.int64), kat[f"fd_{tag}_o{order}a{acc}_start"])
            assert close(F.weights, kat[f"fd_{tag}_o{order}a{acc}_W"])
