"""Host logic of the grid module (validation, registry, getters) — CPU, with the device
expansion replaced by the oracle test double (conftest.cpu_expand).  The cases are the ones the
reference's own suite pins (reference src/pytristan/tests/test_grid.py, cited per test)."""
import warnings

import numpy as np
import pytest
from numpy.polynomial.chebyshev import chebpts2

import pytristan_b200 as pt
from pytristan_b200 import (Grid, _get_grid_manager, cheb, drop_all_grids, drop_grid, drop_last_grid,
                            get_grid, get_polar_grid)

from pytristan_b200.grid import _expand_axes as _REAL_EXPAND  # noqa: E402  (the unpatched device call)

pytestmark = pytest.mark.usefixtures("cpu_expand", "clean_registries")


# -- cheb (reference test_grid.py:21-41) -----------------------------------------------------------
@pytest.mark.parametrize("npts,xmin,xmax", [(10.0, None, None), (-10, None, None), (10, 0.0, None), (10, None, 0.0)])
def test_cheb_rejects(npts, xmin, xmax):
    with pytest.raises((ValueError, TypeError)):
        cheb(npts, xmin, xmax)


def test_cheb_values_and_exact_endpoints():
    np.testing.assert_allclose(cheb(10), chebpts2(10)[::-1], atol=1e-15)
    x = cheb(10, 0.0, 10.0)
    assert x[0] == 0.0 and x[-1] == 10.0
    assert cheb(0).size == 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.isnan(cheb(1)[0])
    assert cheb(np.int64(5)).shape == (5,)


# -- Grid constructors (reference test_grid.py:44-113) -----------------------------------------------
@pytest.mark.parametrize(
    "bounds,axes,mappers",
    [
        (([0.0, 1.0],), [], []),
        (([0.0, 1.0, 2.0],), [], []),
        (([0.0, 1.0, 2.0],), [0], []),
        (([0.0, 1.0, 2.0],), [0.0], [cheb]),
        (([0.0, 1.0, 2.0],), [1], [cheb]),
        (([0.0, 1.0, 2.0],), [-3], [cheb]),
    ],
)
def test_from_bounds_rejects(bounds, axes, mappers):
    with pytest.raises((ValueError, TypeError, IndexError)):
        Grid.from_bounds(*bounds, axes=axes, mappers=mappers)


def test_from_bounds_error_types():
    with pytest.raises(ValueError):
        Grid.from_bounds((0.0, 1.0))
    with pytest.raises(ValueError):
        Grid.from_bounds((0.0, 1.0, 3), axes=[0], mappers=[])
    with pytest.raises(TypeError):
        Grid.from_bounds((0.0, 1.0, 3), axes=0, mappers=cheb)
    with pytest.raises(TypeError):
        Grid.from_bounds((0.0, 1.0, 3), axes=[0.5], mappers=[cheb])
    with pytest.raises(IndexError):
        Grid.from_bounds((0.0, 1.0, 3), axes=[1], mappers=[cheb])
    with pytest.raises(ValueError):
        Grid.from_bounds((0.0, 1.0, 3), (0.0, 1.0, 3), axes=[1, -1], mappers=[cheb, cheb])


def test_mapped_axis_bound_must_be_a_tuple_like_the_reference(cpu_expand):
    """Reference grid.py:203 builds the mapper call as ``(bound[2],) + bound[:2]``: a list bound of a
    MAPPED axis is a TypeError there (tuple + list); unmapped axes accept lists (np.linspace(*bound))."""
    Grid.from_bounds([0.0, 1.0, 3])                                   # unmapped: fine
    Grid.from_bounds((0.0, 1.0, 3), axes=[0], mappers=[cheb])        # mapped tuple: fine
    with pytest.raises(TypeError):
        Grid.from_bounds([0.0, 1.0, 3], axes=[0], mappers=[cheb])
    with pytest.raises(TypeError):
        Grid.from_bounds(np.array([0.0, 1.0, 3.0]), axes=[0], mappers=[cheb])   # npts arrives as a float


def test_from_arrs_rejects():
    with pytest.raises(ValueError):
        Grid.from_arrs(np.arange(4).reshape(2, 2))
    with pytest.raises(ValueError):
        Grid.from_arrs(1, 2, 3)


GOLDEN = np.array([[0.0, 1.0, 0.0, 1.0, 0.0, 1.0], [-1.0, -1.0, 0.0, 0.0, 1.0, 1.0]])


def test_golden_layout():
    np.testing.assert_array_equal(GOLDEN, Grid.from_bounds((0.0, 1.0, 2), (-1.0, 1.0, 3)))
    np.testing.assert_array_equal(GOLDEN, Grid.from_arrs(np.arange(2.0), np.linspace(-1.0, 1.0, 3)))


def test_from_bounds_mappers():
    x = (1.0 - chebpts2(8)[::-1]) / 2.0
    y = (1.0 - chebpts2(6)[::-1]) / 2.0 * 4.0 - 2.0
    g2 = Grid.from_bounds((0.0, 1.0, 8), (-1.0, 1.0, 3), axes=[0], mappers=[cheb])
    np.testing.assert_allclose(Grid.from_arrs(x, np.array([-1.0, 0.0, 1.0])), g2, atol=1e-14)
    g3 = Grid.from_arrs(x, y)
    g4 = Grid.from_bounds((0.0, 1.0, 8), (-2.0, 2.0, 6), axes=[0, 1], mappers=[cheb, cheb])
    g5 = Grid.from_bounds((0.0, 1.0, 8), (-2.0, 2.0, 6), axes=[1, 0], mappers=[cheb, cheb])
    np.testing.assert_allclose(g3, g4, atol=1e-14)
    np.testing.assert_array_equal(g4, g5)
    g7 = Grid.from_bounds((0.0, 1.0, 8), (-2.0, 2.0, 6), axes=[1], mappers=[cheb])
    g8 = Grid.from_bounds((0.0, 1.0, 8), (-2.0, 2.0, 6), axes=[-1], mappers=[cheb])
    g9 = Grid.from_bounds((0.0, 1.0, 8), (-1.0, 1.0, 3), axes=[-2], mappers=[cheb])
    np.testing.assert_array_equal(g7, g8)
    np.testing.assert_array_equal(g2, g9)


def test_attributes_views_and_repr():
    g = Grid.from_bounds((0.0, 1.0, 2), (-1.0, 1.0, 3))
    assert isinstance(g, np.ndarray) and g.shape == (2, 6)
    assert g.num is None and g.num_dim == 2 and g.npts == (2, 3)
    v = g[:, ::2]
    assert isinstance(v, Grid) and v.num is None and v.num_dim is None and v.npts is None
    assert (g + 1.0).npts is None
    assert "no unique identifier" in repr(g)
    with pytest.raises(RuntimeError):
        g.num = 3                                  # only get_grid may name a grid
    m = g.mgrids()
    ref = np.meshgrid(np.linspace(0.0, 1.0, 2), np.linspace(-1.0, 1.0, 3), indexing="ij")
    assert all(np.array_equal(a, b) for a, b in zip(m, ref))
    gi = Grid.from_arrs(np.arange(3), np.arange(2))
    assert gi.dtype == np.int64


# -- axpoints (reference test_grid.py:116-134) ----------------------------------------------------------
def test_axpoints():
    g1 = Grid.from_bounds([0.0, 1.0, 2])
    with pytest.raises(TypeError):
        g1.axpoints(0.0)
    with pytest.raises(IndexError):
        g1.axpoints(1)
    with pytest.raises(IndexError):
        g1.axpoints(-2)
    g = Grid.from_bounds([0.0, 1.0, 2], [-1.0, 1.0, 3], [5.0, 6.0, 4])
    np.testing.assert_array_equal(g.axpoints(0), [0.0, 1.0])
    np.testing.assert_array_equal(g.axpoints(1), [-1.0, 0.0, 1.0])
    np.testing.assert_array_equal(g.axpoints(-1), np.linspace(5.0, 6.0, 4))
    np.testing.assert_array_equal(g.axpoints(-3), [0.0, 1.0])
    assert type(g.axpoints(0)) is np.ndarray


# -- registry (reference test_grid.py:137-258) -------------------------------------------------------------
def test_get_grid_build_and_reuse():
    g1 = Grid.from_bounds((0.0, 1.0, 2), (-2.0, 2.0, 4))
    g2 = get_grid((0.0, 1.0, 2), (-2.0, 2.0, 4), from_bounds=True)
    np.testing.assert_array_equal(g1, g2)
    assert g2.num == 0 and "unique identifier: 0" in repr(g2)
    g3 = get_grid(np.arange(3), np.linspace(-1, 1, 4))
    assert g3.num == 1
    np.testing.assert_array_equal(g3, Grid.from_arrs(np.arange(3), np.linspace(-1, 1, 4)))
    assert get_grid(num=0) is g2 and get_grid(num=1) is g3


def test_get_grid_raises():
    with pytest.raises(TypeError):
        get_grid(num=1.0)
    with pytest.raises(ValueError):
        get_grid(num=1)


def test_get_grid_num_and_overwrite():
    a = get_grid(np.linspace(0.0, 1.0, 5), np.linspace(0.0, 1.0, 5), num=99)
    assert get_grid(num=99) is a and a.num == 99
    drop_last_grid()
    g1 = get_grid((0.0, 1.0, 5), from_bounds=True, num=2)
    assert get_grid(num=2) is g1
    g3 = get_grid((0.0, 2.0, 6), from_bounds=True, num=2, overwrite=True)
    assert g3.num == g1.num and g3 is not g1
    assert get_grid(num=2) is g3
    assert get_grid((5.0, 6.0, 3), from_bounds=True).num == 3     # max + 1


def test_polar_grid():
    g1 = get_polar_grid(4, 6)
    np.testing.assert_allclose(g1.axpoints(0), np.arange(-np.pi, np.pi, 2 * np.pi / 4), atol=1e-15)
    np.testing.assert_array_equal(g1.axpoints(1), np.linspace(0, 1, 6))
    g2 = get_polar_grid(4, 6, axes=[1], mappers=[cheb])
    np.testing.assert_array_equal(g2.axpoints(1), cheb(6, 0.0, 1.0))
    g3 = get_polar_grid(4, 6, fornberg=True, axes=[1], mappers=[cheb])
    np.testing.assert_array_equal(g3.axpoints(1), cheb(12, -1.0, 1.0))
    assert g3.npts == (4, 12) and 0.0 not in g3.axpoints(1)
    assert [g.num for g in (g1, g2, g3)] == [0, 1, 2]
    drop_grid(nitem=3)
    assert _get_grid_manager().nums() == []


@pytest.mark.parametrize("num,nitem", [(None, 1.0), ([1.0, 0], 0), (None, -1), (0, 1), ([[0], [1]], 0)])
def test_drop_grid_rejects(num, nitem):
    with pytest.raises((ValueError, TypeError)):
        drop_grid(num, nitem)


def test_drop_grid_warns():
    with pytest.warns(RuntimeWarning):
        drop_grid()
    with pytest.warns(RuntimeWarning):
        drop_grid(num=10)


def test_drop_sequences():
    for _ in range(5):
        get_grid(np.linspace(0.0, 1.0, 3), np.linspace(0.0, 1.0, 3))
    mgr = _get_grid_manager()
    assert mgr.nums() == [0, 1, 2, 3, 4]
    drop_grid(num=[0, 3])
    assert mgr.nums() == [1, 2, 4]
    drop_last_grid()
    assert mgr.nums() == [1, 2]
    drop_grid(nitem=2)
    assert mgr.nums() == []
    for _ in range(5):
        get_grid(np.linspace(0.0, 1.0, 3))
    drop_grid(2)
    assert mgr.nums() == [0, 1, 3, 4]
    drop_all_grids()
    assert mgr.nums() == []
    with pytest.raises(TypeError):
        drop_all_grids()      # empty registry: drop_grid(num=[]) -> TypeError, as in the reference (grid.py:608-610)


def test_object_manager_composite_keys():
    from pytristan_b200._manager import ObjectManager

    m = ObjectManager()
    setattr(m, "0", 1)
    setattr(m, "7", 1)
    setattr(m, "3-1-2", 1)
    assert m.nums() == [0, 7, [3, 1, 2]]          # reference _manager.py:7-12


# -- operator getters: host-side validation that never reaches the device ------------------------------------------
def test_operator_getter_validation():
    g = get_grid(np.linspace(0.0, 1.0, 5), np.linspace(0.0, 1.0, 4))
    with pytest.raises(ValueError):
        pt.get_cheb_mat(5)                          # grid 5 is not registered
    with pytest.raises(TypeError):
        pt.get_cheb_mat(1.5)
    with pytest.raises(ValueError):
        pt.get_cheb_mat(Grid.from_arrs(np.arange(3.0)))   # unregistered Grid instance
    with pytest.raises(IndexError):
        pt.get_cheb_mat(g, axis=2)
    with pytest.raises(TypeError):
        pt.get_cheb_mat(g, axis=0.0)
    with pytest.raises(ValueError):
        pt.ChebMat(g, axis=0, order=0)
    with pytest.raises(ValueError):
        pt.FourMat(g, axis=0, order=9)
    with pytest.raises(ValueError):
        pt.FinDiffMat(g, axis=0, order=1, accuracy=3)
    with pytest.raises(ValueError):
        pt.FinDiffMat(g, axis=1, order=1, accuracy=6)     # 4 nodes < 7-point stencil
    with pytest.raises(ValueError):
        pt.FourMat(cheb(8), order=1)                      # not equispaced
    with pytest.warns(RuntimeWarning):
        pt.drop_cheb_mat()
    with pytest.warns(RuntimeWarning):
        pt.drop_cheb_mat(key=[0, 0, 1])
    with pytest.raises(TypeError):
        pt.drop_four_mat(key=[0.5, 0, 1])
    with pytest.raises(ValueError):
        pt.drop_findiff_mat(key=[0, 0, 1, 2], nitem=1)


def test_no_cpu_fallback():
    """Without the test double nothing can be expanded on a CPU-only box: the product fails
    loudly instead of computing on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _REAL_EXPAND([np.arange(3.0)], np.dtype(np.float64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pt.ChebMat(cheb(8), order=1)
