"""Test configuration.

* ``gpu`` marker: tests that need a B200; everything else runs on CPU.
* ``cpu_expand`` fixture: CPU-only *test double* for the single device call made while
  constructing a ``Grid`` (``pytristan_b200.grid._expand_axes``), implemented with the oracle, so
  that the host logic (validation, registry, getters) is testable without a GPU.  It is never
  installed outside tests; the product has no CPU path.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


HAS_CUDA = _has_cuda()


def pytest_collection_modifyitems(config, items):
    if HAS_CUDA:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_expand(monkeypatch):
    """Replace the device expansion by the oracle (CPU test double, see module docstring)."""
    from oracle import operators as oracle
    import pytristan_b200.grid as grid_mod

    def fake(axes, dtype):
        return np.ascontiguousarray(oracle.grid_expand(axes), dtype=dtype).reshape(len(axes), -1)

    monkeypatch.setattr(grid_mod, "_expand_axes", fake)
    yield


def clear_grids():
    """Empty the grid registry.  ``drop_all_grids()`` on an empty registry raises TypeError, as in
    the reference (grid.py:608-610), so look before dropping."""
    import pytristan_b200 as pt

    if pt._get_grid_manager().nums():
        pt.drop_all_grids()


@pytest.fixture
def clean_registries():
    import pytristan_b200 as pt

    clear_grids()
    pt.drop_all_mats()
    yield
    clear_grids()
    pt.drop_all_mats()
