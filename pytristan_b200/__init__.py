"""pytristan_b200 — B200-native (sm_100a) implementation of pytristan's differentiation hot path.

Drop-in surface of the reference package (reference src/pytristan/__init__.py:1-5,
grid.py:23-32) plus the operator layer its README announces (README.md:7-10,16,24-26).
Importing the package never touches the GPU; constructing grids/operators and applying them
does, and raises when no CUDA device or no built C-ABI library is present.
"""
from .grid import *  # noqa: F401,F403
from .grid import __all__ as _grid_all
from .operators import *  # noqa: F401,F403
from .operators import __all__ as _op_all
from .host import apply_host, launches_per_apply  # noqa: F401
from .fused import gradient  # noqa: F401
from .solve import AxisSolver, dirichlet, with_rows  # noqa: F401
from .graphs import CapturedApply  # noqa: F401

__version__ = "0.1.0"
__all__ = list(_grid_all) + list(_op_all) + ["apply_host", "launches_per_apply", "gradient", "AxisSolver", "dirichlet", "with_rows", "CapturedApply", "__version__"]
