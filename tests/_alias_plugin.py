"""pytest plugin used by tests/test_reference_parity.py to run the REFERENCE's own test file,
unchanged, against pytristan_b200: it aliases the import name ``pytristan`` to this package.
On a box without CUDA the device expansion is replaced by the oracle test double (host logic
only); on a GPU box the real CUDA path runs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pytristan_b200  # noqa: E402
import pytristan_b200.grid as _grid  # noqa: E402

try:
    import torch

    _cuda = torch.cuda.is_available()
except Exception:  # pragma: no cover
    _cuda = False

if not _cuda:
    from oracle import operators as _oracle

    def _double(axes, dtype):
        return np.ascontiguousarray(_oracle.grid_expand(axes), dtype=dtype).reshape(len(axes), -1)

    _grid._expand_axes = _double

sys.modules["pytristan"] = pytristan_b200
