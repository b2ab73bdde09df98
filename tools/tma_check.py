"""Correctness of the TMA dense kernel vs the cp.async kernel on a few shapes (bitwise equality
expected: same tiles, same summation order)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from pytristan_b200 import _lib
from pytristan_b200.operators import _DiffMat
lib = _lib.load()
ok = True
for npts, axis, batch in [((1024, 64), 0, 2), ((64, 1024), 1, 2), ((130, 64), 0, 7), ((64, 136), 1, 3), ((256, 200), 0, 1), ((128, 520, 3), 1, 1),
                          ((258, 128), 0, 1), ((128, 258), 1, 2), ((1024, 257), 0, 4), ((1024, 257), 1, 4), ((64, 64, 70), 2, 1), ((2048, 64), 0, 1)]:
    n = npts[axis]
    op = _DiffMat._wrap(torch.randn(n, n, dtype=torch.float64, device="cuda"), npts, axis, 1, None)
    U = torch.randn(batch, int(np.prod(npts)), dtype=torch.float64, device="cuda")
    lib.pt_debug_set_dense_tma(ctypes.c_int(0)); y0 = op.apply(U)
    lib.pt_debug_set_dense_tma(ctypes.c_int(int(sys.argv[1]) if len(sys.argv) > 1 else 1)); y1 = op.apply(U)
    torch.cuda.synchronize()
    same = torch.equal(y0, y1)
    err = float((y0 - y1).abs().max())
    print(npts, axis, batch, "equal" if same else f"DIFF {err:.3e}", flush=True)
    ok &= same or err < 1e-9
lib.pt_debug_set_dense_tma(ctypes.c_int(0))
print("ALL OK" if ok else "FAILED")
