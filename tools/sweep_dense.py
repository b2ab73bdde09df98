"""Config sweep over operator sizes: python tools/sweep_dense.py 5,9"""
import ctypes, json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from pytristan_b200 import _lib
from pytristan_b200.operators import _DiffMat
from tools.perf_probe import timeit
lib = _lib.load(); lib.pt_debug_set_dense_config.argtypes = [ctypes.c_int]
cfgs = [int(c) for c in sys.argv[1].split(",")]
for n in (int(v) for v in (sys.argv[2].split(',') if len(sys.argv) > 2 else '64,96,128,129,192,256,257,384,512,513,768,1024,1025,1536,2048,2049'.split(','))):
    for axis in (0, 1):
        cols = max(4096, (1 << 25) // n // 64 * 64)
        npts = (n, cols) if axis == 0 else (cols, n)
        op = _DiffMat._wrap(torch.randn(n, n, dtype=torch.float64, device="cuda"), npts, axis, 1, None)
        U = torch.randn(n * cols, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
        res = {}
        for c in cfgs:
            lib.pt_debug_set_dense_config(c)
            ms = timeit(lambda: op.apply(U, out=Y), reps=4, warm=2)
            res[c] = round(2.0 * n * n * cols / ms * 1e-9 / 37.0, 3)
        print(n, "K2" if axis == 0 else "K1", cols, res, flush=True)
