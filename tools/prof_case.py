"""One dense case, few launches: target for `ncu --set full`.  python tools/prof_case.py k2|k1|k1_257 [cfg]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from pytristan_b200 import _lib
from pytristan_b200.operators import _DiffMat
import ctypes
case = sys.argv[1]
lib = _lib.load()
if len(sys.argv) > 2:
    lib.pt_debug_set_dense_config.argtypes = [ctypes.c_int]
    lib.pt_debug_set_dense_config(int(sys.argv[2]))
npts, axis, batch = {"k2": ((1024, 257), 0, 64), "k1_257": ((1024, 257), 1, 64), "k1": ((8192, 2048), 1, 1)}[case]
n = npts[axis]
op = _DiffMat._wrap(torch.randn(n, n, dtype=torch.float64, device="cuda"), npts, axis, 1, None)
U = torch.randn(batch, int(np.prod(npts)), dtype=torch.float64, device="cuda")
Y = torch.empty_like(U)
for _ in range(3):
    op.apply(U, out=Y)
torch.cuda.synchronize()
print("ok")
