"""One LU axis solve on the config-2 grid along y (n = 257, 64 fields): target for an ncu launch list."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
npts = (1024, 257)
A = pt.dirichlet(pt.ChebMat(pt.cheb(257, -1.0, 1.0), order=2))
s = pt.AxisSolver(A, npts=npts, axis=1, block=blk)
F = torch.randn(64, int(np.prod(npts)), dtype=torch.float64, device="cuda")
Y = torch.empty_like(F)
for _ in range(3):
    s.solve(F, out=Y)
torch.cuda.synchronize()
print("ok")
