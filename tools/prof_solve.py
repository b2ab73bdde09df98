"""One LU axis solve (512^3, y axis, block 128): target for `ncu --set full -k regex:tri_solve`."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
n = 512
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 128
npts = (512, 512, 128)
A = pt.dirichlet(pt.ChebMat(pt.cheb(n, -1.0, 1.0), order=2))
s = pt.AxisSolver(A, npts=npts, axis=1, block=blk)
F = torch.randn(int(np.prod(npts)), dtype=torch.float64, device="cuda")
Y = torch.empty_like(F)
for _ in range(2):
    s.solve(F, out=Y)
torch.cuda.synchronize()
print("ok")
