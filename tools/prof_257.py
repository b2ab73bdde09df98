import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
x = pt.cheb(257, -1.0, 1.0)
op = pt.ChebMat(x, order=2); op.npts, op.axis = (1024, 257), 1
U = torch.randn(64, 1024*257, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
for _ in range(3): op.apply(U, out=Y)
torch.cuda.synchronize()
