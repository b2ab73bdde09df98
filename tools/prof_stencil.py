"""One stencil case for ncu: python tools/prof_stencil.py axis [n]"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
axis = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
npts = (n, n, n)
F = pt.FinDiffMat(np.linspace(0, 1, n), order=1, accuracy=6)
F.npts, F.axis = npts, axis
U = torch.randn(n ** 3, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
for _ in range(3):
    F.apply(U, out=Y)
torch.cuda.synchronize(); print("ok")
