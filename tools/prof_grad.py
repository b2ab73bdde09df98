import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ops = []
for k in range(3):
    F = pt.FinDiffMat(np.linspace(0, 1, n), order=1, accuracy=6); F.npts, F.axis = (n, n, n), k; ops.append(F)
U = torch.randn(n ** 3, dtype=torch.float64, device="cuda"); outs = [torch.empty_like(U) for _ in range(3)]
for _ in range(3):
    pt.gradient(ops, U, outs=outs)
torch.cuda.synchronize(); print("ok")
