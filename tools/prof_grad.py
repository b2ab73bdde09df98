import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nz = int(sys.argv[2]) if len(sys.argv) > 2 else n
ops = []
for k, m in enumerate((n, n, nz)):
    F = pt.FinDiffMat(np.linspace(0, 1, m), order=1, accuracy=6); F.npts, F.axis = (n, n, nz), k; ops.append(F)
U = torch.randn(n * n * nz, dtype=torch.float64, device="cuda"); outs = [torch.empty_like(U) for _ in range(3)]
from pytristan_b200 import _lib
if len(sys.argv) > 3:
    _lib.load().pt_debug_set_grad_ty(int(sys.argv[3]))
for _ in range(3):
    pt.gradient(ops, U, outs=outs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pt.gradient(ops, U, outs=outs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"grad3 cfg={sys.argv[3] if len(sys.argv) > 3 else None} {n}x{n}x{nz}: {ms:.3f} ms, {32.0 * n * n * nz / ms * 1e-6:.0f} GB/s algorithmic, {3 * n * n * nz / ms * 1e-6:.1f} Gpts/s")
