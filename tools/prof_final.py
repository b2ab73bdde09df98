"""One launch each of the round's final kernels at their BASELINE shapes (for ncu --set full captures)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
what = sys.argv[1]
if what == "eo_k1":          # config 2 d2/dy2: ChebMat n = 257 along axis 1, 64 fields
    D = pt.ChebMat(pt.cheb(257, -1.0, 1.0), order=2); D.npts, D.axis = (1024, 257), 1
    U = torch.randn(64, 1024 * 257, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
    run = lambda: D.apply(U, out=Y)
elif what == "eo_k2":        # config 2 d/dx: FourMat n = 1024 along axis 0, 64 fields
    F = pt.FourMat(np.arange(1024) * (2 * np.pi / 1024), order=1); F.npts, F.axis = (1024, 257), 0
    U = torch.randn(64, 1024 * 257, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
    run = lambda: F.apply(U, out=Y)
elif what == "fft":
    F = pt.FourMat(np.arange(1024) * (2 * np.pi / 1024), order=1, method="fft"); F.npts, F.axis = (1024, 257), 0
    U = torch.randn(64, 1024 * 257, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
    run = lambda: F.apply(U, out=Y)
elif what == "grad3":
    n, nz = 1024, 256
    ops = []
    for k, m in enumerate((n, n, nz)):
        G = pt.FinDiffMat(np.linspace(0, 1, m), order=1, accuracy=6); G.npts, G.axis = (n, n, nz), k; ops.append(G)
    U = torch.randn(n * n * nz, dtype=torch.float64, device="cuda"); outs = [torch.empty_like(U) for _ in range(3)]
    run = lambda: pt.gradient(ops, U, outs=outs)
for _ in range(3):
    run()
torch.cuda.synchronize(); print("ok", what)
