import sys
import numpy as np, torch
sys.path.insert(0, ".")
import os
import pytristan_b200 as pt
from pytristan_b200 import _lib
if "PT_FFT_CONFIG" in os.environ:
    _lib.load().pt_debug_set_fft_config(int(os.environ["PT_FFT_CONFIG"]))
n, rows = 1024, 257 * 64
x = np.arange(n) * (2 * np.pi / n)
F = pt.FourMat(x, order=1, method="fft"); F.npts, F.axis = (n, rows), 0
U = torch.randn(n * rows, dtype=torch.float64, device="cuda"); Y = torch.empty_like(U)
for _ in range(3):
    F.apply(U, out=Y)
torch.cuda.synchronize(); print("ok")
