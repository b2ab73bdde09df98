"""Per-tile timeline of the warp-specialised even-odd kernel (clock64 stamps of consumer warp 0 and
producer thread 0 of every CTA; tuning hook pt_debug_set_eo_ws_trace).  Usage: python tools/ws_trace.py"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import pytristan_b200 as pt
from pytristan_b200 import _lib

lib = _lib.load()
lib.pt_debug_set_eo_ws_trace.argtypes = [ctypes.c_void_p]


def run(kind, n, order, npts, axis, batch, label):
    x = np.arange(n) * (2 * np.pi / n) if kind == "four" else pt.cheb(n, -1.0, 1.0)
    op = pt.FourMat(x, order=order) if kind == "four" else pt.ChebMat(x, order=order)
    op.npts, op.axis = tuple(npts), axis
    P = int(np.prod(npts))
    U = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(U)
    for _ in range(3):
        op.apply(U, out=Y)
    buf = torch.zeros(148 * 16 * 8, dtype=torch.int64, device="cuda")
    lib.pt_debug_set_eo_ws_trace(ctypes.c_void_p(buf.data_ptr()))
    op.apply(U, out=Y)
    torch.cuda.synchronize()
    lib.pt_debug_set_eo_ws_trace(None)
    t = buf.cpu().numpy().reshape(148, 16, 8)
    print("==", label)
    for cta in (0, 73, 147):
        rows = t[cta]
        base = rows[0, 0]
        for ti in range(16):
            r = rows[ti]
            if r[0] == 0:
                break
            nxt = rows[ti + 1, 0] if ti + 1 < 16 and rows[ti + 1, 0] else 0
            print(f"cta {cta:3d} tile {ti}: start {r[0] - base:8d}  main {r[1] - r[0]:7d} (waiting for data {r[3]:6d})  "
                  f"epilogue {r[2] - r[1]:6d} (middle row {r[7] - r[1]:5d})  gap to next {nxt - r[2] if nxt else 0:6d} | producer: reached tile at "
                  f"{r[5] - base:8d}, waited for slots {r[4]:7d}")
    main = (t[:, :, 1] - t[:, :, 0])[t[:, :, 0] > 0]
    epi = (t[:, :, 2] - t[:, :, 1])[t[:, :, 0] > 0]
    wait = t[:, :, 3][t[:, :, 0] > 0]
    print(f"all CTAs: main loop mean {main.mean():.0f} cycles (min {main.min()}, max {main.max()}), of which waiting for data "
          f"{wait.mean():.0f}; epilogue mean {epi.mean():.0f} (min {epi.min()}, max {epi.max()})")
    span = t[:, :, 2].max() - t[:, :, 0][t[:, :, 0] > 0].min()
    print(f"kernel span (first tile start -> last epilogue end, SM clocks differ slightly): {span} cycles")


if len(sys.argv) > 1:
    run("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 K1 n=257")
    sys.exit(0)
run("cheb", 256, 2, (1024, 256), 1, 64, "K1 n=256, 64 fields of 1024 x 256")
run("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 K1 n=257")
run("four", 1024, 1, (1024, 257), 0, 64, "cfg2 K2 n=1024")
