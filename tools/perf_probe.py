"""Kernel-only timings at the BASELINE shapes (CUDA events on torch's current stream).
Usage: python tools/perf_probe.py [dense|stencil|all]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import pytristan_b200 as pt
from pytristan_b200 import _lib
from pytristan_b200.operators import _DiffMat

PEAK_TF = 37.0
HBM = 6555.5


def timeit(fn, reps=5, warm=2, inner=10):
    """Best of `reps` timings of `inner` back-to-back launches (a single launch between two events also
    measures the host's launch latency, ~20 us of Python + ctypes, during which the GPU idles)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        fn()                      # keeps the queue non-empty when the first timed launch arrives
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


def dense(npts, axis, batch, label):
    n = npts[axis]
    D = torch.randn(n, n, dtype=torch.float64, device="cuda")
    op = _DiffMat._wrap(D, npts, axis, 1, None)
    P = int(np.prod(npts))
    U = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(U)
    op.apply(U, out=Y)
    ms = timeit(lambda: op.apply(U, out=Y))
    fl = 2.0 * n * P * batch
    print(json.dumps({"case": label, "ms": round(ms, 4), "tflops": round(fl / ms * 1e-9, 2),
                      "frac_fp64_peak": round(fl / ms * 1e-9 / PEAK_TF, 3), "gpts_s": round(P * batch / ms * 1e-6, 2)}))


def stencil(npts, axis, batch, label, order=1, acc=6):
    axes = [np.linspace(0, 1, m) for m in npts]
    F = pt.FinDiffMat(axes[axis], order=order, accuracy=acc)
    F.npts, F.axis = tuple(npts), axis
    P = int(np.prod(npts))
    U = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(U)
    F.apply(U, out=Y)
    ms = timeit(lambda: F.apply(U, out=Y))
    by = 16.0 * P * batch
    print(json.dumps({"case": label, "ms": round(ms, 4), "gbs": round(by / ms * 1e-6, 1),
                      "frac_hbm": round(by / ms * 1e-6 / HBM, 3), "gpts_s": round(P * batch / ms * 1e-6, 2)}))


def fourier(npts, axis, batch, label, order=1):
    n = npts[axis]
    x = np.arange(n) * (2 * np.pi / n)
    P = int(np.prod(npts))
    U = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(U)
    for method in (("fft",) if "cfg" in sys.argv[1] else ("fft", "dense")):
        F = pt.FourMat(x, order=order, method=method)
        F.npts, F.axis = tuple(npts), axis
        F.apply(U, out=Y)
        ms = timeit(lambda: F.apply(U, out=Y))
        by = 16.0 * P * batch
        print(json.dumps({"case": f"{label} [{method}]", "ms": round(ms, 4), "gbs": round(by / ms * 1e-6, 1),
                          "frac_hbm": round(by / ms * 1e-6 / HBM, 3), "gpts_s": round(P * batch / ms * 1e-6, 2),
                          "dense_equiv_tflops": round(2.0 * n * P * batch / ms * 1e-9, 1)}))


def real_op(kind, n, order, npts, axis, batch, label):
    """Real (centrosymmetric) operators: even-odd kernel vs the plain kernel."""
    import pytristan_b200.operators as ops_mod
    x = np.arange(n) * (2 * np.pi / n) if kind == "four" else pt.cheb(n, -1.0, 1.0)
    op = pt.FourMat(x, order=order) if kind == "four" else pt.ChebMat(x, order=order)
    op.npts, op.axis = tuple(npts), axis
    P = int(np.prod(npts))
    U = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(U)
    fl = 2.0 * n * P * batch
    for eo in (True, False):
        ops_mod.EVEN_ODD = eo
        op.apply(U, out=Y)
        ms = timeit(lambda: op.apply(U, out=Y))
        print(json.dumps({"case": f"{label} [{'even-odd' if eo else 'plain'}]", "ms": round(ms, 4),
                          "dense_form_tflops": round(fl / ms * 1e-9, 2), "executed_tflops": round(fl / ms * 1e-9 / (2 if eo else 1), 2),
                          "executed_frac_peak": round(fl / ms * 1e-9 / (2 if eo else 1) / PEAK_TF, 3), "gpts_s": round(P * batch / ms * 1e-6, 2)}))
    ops_mod.EVEN_ODD = True


def eo_variants(kind, n, order, npts, axis, batch, label, which=(True, False)):
    """Centrosymmetric operators: warp-specialised kernel vs the cp.async even-odd kernel."""
    import pytristan_b200.operators as ops_mod
    x = np.arange(n) * (2 * np.pi / n) if kind == "four" else pt.cheb(n, -1.0, 1.0)
    op = pt.FourMat(x, order=order) if kind == "four" else pt.ChebMat(x, order=order)
    op.npts, op.axis = tuple(npts), axis
    P = int(np.prod(npts))
    nb = P * batch * 8
    ring = [torch.randn(batch, P, dtype=torch.float64, device="cuda") for _ in range(max(1, min(4, int(6e8 // nb))))]
    Y = torch.empty_like(ring[0])
    fl = 2.0 * n * P * batch
    for ws in which:
        ops_mod.EVEN_ODD_WS = ws
        i = [0]

        def run():
            op.apply(ring[i[0] % len(ring)], out=Y)
            i[0] += 1

        ms = timeit(run, reps=8, warm=3)
        print(json.dumps({"case": f"{label} [{'ws' if ws else 'cp.async'}]", "ms": round(ms, 4),
                          "executed_tflops": round(fl / ms * 1e-9 / 2, 2),
                          "executed_frac_peak": round(fl / ms * 1e-9 / 2 / PEAK_TF, 3), "gpts_s": round(P * batch / ms * 1e-6, 2)}))
    ops_mod.EVEN_ODD_WS = True
    del ring, Y


def solve_probe(n, npts, axis, batch, label, blocks=(32, 64, 128)):
    """AxisSolver: one pass with the explicit inverse against the blocked LU substitution (executed flops of the
    LU form: n^2 per line for the two triangles, x2 for FMA)."""
    x = pt.cheb(n, -1.0, 1.0)
    A = pt.dirichlet(pt.ChebMat(x, order=2))
    P = int(np.prod(npts))
    F = torch.randn(batch, P, dtype=torch.float64, device="cuda")
    Y = torch.empty_like(F)
    inv = pt.AxisSolver(A, npts=npts, axis=axis, method="inverse")
    ms = timeit(lambda: inv.solve(F, out=Y))
    lines = P * batch / n
    print(json.dumps({"case": label + " inverse", "ms": round(ms, 4), "tflops": round(2.0 * n * n * lines / ms * 1e-9, 2)}))
    for blk in blocks:
        t0 = time.perf_counter()
        lu = pt.AxisSolver(A, npts=npts, axis=axis, block=blk)
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        ms = timeit(lambda: lu.solve(F, out=Y))
        fl = 2.0 * n * n * lines
        print(json.dumps({"case": label + f" lu block={blk}", "ms": round(ms, 4), "tflops_executed": round(fl / ms * 1e-9, 2),
                          "frac_fp64_peak": round(fl / ms * 1e-9 / PEAK_TF, 3), "setup_s": round(setup, 3)}))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "solve":
        if len(sys.argv) > 2:
            _lib.load().pt_debug_set_tri_lines(int(sys.argv[2]))
        solve_probe(257, (1024, 257), 1, 64, "cfg2 y-solve n=257 (strided axis)")
        solve_probe(1024, (1024, 257), 0, 64, "cfg2 x-solve n=1024 (contiguous axis)")
        solve_probe(512, (512, 512, 512), 1, 1, "512^3 y-solve n=512")
        solve_probe(2048, (8192, 2048), 1, 1, "n=2048 8192 lines", blocks=(64, 128))
        sys.exit(0)
    if what == "wsdbg":
        lib = _lib.load()
        modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else (1, 6, 1, 6)
        for mode in modes:
            lib.pt_debug_set_eo_ws(mode)
            print("== ws debug mode", mode, "(1 default, 6 no consumer stagger)")
            eo_variants("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024", (True,))
            eo_variants("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257", (True,))
            eo_variants("cheb", 256, 2, (1024, 256), 1, 64, "n=256", (True,))
            eo_variants("cheb", 2048, 1, (8192, 2048), 1, 1, "cfg5a-1/8 ChebMat K1 n=2048", (True,))
            eo_variants("cheb", 512, 2, (512, 512, 512), 1, 1, "cfg5b y K1 n=512", (True,))
        lib.pt_debug_set_eo_ws(1)
        sys.exit(0)
    if what == "ws257":
        for _ in range(2):
            eo_variants("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257", (True,))
            eo_variants("cheb", 255, 2, (1024, 255), 1, 64, "n=255", (True,))
            eo_variants("cheb", 513, 2, (1024, 513), 1, 32, "n=513", (True,))
            eo_variants("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024", (True,))
        sys.exit(0)
    if what == "ws":
        eo_variants("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024")
        eo_variants("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257")
        eo_variants("cheb", 256, 2, (1024, 256), 1, 64, "like cfg2 but n=256 (no middle row/column)")
        eo_variants("cheb", 512, 2, (1024, 512), 1, 64, "cfg3 radial K1 n=512 b64")
        eo_variants("four", 1024, 2, (1024, 512), 0, 64, "cfg3 azimuthal K2 n=1024 b64")
        eo_variants("cheb", 2048, 1, (8192, 2048), 1, 1, "cfg5a-1/8 ChebMat K1 n=2048")
        eo_variants("cheb", 512, 2, (512, 512, 512), 0, 1, "cfg5b x K2 n=512")
        eo_variants("cheb", 512, 2, (512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        eo_variants("cheb", 512, 2, (512, 512, 512), 2, 1, "cfg5b z K1 n=512")
        sys.exit(0)
    if what == "eosmall":
        lib = _lib.load()
        for cfg in (1, 2, 0):
            lib.pt_debug_set_eo_config(cfg)
            print("== eo config", cfg)
            real_op("cheb", 512, 2, (1024, 512), 1, 1, "cfg3 B=1 radial K1 n=512")
            real_op("four", 1024, 2, (1024, 512), 0, 1, "cfg3 B=1 azimuthal K2 n=1024")
            real_op("cheb", 257, 2, (1024, 257), 1, 1, "cfg2 B=1 K1 n=257")
            real_op("four", 1024, 1, (1024, 257), 0, 4, "cfg2 4 fields K2 n=1024 (e2e chunk)")
        lib.pt_debug_set_eo_config(-1)
        sys.exit(0)
    if what == "eopersist":
        lib = _lib.load()
        for mode in (50, 51):
            lib.pt_debug_set_eo_config(mode)
            print("== eo persistent", mode - 50)
            real_op("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024")
            real_op("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257")
            real_op("cheb", 2048, 1, (8192, 2048), 1, 1, "cfg5a-1/8 ChebMat K1 n=2048")
            real_op("cheb", 512, 2, (512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        lib.pt_debug_set_eo_config(51)
        sys.exit(0)
    if what == "eocfg":
        lib = _lib.load()
        for cfg in (0, 2):
            lib.pt_debug_set_eo_config(cfg)
            print("== eo config", cfg)
            real_op("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024")
            real_op("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257")
            real_op("cheb", 2048, 1, (8192, 2048), 1, 1, "cfg5a-1/8 ChebMat K1 n=2048")
            real_op("cheb", 512, 2, (512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        sys.exit(0)
    if what == "eo":
        real_op("four", 1024, 1, (1024, 257), 0, 64, "cfg2 d/dx FourMat K2 n=1024")
        real_op("cheb", 257, 2, (1024, 257), 1, 64, "cfg2 d2/dy2 ChebMat K1 n=257")
        real_op("cheb", 512, 2, (1024, 512), 1, 16, "cfg3 radial-like K1 n=512")
        real_op("cheb", 2048, 1, (8192, 2048), 1, 1, "cfg5a-1/8 ChebMat K1 n=2048")
        real_op("cheb", 512, 2, (512, 512, 512), 0, 1, "cfg5b x K2 n=512")
        real_op("cheb", 512, 2, (512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        real_op("four", 1024, 2, (1024, 512), 0, 1, "cfg3 B=1 azimuthal K2 n=1024")
        sys.exit(0)
    if what == "fftcfg":
        lib = _lib.load()
        cfgs = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else (0, 1, 2, 3, 4, 5, 6, 7, 8)
        for cfg in cfgs:
            lib.pt_debug_set_fft_config(cfg)
            print("== fft config", cfg)
            fourier((1024, 257), 0, 64, "cfg2 d/dx FourMat n=1024 axis0")
            if len(sys.argv) > 3:
                continue
            fourier((64, 65536), 0, 4, "n=64 axis0")
            fourier((512, 8192), 0, 4, "n=512 axis0")
            fourier((2048, 2048), 0, 4, "n=2048 axis0")
            fourier((16, 2048, 64), 1, 4, "n=2048 axis1 before=16")
            fourier((256, 4096), 0, 16, "n=256 axis0")
            fourier((4096, 1024), 0, 4, "n=4096 axis0")
            fourier((64, 1024, 64), 1, 4, "n=1024 axis1 before=64")
            fourier((512, 512, 64), 1, 1, "n=512 axis1 before=512")
        sys.exit(0)
    if what == "fft":
        fourier((1024, 257), 0, 64, "cfg2 d/dx FourMat n=1024 axis0")
        fourier((1024, 512), 0, 64, "cfg3 D2_theta n=1024 axis0", order=2)
        fourier((256, 4096), 0, 16, "n=256 axis0")
        fourier((4096, 1024), 0, 4, "n=4096 axis0")
        fourier((64, 1024, 64), 1, 4, "n=1024 axis1 before=64")
        fourier((512, 512, 64), 1, 1, "n=512 axis1 before=512")
        sys.exit(0)
    if what == "tma":
        import ctypes
        lib = _lib.load()
        for on in (0, 2):
            lib.pt_debug_set_dense_tma(ctypes.c_int(on))
            print("== tma", on)
            dense((1024, 257), 0, 64, "cfg2 d/dx K2 n=1024")
            dense((1024, 257), 1, 64, "cfg2 d2/dy2 K1 n=257")
            dense((1024, 512), 1, 16, "cfg3 radial K1 n=512")
            dense((8192, 2048), 1, 1, "cfg5a-1/8 K1 n=2048")
            dense((512, 512, 512), 0, 1, "cfg5b x K2 n=512")
            dense((512, 512, 512), 1, 1, "cfg5b y K1 n=512")
            dense((64, 4096), 0, 1, "K2 n=64")
        lib.pt_debug_set_dense_tma(ctypes.c_int(0))
    if what == "small":
        lib = _lib.load()
        lib.pt_debug_set_dense_config.argtypes = [__import__("ctypes").c_int]
        for cfg in (9, 5, 2, 3, -1):
            lib.pt_debug_set_dense_config(cfg)
            print("== dense config", cfg)
            dense((1024, 512), 1, 1, "cfg3 B=1 radial K1 n=512")
            dense((1024, 512), 0, 1, "cfg3 B=1 azimuthal K2 n=1024")
            dense((1024, 1024), 1, 1, "cfg3 fornberg B=1 radial K1 n=1024")
            dense((1024, 257), 1, 1, "cfg2 B=1 K1 n=257")
            dense((1024, 257), 0, 1, "cfg2 B=1 K2 n=1024")
            dense((64,), 0, 1, "cfg1 n=64 one field")
            dense((256, 256), 0, 1, "256x256 K2")
        sys.exit(0)
    if what == "tune":
        lib = _lib.load()
        lib.pt_debug_set_dense_config.argtypes = [__import__("ctypes").c_int]
        for cfg in [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "0,1,2,3,4").split(",")]:
            lib.pt_debug_set_dense_config(cfg)
            print("== dense config", cfg)
            dense((1024, 257), 0, 64, "cfg2 d/dx K2 n=1024")
            dense((1024, 257), 1, 64, "cfg2 d2/dy2 K1 n=257")
            dense((1024, 512), 1, 16, "cfg3 radial K1 n=512")
            dense((8192, 2048), 1, 1, "cfg5a-1/8 K1 n=2048")
            dense((512, 512, 512), 0, 1, "cfg5b x K2 n=512")
            dense((512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        lib.pt_debug_set_dense_config(0)
    if what == "grad":
        import ctypes
        _l = _lib.load(); _l.pt_debug_set_grad_ty.argtypes = [ctypes.c_int]
        for n, ty in ((1024, 108), (1024, 116), (512, 108)):
            _l.pt_debug_set_grad_ty(ty)
            ops = []
            for k in range(3):
                F = pt.FinDiffMat(np.linspace(0, 1, n), order=1, accuracy=6); F.npts, F.axis = (n, n, n), k; ops.append(F)
            U = torch.randn(n ** 3, dtype=torch.float64, device="cuda"); outs = [torch.empty_like(U) for _ in range(3)]
            pt.gradient(ops, U, outs=outs)
            ms = timeit(lambda: pt.gradient(ops, U, outs=outs))
            print(json.dumps({"case": f"fused gradient {n}^3 tile={ty}", "ms": round(ms, 4), "gbs_32B": round(32.0 * n ** 3 / ms * 1e-6, 1),
                              "frac_hbm": round(32.0 * n ** 3 / ms * 1e-6 / HBM, 3), "gpts_s_x3": round(3 * n ** 3 / ms * 1e-6, 2)}))
            del U, outs
        sys.exit(0)
    if what in ("dense", "all"):
        dense((1024, 257), 0, 64, "cfg2 d/dx K2 n=1024")
        dense((1024, 257), 1, 64, "cfg2 d2/dy2 K1 n=257")
        dense((1024, 512), 1, 16, "cfg3 radial K1 n=512")
        dense((1024, 1024), 1, 8, "cfg3 radial K1 n=1024")
        dense((8192, 2048), 1, 1, "cfg5a-1/8 K1 n=2048 before=8192")
        dense((2048, 8192), 0, 1, "K2 n=2048 after=8192")
        dense((512, 512, 512), 0, 1, "cfg5b x K2 n=512")
        dense((512, 512, 512), 1, 1, "cfg5b y K1 n=512")
        dense((512, 512, 512), 2, 1, "cfg5b z K1 n=512")
        dense((64,), 0, 1, "cfg1 n=64 single field")
    if what in ("stencil", "all"):
        stencil((512, 512, 512), 0, 1, "512^3 x (K3)")
        stencil((512, 512, 512), 1, 1, "512^3 y (K4)")
        stencil((512, 512, 512), 2, 1, "512^3 z (K4)")
        stencil((1024, 1024, 512), 0, 1, "1024x1024x512 x (K3)")
        stencil((1024, 1024, 512), 1, 1, "1024x1024x512 y (K4)")
        stencil((1024, 1024, 512), 2, 1, "1024x1024x512 z (K4)")
        import ctypes
        _l = _lib.load(); _l.pt_debug_set_grad_ty.argtypes = [ctypes.c_int]
        for n, ty in ((512, 16), (512, 32), (1024, 16), (1024, 32)):
            _l.pt_debug_set_grad_ty(ty)
            ops = []
            for k in range(3):
                F = pt.FinDiffMat(np.linspace(0, 1, n), order=1, accuracy=6); F.npts, F.axis = (n, n, n), k; ops.append(F)
            U = torch.randn(n ** 3, dtype=torch.float64, device="cuda"); outs = [torch.empty_like(U) for _ in range(3)]
            ms = timeit(lambda: pt.gradient(ops, U, outs=outs))
            print(json.dumps({"case": f"fused gradient {n}^3 TY={ty}", "ms": round(ms, 4), "gbs_32B": round(32.0 * n ** 3 / ms * 1e-6, 1),
                              "frac_hbm": round(32.0 * n ** 3 / ms * 1e-6 / HBM, 3), "gpts_s_x3": round(3 * n ** 3 / ms * 1e-6, 2)}))
            del U, outs
        x = torch.randn(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
        ms = timeit(lambda: y.copy_(x))
        print(json.dumps({"case": "torch copy 2 GiB", "ms": round(ms, 4), "gbs": round(2 * x.numel() * 8 / ms * 1e-6, 1)}))
