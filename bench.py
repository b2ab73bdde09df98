#!/usr/bin/env python
"""bench.py — headline benchmark of the differentiation hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Metric: derivative grid-points per second (= output points per applied operator, summed over the
operators of one step) — BASELINE.json:metric.  Default workload = BASELINE configs[1]:
2-D periodic-x Fourier x Chebyshev-y channel grid 1024 x 257, d/dx (FourMat, axis 0, n = 1024)
and d2/dy2 (ChebMat order 2, axis 1, n = 257) on a batch of 64 FP64 fields.  One step = both
operators applied to one batch.  Fields are synthetic N(0,1) (seed 1234), operators are the real
ones generated on the device.

Timed region: K steps bracketed by barrier + cuda synchronize, CUDA events on the launching
stream, max over ranks.  `value` has inputs resident in HBM (a ring of input batches larger than
L2 so that no step finds its input in cache); `e2e` goes through the public host API
(pinned host buffers, H2D of the batch and D2H of both derivative fields inside the region).
`roofline` is for the dominant kernel (the K2 contiguous-axis DMMA kernel of d/dx): algorithmic
flops / average launch duration measured live with CUDA events, against the FP64 tensor-core
peak measured on this GPU by the library's register-resident DMMA loop.
`cpu_baseline` is the NumPy oracle (kind "port": the reference has no operator code) on the host.

Under torchrun (N > 1) every rank differentiates its own batch — the path shards over fields with
no data-path collective — so scaling is "weak".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, BATCH = 1024, 257, 64
SEED = 1234


# --------------------------------------------------------------------------------------------------
# workloads: each returns dict(name, npts, batch, ops=[(label, kind, axis, order, flops)], points_per_step)
# --------------------------------------------------------------------------------------------------
def workload_spec(name):
    if name == "channel_1024x257_b64":
        P = NX * NY
        return dict(name=name, npts=(NX, NY), batch=BATCH, bound="tensor",
                    ops=[("d/dx FourMat n=1024 axis0 (K2)", "four", 0, 1, 2.0 * NX * P * BATCH),
                         ("d2/dy2 ChebMat n=257 axis1 (K1)", "cheb", 1, 2, 2.0 * NY * P * BATCH)],
                    points_per_step=2 * P * BATCH)
    if name == "channel_1024x257_b64_fft":
        # same operators and fields as the default workload; d/dx evaluated by FourMat(method="fft")
        # (pt_fourier_apply: the same linear map in O(n log n), HBM-bound) — SURVEY.md 8f rank 2
        P = NX * NY
        return dict(name=name, npts=(NX, NY), batch=BATCH, bound="tensor",
                    ops=[("d/dx FourMat n=1024 axis0 (FFT)", "four_fft", 0, 1, 16.0 * P * BATCH),
                         ("d2/dy2 ChebMat n=257 axis1 (K1)", "cheb", 1, 2, 2.0 * NY * P * BATCH)],
                    points_per_step=2 * P * BATCH)
    if name in ("polar_1024x512_b64", "polar_1024x512_b1", "polar_1024x512_b64_fft"):
        # BASELINE configs[2]: polar grid, theta Fourier (1024) x r Chebyshev (get_polar_grid(1024, 256,
        # fornberg=True): 512 mapped CGL nodes on [-1, 1], none at r = 0); one step = one Laplacian
        # (D2_r + D_r/r) u + (1/r^2) D2_theta u = two operator applications per point
        nt, nr = 1024, 512
        P = nt * nr
        b = 1 if name.endswith("_b1") else BATCH
        kind = "polar_fft" if name.endswith("_fft") else "polar"
        return dict(name=name, npts=(nt, nr), batch=b, bound="tensor",
                    ops=[("polar Laplacian: radial n=512 (K1) + azimuthal n=1024 (" + ("FFT" if kind == "polar_fft" else "K2") + ", 1/r^2 fused)",
                          kind, 0, 2, 2.0 * (nr + (0 if kind == "polar_fft" else nt)) * P * b)],
                    points_per_step=2 * P * b)
    if name == "fd6_512cube":
        P = 512 ** 3
        return dict(name=name, npts=(512, 512, 512), batch=1, bound="hbm",
                    ops=[(f"d/dx{k} FinDiffMat 7-pt axis{k}", "findiff", k, 1, 16.0 * P) for k in range(3)],
                    points_per_step=3 * P)
    if name == "fd6_1024cube":
        P = 1024 ** 3
        return dict(name=name, npts=(1024, 1024, 1024), batch=1, bound="hbm",
                    ops=[(f"d/dx{k} FinDiffMat 7-pt axis{k}", "findiff", k, 1, 16.0 * P) for k in range(3)],
                    points_per_step=3 * P)
    if name == "cheb2048_b8192":
        P = 2048
        return dict(name=name, npts=(8192, 2048), batch=1, bound="tensor",
                    ops=[("ChebMat n=2048 on 8192 fields (K1)", "cheb", 1, 1, 2.0 * 2048 * 2048 * 8192)],
                    points_per_step=2048 * 8192)
    if name == "fd6_1024cube_slab" or name == "fd6_512cube_slab":
        n = 1024 if "1024" in name else 512
        P = n ** 3
        return dict(name=name, kind="slab", n=n, npts=(n, n, n), batch=1, bound="hbm", scaling="strong",
                    ops=[("d/dx + d/dy + d/dz FinDiffMat 7-pt, z slab-sharded, halo exchange", "findiff", 2, 1, 3 * 16.0 * P)],
                    points_per_step=3 * P)
    if name == "cheb512cube_pencil" or name == "cheb256cube_pencil":
        n = 512 if "512" in name else 256
        P = n ** 3
        return dict(name=name, kind="pencil", n=n, npts=(n, n, n), batch=1, bound="tensor", scaling="strong",
                    ops=[("D2x + D2y + D2z ChebMat on pencils (2 all-to-all rounds)", "cheb", 0, 2, 3 * 2.0 * n * P)],
                    points_per_step=3 * P)
    raise SystemExit(f"unknown workload {name}")


def axis_nodes(spec):
    import pytristan_b200 as pt

    if spec["name"].startswith("channel_1024x257_b64"):
        return [np.linspace(0.0, 2 * np.pi - 2 * np.pi / NX, NX), pt.cheb(NY, -1.0, 1.0)]
    if spec["name"] == "cheb2048_b8192":
        return [np.linspace(0.0, 1.0, 8192), pt.cheb(2048, -1.0, 1.0)]
    return [np.linspace(0.0, 1.0, n) for n in spec["npts"]]


# --------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------------
# CPU side: the NumPy oracle (the reference ships no operator code; kind = "port")
# --------------------------------------------------------------------------------------------------
def cpu_operators(spec, nodes):
    from oracle import operators as oracle

    if spec["ops"][0][1] in ("polar", "polar_fft"):
        import pytristan_b200 as pt

        theta = np.linspace(-np.pi, np.pi - 2 * np.pi / 1024, 1024)
        r = pt.cheb(512, -1.0, 1.0)
        h = (theta[-1] - theta[0]) / 1023
        return [("polar", (oracle.four_mat(1024, 1024 * h, 2), oracle.cheb_mat(r, 2) + oracle.cheb_mat(r, 1) / r[:, None],
                           1.0 / (r * r)))]
    mats = []
    for (_, kind, axis, order, _) in spec["ops"]:
        x = nodes[axis]
        if kind in ("four", "four_fft"):
            n = x.size
            mats.append(("dense", oracle.four_mat(n, n * ((x[-1] - x[0]) / (n - 1)), order)))
        elif kind == "cheb":
            mats.append(("dense", oracle.cheb_mat(x, order)))
        else:
            mats.append(("band", oracle.findiff_band(x, order, 6)))
    return mats


def cpu_step(spec, mats, U):
    from oracle import operators as oracle

    outs = []
    for (kind, M), (_, _, axis, _, _) in zip(mats, spec["ops"]):
        if kind == "polar":     # oracle.polar_laplacian with the matrices built once
            D2t, Ar, inv_r2 = M
            nt, nr = spec["npts"]
            Y = oracle.apply_along_axis(Ar, U, (nt, nr), 1)
            T = oracle.apply_along_axis(D2t, U, (nt, nr), 0)
            outs.append(Y + np.tile(np.repeat(inv_r2, nt), U.size // (nt * nr)).reshape(U.shape) * T)
        elif kind == "dense":
            outs.append(oracle.apply_along_axis(M, U, spec["npts"], axis))
        else:
            outs.append(oracle.stencil_apply(M, U, spec["npts"], axis))
    return outs


def time_cpu(spec, nodes, steps, warmup, sample_batch):
    rng = np.random.default_rng(SEED)
    P = int(np.prod(spec["npts"]))
    U = rng.standard_normal((sample_batch, P))
    mats = cpu_operators(spec, nodes)
    for _ in range(warmup):
        cpu_step(spec, mats, U)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(spec, mats, U)
    dt = (time.perf_counter() - t0) / steps
    pts = spec["points_per_step"] * sample_batch / spec["batch"]
    return pts / dt, dt


def run_reference(args, spec):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference
    snapshot has no operator code, so this is the NumPy oracle (OpenBLAS, all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pytristan_b200 as pt  # noqa: F401  (cheb nodes come from the host formula)

    nodes = axis_nodes(spec)
    cores = os.cpu_count() or 1
    value, dt = time_cpu(spec, nodes, args.steps, args.warmup, spec["batch"])
    line = {
        "impl": "reference", "metric": "derivative grid-points/sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "npts": list(spec["npts"]), "batch": spec["batch"],
                   "ops": [o[0] for o in spec["ops"]]},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port",
                         "sample": f"full workload, {args.steps} steps (NumPy oracle, OpenBLAS on {cores} threads; "
                                   "the reference snapshot ships no operator code)"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------
def build_ops(spec, nodes):
    import pytristan_b200 as pt

    if spec["ops"][0][1] in ("polar", "polar_fft"):
        pt.drop_all_grids()
        g = pt.get_polar_grid(1024, 256, axes=[1], mappers=[pt.cheb], fornberg=True)
        return [pt.PolarLaplacian(g, azimuthal="fft" if spec["ops"][0][1] == "polar_fft" else "dense")]
    ops = []
    for (_, kind, axis, order, _) in spec["ops"]:
        x = nodes[axis]
        if kind == "four":
            op = pt.FourMat(x, order=order)
        elif kind == "four_fft":
            op = pt.FourMat(x, order=order, method="fft")
        elif kind == "cheb":
            op = pt.ChebMat(x, order=order)
        else:
            op = pt.FinDiffMat(x, order=order, accuracy=6)
        op.npts, op.axis = tuple(spec["npts"]), axis      # bind to the (implicit) tensor-product grid
        ops.append(op)
    return ops


def run_gpu(args, spec):
    import torch
    import torch.distributed as dist

    import pytristan_b200 as pt
    from pytristan_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    nodes = axis_nodes(spec)
    ops = build_ops(spec, nodes)
    P = int(np.prod(spec["npts"]))
    B = spec["batch"]
    nbytes = P * B * 8
    # ring of input batches: > L2 (126 MB) in total so that no step finds its input cached
    ring_len = max(2, int(np.ceil(400e6 / nbytes)) + 1) if nbytes < 4e9 else 1
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    ring = [torch.randn(B, P, dtype=torch.float64, device="cuda", generator=gen) for _ in range(ring_len)]
    outs = [torch.empty(B, P, dtype=torch.float64, device="cuda") for _ in ops]

    def step(i):
        u = ring[i % ring_len]
        for op, y in zip(ops, outs):
            op.apply(u, out=y)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()

    # ---- timed region (device-resident) -----------------------------------------------------------
    nops = len(ops)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(nops + 1)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    with sampler:
        barrier()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(args.steps):
            u = ring[i % ring_len]
            ev[i][0].record()
            for k, (op, y) in enumerate(zip(ops, outs)):
                op.apply(u, out=y)
                ev[i][k + 1].record()
        t_end.record()
        barrier()
    total_ms = t_start.elapsed_time(t_end)
    per_op_ms = [float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(args.steps)])) for k in range(nops)]
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = spec["points_per_step"] * world / (ms_per_step * 1e-3)

    # ---- end to end through the public host API (pinned host buffers) ------------------------------
    e2e_steps = max(3, min(args.steps, 10))
    host_in = torch.randn(B, P, dtype=torch.float64).pin_memory()
    host_out = [torch.empty(B, P, dtype=torch.float64).pin_memory() for _ in ops]

    def e2e_step():
        pt.apply_host(ops, host_in, host_out)

    e2e_step()
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(e2e_steps):
        e2e_step()
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = spec["points_per_step"] * world / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    dom = int(np.argmax(per_op_ms))
    label, kind, axis, order, alg = spec["ops"][dom]
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:
        pass
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(spec["name"], {}).get(label)
    except Exception:
        pass
    if spec["bound"] == "tensor":
        import ctypes

        tf = ctypes.c_double(0.0)
        _lib.check(lib.pt_measure_dmma_peak(ctypes.byref(tf), 4000), "pt_measure_dmma_peak")
        dense_form = alg / (per_op_ms[dom] * 1e-3) * 1e-12
        # a centrosymmetric operator runs as two half-size products (even-odd decomposition): the
        # kernel EXECUTES half the flops of the dense form; the roofline fraction is on executed flops
        eo = bool(getattr(ops[dom], "even_odd", False))
        ratio = ops[dom].executed_flop_ratio() if hasattr(ops[dom], "executed_flop_ratio") else (0.5 if eo else 1.0)
        eo = eo or ratio < 1.0
        achieved = dense_form * ratio
        roof = {"bound": "tensor", "kernel": label + (" [even-odd: two half-size DMMA products]" if eo else ""),
                "achieved": achieved, "peak": tf.value, "unit": "TFLOP/s",
                "frac": achieved / tf.value, "traffic": traffic,
                "peak_source": "FP64 DMMA m8n8k4 register-resident loop measured live on this GPU "
                               "(pt_measure_dmma_peak); MEASURED_PEAKS.json has no FP64 entry",
                "alg_flops_per_launch": alg, "executed_flops_per_launch": alg * ratio,
                "dense_form_tflops": dense_form, "even_odd": eo, "avg_launch_ms": per_op_ms[dom]}
    else:
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = alg / (per_op_ms[dom] * 1e-3) * 1e-9
        roof = {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "alg_bytes_per_launch": alg, "avg_launch_ms": per_op_ms[dom]}
    roof["per_op"] = [{"op": o[0], "avg_ms": m, "alg": o[4]} for o, m in zip(spec["ops"], per_op_ms)]

    # ---- CPU baseline on a bounded sample (rank 0, N = 1 only) ------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample_b = B if spec["name"].startswith(("channel_1024x257_b64", "polar_")) else 1
        if spec["name"].startswith("fd6_1024"):
            cpu = {"value": None, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": "skipped: 1024^3 slice-sum oracle needs > 60 s"}
        else:
            v, dt = time_cpu(spec, nodes, 3, 1, sample_b)
            cpu = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"batch of {sample_b} fields x 3 steps, {dt * 1e3:.1f} ms/step "
                             f"(NumPy oracle, OpenBLAS on {cores} threads)"}

    line = {
        "metric": "derivative grid-points/sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec["name"], "npts": list(spec["npts"]), "batch_per_gpu": B,
                   "ops": [o[0] for o in spec["ops"]], "sharding": "fields (batch) per rank, no collective",
                   "l2": f"ring of {ring_len} input batches ({ring_len * nbytes / 1e6:.0f} MB > 126 MB L2)"},
        "roofline": roof,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": nbytes,
                "d2h_bytes_per_step": nbytes * len(ops), "ms_per_step": e2e_ms, "steps": e2e_steps,
                "api": "pytristan_b200.apply_host(ops, pinned_host_fields, pinned_host_outputs)"},
        "gpu_launches": args.steps * pt.launches_per_apply(ops) * world,
        "clocks": sampler.summary(),
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_gpu_sharded(args, spec):
    """Strong-scaling workloads where the differentiated axes are sharded (SURVEY.md 8e ii/iii):
    slab-sharded 3-axis stencil with halo exchange, and the three-axis Chebyshev operator on pencils."""
    import torch
    import torch.distributed as dist

    import pytristan_b200 as pt
    from pytristan_b200 import dist as ptd

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 or spec["kind"] == "pencil":
        if "MASTER_ADDR" not in os.environ:
            os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", RANK="0", WORLD_SIZE="1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = spec["n"]
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    peer = args.comm == "peer"
    fab = None
    if peer:
        from pytristan_b200.peer import PeerFabric

        if not dist.is_initialized():
            os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", RANK="0", WORLD_SIZE="1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        fab = PeerFabric()
    if spec["kind"] == "slab":
        x = np.linspace(0.0, 1.0, n)
        F = pt.FinDiffMat(x, order=1, accuracy=6)
        plan = ptd.SlabStencil.from_operator(F, rank, world)
        fx, fy = pt.FinDiffMat(x, order=1, accuracy=6), pt.FinDiffMat(x, order=1, accuracy=6)
        fx.npts, fx.axis = (n, n, plan.nloc), 0
        fy.npts, fy.axis = (n, n, plan.nloc), 1
        outs = [torch.empty(plan.nloc * n * n, dtype=torch.float64, device="cuda") for _ in range(3)]
        field = torch.randn(1, plan.nloc, n * n, dtype=torch.float64, device="cuda", generator=gen)
        if peer:
            buf = plan.peer_field(fab, 1, n * n)       # the field lives in peer-mapped memory
            own = buf.local[:plan.nloc * n * n]
            own.view(1, plan.nloc, n * n).copy_(field)

            fused = args.fused_gradient

            def step():
                fab.barrier()                           # neighbours' fields are current (device flag barrier)
                if fused and plan.gradient_peer(buf, fx, fy, n, n, outs):
                    return                              # one read of the slab, z halo planes over NVLink
                fx.apply(own, out=outs[0])
                fy.apply(own, out=outs[1])
                plan.apply_peer(buf, 1, n * n, outs[2].view(1, plan.nloc, n * n))   # halo rows loaded over NVLink

            launches = 5 if fused else 6                # barrier + kernels (+ boundary-row kernels)
        else:
            ext = torch.zeros(plan.ext_shape(1, n * n), dtype=torch.float64, device="cuda")
            plan.interior(ext).copy_(field)
            own = plan.interior(ext)                       # contiguous: slabs of the slowest axis

            def step():
                plan.exchange(ext)                          # halo planes over NVLink (NCCL send/recv)
                fx.apply(own.reshape(-1), out=outs[0])
                fy.apply(own.reshape(-1), out=outs[1])
                plan.apply(ext, out=outs[2].view(1, plan.nloc, n * n))

            launches = 5
        del field
        local_alg = 3 * 16.0 * plan.nloc * n * n
        host_bytes = plan.nloc * n * n * 8
    else:
        py = 2 if world % 2 == 0 else 1
        pz = world // py
        xs = pt.cheb(n)
        ops = [pt.ChebMat(xs, order=2) for _ in range(3)]
        if peer:
            pen = ptd.PencilPeer((n, n, n), py, pz, fab)
            u = torch.randn(pen.nzl, pen.nyl, n, dtype=torch.float64, device="cuda", generator=gen)

            def step():
                return pen.apply_sum(ops, u)

            launches = 3 + 2 + 2              # 3 DMMA kernels (two with scatter epilogue), 2 peer scatters, 2 barriers
        else:
            pen = ptd.PencilTranspose((n, n, n), py, pz, rank, world)
            u = torch.randn(pen.nzl, pen.nyl, n, dtype=torch.float64, device="cuda", generator=gen)

            def step():
                return ptd.pencil_apply_sum(ops, u, pen, back=False)

            launches = 3 + 2 + 2 + 1          # 3 DMMA kernels, pack/unpack of two redistributions (z unpack only for batch)
        local_alg = 3 * 2.0 * n * (n ** 3) / world
        host_bytes = u.numel() * 8

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    with sampler:
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            step()
        t1.record()
        barrier()
    ms = t0.elapsed_time(t1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = spec["points_per_step"] / (ms * 1e-3)
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        if spec["bound"] == "hbm":
            peak = peaks.get("hbm_gbs", 6650.0)
            ach = local_alg / (ms * 1e-3) * 1e-9
            if getattr(args, "fused_gradient", False) and peer:
                ach *= 32.0 / 48.0                      # fused gradient: 8 B/pt read + 24 B/pt written
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "note": "per-GPU algorithmic bytes of the whole step (3 derivatives: 16 B/pt each, or 32 B/pt for the fused gradient) over the step time, halo traffic included"}
        else:
            dense_form = local_alg / (ms * 1e-3) * 1e-12
            eo = all(bool(getattr(o, "even_odd", False)) for o in ops)
            ach = dense_form * (0.5 if eo else 1.0)
            roof = {"bound": "tensor", "achieved": ach, "peak": 37.0, "unit": "TFLOP/s", "frac": ach / 37.0, "traffic": None,
                    "dense_form_tflops": dense_form, "even_odd": eo,
                    "note": "per-GPU EXECUTED tensor flops of the three applies (half the dense form when the operators run the even-odd kernel) over the whole step time (redistribution included); peak = measured FP64 DMMA 37.0 (profiles/r01_fp64_peak.json)"}
        line = {"metric": "derivative grid-points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": spec["name"], "npts": list(spec["npts"]), "ops": [o[0] for o in spec["ops"]],
                           "sharding": ("z slabs, " + ("halo rows loaded from the neighbours' memory (peer loads, device flag barrier)" if peer else "halo exchange (NCCL send/recv)")) if spec["kind"] == "slab"
                           else ("2-D pencils, " + ("redistribution fused into the DMMA epilogue / peer-store scatter over NVLink, device flag barriers" if peer else "pack + all_to_all + unpack")),
                           "comm": args.comm, "fused_gradient": bool(args.fused_gradient and peer and spec["kind"] == "slab"),
                           "l2": "per-rank field larger than L2" if host_bytes > 126e6 else "per-rank field fits L2 (strong scaling)"},
                "roofline": roof, "cpu_baseline": None,
                "e2e": None, "gpu_launches": args.steps * launches * world, "clocks": sampler.summary()}
        emit(line)
    if fab is not None:
        fab.check()
        fab.close()
    if dist.is_initialized():
        dist.destroy_process_group()


_REAL_STDOUT = None


def reserve_stdout():
    """stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's
    back (NCCL prints its version banner there when the first communicator is created), so the
    descriptor is pointed at stderr for the whole run and the line goes to the saved original."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="channel_1024x257_b64")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-even-odd", action="store_true",
                    help="apply centrosymmetric operators with the plain full-size kernel (A/B against the default)")
    ap.add_argument("--fused-gradient", action="store_true",
                    help="slab workload: the three derivatives in one read of the field (pt_stencil_grad3_halo)")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="sharded workloads: peer-memory kernels (default) or the NCCL exchange/all_to_all path")
    args = ap.parse_args()
    spec = workload_spec(args.workload)
    if args.no_even_odd and args.impl != "reference":
        import pytristan_b200.operators as ops_mod

        ops_mod.EVEN_ODD = False
    if args.impl == "reference":
        run_reference(args, spec)
    elif spec.get("kind") in ("slab", "pencil"):
        run_gpu_sharded(args, spec)
    else:
        run_gpu(args, spec)


if __name__ == "__main__":
    main()
