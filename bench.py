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
stream, max over ranks (every rank's own time and clock record is kept in `per_rank`).  `value`
has inputs resident in HBM (a ring of input batches larger than L2 so that no step finds its input
in cache); `e2e` goes through the public host API (pinned host buffers, H2D of the batch and D2H of
both derivative fields inside the region) and carries the host-link ceiling measured beside it.
`roofline` is for the dominant kernel: algorithmic flops / average launch duration measured live
with CUDA events, against the FP64 tensor-core peak measured on this GPU by the library's
register-resident DMMA loop (cuBLAS DGEMM beside it as a cross-check).  `sustained` repeats the
step for >= 2 s.  `cpu_baseline` is the NumPy oracle (kind "port": the reference has no operator
code) on the host, on the BLAS thread pool it actually got.

Under torchrun (N > 1) every rank differentiates its own batch — the path shards over fields with
no data-path collective — so the headline scaling is "weak".  The default run also measures, in
the same invocation, the other BASELINE configs (`workloads`): at every N the strong-scaling
sharded paths (3-D Chebyshev operator on pencils, slab-sharded 1024^3 stencil; peer-memory and
NCCL forms), at N = 1 the single-GPU configs 3, 4 and 5a at their full sizes.
"""
import argparse
import json
import os
import socket
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NX, NY, BATCH = 1024, 257, 64
SEED = 1234
DEFAULT_WORKLOAD = "channel_1024x257_b64"
L2_BYTES = 126e6


# --------------------------------------------------------------------------------------------------
# workloads: each returns dict(name, npts, batch, ops=[(label, kind, axis, order, alg)], points_per_step)
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
    if name in ("fd6_512cube", "fd6_1024cube"):
        n = 1024 if "1024" in name else 512
        P = n ** 3
        return dict(name=name, npts=(n, n, n), batch=1, bound="hbm", cpu_npts=(256, 256, 256),
                    ops=[(f"d/dx{k} FinDiffMat 7-pt axis{k}", "findiff", k, 1, 16.0 * P) for k in range(3)],
                    points_per_step=3 * P)
    if name in ("fd6_512cube_grad3", "fd6_1024cube_grad3"):
        # the same three derivatives in ONE pass over the field (pt_stencil_grad3): 8 B/pt read + 24 B/pt written
        n = 1024 if "1024" in name else 512
        P = n ** 3
        return dict(name=name, npts=(n, n, n), batch=1, bound="hbm", cpu_npts=(256, 256, 256),
                    ops=[("fused gradient (d/dx, d/dy, d/dz) FinDiffMat 7-pt, one read of the field", "grad3", 0, 1, 32.0 * P)],
                    points_per_step=3 * P)
    if name in ("cheb2048_b8192", "cheb2048_b65536"):
        # BASELINE configs[4] first half: ChebMat N = 2048 applied to 65 536 fields, batch-contiguous
        # (a grid (nb, 2048) differentiated along axis 1: K1 with before = nb)
        nb = 65536 if name.endswith("65536") else 8192
        return dict(name=name, npts=(nb, 2048), batch=1, bound="tensor", cpu_npts=(4096, 2048),
                    ops=[(f"ChebMat n=2048 on {nb} fields (K1)", "cheb", 1, 1, 2.0 * 2048 * 2048 * nb)],
                    points_per_step=2048 * nb)
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


def ring_length(spec):
    """Input batches kept in rotation: more than L2 (126 MB) in total, so that no step finds its input cached."""
    nbytes = int(np.prod(spec["npts"])) * spec["batch"] * 8
    return max(2, int(np.ceil(400e6 / nbytes)) + 1) if nbytes < 4e9 else 1


def config_dict(spec):
    """The `config` object of the JSON line — the SAME dict in the GPU arm and in the reference arm
    (the driver compares them to decide that both arms ran one workload)."""
    nbytes = int(np.prod(spec["npts"])) * spec["batch"] * 8
    ring = ring_length(spec)
    return {"workload": spec["name"], "npts": list(spec["npts"]), "batch_per_gpu": spec["batch"],
            "ops": [o[0] for o in spec["ops"]], "sharding": "fields (batch) per rank, no collective",
            "l2": (f"ring of {ring} input batches ({ring * nbytes / 1e6:.0f} MB > 126 MB L2)" if ring > 1
                   else f"one input of {nbytes / 1e6:.0f} MB > 126 MB L2")}


def axis_nodes(spec, npts=None):
    import pytristan_b200 as pt

    npts = tuple(spec["npts"] if npts is None else npts)
    if spec["name"].startswith("channel_1024x257_b64"):
        return [np.linspace(0.0, 2 * np.pi - 2 * np.pi / NX, NX), pt.cheb(NY, -1.0, 1.0)]
    if spec["name"].startswith("cheb2048_b"):
        return [np.linspace(0.0, 1.0, npts[0]), pt.cheb(2048, -1.0, 1.0)]
    return [np.linspace(0.0, 1.0, n) for n in npts]


# --------------------------------------------------------------------------------------------------
# clocks / power sampler (NVML), runs during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index, period=0.002):
        self.samples, self.power, self.reasons, self.max_mhz = [], [], set(), None
        self.period = period
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
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(self.period)

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
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.samples:
            out["sm_mhz_min"] = float(np.min(self.samples))
        if self.power:
            out["power_w_max"] = float(np.max(self.power))
        return out


# --------------------------------------------------------------------------------------------------
# CPU side: the NumPy oracle (the reference ships no operator code; kind = "port")
# --------------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_blas_threads():
    """Give the BLAS behind NumPy every core this process may run on and return the pool size it
    REALLY has.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, which OpenBLAS
    honours when it loads: without this call the CPU arm under torchrun would run on one thread."""
    want = host_cores()
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=want, user_api="blas")
        a = np.ones((256, 256))
        a @ a                                   # spawns the pool's threads outside the timed region
        pools = [d["num_threads"] for d in threadpool_info() if d.get("user_api") == "blas"]
        return int(max(pools)) if pools else 1
    except Exception:
        return int(os.environ.get("OMP_NUM_THREADS", "1") or 1)


def cpu_operators(spec, nodes):
    from oracle import operators as oracle

    if spec["ops"][0][1] in ("polar", "polar_fft"):
        import pytristan_b200 as pt

        theta = np.linspace(-np.pi, np.pi - 2 * np.pi / 1024, 1024)
        r = pt.cheb(512, -1.0, 1.0)
        h = (theta[-1] - theta[0]) / 1023
        return [("polar", (oracle.four_mat(1024, 1024 * h, 2), oracle.cheb_mat(r, 2) + oracle.cheb_mat(r, 1) / r[:, None],
                           1.0 / (r * r)))]
    if spec["ops"][0][1] == "grad3":
        return [("grad3", [oracle.findiff_band(nodes[k], 1, 6) for k in range(3)])]
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


def cpu_step(spec, mats, U, npts):
    from oracle import operators as oracle

    outs = []
    for (kind, M), (_, _, axis, _, _) in zip(mats, spec["ops"]):
        if kind == "polar":     # oracle.polar_laplacian with the matrices built once
            D2t, Ar, inv_r2 = M
            nt, nr = npts
            Y = oracle.apply_along_axis(Ar, U, (nt, nr), 1)
            T = oracle.apply_along_axis(D2t, U, (nt, nr), 0)
            outs.append(Y + np.tile(np.repeat(inv_r2, nt), U.size // (nt * nr)).reshape(U.shape) * T)
        elif kind == "grad3":
            outs.extend(oracle.stencil_apply(M[k], U, npts, k) for k in range(3))
        elif kind == "dense":
            outs.append(oracle.apply_along_axis(M, U, npts, axis))
        else:
            outs.append(oracle.stencil_apply(M, U, npts, axis))
    return outs


def time_cpu(spec, steps, warmup, sample_batch, npts=None):
    """Oracle timing on `sample_batch` fields of a grid `npts` (default: the spec's own grid).
    Returns (points/s, seconds per step)."""
    npts = tuple(spec["npts"] if npts is None else npts)
    nodes = axis_nodes(spec, npts)
    rng = np.random.default_rng(SEED)
    P = int(np.prod(npts))
    U = rng.standard_normal((sample_batch, P))
    mats = cpu_operators(spec, nodes)
    for _ in range(warmup):
        cpu_step(spec, mats, U, npts)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(spec, mats, U, npts)
    dt = (time.perf_counter() - t0) / steps
    per_point = spec["points_per_step"] / (int(np.prod(spec["npts"])) * spec["batch"])   # derivatives per point
    return per_point * P * sample_batch / dt, dt


def cpu_baseline_record(spec, threads, steps=3, warmup=1):
    """Bounded sample of the workload on the host cores (NumPy oracle)."""
    full = spec["name"].startswith(("channel_1024x257_b64", "polar_"))
    npts = tuple(spec["npts"]) if full else tuple(spec.get("cpu_npts", spec["npts"]))
    sample_b = spec["batch"] if full else 1
    v, dt = time_cpu(spec, steps, warmup, sample_b, npts)
    what = (f"full batch of {sample_b} fields" if full else
            f"grid {'x'.join(map(str, npts))} (the workload's grid is {'x'.join(map(str, spec['npts']))})")
    return {"value": v, "unit": "points/s", "cores": threads, "kind": "port",
            "sample": f"{what} x {steps} steps, {dt * 1e3:.1f} ms/step (NumPy oracle, OpenBLAS pool of {threads} threads, "
                      f"{host_cores()} cores visible)"}


def run_reference(args, spec):
    """--impl reference: the CPU implementation of the path on the host cores.  The reference
    snapshot has no operator code, so this is the NumPy oracle (OpenBLAS, all cores).  Under
    torchrun rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pytristan_b200 as pt  # noqa: F401  (cheb nodes come from the host formula)

    threads = set_blas_threads()
    full = spec["name"].startswith(("channel_1024x257_b64", "polar_")) or "cpu_npts" not in spec
    npts = tuple(spec["npts"]) if full else tuple(spec["cpu_npts"])
    value, dt = time_cpu(spec, args.steps, args.warmup, spec["batch"], npts)
    sample = ("full workload" if full else f"bounded sample: grid {'x'.join(map(str, npts))}") + \
             f", {args.steps} steps (NumPy oracle, OpenBLAS pool of {threads} threads, {host_cores()} cores visible; " \
             "the reference snapshot ships no operator code)"
    line = {
        "impl": "reference", "metric": "derivative grid-points/sec", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(spec),
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# GPU side: process context
# --------------------------------------------------------------------------------------------------
class Ctx:
    """rank / world / device of this process; torch.distributed initialised when world > 1 (or on
    demand with world = 1 for the sharded workloads, whose classes take a process group)."""

    def __init__(self):
        import torch

        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            from pytristan_b200.host import bind_rank_to_cores

            self.cores = bind_rank_to_cores(self.local, int(os.environ.get("LOCAL_WORLD_SIZE", self.world)))
            self.need_group()
        else:
            self.cores = sorted(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else []

    def need_group(self):
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            if "MASTER_ADDR" not in os.environ:
                with socket.socket() as s:
                    s.bind(("127.0.0.1", 0))
                    port = s.getsockname()[1]
                os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        import torch
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(self, obj):
        """list over ranks of a small Python object."""
        import torch.distributed as dist

        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def close(self):
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


_DMMA_PEAK = {}


def dmma_peak():
    """FP64 tensor-core peak of this GPU: the library's register-resident DMMA loop, measured live
    (MEASURED_PEAKS.json has no FP64 entry), and cuBLAS DGEMM 8192^3 beside it as a cross-check."""
    if not _DMMA_PEAK:
        import ctypes

        import torch

        from pytristan_b200 import _lib

        lib = _lib.load()
        tf = ctypes.c_double(0.0)
        _lib.check(lib.pt_measure_dmma_peak(ctypes.byref(tf), 4000), "pt_measure_dmma_peak")
        _DMMA_PEAK["tflops"] = tf.value
        try:
            n = 8192
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            torch.matmul(a, b)
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            _DMMA_PEAK["cublas_dgemm_8192_tflops"] = 2.0 * n ** 3 / (best * 1e-3) * 1e-12
            del a, b
            torch.cuda.empty_cache()
        except Exception as exc:                 # the cross-check is optional
            _DMMA_PEAK["cublas_dgemm_8192_tflops"] = None
            _DMMA_PEAK["cublas_error"] = str(exc)[:200]
    return _DMMA_PEAK


def build_ops(spec, nodes):
    import pytristan_b200 as pt

    kind0 = spec["ops"][0][1]
    if kind0 in ("polar", "polar_fft"):
        if pt._get_grid_manager().nums():
            pt.drop_all_grids()
        g = pt.get_polar_grid(1024, 256, axes=[1], mappers=[pt.cheb], fornberg=True)
        return [pt.PolarLaplacian(g, azimuthal="fft" if kind0 == "polar_fft" else "dense")]
    if kind0 == "grad3":
        fds = []
        for k in range(3):
            op = pt.FinDiffMat(nodes[k], order=1, accuracy=6)
            op.npts, op.axis = tuple(spec["npts"]), k
            fds.append(op)
        return [Gradient3(fds)]
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


class Gradient3:
    """pt.gradient as one 'operator' with three outputs (the fused one-read stencil pass)."""

    launches = 1
    n_out = 3

    def __init__(self, fds):
        self.fds = fds

    def run(self, u, outs):
        import pytristan_b200 as pt

        pt.gradient(self.fds, u, outs=outs)


def time_steps(ctx, step, steps, warmup):
    """warm up, then time `steps` calls of step(i) on the current stream: (mean ms of this rank, ClockSampler)."""
    import torch

    for i in range(warmup):
        step(i)
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    with sampler:
        ctx.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(steps):
            step(i)
        t1.record()
        ctx.barrier()
    return t0.elapsed_time(t1) / steps, sampler


# --------------------------------------------------------------------------------------------------
# device-resident measurement of an un-sharded workload (fields split over ranks, no collective)
# --------------------------------------------------------------------------------------------------
def measure_plain(ctx, spec, steps, warmup, e2e=True, sustained_s=0.0, graph=False):
    import torch

    import pytristan_b200 as pt

    nodes = axis_nodes(spec)
    ops = build_ops(spec, nodes)
    P = int(np.prod(spec["npts"]))
    B = spec["batch"]
    nbytes = P * B * 8
    ring_len = ring_length(spec)
    gen = torch.Generator(device="cuda").manual_seed(SEED + ctx.rank)
    ring = [torch.randn(B, P, dtype=torch.float64, device="cuda", generator=gen) for _ in range(ring_len)]
    outs = [[torch.empty(B, P, dtype=torch.float64, device="cuda") for _ in range(getattr(op, "n_out", 1))] for op in ops]

    def run_op(op, u, ys):
        if hasattr(op, "run"):
            op.run(u, ys)
        else:
            op.apply(u, out=ys[0])

    def step(i):
        u = ring[i % ring_len]
        for op, ys in zip(ops, outs):
            run_op(op, u, ys)

    for i in range(warmup):
        step(i)
    ctx.barrier()

    # ---- timed region (device-resident) -----------------------------------------------------------
    nops = len(ops)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(nops + 1)] for _ in range(steps)]
    sampler = ClockSampler(ctx.local)
    with sampler:
        ctx.barrier()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(steps):
            u = ring[i % ring_len]
            ev[i][0].record()
            for k, (op, ys) in enumerate(zip(ops, outs)):
                run_op(op, u, ys)
                ev[i][k + 1].record()
        t_end.record()
        ctx.barrier()
    my_ms = t_start.elapsed_time(t_end) / steps
    per_op_ms = [float(np.mean([ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(steps)])) for k in range(nops)]
    mine = {"rank": ctx.rank, "gpu": ctx.local, "ms_per_step": my_ms, "per_op_ms": per_op_ms, **sampler.summary()}

    # ---- launch-bound steps: the same launches captured once per ring slot and replayed as CUDA graphs ----
    graph_ms = None
    if graph and all(hasattr(op, "apply") for op in ops):
        from pytristan_b200.graphs import CapturedApply

        caps = [CapturedApply(ops, ring[k], [ys[0] for ys in outs]) for k in range(ring_len)]
        for k in range(min(warmup, ring_len)):
            caps[k].replay()
        ctx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(steps, 50)
        g0.record()
        for i in range(reps):
            caps[i % ring_len].replay()
        g1.record()
        ctx.barrier()
        graph_ms = g0.elapsed_time(g1) / reps
        del caps

    # ---- sustained: the same step back to back for >= sustained_s seconds --------------------------------
    sus = None
    if sustained_s > 0:
        chunk = max(50, int(0.25 / max(my_ms * 1e-3, 1e-6)))
        ssam = ClockSampler(ctx.local, period=0.02)
        done, ms_total = 0, 0.0
        with ssam:
            ctx.barrier()
            wall0 = time.perf_counter()
            while time.perf_counter() - wall0 < sustained_s:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(chunk):
                    step(done + i)
                e1.record()
                torch.cuda.synchronize()
                ms_total += e0.elapsed_time(e1)
                done += chunk
            ctx.barrier()
        sus = {"rank": ctx.rank, "steps": done, "seconds": ms_total * 1e-3, "ms_per_step": ms_total / done, **ssam.summary()}

    # ---- end to end through the public host API (pinned host buffers) ------------------------------
    e2e_rec = None
    if e2e:
        e2e_steps = max(3, min(steps, 10))
        host_in = torch.randn(B, P, dtype=torch.float64).pin_memory()
        host_out = [torch.empty(B, P, dtype=torch.float64).pin_memory() for _ in ops]
        pt.apply_host(ops, host_in, host_out)
        ctx.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(e2e_steps):
            pt.apply_host(ops, host_in, host_out)
        t1.record()
        ctx.barrier()
        e2e_ms = t0.elapsed_time(t1) / e2e_steps
        link = host_link_ceiling(ctx, host_in, host_out)
        e2e_rec = {"ms": e2e_ms, "steps": e2e_steps, "link": link}
        del host_in, host_out

    ranks = ctx.gather(mine)
    sus_ranks = ctx.gather(sus) if sus is not None else None
    e2e_ranks = ctx.gather(e2e_rec) if e2e_rec is not None else None
    launches = pt.launches_per_apply(ops)
    eo_ratio = [(op.executed_flop_ratio() if hasattr(op, "executed_flop_ratio")
                 else (0.5 if bool(getattr(op, "even_odd", False)) else 1.0)) for op in ops]
    del ring, outs, ops
    torch.cuda.empty_cache()
    return dict(ranks=ranks, sustained=sus_ranks, e2e=e2e_ranks, launches=launches, eo_ratio=eo_ratio, nbytes=nbytes,
                graph_ms=graph_ms)


def host_link_ceiling(ctx, host_in, host_outs):
    """What the host link gives this rank while ALL ranks copy at once: the step's H2D bytes alone,
    its D2H bytes alone, and both directions together (pinned buffers, two copy streams) — the
    ceiling of the e2e figure, measured rather than assumed."""
    import torch

    d_in = torch.empty_like(host_in, device="cuda")
    d_outs = [torch.empty_like(h, device="cuda") for h in host_outs]
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h2d_bytes = host_in.numel() * 8
    d2h_bytes = sum(h.numel() for h in host_outs) * 8

    def run(do_in, do_out, reps=3):
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cur = torch.cuda.current_stream()
        e0.record()
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        for _ in range(reps):
            if do_in:
                with torch.cuda.stream(s1):
                    d_in.copy_(host_in, non_blocking=True)
            if do_out:
                with torch.cuda.stream(s2):
                    for h, d in zip(host_outs, d_outs):
                        h.copy_(d, non_blocking=True)
        cur.wait_stream(s1)
        cur.wait_stream(s2)
        e1.record()
        ctx.barrier()
        return e0.elapsed_time(e1) / reps

    run(True, True, 1)
    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_gbs": h2d_bytes / t_in * 1e-6, "d2h_gbs": d2h_bytes / t_out * 1e-6,
            "both_ms": t_both, "both_gbs": (h2d_bytes + d2h_bytes) / t_both * 1e-6}


def roofline_record(spec, per_op_ms, eo_ratio, label_suffix=""):
    """Roofline of the dominant kernel of an un-sharded workload."""
    dom = int(np.argmax(per_op_ms))
    label, kind, axis, order, alg = spec["ops"][dom]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(spec["name"], {}).get(label)
    except Exception:
        pass
    if spec["bound"] == "tensor":
        pk = dmma_peak()
        dense_form = alg / (per_op_ms[dom] * 1e-3) * 1e-12
        # a centrosymmetric operator runs as two half-size products (even-odd decomposition): the
        # kernel EXECUTES half the flops of the dense form.  `frac` is on executed flops (what the
        # tensor pipe did); `frac_algorithmic` is SURVEY 8(d)'s definition on dense-form flops
        # (it exceeds the executed figure, and may exceed 1, exactly because half the work is avoided)
        ratio = eo_ratio[dom]
        achieved = dense_form * ratio
        roof = {"bound": "tensor", "kernel": label + (" [even-odd: two half-size DMMA products]" if ratio < 1.0 else ""),
                "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"],
                "frac_algorithmic": dense_form / pk["tflops"], "traffic": traffic,
                "peak_source": "FP64 DMMA m8n8k4 register-resident loop measured live on this GPU (pt_measure_dmma_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM 8192^3 measured beside it",
                "peak_crosscheck": {"cublas_dgemm_8192_tflops": pk.get("cublas_dgemm_8192_tflops"), "nominal_tflops": 37.0},
                "alg_flops_per_launch": alg, "executed_flops_per_launch": alg * ratio,
                "dense_form_tflops": dense_form, "even_odd": ratio < 1.0, "avg_launch_ms": per_op_ms[dom]}
        # whole step on executed flops
        tot = sum(o[4] * r for o, r in zip(spec["ops"], eo_ratio))
        roof["step_frac"] = tot / (sum(per_op_ms) * 1e-3) * 1e-12 / pk["tflops"]
    else:
        peaks = load_peaks()
        hbm = peaks.get("hbm_gbs", 6650.0)
        achieved = alg / (per_op_ms[dom] * 1e-3) * 1e-9
        roof = {"bound": "hbm", "kernel": label, "achieved": achieved, "peak": hbm, "unit": "GB/s",
                "frac": achieved / hbm, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "alg_bytes_per_launch": alg, "avg_launch_ms": per_op_ms[dom]}
    roof["per_op"] = [{"op": o[0], "avg_ms": m, "alg": o[4]} for o, m in zip(spec["ops"], per_op_ms)]
    return roof


def headline_line(ctx, args, spec):
    m = measure_plain(ctx, spec, args.steps, args.warmup, e2e=True, sustained_s=args.sustained)
    ranks = m["ranks"]
    world = ctx.world
    ms_per_step = max(r["ms_per_step"] for r in ranks)
    value = spec["points_per_step"] * world / (ms_per_step * 1e-3)
    e2e_ms = max(r["ms"] for r in m["e2e"])
    e2e_value = spec["points_per_step"] * world / (e2e_ms * 1e-3)
    if ctx.rank != 0:
        return None
    r0 = ranks[0]
    roof = roofline_record(spec, r0["per_op_ms"], m["eo_ratio"])
    slow = max(ranks, key=lambda r: r["ms_per_step"])
    nbytes = m["nbytes"]
    n_out = len(spec["ops"]) if spec["ops"][0][1] != "grad3" else 3
    links = [r["link"] for r in m["e2e"]]
    ideal_ms = max(l["both_ms"] for l in links)
    line = {
        "metric": "derivative grid-points/sec", "value": value, "unit": "points/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(spec),
        "roofline": roof,
        "cpu_baseline": None,
        "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": nbytes,
                "d2h_bytes_per_step": nbytes * n_out, "ms_per_step": e2e_ms, "steps": m["e2e"][0]["steps"],
                "api": "pytristan_b200.apply_host(ops, pinned_host_fields, pinned_host_outputs)",
                "per_rank_ms": [r["ms"] for r in m["e2e"]],
                "host_link": {"what": "the step's H2D and D2H bytes copied with NO kernels, all ranks at once, pinned buffers "
                                      "(per-rank GB/s; both_ms = both directions together = the floor of e2e ms_per_step)",
                              "h2d_gbs_per_rank": [round(l["h2d_gbs"], 2) for l in links],
                              "d2h_gbs_per_rank": [round(l["d2h_gbs"], 2) for l in links],
                              "both_gbs_per_rank": [round(l["both_gbs"], 2) for l in links],
                              "both_ms_max": ideal_ms, "aggregate_both_gbs": sum(l["both_gbs"] for l in links),
                              "e2e_over_link_floor": e2e_ms / ideal_ms}},
        "gpu_launches": args.steps * m["launches"] * world,
        "clocks": {k: r0.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "sm_mhz_min", "power_w_max")},
        "per_rank": ranks,
        "slowest_rank": {"rank": slow["rank"], "gpu": slow["gpu"], "ms_per_step": slow["ms_per_step"],
                         "sm_mhz": slow.get("sm_mhz"), "reasons": slow.get("reasons"),
                         "over_fastest": slow["ms_per_step"] / min(r["ms_per_step"] for r in ranks)},
        "host": {"cores_visible": host_cores(), "cores_of_rank0": len(ctx.cores)},
    }
    if m["sustained"]:
        sus = m["sustained"]
        sms = max(s["ms_per_step"] for s in sus)
        exec_flops = sum(o[4] * r for o, r in zip(spec["ops"], m["eo_ratio"]))
        rec = {"seconds": max(s["seconds"] for s in sus), "steps": min(s["steps"] for s in sus), "ms_per_step": sms,
               "value": spec["points_per_step"] * world / (sms * 1e-3), "over_burst": sms / ms_per_step,
               "per_rank": sus}
        if spec["bound"] == "tensor":
            rec["step_frac_executed"] = exec_flops / (sms * 1e-3) * 1e-12 / dmma_peak()["tflops"]
        line["sustained"] = rec
    return line


# --------------------------------------------------------------------------------------------------
# strong-scaling workloads whose differentiated axes are sharded (SURVEY.md 8e ii/iii)
# --------------------------------------------------------------------------------------------------
def measure_sharded(ctx, spec, comm, steps, warmup, fused=False):
    """Slab-sharded 3-axis stencil / three-axis Chebyshev operator on pencils over all ranks.
    Returns the record of rank 0 (None elsewhere)."""
    import torch

    import pytristan_b200 as pt
    from pytristan_b200 import dist as ptd

    ctx.need_group()
    rank, world = ctx.rank, ctx.world
    n = spec["n"]
    gen = torch.Generator(device="cuda").manual_seed(SEED + rank)
    peer = comm == "peer"
    fab = None
    if peer:
        from pytristan_b200.peer import PeerFabric

        fab = PeerFabric()
    compute_only = None
    if spec["kind"] == "slab":
        x = np.linspace(0.0, 1.0, n)
        F = pt.FinDiffMat(x, order=1, accuracy=6)
        plan = ptd.SlabStencil.from_operator(F, rank, world)
        fx, fy = pt.FinDiffMat(x, order=1, accuracy=6), pt.FinDiffMat(x, order=1, accuracy=6)
        fx.npts, fx.axis = (n, n, plan.nloc), 0
        fy.npts, fy.axis = (n, n, plan.nloc), 1
        fz = pt.FinDiffMat(np.linspace(0.0, 1.0, max(plan.nloc, 8)), order=1, accuracy=6)   # local stand-in: no halo
        fz.npts, fz.axis = (n, n, max(plan.nloc, 8)), 2
        outs = [torch.empty(plan.nloc * n * n, dtype=torch.float64, device="cuda") for _ in range(3)]
        field = torch.randn(1, plan.nloc, n * n, dtype=torch.float64, device="cuda", generator=gen)
        if peer:
            buf = plan.peer_field(fab, 1, n * n)       # the field lives in peer-mapped memory
            own = buf.local[:plan.nloc * n * n]
            own.view(1, plan.nloc, n * n).copy_(field)

            def step(i):
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

            def step(i):
                plan.exchange(ext)                          # halo planes over NVLink (NCCL send/recv)
                fx.apply(own.reshape(-1), out=outs[0])
                fy.apply(own.reshape(-1), out=outs[1])
                plan.apply(ext, out=outs[2].view(1, plan.nloc, n * n))

            launches = 5
        del field
        if plan.nloc >= 8:
            def compute_only(i):                        # the same kernels on the local slab: no halo, no barrier
                if fused and peer:
                    pt.gradient([fx, fy, fz], own.reshape(-1), outs=outs)
                    return
                fx.apply(own.reshape(-1), out=outs[0])
                fy.apply(own.reshape(-1), out=outs[1])
                fz.apply(own.reshape(-1), out=outs[2])
        per_pt = 32.0 if (fused and peer) else 48.0
        local_alg = per_pt * plan.nloc * n * n
        link_bytes = (plan.halo_lo + plan.halo_hi) * n * n * 8
        block_bytes = plan.nloc * n * n * 8
    else:
        py = 2 if world % 2 == 0 else 1
        pz = world // py
        xs = pt.cheb(n)
        ops = [pt.ChebMat(xs, order=2) for _ in range(3)]
        if peer:
            pen = ptd.PencilPeer((n, n, n), py, pz, fab)
            u = torch.randn(pen.nzl, pen.nyl, n, dtype=torch.float64, device="cuda", generator=gen)

            def step(i):
                return pen.apply_sum(ops, u)

            launches = 3 + 2 + 2              # 3 DMMA kernels (two with scatter epilogue), 2 peer scatters, 2 barriers
        else:
            pen = ptd.PencilTranspose((n, n, n), py, pz, rank, world)
            u = torch.randn(pen.nzl, pen.nyl, n, dtype=torch.float64, device="cuda", generator=gen)

            def step(i):
                return ptd.pencil_apply_sum(ops, u, pen, back=False)

            launches = 3 + 2 + 2 + 1          # 3 DMMA kernels, pack/unpack of two redistributions (z unpack only for batch)
        acc = torch.empty_like(u)
        u2 = torch.randn_like(u)

        def compute_only(i):                  # the three local applies on blocks of the pencil shapes, nothing moved
            ops[0]._apply_device(u.reshape(-1), pen.nzl * pen.nyl, pen.nx, 1, acc.reshape(-1), 1.0, 0.0, None, None)
            ops[1]._apply_device(u2.reshape(-1), pen.nzl, pen.ny, pen.nxl, acc.reshape(-1), 1.0, 1.0, None, None)
            ops[2]._apply_device(u.reshape(-1), 1, pen.nz, pen.nyl2 * pen.nxl, acc.reshape(-1), 1.0, 1.0, None, None)

        eo = all(bool(getattr(o, "even_odd", False)) for o in ops)
        local_alg = 3 * 2.0 * n * (n ** 3) / world * (0.5 if eo else 1.0)
        block_bytes = u.numel() * 8
        link_bytes = int(2 * block_bytes * ((py - 1) / py + (pz - 1) / pz))    # u and the running sum, two redistributions

    my_ms, sampler = time_steps(ctx, step, steps, warmup)
    comp_ms = None
    if compute_only is not None:
        comp_ms, _ = time_steps(ctx, compute_only, steps, 2)
    mine = {"rank": rank, "gpu": ctx.local, "ms_per_step": my_ms, "compute_only_ms": comp_ms, **sampler.summary()}
    ranks = ctx.gather(mine)
    if fab is not None:
        fab.check()
        fab.close()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    ms = max(r["ms_per_step"] for r in ranks)
    comp = max(r["compute_only_ms"] for r in ranks) if comp_ms is not None else None
    value = spec["points_per_step"] / (ms * 1e-3)
    if spec["bound"] == "hbm":
        peaks = load_peaks()
        peak = peaks.get("hbm_gbs", 6650.0)
        ach = local_alg / (ms * 1e-3) * 1e-9
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "note": "per-GPU algorithmic bytes of the whole step (3 derivatives: 16 B/pt each, or 32 B/pt for the fused "
                        "gradient) over the step time, halo traffic and barrier included"}
    else:
        pk = dmma_peak()["tflops"]
        ach = local_alg / (ms * 1e-3) * 1e-12
        roof = {"bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk, "traffic": None,
                "even_odd": eo, "dense_form_tflops": ach * (2.0 if eo else 1.0),
                "note": "per-GPU EXECUTED tensor flops of the three applies (half the dense form: even-odd kernel) over the "
                        "whole step time (redistribution and barriers included); peak = FP64 DMMA measured live"}
    how = {("slab", True): "halo planes loaded from the neighbours' memory inside the stencil kernel + 1 device flag barrier",
           ("slab", False): "NCCL send/recv halo exchange into an extended buffer",
           ("pencil", True): "redistribution fused into the DMMA epilogue + peer-store scatter over NVLink + 2 device flag barriers",
           ("pencil", False): "pack + all_to_all_single + unpack, twice"}[(spec["kind"], peer)]
    rec = {"workload": spec["name"], "comm": comm, "fused_gradient": bool(fused and peer and spec["kind"] == "slab"),
           "n_gpus": world, "scaling": "strong", "steps": steps, "ms_per_step": ms, "value": value, "unit": "points/s",
           "roofline": roof, "exchange": how, "nvlink_bytes_per_gpu_per_step": link_bytes if world > 1 else 0,
           "per_gpu_block_bytes": block_bytes,
           "l2": "per-rank field larger than L2" if block_bytes > L2_BYTES else "per-rank field fits L2 (strong scaling)",
           "per_rank_ms": [round(r["ms_per_step"], 4) for r in ranks], "gpu_launches": steps * launches * world,
           "clocks": {k: ranks[0].get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")}}
    if comp is not None:
        exposed = max(0.0, ms - comp)
        rec["compute_only_ms"] = comp
        rec["exposed_exchange_ms"] = exposed
        if world == 1:
            rec["limiter"] = "single GPU: no exchange; kernels only"
        elif exposed > 0.15 * ms:
            rec["limiter"] = f"the exchange ({how}): {exposed:.3f} of {ms:.3f} ms are not hidden behind the kernels " \
                             f"({link_bytes / 1e6:.0f} MB per GPU per step over NVLink = {link_bytes / max(exposed, 1e-6) * 1e-6:.0f} GB/s if it were all transfer)"
        else:
            rec["limiter"] = f"the kernels ({comp:.3f} of {ms:.3f} ms); the exchange is hidden to within {exposed:.3f} ms"
    return rec


def sharded_line(ctx, args, spec):
    rec = measure_sharded(ctx, spec, args.comm, args.steps, args.warmup, fused=args.fused_gradient)
    if rec is None:
        return None
    peer = args.comm == "peer"
    return {"metric": "derivative grid-points/sec", "value": rec["value"], "unit": "points/s", "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": spec["name"], "npts": list(spec["npts"]), "ops": [o[0] for o in spec["ops"]],
                       "sharding": ("z slabs, " if spec["kind"] == "slab" else "2-D pencils, ") + rec["exchange"],
                       "comm": args.comm, "fused_gradient": bool(args.fused_gradient and peer and spec["kind"] == "slab"),
                       "l2": rec["l2"]},
            "roofline": rec["roofline"], "cpu_baseline": None, "e2e": None, "gpu_launches": rec["gpu_launches"],
            "clocks": rec["clocks"], "detail": rec}


def extra_workloads(ctx, args, threads, recs):
    """The other BASELINE configs, measured in the same invocation: sharded paths at
    every N; configs 3, 4, 5a at full size on one GPU."""
    import torch

    steps = max(3, min(args.steps, 10))

    def guarded(label, fn):
        try:
            r = fn()
            if r is not None:
                recs.append(r)
        except Exception as exc:          # one failing extra must not lose the headline
            torch.cuda.empty_cache()
            if ctx.rank == 0:
                recs.append({"workload": label, "error": f"{type(exc).__name__}: {exc}"[:400]})

    for name, comm, fused in (("cheb512cube_pencil", "peer", False), ("cheb512cube_pencil", "nccl", False),
                              ("fd6_1024cube_slab", "peer", False), ("fd6_1024cube_slab", "peer", True),
                              ("fd6_1024cube_slab", "nccl", False)):
        spec = workload_spec(name)
        guarded(f"{name}/{comm}", lambda: measure_sharded(ctx, spec, comm, steps, 3, fused=fused))
    if ctx.world == 1:
        for name in ("polar_1024x512_b64", "polar_1024x512_b1", "fd6_1024cube", "fd6_1024cube_grad3", "cheb2048_b65536",
                     "channel_1024x257_b64_fft", "polar_1024x512_b64_fft"):
            spec = workload_spec(name)

            def one(spec=spec):
                # a second of idle first: the power-cap window of the workload before must not be charged to this one
                # (seen: sw_power_cap at 1680 MHz on a 20 ms polar measurement that followed the 1024^3 slabs)
                torch.cuda.synchronize()
                time.sleep(1.0)
                m = measure_plain(ctx, spec, steps, 3, e2e=False, graph=spec["name"].endswith("_b1"))
                r0 = m["ranks"][0]
                ms = r0["ms_per_step"]
                rec = {"workload": spec["name"], "n_gpus": 1, "scaling": "single GPU", "npts": list(spec["npts"]),
                       "batch": spec["batch"], "steps": steps, "ms_per_step": ms,
                       "value": spec["points_per_step"] / (ms * 1e-3), "unit": "points/s",
                       "roofline": roofline_record(spec, r0["per_op_ms"], m["eo_ratio"]),
                       "clocks": {k: r0.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples")},
                       "l2": config_dict(spec)["l2"], "gpu_launches": steps * m["launches"]}
                if m["graph_ms"] is not None:
                    # the eager step is bound by the host (~25 us of Python + ctypes per launch): same kernels,
                    # captured once per input buffer and replayed (pytristan_b200.graphs.CapturedApply)
                    rec["cuda_graph"] = {"ms_per_step": m["graph_ms"], "value": spec["points_per_step"] / (m["graph_ms"] * 1e-3),
                                         "unit": "points/s", "api": "pytristan_b200.graphs.CapturedApply(ops, u, outs).replay()"}
                if not args.no_cpu:
                    rec["cpu_baseline"] = cpu_baseline_record(spec, threads, steps=2, warmup=1)
                return rec

            guarded(name, one)

        def axis_solve():
            """SURVEY 8f rank 4 under the same clock: (D2 - k^2) y = f with Dirichlet rows along y of the config-2 grid
            (n = 257, 64 fields = 65 536 grid lines), factorisation form and one-pass inverse form."""
            import numpy as np
            import pytristan_b200 as pt
            npts, batch, k2 = (1024, 257), 64, 4.0
            x = pt.cheb(npts[1], -1.0, 1.0)
            A = np.asarray(pt.ChebMat(x, order=2)) - k2 * np.eye(npts[1])
            A[0] = 0.0; A[0, 0] = 1.0
            A[-1] = 0.0; A[-1, -1] = 1.0
            F = torch.randn(batch, npts[0] * npts[1], dtype=torch.float64, device="cuda")
            Y = torch.empty_like(F)
            R = torch.empty_like(F)
            rec = {"workload": "axis_solve_1024x257_b64", "n_gpus": 1, "scaling": "single GPU", "npts": list(npts),
                   "batch": batch, "what": "(D2 - 4) y = f, Dirichlet rows, along y for every grid line", "steps": steps}
            a_norm = float(np.max(np.sum(np.abs(A), axis=1)))
            torch.cuda.synchronize()
            time.sleep(1.0)
            for method in ("lu", "inverse"):
                t0 = time.perf_counter()
                solver = pt.AxisSolver(A, npts=npts, axis=1, method=method)
                torch.cuda.synchronize()
                setup = time.perf_counter() - t0
                for _ in range(3):
                    solver.solve(F, out=Y)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(steps):
                    solver.solve(F, out=Y)
                e1.record(); e1.synchronize()
                ms = e0.elapsed_time(e1) / steps
                R.copy_(F)
                solver.matrix.apply(Y, out=R, alpha=-1.0, beta=1.0)                                  # r = f - A y
                eta = float(R.abs().max() / (a_norm * Y.abs().max() + F.abs().max()))
                rec[method] = {"ms_per_solve": ms, "value": F.numel() / (ms * 1e-3), "unit": "points/s", "setup_s": setup,
                               "normwise_backward_error": eta,
                               "api": f"pytristan_b200.AxisSolver(A, npts, axis=1, method='{method}').solve(f, out=y)"}
            if not args.no_cpu:
                import scipy.linalg as sla
                set_blas_threads()
                lu = sla.lu_factor(A)
                sample = 4
                Fh = F[:sample].cpu().numpy().reshape(sample, npts[1], npts[0]).transpose(1, 0, 2).reshape(npts[1], -1)
                t0 = time.perf_counter()
                sla.lu_solve(lu, Fh)
                dt = time.perf_counter() - t0
                rec["cpu_baseline"] = {"value": sample * npts[0] * npts[1] / dt, "unit": "points/s", "cores": threads,
                                       "kind": "port", "sample": f"scipy.linalg.lu_solve (LAPACK getrs) on {sample} of the {batch} fields"}
            return rec

        guarded("axis_solve_1024x257_b64", axis_solve)


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
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no `workloads` record)")
    ap.add_argument("--extras-budget", type=float, default=420.0, help="wall-clock limit of the extra workloads, seconds")
    ap.add_argument("--sustained", type=float, default=2.0, help="seconds of the back-to-back `sustained` record (0: skip)")
    ap.add_argument("--no-even-odd", action="store_true",
                    help="apply centrosymmetric operators with the plain full-size kernel (A/B against the default)")
    ap.add_argument("--fused-gradient", action="store_true",
                    help="slab workload: the three derivatives in one read of the field (pt_stencil_grad3_halo)")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="sharded workloads: peer-memory kernels (default) or the NCCL exchange/all_to_all path")
    ap.add_argument("--deadline", type=float, default=1200.0,
                    help="seconds after which the process gives up (a stalled rank must not eat the caller's whole budget)")
    args = ap.parse_args()
    if args.deadline > 0:
        def give_up():
            sys.stderr.write(f"bench.py: no result after {args.deadline:.0f} s, giving up\n")
            sys.stderr.flush()
            os._exit(3)

        t = threading.Timer(args.deadline, give_up)
        t.daemon = True
        t.start()
    spec = workload_spec(args.workload)
    if args.impl == "reference":
        run_reference(args, spec)
        return
    if args.no_even_odd:
        import pytristan_b200.operators as ops_mod

        ops_mod.EVEN_ODD = False
    ctx = Ctx()
    if spec.get("kind") in ("slab", "pencil"):
        line = sharded_line(ctx, args, spec)
    else:
        line = headline_line(ctx, args, spec)
        threads = None
        if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
            threads = set_blas_threads()
            line["cpu_baseline"] = cpu_baseline_record(spec, threads)
        if args.workload == DEFAULT_WORKLOAD and not args.no_extras:
            # the extras contain collectives: if they ever stall, the headline must still be printed
            partial = []

            def bail():
                if line is not None:
                    line["workloads"] = partial + [{"error": f"extra workloads exceeded {args.extras_budget:.0f} s; stopped"}]
                    emit(line)
                os._exit(0)

            dog = threading.Timer(args.extras_budget, bail)
            dog.daemon = True
            dog.start()
            extra_workloads(ctx, args, threads, partial)
            dog.cancel()
            if line is not None:
                line["workloads"] = partial
    if line is not None:
        emit(line)
    ctx.close()


if __name__ == "__main__":
    main()
