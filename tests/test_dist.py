"""Multi-rank paths: world-size-2 (and 4) gloo runs on CPU for the host-side logic (halo
bookkeeping, pencil redistributions, collectives), and the same checks under nccl when at least two
GPUs are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import _dist_worker as worker


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_world(world):
    import torch.multiprocessing as mp

    mp.spawn(worker.cpu_entry, args=(world, free_port()), nprocs=world, join=True)


def test_shard_range_and_plan_shapes():
    from oracle import operators as oracle
    from pytristan_b200.dist import SlabStencil, shard_range

    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    band = oracle.findiff_band(np.linspace(0, 1, 1024), 1, 6)
    plans = [SlabStencil(band["start"], band["W"], r, 8, wuni=band["wuni"], w=7, lo=band["lo"], hi=band["hi"], device=False)
             for r in range(8)]
    assert [(p.halo_lo, p.halo_hi) for p in plans] == [(0, 3)] + [(3, 3)] * 6 + [(3, 0)]
    assert [(p.send_lo, p.send_hi) for p in plans] == [(0, 3)] + [(3, 3)] * 6 + [(3, 0)]
    assert all(p.nloc == 128 for p in plans)
    assert (plans[0].band["lo"], plans[0].band["hi"]) == (3, 128) and (plans[7].band["lo"], plans[7].band["hi"]) == (0, 125)
    assert (plans[3].band["lo"], plans[3].band["hi"]) == (0, 128) and plans[3].shift == 3
    with pytest.raises(ValueError):
        SlabStencil(band["start"][:16], band["W"][:16], 0, 8, device=False)        # 2-row slabs < halo


@pytest.mark.gpu
def test_nccl_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    n = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(os.path.dirname(__file__), "_dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "multi-GPU checks ok" in res.stdout
