"""Peer-memory multi-GPU path (pytristan_b200.peer, SlabStencil.apply_peer, PencilPeer).

CPU: the destination maps of the pencil redistributions (pt_scatter_map arithmetic restated in
NumPy, all ranks emulated in one process) against the global transposed layouts.
GPU: two worker processes sharing ONE GPU map each other's buffers through CUDA IPC and run the
real kernels (scatter epilogue, peer scatter, halo loads, flag barrier); with >= 2 GPUs the same
worker runs one rank per GPU."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

WORKER = os.path.join(os.path.dirname(__file__), "_peer_worker.py")


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeFabric:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world


@pytest.mark.parametrize("py,pz,batch", [(2, 1, 1), (1, 2, 2), (2, 2, 1), (2, 4, 2), (4, 2, 1)])
def test_scatter_maps_reproduce_the_pencil_layouts(py, pz, batch):
    from pytristan_b200.dist import PencilPeer, PencilTranspose
    from pytristan_b200.peer import scatter_dest

    world = py * pz
    NZ, NY, NX = 2 * pz, 3 * py * pz, 5 * py
    G = np.arange(batch * NZ * NY * NX, dtype=np.float64).reshape(batch, NZ, NY, NX)
    plans = []
    for r in range(world):
        pp = PencilPeer.__new__(PencilPeer)
        PencilTranspose.__init__(pp, (NZ, NY, NX), py, pz, rank=r, world=world, groups=(None, None))
        pp.batch = batch
        plans.append(pp)
    count = batch * NZ * NY * NX // world
    bufB = [np.full(count, np.nan) for _ in range(world)]
    bufC = [np.full(count, np.nan) for _ in range(world)]

    # A -> B: every rank scatters its layout-A block into the row peers' layout-B buffers
    for r, pp in enumerate(plans):
        blockA = G[:, pp.rz * pp.nzl:(pp.rz + 1) * pp.nzl, pp.ry * pp.nyl:(pp.ry + 1) * pp.nyl, :]
        after, n = batch * pp.nzl * pp.nyl, NX
        a, i = np.meshgrid(np.arange(after), np.arange(n), indexing="ij")
        q, off = scatter_dest(a, i, 0, *pp.map_ab())
        row = [pp.rz * py + y for y in range(py)]
        flat = blockA.reshape(after, n)
        for dst in range(py):
            sel = q == dst
            bufB[row[dst]][off[sel]] = flat[sel]
    for r, pp in enumerate(plans):
        want = G[:, pp.rz * pp.nzl:(pp.rz + 1) * pp.nzl, :, pp.ry * pp.nxl:(pp.ry + 1) * pp.nxl]
        np.testing.assert_array_equal(bufB[r].reshape(want.shape), want)

    # B -> C
    for r, pp in enumerate(plans):
        blockB = bufB[r].reshape(batch * pp.nzl, NY, pp.nxl)
        a, i, b = np.meshgrid(np.arange(batch * pp.nzl), np.arange(NY), np.arange(pp.nxl), indexing="ij")
        q, off = scatter_dest(a, i, b, *pp.map_bc())
        col = [z * py + pp.ry for z in range(pz)]
        for dst in range(pz):
            sel = q == dst
            bufC[col[dst]][off[sel]] = blockB[sel]
    for r, pp in enumerate(plans):
        want = G[:, :, pp.rz * pp.nyl2:(pp.rz + 1) * pp.nyl2, pp.ry * pp.nxl:(pp.ry + 1) * pp.nxl]
        np.testing.assert_array_equal(bufC[r].reshape(want.shape), want)

    # the way back: C -> B -> A must reproduce the layout-B and layout-A blocks
    backB = [np.full(count, np.nan) for _ in range(world)]
    backA = [np.full(count, np.nan) for _ in range(world)]
    for r, pp in enumerate(plans):
        blockC = bufC[r].reshape(batch, NZ, pp.nyl2 * pp.nxl)
        a, i, b = np.meshgrid(np.arange(batch), np.arange(NZ), np.arange(pp.nyl2 * pp.nxl), indexing="ij")
        q, off = scatter_dest(a, i, b, *pp.map_cb())
        col = [z * py + pp.ry for z in range(pz)]
        for dst in range(pz):
            sel = q == dst
            backB[col[dst]][off[sel]] = blockC[sel]
    for r in range(world):
        np.testing.assert_array_equal(backB[r], bufB[r])
    for r, pp in enumerate(plans):
        blockB = backB[r].reshape(batch * pp.nzl, NY, pp.nxl)
        a, i, b = np.meshgrid(np.arange(batch * pp.nzl), np.arange(NY), np.arange(pp.nxl), indexing="ij")
        q, off = scatter_dest(a, i, b, *pp.map_ba())
        row = [pp.rz * py + y for y in range(py)]
        for dst in range(py):
            sel = q == dst
            backA[row[dst]][off[sel]] = blockB[sel]
    for r, pp in enumerate(plans):
        want = G[:, pp.rz * pp.nzl:(pp.rz + 1) * pp.nzl, pp.ry * pp.nyl:(pp.ry + 1) * pp.nyl, :]
        np.testing.assert_array_equal(backA[r].reshape(want.shape), want)


def _run_workers(world, same_gpu):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port)] + (["same-gpu"] if same_gpu else []),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            out, _ = p.communicate(timeout=420)
            outs.append(out)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2500:] for o in outs)
    assert "peer-memory checks ok" in outs[0]


@pytest.mark.gpu
def test_peer_paths_two_processes_one_gpu():
    _run_workers(2, same_gpu=True)


@pytest.mark.gpu
def test_peer_paths_one_rank_per_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    _run_workers(4 if n >= 4 else 2, same_gpu=False)
