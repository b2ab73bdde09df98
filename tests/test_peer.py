"""Peer-memory multi-GPU path (pytristan_b200.peer, SlabStencil.apply_peer, PencilPeer).

CPU: the destination maps of the pencil redistributions (pt_scatter_map arithmetic restated in
NumPy, all ranks emulated in one process) against the global transposed layouts.
GPU, one device: all ranks of the peer AND of the NCCL plans emulated in this process (threads, real
kernels on real device buffers; tests/_emulated_fabric.py) — no kernel ever spins on another launch.
GPU, >= 2 devices: one worker process per GPU over real CUDA IPC mappings and the device flag barrier."""
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


def _run_workers(world):
    port = free_port()
    procs = [subprocess.Popen([sys.executable, WORKER, str(r), str(world), str(port)],
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
@pytest.mark.parametrize("world", [2, 4])
def test_peer_paths_all_ranks_emulated_on_one_gpu(world):
    """SURVEY 8e (ii)/(iii) on a 1-GPU box: every rank of the slab stencil, the fused slab gradient and the
    pencil operator (scatter epilogue, pt_peer_scatter, way back) runs in this process — the real
    kernels on real device buffers; only the CUDA IPC mapping and the flag barrier are replaced
    (tests/_emulated_fabric.py) — against the oracle and, bit for bit, the single-device kernels."""
    from _emulated_fabric import run_ranks
    from _peer_checks import run_checks

    failures = {}

    def body(rank, w, fab):
        failures[rank] = run_checks(rank, w, fab)

    run_ranks(world, body)
    assert all(not failures[r] for r in range(world)), failures


@pytest.mark.gpu
@pytest.mark.parametrize("py,pz", [(2, 1), (1, 2), (2, 2), (4, 2)])
def test_nccl_pencil_plan_all_ranks_emulated_on_one_gpu(py, pz):
    """The pack -> all_to_all -> unpack form (PencilTranspose, pencil_apply_sum) with all ranks in this
    process: pt_permute3d and the DMMA kernels are real, only all_to_all_single is replaced by copies
    between the emulated ranks' buffers.  Result = the single-device sum, bit for bit, and the way back
    (z_to_y, y_to_x) restores the layout."""
    import threading

    import torch

    import pytristan_b200 as pt
    import pytristan_b200.dist as D
    from _emulated_fabric import current_rank, run_ranks

    world = py * pz
    NZ, NY, NX = 4 * pz, 6 * py * pz, 10 * py
    rng = np.random.default_rng(7)
    G = rng.standard_normal((NZ, NY, NX))
    mats = [pt.ChebMat(pt.cheb(n), order=1 + (k % 2)) for k, n in enumerate((NX, NY, NZ))]
    Gd = torch.from_numpy(G).cuda()
    full = mats[0]._apply_device(Gd.reshape(-1), NZ * NY, NX, 1, None, 1.0, 0.0, None, None)
    mats[1]._apply_device(Gd.reshape(-1), NZ, NY, NX, full, 1.0, 1.0, None, None)
    mats[2]._apply_device(Gd.reshape(-1), 1, NZ, NY * NX, full, 1.0, 1.0, None, None)
    full = full.view(NZ, NY, NX).cpu().numpy()

    posted, host = {}, threading.Barrier(world)

    def a2a(recv, send, group):          # group: the world ranks of the row / column, in order
        me = current_rank()
        idx, p = group.index(me), len(group)
        posted[(tuple(group), me)] = send.reshape(-1)
        host.wait()
        chunk = send.numel() // p
        out = recv.view(-1)
        for j, src in enumerate(group):
            out[j * chunk:(j + 1) * chunk].copy_(posted[(tuple(group), src)][idx * chunk:(idx + 1) * chunk])
        host.wait()                      # everybody has read before the send buffers are reused

    results = {}

    def body(rank, w, fab):
        ry, rz = rank % py, rank // py
        row = [rz * py + y for y in range(py)]
        col = [z * py + ry for z in range(pz)]
        pen = D.PencilTranspose((NZ, NY, NX), py, pz, rank=rank, world=w, groups=(row, col))
        u = torch.from_numpy(np.ascontiguousarray(G[rz * pen.nzl:(rz + 1) * pen.nzl, ry * pen.nyl:(ry + 1) * pen.nyl])).cuda()
        res = D.pencil_apply_sum(mats, u, pen, back=False)
        back = D.pencil_apply_sum(mats, u, pen, back=True)
        results[rank] = (res.cpu().numpy(), back.cpu().numpy(), pen)

    old = D.PencilTranspose._a2a
    D.PencilTranspose._a2a = staticmethod(a2a)
    try:
        run_ranks(world, body)
    except BaseException:
        host.abort()
        raise
    finally:
        D.PencilTranspose._a2a = old
    for rank in range(world):
        res, back, pen = results[rank]
        ry, rz = rank % py, rank // py
        slc = (slice(None), slice(rz * pen.nyl2, (rz + 1) * pen.nyl2), slice(ry * pen.nxl, (ry + 1) * pen.nxl))
        assert np.array_equal(res.reshape(full[slc].shape), full[slc]), ("layout C", rank)
        sla = (slice(rz * pen.nzl, (rz + 1) * pen.nzl), slice(ry * pen.nyl, (ry + 1) * pen.nyl), slice(None))
        assert np.array_equal(back.reshape(full[sla].shape), full[sla]), ("layout A", rank)


@pytest.mark.gpu
def test_peer_paths_one_rank_per_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    _run_workers(4 if n >= 4 else 2)
