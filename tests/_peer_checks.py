"""The peer-memory checks themselves, shared by the real multi-process worker (tests/_peer_worker.py:
CUDA IPC mappings + device flag barrier, one rank per GPU) and the single-GPU emulation
(tests/test_peer.py::test_peer_paths_all_ranks_emulated_on_one_gpu, tests/_emulated_fabric.py):
SlabStencil.apply_peer / gradient_peer and PencilPeer.apply_sum / to_layout_a against the NumPy
oracle and, bit for bit, against the single-device kernels."""
import numpy as np


def run_checks(rank, world, fab):
    """Returns the list of failures of this rank (empty = all good).  Collective: every rank calls it."""
    import torch

    import pytristan_b200 as pt
    import pytristan_b200.dist as D
    from oracle import operators as oracle

    rng = np.random.default_rng(4321)          # same stream on every rank: global data is identical
    failures = []

    # ---- barrier + raw peer stores --------------------------------------------------------------------
    import ctypes

    from pytristan_b200 import _lib

    box = fab.alloc(world * 4)
    src = torch.tensor([100.0 * rank + r for r in range(world) for _ in range(4)], dtype=torch.float64, device="cuda")
    smap = fab.scatter_map(box, list(range(world)), 4, 1, 0, 0, 1, 4 * rank)
    _lib.check(_lib.load().pt_peer_scatter(_lib.dptr(src), 4 * world, 1, 1, ctypes.byref(smap), _lib.stream_ptr()),
               "pt_peer_scatter")
    fab.barrier()
    got = box.local.cpu().numpy().reshape(world, 4)
    if not np.array_equal(got, np.array([[100.0 * s + rank] * 4 for s in range(world)])):
        failures.append(("barrier/peer stores", got.tolist()))
    fab.barrier()

    # ---- slab-sharded stencil, halo rows loaded from the neighbours ---------------------------------------
    nx, ny, nz, nf = 6, 5, 32 if world <= 4 else 64, 2
    for (order, acc, periodic, mapped) in [(1, 6, False, False), (2, 6, False, False), (1, 4, True, False),
                                           (1, 4, False, True)]:
        z = np.linspace(0.0, 1.0, nz)
        if mapped:
            z = oracle.cheb_nodes(nz, 0.0, 1.0)
        if periodic:
            z = np.linspace(0.0, 1.0 - 1.0 / nz, nz)
        band = oracle.findiff_band(z, order, acc, periodic)
        U = rng.standard_normal((nf, nz, ny * nx))
        ref = oracle.stencil_apply(band, U.reshape(nf, -1), (nx, ny, nz), 2).reshape(nf, nz, ny * nx)
        plan = D.SlabStencil(band["start"], band["W"], rank, world, periodic=periodic, wuni=band["wuni"], w=band["w"],
                             lo=band["lo"], hi=band["hi"])
        buf = plan.peer_field(fab, nf, ny * nx)
        buf.local[:nf * plan.nloc * ny * nx].view(nf, plan.nloc, ny * nx).copy_(
            torch.from_numpy(U[:, plan.z0:plan.z1]).cuda())
        fab.barrier()
        out = torch.empty(nf, plan.nloc, ny * nx, dtype=torch.float64, device="cuda")
        plan.apply_peer(buf, nf, ny * nx, out)
        got = out.cpu().numpy()
        err = np.max(np.abs(got - ref[:, plan.z0:plan.z1])) / np.max(np.abs(ref))
        if not err <= 1e-12:
            failures.append(("slab_peer", order, acc, periodic, mapped, float(err)))
        # the exchange-based path must give the same bits
        ext = torch.zeros(plan.ext_shape(nf, ny * nx), dtype=torch.float64, device="cuda")
        plan.interior(ext).copy_(torch.from_numpy(U[:, plan.z0:plan.z1]).cuda())
        idx = (np.arange(plan.z0 - plan.halo_lo, plan.z1 + plan.halo_hi)) % nz
        ext.copy_(torch.from_numpy(U[:, idx]).cuda())
        if not torch.equal(plan.apply(ext), out):
            failures.append(("slab_peer != slab_exchange bits", order, acc, periodic, mapped))
        fab.barrier()

    # ---- fused gradient of a slab-sharded field (one read, z halos from the neighbours) -----------------
    gx_n, gy_n, gz_n = 16, 18, 14 * world
    axes = [np.linspace(0.0, 1.0, m) for m in (gx_n, gy_n, gz_n)]
    fops = []
    for k in range(3):
        F = pt.FinDiffMat(axes[k], order=1, accuracy=6)
        F.npts, F.axis = (gx_n, gy_n, gz_n), k
        fops.append(F)
    G = rng.standard_normal((gz_n, gy_n * gx_n))
    plan = D.SlabStencil.from_operator(fops[2], rank, world)
    buf = plan.peer_field(fab, 1, gy_n * gx_n)
    buf.local[:plan.nloc * gy_n * gx_n].view(plan.nloc, gy_n * gx_n).copy_(torch.from_numpy(G[plan.z0:plan.z1]).cuda())
    fab.barrier()
    outs = [torch.full((plan.nloc * gy_n * gx_n,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3)]
    if not plan.gradient_peer(buf, fops[0], fops[1], gx_n, gy_n, outs):
        failures.append(("gradient_peer declined",))
    full = pt.gradient(fops, torch.from_numpy(G).cuda().reshape(-1))          # single-device fused kernel
    for k in range(3):
        band = oracle.findiff_band(axes[k], 1, 6)
        ref = oracle.stencil_apply(band, G.reshape(1, -1), (gx_n, gy_n, gz_n), k).reshape(gz_n, -1)[plan.z0:plan.z1]
        got = outs[k].cpu().numpy().reshape(plan.nloc, -1)
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        if not err <= 1e-12:
            failures.append(("gradient_peer", k, float(err)))
        if not np.array_equal(got, full[k].cpu().numpy().reshape(gz_n, -1)[plan.z0:plan.z1]):
            failures.append(("gradient_peer != single-device bits", k))
    fab.barrier()

    # ---- pencils over peer memory -----------------------------------------------------------------------
    for (py, pz) in sorted({(world, 1), (1, world), (2 if world % 2 == 0 else 1, world // (2 if world % 2 == 0 else 1))}):
        for batch in (1, 2):
            NZ, NY, NX = 4 * pz, 6 * py * pz, 10 * py
            G = rng.standard_normal((batch, NZ, NY, NX))
            mats = [pt.ChebMat(pt.cheb(n), order=1 + (k % 2)) for k, n in enumerate((NX, NY, NZ))]
            pp = D.PencilPeer((NZ, NY, NX), py, pz, fab, batch=batch)
            ry, rz = rank % py, rank // py
            u = torch.from_numpy(np.ascontiguousarray(
                G[:, rz * pp.nzl:(rz + 1) * pp.nzl, ry * pp.nyl:(ry + 1) * pp.nyl, :])).cuda()
            for rep in range(2):                     # twice: buffer reuse across steps
                res = pp.apply_sum(mats, u).cpu().numpy()
            fab.check()
            Dx, Dy, Dz = (np.asarray(m) for m in mats)
            ref = (np.einsum("ij,bzyj->bzyi", Dx, G) + np.einsum("ij,bzjx->bzix", Dy, G)
                   + np.einsum("ij,bjyx->biyx", Dz, G))
            scale = (np.einsum("ij,bzyj->bzyi", np.abs(Dx), np.abs(G)) + np.einsum("ij,bzjx->bzix", np.abs(Dy), np.abs(G))
                     + np.einsum("ij,bjyx->biyx", np.abs(Dz), np.abs(G)))
            sl = (slice(None), slice(None), slice(rz * pp.nyl2, (rz + 1) * pp.nyl2), slice(ry * pp.nxl, (ry + 1) * pp.nxl))
            err = np.max(np.abs(res - ref[sl]) / scale[sl])
            if not err <= 1e-12:
                failures.append(("pencil_peer", py, pz, batch, float(err)))
            # the way back: layout C -> B -> A
            back = pp.to_layout_a(pp.apply_sum(mats, u)).cpu().numpy()
            sla = (slice(None), slice(rz * pp.nzl, (rz + 1) * pp.nzl), slice(ry * pp.nyl, (ry + 1) * pp.nyl), slice(None))
            if not np.max(np.abs(back - ref[sla]) / scale[sla]) <= 1e-12:
                failures.append(("pencil_peer to_layout_a", py, pz, batch))
            # single-device computation of the same sum: must agree bit for bit (same summation order)
            Gd = torch.from_numpy(G).cuda()
            full = mats[0]._apply_device(Gd.reshape(-1), batch * NZ * NY, NX, 1, None, 1.0, 0.0, None, None)
            mats[1]._apply_device(Gd.reshape(-1), batch * NZ, NY, NX, full, 1.0, 1.0, None, None)
            mats[2]._apply_device(Gd.reshape(-1), batch, NZ, NY * NX, full, 1.0, 1.0, None, None)
            if not np.array_equal(full.view(batch, NZ, NY, NX).cpu().numpy()[sl], res):
                failures.append(("pencil_peer != single-device bits", py, pz, batch))
            fab.barrier()

    # ---- pencils large enough for the warp-specialised kernel: the field redistribution is fused into
    # the DMMA kernel (pt_dense_apply_eo_ws_scatter_copy); same bits as the single-device kernels -------------
    if world == 2:
        import pytristan_b200.operators as ops_mod

        for (py, pz) in ((2, 1), (1, 2)):
            NZ, NY, NX = 128, 256, 256
            G = rng.standard_normal((1, NZ, NY, NX))
            mats = [pt.ChebMat(pt.cheb(n), order=1 + (k % 2)) for k, n in enumerate((NX, NY, NZ))]
            pp = D.PencilPeer((NZ, NY, NX), py, pz, fab, batch=1)
            ry, rz = rank % py, rank // py
            u = torch.from_numpy(np.ascontiguousarray(
                G[:, rz * pp.nzl:(rz + 1) * pp.nzl, ry * pp.nyl:(ry + 1) * pp.nyl, :])).cuda()
            before_count = ops_mod.STATS["fused_field_copies"]
            for rep in range(2):
                res = pp.apply_sum(mats, u).cpu().numpy()
            fab.check()
            fused = ops_mod.STATS["fused_field_copies"] - before_count
            if torch.cuda.get_device_properties(0).multi_processor_count <= 148 and fused < 2:
                failures.append(("large pencil: the field copy was not fused", py, pz, fused))
            Gd = torch.from_numpy(G).cuda()
            full = mats[0]._apply_device(Gd.reshape(-1), NZ * NY, NX, 1, None, 1.0, 0.0, None, None)
            mats[1]._apply_device(Gd.reshape(-1), NZ, NY, NX, full, 1.0, 1.0, None, None)
            mats[2]._apply_device(Gd.reshape(-1), 1, NZ, NY * NX, full, 1.0, 1.0, None, None)
            sl = (slice(None), slice(None), slice(rz * pp.nyl2, (rz + 1) * pp.nyl2), slice(ry * pp.nxl, (ry + 1) * pp.nxl))
            if not np.array_equal(full.view(1, NZ, NY, NX).cpu().numpy()[sl], res):
                failures.append(("large pencil (fused field copy) != single-device bits", py, pz))
            # the moved field itself: layout C holds u
            uC = pp.uC.local.view(1, NZ, pp.nyl2, pp.nxl).cpu().numpy()
            if not np.array_equal(uC, G[sl]):
                failures.append(("large pencil: redistributed field differs", py, pz))
            fab.barrier()
    return failures
