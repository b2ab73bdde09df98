"""Worker run by the distributed tests (gloo on CPU, nccl on GPUs): checks SlabStencil and
PencilTranspose against single-process oracle results.  Launched with torch.multiprocessing.spawn
(CPU) or torchrun (GPU)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def install_cpu_doubles():
    """CPU stand-ins for the two GPU entry points of pytristan_b200.dist (tests only)."""
    import torch

    import pytristan_b200.dist as D
    from oracle import operators as oracle

    def permute3d(src, dst, d2, d1, d0, perm):
        dst.copy_(src.view(d2, d1, d0).permute(*perm).contiguous().view(-1))

    def stencil_local(band, ext, out, n_in, n_out, shift, after, before, alpha, beta):
        e = ext.numpy().reshape(after, n_in, before)
        res = np.zeros((after, n_out, before))
        for s in range(band["wmax"]):
            j = band["start"].astype(np.int64) + s
            res += band["weights"][:, s][None, :, None] * e[:, j, :]
        o = out.numpy().reshape(after, n_out, before)
        o[...] = alpha * res + (beta * o if beta != 0.0 else 0.0)

    D._permute3d = permute3d
    D._stencil_local = stencil_local


def run_checks(rank, world, device):
    import torch
    import torch.distributed as dist

    import pytristan_b200.dist as D
    from oracle import operators as oracle

    cuda = device == "cuda"
    rng = np.random.default_rng(1234)          # same stream on every rank: global data is identical
    failures = []

    # ---- slab-sharded stencil along the slowest axis ------------------------------------------------
    nx, ny, nz, nf = 6, 5, 32 if world <= 4 else 64, 2
    for (order, acc, periodic, mapped) in [(1, 6, False, False), (2, 6, False, False), (1, 4, True, False), (1, 4, False, True)]:
        z = np.linspace(0.0, 1.0, nz)
        if mapped:
            z = oracle.cheb_nodes(nz, 0.0, 1.0)
        if periodic:
            z = np.linspace(0.0, 1.0 - 1.0 / nz, nz)
        band = oracle.findiff_band(z, order, acc, periodic)
        U = rng.standard_normal((nf, nz, ny * nx))
        ref = oracle.stencil_apply(band, U.reshape(nf, -1), (nx, ny, nz), 2).reshape(nf, nz, ny * nx)
        plan = D.SlabStencil(band["start"], band["W"], rank, world, periodic=periodic, wuni=band["wuni"], w=band["w"],
                             lo=band["lo"], hi=band["hi"], device=cuda)
        ext = torch.zeros(plan.ext_shape(nf, ny * nx), dtype=torch.float64, device=device)
        plan.interior(ext).copy_(torch.from_numpy(U[:, plan.z0:plan.z1]).to(device))
        plan.exchange(ext)
        out = plan.apply(ext)
        got = out.cpu().numpy()
        err = np.max(np.abs(got - ref[:, plan.z0:plan.z1])) / np.max(np.abs(ref))
        if not err <= 1e-12:
            failures.append(("slab", order, acc, periodic, mapped, float(err)))
        # the halo rows must hold the neighbours' rows exactly
        e = ext.cpu().numpy()
        idx = (np.arange(plan.z0 - plan.halo_lo, plan.z1 + plan.halo_hi)) % nz
        if not np.array_equal(e, U[:, idx]):
            failures.append(("halo", order, acc, periodic, mapped))

    # ---- pencil transposes ------------------------------------------------------------------------------
    py = 2 if world % 2 == 0 else 1
    pz = world // py
    NZ, NY, NX = 4 * pz, 2 * py * pz, 6 * py
    G = rng.standard_normal((3, NZ, NY, NX))             # batch of 3 global fields
    pt = D.PencilTranspose((NZ, NY, NX), py, pz, rank, world)
    ry, rz = pt.ry, pt.rz
    A = G[:, rz * pt.nzl:(rz + 1) * pt.nzl, ry * pt.nyl:(ry + 1) * pt.nyl, :]
    B = G[:, rz * pt.nzl:(rz + 1) * pt.nzl, :, ry * pt.nxl:(ry + 1) * pt.nxl]
    C = G[:, :, rz * pt.nyl2:(rz + 1) * pt.nyl2, ry * pt.nxl:(ry + 1) * pt.nxl]
    tA = torch.from_numpy(np.ascontiguousarray(A)).to(device)
    tB = pt.x_to_y(tA)
    if not np.array_equal(tB.cpu().numpy(), B):
        failures.append(("x_to_y",))
    tC = pt.y_to_z(tB)
    if not np.array_equal(tC.cpu().numpy(), C):
        failures.append(("y_to_z",))
    if not np.array_equal(pt.z_to_y(tC).cpu().numpy(), B):
        failures.append(("z_to_y",))
    if not np.array_equal(pt.y_to_x(tB).cpu().numpy(), A):
        failures.append(("y_to_x",))
    one = pt.y_to_z(pt.x_to_y(tA[0].contiguous()))           # no batch dimension
    if not np.array_equal(one.cpu().numpy(), C[0]):
        failures.append(("single field",))

    # ---- three-axis operator sum on pencils (GPU only: needs the dense kernels) ---------------------------------
    if cuda:
        import pytristan_b200 as ptb
        from pytristan_b200.operators import _DiffMat

        mats = [rng.standard_normal((n, n)) for n in (NX, NY, NZ)]
        ops = [_DiffMat._wrap(torch.from_numpy(m).cuda(), (n,), 0, 1, None) for m, n in zip(mats, (NX, NY, NZ))]
        g = G[0]
        ref = (np.einsum("ij,zyj->zyi", mats[0], g) + np.einsum("ij,zjx->zix", mats[1], g) + np.einsum("ij,jyx->iyx", mats[2], g))
        res = D.pencil_apply_sum(ops, tA[0].contiguous(), pt, back=True)
        err = np.max(np.abs(res.cpu().numpy() - ref[rz * pt.nzl:(rz + 1) * pt.nzl, ry * pt.nyl:(ry + 1) * pt.nyl])) / np.max(np.abs(ref))
        if not err <= 1e-12:
            failures.append(("pencil_apply_sum", float(err)))

    # ---- field sharding is a partition -------------------------------------------------------------------------
    cover = []
    for r in range(world):
        cover += list(range(*D.shard_range(37, r, world)))
    if cover != list(range(37)):
        failures.append(("shard_range",))

    flag = torch.tensor([len(failures)], dtype=torch.int64, device=device)
    dist.all_reduce(flag)
    return failures, int(flag.item())


def cpu_entry(rank, world, port):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    install_cpu_doubles()
    failures, total = run_checks(rank, world, "cpu")
    dist.destroy_process_group()
    if failures or total:
        print(f"rank {rank}: {failures}", flush=True)
        sys.exit(1)


def main():
    """torchrun entry (GPU, nccl)."""
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures, total = run_checks(rank, world, "cuda")
    dist.destroy_process_group()
    if failures or total:
        print(f"rank {rank}: FAIL {failures}", flush=True)
        sys.exit(1)
    if rank == 0:
        print(f"multi-GPU checks ok on {world} ranks", flush=True)


if __name__ == "__main__":
    main()
