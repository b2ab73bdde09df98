"""Worker of the multi-GPU peer-memory test: the checks of tests/_peer_checks.py over REAL peer mappings
(CUDA IPC) and the device flag barrier, one rank per GPU, NCCL for the handle exchange.

Usage: python _peer_worker.py RANK WORLD PORT
(Several ranks are never run as processes on ONE GPU: kernels that spin on a flag another launch
writes must not share a GPU — B200_PROFILING.md.  On one GPU all ranks are emulated in one
process instead: tests/_emulated_fabric.py.)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))

    from pytristan_b200.peer import PeerFabric

    from _peer_checks import run_checks

    fab = PeerFabric(timeout_s=30.0)
    failures = run_checks(rank, world, fab)
    torch.cuda.synchronize()
    fab.check()
    if failures:
        print(f"rank {rank} FAILURES: {failures}", flush=True)
    fab.close()
    dist.destroy_process_group()
    if failures:
        sys.exit(1)
    if rank == 0:
        print("peer-memory checks ok", flush=True)


if __name__ == "__main__":
    main()
