"""All ranks of a peer-memory job emulated in ONE process on ONE GPU (test infrastructure).

``EmulatedFabric.create(world)`` returns one fabric object per emulated rank with the interface of
``pytristan_b200.peer.PeerFabric`` (``rank``, ``world``, ``alloc``, ``scatter_map``, ``barrier``,
``check``, ``close``).  Every rank runs in its own Python thread; the "peer" buffers are plain
device allocations visible to all of them, and ``barrier()`` is a host barrier: every rank has
ENQUEUED its kernels of the phase before any rank enqueues the next phase, and all threads launch
on the same (legacy default) stream, so stream order gives the visibility the device flag barrier
gives between real ranks.  No kernel ever waits for another kernel — the multi-rank spin barrier
must not be run as several launches on one GPU (B200_PROFILING.md) — while the scatter maps, the DMMA
scatter epilogue, pt_peer_scatter and the halo-loading stencil kernels are the real ones.
"""
import threading

from pytristan_b200.peer import PeerFabric


class _Buffer:
    def __init__(self, tensors, rank):
        import torch

        self._tensors = tensors
        self.local = tensors[rank]
        self.ptrs = [t.data_ptr() for t in tensors]
        self.count = tensors[rank].numel()
        self._tables = {}
        self._torch = torch

    def table(self, ranks=None):
        key = tuple(range(len(self.ptrs))) if ranks is None else tuple(int(r) for r in ranks)
        if key not in self._tables:
            self._tables[key] = self._torch.tensor([self.ptrs[r] for r in key], dtype=self._torch.int64, device="cuda")
        return self._tables[key]

    def peer_view(self, rank, count=None):
        t = self._tensors[rank]
        return t if count is None else t[:count]

    def close(self):
        pass


class _Shared:
    def __init__(self, world):
        self.world = world
        self.lock = threading.Lock()
        self.host_barrier = threading.Barrier(world)
        self.allocs = []          # k-th collective allocation -> list of per-rank tensors


class EmulatedFabric:
    def __init__(self, shared, rank):
        self._s = shared
        self.rank, self.world = rank, shared.world
        self._n = 0

    @classmethod
    def create(cls, world):
        shared = _Shared(world)
        return [cls(shared, r) for r in range(world)]

    def alloc(self, count):
        import torch

        with self._s.lock:
            if len(self._s.allocs) <= self._n:
                self._s.allocs.append([torch.zeros(max(1, int(count)), dtype=torch.float64, device="cuda")
                                       for _ in range(self.world)])
            tensors = self._s.allocs[self._n]
        self._n += 1
        return _Buffer(tensors, self.rank)

    scatter_map = PeerFabric.scatter_map

    def barrier(self):
        self._s.host_barrier.wait()

    def check(self):
        pass

    def close(self):
        pass


_local = threading.local()


def current_rank():
    """Emulated rank of the calling thread (inside run_ranks)."""
    return _local.rank


def run_ranks(world, fn):
    """Run ``fn(rank, world, fabric)`` for every emulated rank in its own thread; re-raise the first
    failure.  A rank that dies breaks the host barrier so the others do not hang."""
    import torch

    fabs = EmulatedFabric.create(world)
    errors = [None] * world
    dev = torch.cuda.current_device()

    def body(r):
        try:
            torch.cuda.set_device(dev)
            _local.rank = r
            fn(r, world, fabs[r])
        except BaseException as exc:      # noqa: BLE001
            errors[r] = exc
            fabs[r]._s.host_barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    torch.cuda.synchronize()
    real = [e for e in errors if e is not None and not isinstance(e, threading.BrokenBarrierError)]
    if real:
        raise real[0]
    broken = [e for e in errors if e is not None]
    if broken:
        raise broken[0]
