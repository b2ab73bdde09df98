"""Peer-memory fabric of the multi-GPU path (one process per GPU on one NVSwitch box).

Every rank allocates its exchange buffers through the C-ABI (``pt_peer_alloc``), the 64-byte CUDA
IPC handles travel once through ``torch.distributed.all_gather_object`` and every rank maps the
buffers of all the others.  After that the data path has no collective call at all: kernels store
into / load from the peers' buffers over NVLink (``pt_dense_apply_scatter``, ``pt_peer_scatter``,
``pt_stencil_apply_halo``) and synchronise with a device-side flag barrier (``pt_peer_barrier``).

``torch.distributed`` is used for the handle exchange only, so any backend works (NCCL on a
multi-GPU box; gloo lets two processes share ONE GPU, which is how the single-GPU test exercises
the real peer kernels).
"""
import ctypes

import numpy as np

from . import _lib

__all__ = ["PeerFabric", "PeerBuffer"]


class _CudaArray:
    """Minimal ``__cuda_array_interface__`` carrier for a raw device pointer."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _as_tensor(ptr, count, typestr="<f8"):
    torch = _lib.torch_mod()
    return torch.as_tensor(_CudaArray(ptr, count, typestr), device="cuda")


class PeerBuffer:
    """One symmetric allocation: ``count`` doubles on every rank.

    ``local``   torch view of this rank's buffer,
    ``ptrs``    device address of every rank's buffer as mapped in THIS process,
    ``table(ranks)`` device array of those addresses for a subset of ranks (cached).
    """

    def __init__(self, fabric, count, typestr="<f8", itemsize=8):
        lib = _lib.load()
        self.fabric, self.count = fabric, int(count)
        nbytes = max(256, self.count * itemsize)
        raw = ctypes.c_void_p()
        _lib.check(lib.pt_peer_alloc(nbytes, ctypes.byref(raw)), "pt_peer_alloc")
        self._own = raw.value
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.pt_ipc_export(ctypes.c_void_p(self._own), handle), "pt_ipc_export")
        handles = fabric._all_gather(bytes(handle.raw))
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == fabric.rank:
                self.ptrs.append(self._own)
                continue
            q = ctypes.c_void_p()
            _lib.check(lib.pt_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)), "pt_ipc_open")
            self.ptrs.append(q.value)
            self._opened.append(q.value)
        self.local = _as_tensor(self._own, self.count, typestr)
        self._tables = {}

    def table(self, ranks=None):
        torch = _lib.torch_mod()
        key = tuple(range(self.fabric.world)) if ranks is None else tuple(int(r) for r in ranks)
        if key not in self._tables:
            self._tables[key] = torch.tensor([self.ptrs[r] for r in key], dtype=torch.int64, device="cuda")
        return self._tables[key]

    def peer_view(self, rank, count=None):
        """Torch view of another rank's buffer (peer loads/stores through torch ops; tests only)."""
        return _as_tensor(self.ptrs[rank], self.count if count is None else count)

    def close(self):
        lib = _lib.load()
        for q in self._opened:
            lib.pt_ipc_close(ctypes.c_void_p(q))
        self._opened = []
        if self._own:
            lib.pt_peer_free(ctypes.c_void_p(self._own))
            self._own = None


class PeerFabric:
    """Handle exchange + device barrier among the ranks of ``group`` (default: the world)."""

    def __init__(self, group=None, timeout_s=20.0):
        import torch.distributed as dist

        torch = _lib.require_cuda()
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.timeout_ns = int(timeout_s * 1e9)
        self._buffers = []
        self._epoch = 0
        self._flags = PeerBuffer(self, self.world, typestr="<i8")
        self._buffers.append(self._flags)
        self._status = torch.zeros(1, dtype=torch.int32, device="cuda")
        # every rank has mapped every flag array before anybody signals
        dist.barrier(group)

    def _all_gather(self, payload):
        import torch.distributed as dist

        out = [None] * self.world
        dist.all_gather_object(out, payload, group=self.group)
        return out

    def alloc(self, count):
        """Symmetric buffer of ``count`` doubles (collective: every rank must call it)."""
        buf = PeerBuffer(self, count)
        self._buffers.append(buf)
        return buf

    def barrier(self):
        """Stream-ordered device barrier: everything enqueued before it on any rank is visible to
        everything enqueued after it on every rank.  No host synchronisation."""
        lib = _lib.load()
        self._epoch += 1
        _lib.check(lib.pt_peer_barrier(_lib.dptr(self._flags.table()), self.rank, self.world, self._epoch,
                                       self.timeout_ns, _lib.dptr(self._status), _lib.stream_ptr()),
                   "pt_peer_barrier")

    def check(self):
        """Host check (synchronises): raises if a barrier timed out waiting for a peer."""
        st = int(self._status.item())
        if st:
            raise RuntimeError(f"pt_peer_barrier timed out waiting for rank {st - 1}")

    def scatter_map(self, buf, ranks, chunk, a_inner, s_outer, s_inner, s_i, offset):
        """``pt_scatter_map`` whose destinations are ``buf`` on ``ranks`` (world ranks, in
        destination order)."""
        m = _lib.ScatterMap()
        m.nranks, m.chunk = len(ranks), int(chunk)
        m.a_inner, m.s_outer, m.s_inner, m.s_i, m.offset = (int(a_inner), int(s_outer), int(s_inner), int(s_i),
                                                            int(offset))
        m._table = buf.table(ranks)
        m.base = m._table.data_ptr()
        return m

    def close(self):
        torch = _lib.torch_mod()
        torch.cuda.synchronize()
        for b in self._buffers:
            b.close()
        self._buffers = []


def scatter_dest(a, i, b, chunk, a_inner, s_outer, s_inner, s_i, offset):
    """(destination index, destination offset) of element (a, i, b) under a pt_scatter_map —
    the host restatement of the device address arithmetic (used by the CPU tests)."""
    a, i, b = np.asarray(a), np.asarray(i), np.asarray(b)
    q = i // chunk
    return q, offset + (a // a_inner) * s_outer + (a % a_inner) * s_inner + (i % chunk) * s_i + b
