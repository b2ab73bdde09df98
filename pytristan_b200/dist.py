"""Multi-GPU partitioning of the hot path over one NVSwitch box (one process per GPU,
``torch.distributed`` for the plumbing).

* ``shard_range``        — fields / non-differentiated axes split contiguously, **no communication**.
* ``SlabStencil``        — a ``FinDiffMat`` along the slab-sharded (slowest) grid axis: halo planes are
                           exchanged with the two neighbours (NCCL send/recv straight out of / into the
                           extended field buffer — planes of the slowest axis are contiguous, nothing is
                           packed), then the rectangular stencil kernel (``pt_stencil_apply_ex``) runs on
                           the local rows; boundary slabs keep the one-sided closures.
* ``PencilTranspose``    — 2-D (py x pz) pencil decomposition of a 3-D field; the x<->y and y<->z
                           redistributions are pack (``pt_permute3d``, fastest axis untouched) ->
                           ``all_to_all_single`` -> unpack.  Because the apply kernels address any axis
                           in place, a "transpose" only redistributes; no local axis permutation that
                           moves the fastest axis is ever needed.
* ``pencil_apply_sum``   — sum of one operator per axis (e.g. a 3-D Chebyshev Laplacian) on pencils.
* ``SlabStencil.apply_peer`` / ``PencilPeer`` — the same two paths over **peer memory**
                           (``pytristan_b200.peer``): the slab stencil loads its halo rows straight
                           from the neighbours' fields (no exchange, no extended buffer), and the
                           pencil redistributions are the store phase of the kernels that produce
                           the data (``pt_dense_apply_scatter`` epilogue + ``pt_peer_scatter`` on a
                           second stream), separated by device-side flag barriers.  No collective
                           call on the data path.

The two module-level functions ``_permute3d`` and ``_stencil_local`` are the only places that touch
the GPU library; the world-size-2 gloo tests replace them with CPU doubles to exercise the
bookkeeping and the collectives without a GPU.
"""
import ctypes

import numpy as np

from . import _lib

__all__ = ["shard_range", "SlabStencil", "PencilTranspose", "pencil_apply_sum", "PencilPeer"]


def shard_range(total, rank, world):
    """Contiguous balanced split of ``total`` items: returns ``(lo, hi)`` of ``rank``."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# ---- the two device entry points (replaced by CPU doubles in the gloo tests) -------------------
def _permute3d(src, dst, d2, d1, d0, perm):
    """dst = src viewed as (d2, d1, d0) with axes reordered to ``perm`` (slowest first)."""
    lib = _lib.load()
    code = perm[0] + 3 * perm[1] + 9 * perm[2]
    _lib.check(lib.pt_permute3d(_lib.dptr(src), _lib.dptr(dst), d2, d1, d0, code, _lib.stream_ptr()),
               "pt_permute3d")


def _stencil_local(band, ext, out, n_in, n_out, shift, after, before, alpha, beta):
    lib = _lib.load()
    if "uni" not in band:
        band["uni"] = _lib.uniform_rows(band, n_out)
    uni = band["uni"]
    _lib.check(
        lib.pt_stencil_apply_ex(_lib.dptr(band["d_w"]), _lib.dptr(band["d_start"]), band["wmax"], n_in,
                                n_out, shift, _lib.dptr(ext), _lib.dptr(out), after, before, 0,
                                ctypes.byref(uni) if uni is not None else None, float(alpha), float(beta),
                                _lib.stream_ptr()),
        "pt_stencil_apply_ex",
    )


# ---- slab-sharded stencil --------------------------------------------------------------------------
class SlabStencil:
    """``FinDiffMat`` along the slowest grid axis when that axis is split into contiguous slabs.

    ``plan = SlabStencil(start, weights, meta, rank, world)`` is built from the *global* band of
    the operator (``FinDiffMat.start`` / ``.weights``).  The rank owns rows ``[z0, z1)``; its
    extended buffer has ``halo_lo + (z1 - z0) + halo_hi`` rows.  ``exchange`` fills the halo rows
    from the neighbours, ``apply`` differentiates the local rows.
    """

    def __init__(self, start, weights, rank, world, periodic=False, wuni=None, w=None, lo=0, hi=0,
                 group=None, device=True):
        start = np.asarray(start, dtype=np.int64)
        n, wmax = weights.shape
        self.n, self.wmax, self.rank, self.world, self.group = n, wmax, rank, world, group
        self.periodic = bool(periodic)
        self.z0, self.z1 = shard_range(n, rank, world)
        self.nloc = self.z1 - self.z0
        if self.nloc <= 0:
            raise ValueError("SlabStencil: more ranks than rows")
        first = start[self.z0:self.z1]
        self.halo_lo = int(max(0, self.z0 - first.min()))
        self.halo_hi = int(max(0, (first.max() + wmax) - self.z1))
        if not self.periodic:
            self.halo_lo = min(self.halo_lo, self.z0)
            self.halo_hi = min(self.halo_hi, n - self.z1)
        # what the neighbours need from this rank (their halos), computed from the same band
        self.send_lo = self._neighbour_need(start, rank - 1, upper=True)    # rows sent to rank-1
        self.send_hi = self._neighbour_need(start, rank + 1, upper=False)   # rows sent to rank+1
        def rows_of(r):
            lo_r, hi_r = shard_range(n, r % world, world)
            return hi_r - lo_r

        if (max(self.send_lo, self.send_hi) > self.nloc or self.halo_lo > rows_of(rank - 1)
                or self.halo_hi > rows_of(rank + 1)):
            raise ValueError("SlabStencil: slabs are thinner than the stencil halo")
        self.n_ext = self.halo_lo + self.nloc + self.halo_hi
        self.shift = self.halo_lo
        origin = self.z0 - self.halo_lo                       # global row of ext row 0
        start_local = (first - origin).astype(np.int32)
        w_local = np.ascontiguousarray(weights[self.z0:self.z1])
        hwu = (w // 2) if w else 0
        if wuni is not None:
            lo_l = max(lo, self.z0) - self.z0
            hi_l = min(hi, self.z1) - self.z0
            if self.periodic:
                lo_l, hi_l = 0, self.nloc
        else:
            lo_l = hi_l = 0
        self.band = dict(wmax=wmax, w=w or 0, wuni=None if wuni is None else np.asarray(wuni, dtype=np.float64),
                         lo=int(lo_l), hi=int(max(hi_l, lo_l)), start=start_local, weights=w_local, hw=hwu)
        if device:
            torch = _lib.require_cuda()
            self.band["d_start"] = torch.from_numpy(start_local).cuda()
            self.band["d_w"] = torch.from_numpy(w_local).cuda()

    @classmethod
    def from_operator(cls, op, rank, world, group=None):
        """Plan for a ``FinDiffMat`` bound to the slowest axis of its grid."""
        band = op._band
        return cls(band["start"], band["weights"], rank, world, periodic=band["periodic"], wuni=band["wuni"],
                   w=band["w"], lo=band["lo"], hi=band["hi"], group=group)

    def _neighbour_need(self, start, r, upper):
        if self.periodic:
            r %= self.world
        elif r < 0 or r >= self.world:
            return 0
        z0, z1 = shard_range(self.n, r, self.world)
        first = start[z0:z1]
        if upper:   # rank r sits below: it needs rows above its z1, i.e. this rank's first rows
            need = int(max(0, (first.max() + self.wmax) - z1))
            return need if self.periodic else min(need, self.n - z1)
        need = int(max(0, z0 - first.min()))
        return need if self.periodic else min(need, z0)

    # -- buffers -----------------------------------------------------------------------------------
    def ext_shape(self, nfields, plane):
        return (nfields, self.n_ext, plane)

    def interior(self, ext):
        """View of the rank's own rows inside an extended buffer ``(nfields, n_ext, plane)``."""
        return ext[:, self.halo_lo:self.halo_lo + self.nloc, :]

    def exchange(self, ext):
        """Fill the halo rows of ``ext`` from the neighbouring ranks (grouped send/recv)."""
        import torch.distributed as dist

        if self.world == 1:
            if self.periodic:
                if self.halo_lo:
                    ext[:, :self.halo_lo] = ext[:, self.halo_lo + self.nloc - self.halo_lo:self.halo_lo + self.nloc]
                if self.halo_hi:
                    ext[:, self.halo_lo + self.nloc:] = ext[:, self.halo_lo:self.halo_lo + self.halo_hi]
            return
        ops, keep = [], []
        lo_peer, hi_peer = self.rank - 1, self.rank + 1
        if self.periodic:
            lo_peer %= self.world
            hi_peer %= self.world
        nf = ext.shape[0]
        for f in range(nf):       # one contiguous block per field
            own = ext[f, self.halo_lo:self.halo_lo + self.nloc]
            if self.send_lo and 0 <= lo_peer < self.world:
                ops.append(dist.P2POp(dist.isend, own[:self.send_lo], self._global(lo_peer), self.group))
            if self.send_hi and 0 <= hi_peer < self.world:
                ops.append(dist.P2POp(dist.isend, own[self.nloc - self.send_hi:], self._global(hi_peer), self.group))
            # receives are posted upper halo first: with two ranks and a periodic axis both
            # neighbours are the same peer, and messages between a pair match in posting order
            # (the peer sends its first rows, which are this rank's upper halo, first)
            if self.halo_hi and 0 <= hi_peer < self.world:
                ops.append(dist.P2POp(dist.irecv, ext[f, self.halo_lo + self.nloc:], self._global(hi_peer), self.group))
            if self.halo_lo and 0 <= lo_peer < self.world:
                ops.append(dist.P2POp(dist.irecv, ext[f, :self.halo_lo], self._global(lo_peer), self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def _global(self, r):
        import torch.distributed as dist

        return r if self.group is None else dist.get_global_rank(self.group, r)

    def apply(self, ext, out=None, alpha=1.0, beta=0.0):
        """Differentiate the local rows of ``ext`` (halos must be current) into ``out``
        ``(nfields, nloc, plane)``."""
        nf, n_ext, plane = ext.shape
        if n_ext != self.n_ext:
            raise ValueError("SlabStencil.apply: buffer does not have n_ext rows")
        if out is None:
            out = ext.new_empty((nf, self.nloc, plane))
        _stencil_local(self.band, ext, out, self.n_ext, self.nloc, self.shift, nf, plane, alpha, beta)
        return out


    # -- peer-memory form: no exchange, the kernel reads the neighbours' rows over NVLink ---------
    def peer_field(self, fabric, nfields, plane):
        """Symmetric field buffer ``(nfields, nloc, plane)`` for ``apply_peer`` (collective)."""
        rows = [shard_range(self.n, r, self.world)[1] - shard_range(self.n, r, self.world)[0]
                for r in range(self.world)]
        buf = fabric.alloc(nfields * max(rows) * plane)
        buf.rows = rows
        return buf

    def apply_peer(self, buf, nfields, plane, out, alpha=1.0, beta=0.0):
        """Differentiate this rank's rows of the field held in ``buf`` (``peer_field``) into
        ``out`` ``(nfields, nloc, plane)``.  The neighbours' fields must be current: separate the
        writes of a field from this call by ``fabric.barrier()``."""
        lib = _lib.load()
        band = self.band
        if "uni" not in band:
            band["uni"] = _lib.uniform_rows(band, self.nloc)
        uni = band["uni"]
        lo_peer, hi_peer = self.rank - 1, self.rank + 1
        if self.periodic:
            lo_peer %= self.world
            hi_peer %= self.world
        u_lo = u_hi = None
        lo_stride = hi_stride = 0
        if self.halo_lo:
            lo_stride = buf.rows[lo_peer] * plane
            u_lo = ctypes.c_void_p(buf.ptrs[lo_peer] + 8 * (buf.rows[lo_peer] - self.halo_lo) * plane)
        if self.halo_hi:
            hi_stride = buf.rows[hi_peer] * plane
            u_hi = ctypes.c_void_p(buf.ptrs[hi_peer])
        _lib.check(
            lib.pt_stencil_apply_halo(_lib.dptr(band["d_w"]), _lib.dptr(band["d_start"]), band["wmax"], self.nloc,
                                      self.halo_lo, self.halo_hi, ctypes.c_void_p(buf.ptrs[self.rank]), u_lo,
                                      lo_stride, u_hi, hi_stride, _lib.dptr(out), nfields, plane,
                                      ctypes.byref(uni) if uni is not None else None, float(alpha), float(beta),
                                      _lib.stream_ptr()),
            "pt_stencil_apply_halo",
        )
        return out


    def gradient_peer(self, buf, fx, fy, nx, ny, outs):
        """Fused ``(d/dx, d/dy, d/dz)`` of this rank's slab of the field in ``buf`` (``peer_field``
        with one field): one read of the slab (``pt_stencil_grad3_halo``), the z halo planes loaded
        from the neighbours.  ``fx``/``fy``: uniform non-periodic ``FinDiffMat``s of the x and y axes
        with the same stencil width as this plan's; ``outs``: three CUDA tensors of the slab size.
        Returns False (nothing done) when the fused kernel does not cover the configuration."""
        lib = _lib.load()
        bz = self.band
        if self.periodic or bz.get("wuni") is None:
            return False
        if "uni" not in bz:
            bz["uni"] = _lib.uniform_rows(bz, self.nloc)
        bands = []
        for op, n in ((fx, nx), (fy, ny)):
            b = op._band
            if b is None or b["wuni"] is None or b["periodic"]:
                return False
            if "uni" not in b:
                b["uni"] = _lib.uniform_rows(b, n)
            bands.append(b)
        bands.append(bz)
        plane = nx * ny
        u_lo = u_hi = None
        if self.halo_lo:
            u_lo = ctypes.c_void_p(buf.ptrs[self.rank - 1] + 8 * (buf.rows[self.rank - 1] - self.halo_lo) * plane)
        if self.halo_hi:
            u_hi = ctypes.c_void_p(buf.ptrs[self.rank + 1])
        w_arr = (ctypes.c_void_p * 3)(*[b["d_w"].data_ptr() for b in bands])
        s_arr = (ctypes.c_void_p * 3)(*[b["d_start"].data_ptr() for b in bands])
        m_arr = (ctypes.c_int * 3)(*[int(b["wmax"]) for b in bands])
        u_arr = (ctypes.POINTER(_lib.UniformRows) * 3)(*[ctypes.pointer(b["uni"]) for b in bands])
        rc = lib.pt_stencil_grad3_halo(ctypes.c_void_p(buf.ptrs[self.rank]), u_lo, u_hi, _lib.dptr(outs[0]),
                                       _lib.dptr(outs[1]), _lib.dptr(outs[2]), nx, ny, self.nloc, w_arr, s_arr,
                                       m_arr, u_arr, _lib.stream_ptr())
        if rc == -3:
            return False
        _lib.check(rc, "pt_stencil_grad3_halo")
        return True


# ---- pencil decomposition -----------------------------------------------------------------------------
class PencilTranspose:
    """Redistributions of a 3-D field ``(nz, ny, nx)`` (x fastest) over a ``py x pz`` process grid.

    Layout A: ``(nz/pz, ny/py, nx)``   — x local            (rank = rz*py + ry)
    Layout B: ``(nz/pz, ny, nx/py)``   — y local
    Layout C: ``(nz, ny/pz, nx/py)``   — z local
    Leading batch dimensions are allowed (they ride along).
    """

    def __init__(self, shape, py, pz, rank=None, world=None, groups=None):
        import torch.distributed as dist

        self.nz, self.ny, self.nx = (int(v) for v in shape)
        self.py, self.pz = int(py), int(pz)
        self.world = dist.get_world_size() if world is None else world
        self.rank = dist.get_rank() if rank is None else rank
        if self.py * self.pz != self.world:
            raise ValueError("PencilTranspose: py*pz must equal the world size")
        for a, b, what in ((self.nx, self.py, "nx % py"), (self.ny, self.py, "ny % py"),
                           (self.ny, self.pz, "ny % pz"), (self.nz, self.pz, "nz % pz")):
            if a % b:
                raise ValueError(f"PencilTranspose: {what} must be 0")
        self.ry, self.rz = self.rank % self.py, self.rank // self.py
        self.nzl, self.nyl, self.nxl, self.nyl2 = self.nz // self.pz, self.ny // self.py, self.nx // self.py, self.ny // self.pz
        if groups is None:
            rows = [dist.new_group([z * self.py + y for y in range(self.py)]) for z in range(self.pz)]
            cols = [dist.new_group([z * self.py + y for z in range(self.pz)]) for y in range(self.py)]
            groups = (rows[self.rz], cols[self.ry])
        self.row_group, self.col_group = groups

    @staticmethod
    def _a2a(recv, send, group):
        import torch.distributed as dist

        dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)

    def _swap01(self, t, d2, d1, d0):
        out = t.new_empty(t.numel())
        _permute3d(t.reshape(-1), out, d2, d1, d0, (1, 0, 2))
        return out

    def x_to_y(self, t):
        """Layout A -> B."""
        lead = t.numel() // (self.nyl * self.nx)          # batch * nzl
        send = self._swap01(t, lead * self.nyl, self.py, self.nxl)       # (py, lead*nyl, nxl)
        recv = send.new_empty(send.numel())
        self._a2a(recv, send, self.row_group)                             # (py[src], lead, nyl*nxl)
        out = self._swap01(recv, self.py, lead, self.nyl * self.nxl)     # (lead, py*nyl, nxl)
        return out.view(*t.shape[:-3], self.nzl, self.ny, self.nxl)

    def y_to_x(self, t):
        """Layout B -> A."""
        lead = t.numel() // (self.ny * self.nxl)
        send = self._swap01(t, lead, self.py, self.nyl * self.nxl)       # (py, lead, nyl*nxl)
        recv = send.new_empty(send.numel())
        self._a2a(recv, send, self.row_group)                             # (py[src x-chunk], lead*nyl, nxl)
        out = self._swap01(recv, self.py, lead * self.nyl, self.nxl)     # (lead*nyl, py, nxl)
        return out.view(*t.shape[:-3], self.nzl, self.nyl, self.nx)

    def y_to_z(self, t):
        """Layout B -> C."""
        batch = t.numel() // (self.nzl * self.ny * self.nxl)
        blk = self.nyl2 * self.nxl
        send = self._swap01(t, batch * self.nzl, self.pz, blk)           # (pz, batch*nzl, blk)
        recv = send.new_empty(send.numel())
        self._a2a(recv, send, self.col_group)                             # (pz[src], batch, nzl*blk)
        out = recv if batch == 1 else self._swap01(recv, self.pz, batch, self.nzl * blk)
        return out.view(*t.shape[:-3], self.nz, self.nyl2, self.nxl)

    def z_to_y(self, t):
        """Layout C -> B."""
        batch = t.numel() // (self.nz * self.nyl2 * self.nxl)
        blk = self.nyl2 * self.nxl
        send = t.reshape(-1) if batch == 1 else self._swap01(t, batch, self.pz, self.nzl * blk)  # (pz, batch, nzl*blk)
        recv = send.new_empty(send.numel())
        self._a2a(recv, send.contiguous(), self.col_group)                # (pz[src y-chunk], batch*nzl, blk)
        out = self._swap01(recv, self.pz, batch * self.nzl, blk)         # (batch*nzl, pz, blk)
        return out.view(*t.shape[:-3], self.nzl, self.ny, self.nxl)


def pencil_apply_sum(ops, u, pencils, back=False):
    """``sum_k ops[k] applied along axis k`` of a 3-D field distributed in layout A.

    ``ops = (Dx, Dy, Dz)`` dense operators (any may be None); ``u`` is the local block in layout A
    ``(nz/pz, ny/py, nx)``.  Returns the result in layout C, or in layout A if ``back``.
    Two all-to-all rounds carry (u, partial sum) together; each partial sum is accumulated in place
    by the kernels' fused ``beta = 1`` epilogue.
    """
    import torch

    pt = pencils
    dx, dy, dz = ops
    acc = torch.zeros_like(u) if dx is None else dx._apply_device(
        u.reshape(-1), pt.nzl * pt.nyl, pt.nx, 1, None, 1.0, 0.0, None, None).view(u.shape)
    pair = torch.stack([u, acc])                                   # (2, nzl, nyl, nx)
    pair = pt.x_to_y(pair)                                         # (2, nzl, ny, nxl)
    if dy is not None:
        dy._apply_device(pair[0].reshape(-1), pt.nzl, pt.ny, pt.nxl, pair[1].reshape(-1), 1.0, 1.0, None, None)
    pair = pt.y_to_z(pair)                                         # (2, nz, nyl2, nxl)
    if dz is not None:
        dz._apply_device(pair[0].reshape(-1), 1, pt.nz, pt.nyl2 * pt.nxl, pair[1].reshape(-1), 1.0, 1.0, None, None)
    res = pair[1]
    if back:
        res = pt.y_to_x(pt.z_to_y(res.contiguous()))
    return res


class PencilPeer(PencilTranspose):
    """Pencil decomposition whose redistributions run over peer memory (``pytristan_b200.peer``).

    Four symmetric buffers of the local block size hold the field and the running sum in layouts
    B and C.  ``apply_sum`` is then, per rank and with no collective call:

        ONE kernel  : Dx u         --scatter epilogue-->         peers' accB
                      u            --stored by the loading warps-> peers' uB    (pt_dense_apply_eo_ws_scatter_copy)
        barrier
        ONE kernel  : Dy uB + accB --scatter epilogue-->         peers' accC
                      uB           --stored by the loading warps-> peers' uC
        barrier
        main stream : accC += Dz uC                                              (local)

    (shapes the warp-specialised kernel declines move the field with ``pt_peer_scatter`` instead)
    """

    def __init__(self, shape, py, pz, fabric, batch=1):
        torch = _lib.require_cuda()
        super().__init__(shape, py, pz, rank=fabric.rank, world=fabric.world, groups=(None, None))
        self.fabric, self.batch = fabric, int(batch)
        self.count = self.batch * self.nzl * self.nyl * self.nx
        self.uB, self.accB, self.uC, self.accC = (fabric.alloc(self.count) for _ in range(4))
        row = [self.rz * self.py + y for y in range(self.py)]
        col = [z * self.py + self.ry for z in range(self.pz)]
        self.row_ranks, self.col_ranks = row, col
        self._side = torch.cuda.Stream()
        self.maps = {}
        for name, bufB, bufC in (("u", self.uB, self.uC), ("acc", self.accB, self.accC)):
            self.maps[name + "_AB"] = fabric.scatter_map(bufB, row, *self.map_ab())
            self.maps[name + "_BC"] = fabric.scatter_map(bufC, col, *self.map_bc())
        self._back = None     # buffers of the way back (layout C -> B -> A), allocated on first use

    # destination maps as (chunk, a_inner, s_outer, s_inner, s_i, offset); see pt_scatter_map
    def map_ab(self):
        """Layout A ``(lead*nyl, nx)`` -> layout B ``(lead, ny, nxl)`` of the rank owning x // nxl."""
        return (self.nxl, self.nyl, self.ny * self.nxl, self.nxl, 1, self.ry * self.nyl * self.nxl)

    def map_bc(self):
        """Layout B ``(batch*nzl, ny, nxl)`` -> layout C ``(batch, nz, nyl2, nxl)`` of the rank owning y // nyl2."""
        blk = self.nyl2 * self.nxl
        return (self.nyl2, self.nzl, self.nz * blk, blk, self.nxl, self.rz * self.nzl * blk)

    def map_cb(self):
        """Layout C ``(batch, nz, nyl2*nxl)`` -> layout B ``(batch*nzl, ny, nxl)`` of the rank owning z // nzl."""
        return (self.nzl, 1, self.nzl * self.ny * self.nxl, 0, self.ny * self.nxl, self.rz * self.nyl2 * self.nxl)

    def map_ba(self):
        """Layout B ``(lead, ny, nxl)`` -> layout A ``(lead, nyl, nx)`` of the rank owning y // nyl."""
        return (self.nyl, 1, self.nyl * self.nx, 0, self.nx, self.ry * self.nxl)

    def to_layout_a(self, c_block):
        """Redistribute a layout-C block (e.g. the result of ``apply_sum``) back to layout A
        ``(batch, nzl, nyl, nx)``: two peer-store scatters and two barriers (collective: every rank
        must call it).  Returns a view of a symmetric buffer, valid until the next call."""
        lib = _lib.load()
        fab = self.fabric
        if self._back is None:                       # collective allocation on first use
            bB, bA = fab.alloc(self.count), fab.alloc(self.count)
            self._back = (bB, bA, fab.scatter_map(bB, self.col_ranks, *self.map_cb()),
                          fab.scatter_map(bA, self.row_ranks, *self.map_ba()))
        bB, bA, m_cb, m_ba = self._back
        lead = self.batch * self.nzl
        _lib.check(lib.pt_peer_scatter(_lib.dptr(c_block.reshape(-1)), self.nz, self.batch, self.nyl2 * self.nxl,
                                       ctypes.byref(m_cb), _lib.stream_ptr()), "pt_peer_scatter")
        fab.barrier()
        _lib.check(lib.pt_peer_scatter(_lib.dptr(bB.local), self.ny, lead, self.nxl, ctypes.byref(m_ba),
                                       _lib.stream_ptr()), "pt_peer_scatter")
        fab.barrier()
        return bA.local.view(self.batch, self.nzl, self.nyl, self.nx)

    def apply_sum(self, ops, u):
        """``sum_k ops[k] along axis k`` of the field whose layout-A block is ``u``
        ``(batch, nzl, nyl, nx)``; returns the layout-C block ``(batch, nz, nyl2, nxl)`` (a view
        of the symmetric buffer ``accC``, valid until the next call)."""
        torch = _lib.torch_mod()
        lib = _lib.load()
        dx, dy, dz = ops
        fab = self.fabric
        main = torch.cuda.current_stream()
        side = self._side
        lead = self.batch * self.nzl
        u = u.reshape(-1)
        # ---- A -> B: Dx u leaves through the scatter epilogue; u itself rides in the same kernel (stored by
        # the warps that load it) when the warp-specialised kernel covers the shape, else it is moved by
        # pt_peer_scatter on a side stream that only waits for what preceded the DMMA kernel
        ready = torch.cuda.Event()
        ready.record(main)
        if not dx._apply_scatter(u, lead * self.nyl, self.nx, 1, None, 1.0, 0.0, self.maps["acc_AB"],
                                 field_map=self.maps["u_AB"]):
            side.wait_event(ready)
            with torch.cuda.stream(side):
                _lib.check(lib.pt_peer_scatter(_lib.dptr(u), self.nx, lead * self.nyl, 1,
                                               ctypes.byref(self.maps["u_AB"]), _lib.stream_ptr()), "pt_peer_scatter")
            main.wait_stream(side)
        fab.barrier()
        # ---- B -> C
        uB, accB = self.uB.local, self.accB.local
        ready = torch.cuda.Event()
        ready.record(main)
        if not dy._apply_scatter(uB, lead, self.ny, self.nxl, accB, 1.0, 1.0, self.maps["acc_BC"],
                                 field_map=self.maps["u_BC"]):
            side.wait_event(ready)
            with torch.cuda.stream(side):
                _lib.check(lib.pt_peer_scatter(_lib.dptr(uB), self.ny, lead, self.nxl,
                                               ctypes.byref(self.maps["u_BC"]), _lib.stream_ptr()), "pt_peer_scatter")
            main.wait_stream(side)
        fab.barrier()
        # ---- C: local
        dz._apply_device(self.uC.local, self.batch, self.nz, self.nyl2 * self.nxl, self.accC.local, 1.0, 1.0,
                         None, None)
        return self.accC.local.view(self.batch, self.nz, self.nyl2, self.nxl)
