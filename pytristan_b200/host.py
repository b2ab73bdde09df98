"""Host-buffer entry point: differentiate fields that live in host memory.

``apply_host(ops, fields, outs)`` is the call a host-side user of the reference API makes when
the data is in NumPy/pinned memory: it streams the stack of fields through the GPU in chunks of
whole fields — H2D on one copy stream, the operator kernels on the caller's stream, D2H on a
second copy stream, double buffered — so that the PCIe transfers of neighbouring chunks overlap
each other and the kernels.  This is what bench.py's ``e2e`` figure measures.
"""
import os

import numpy as np

from . import _lib

__all__ = ["apply_host", "launches_per_apply", "bind_rank_to_cores"]


def bind_rank_to_cores(local_rank, local_world):
    """One process per GPU: give this rank its own contiguous share of the host cores this job may
    run on (``sched_setaffinity``), so that the copy-issuing threads and the pinned buffers they
    first touch of different ranks do not migrate over each other.  Call it before allocating
    pinned memory.  Returns the list of cores now owned (all of them if there are fewer cores than
    ranks or the platform has no affinity call)."""
    if not hasattr(os, "sched_getaffinity"):
        return []
    cores = sorted(os.sched_getaffinity(0))
    per = len(cores) // max(1, int(local_world))
    if per < 1:
        return cores
    mine = cores[int(local_rank) * per:(int(local_rank) + 1) * per]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return cores
    return mine

_streams = {}


def _copy_streams(torch):
    dev = torch.cuda.current_device()
    if dev not in _streams:
        _streams[dev] = (torch.cuda.Stream(), torch.cuda.Stream())
    return _streams[dev]


def _as_host_tensor(torch, a):
    if isinstance(a, torch.Tensor):
        if a.is_cuda or a.dtype != torch.float64 or not a.is_contiguous():
            raise TypeError("host fields must be contiguous float64 CPU tensors or ndarrays")
        return a
    arr = np.ascontiguousarray(a, dtype=np.float64)
    return torch.from_numpy(arr)


def apply_host(ops, fields, outs=None, chunk_fields=None):
    """Apply every operator of ``ops`` to the host-resident stack ``fields``.

    ``fields``: ndarray or CPU torch tensor holding B flat fields (pinned memory makes the
    copies asynchronous).  ``outs``: optional list of host buffers of the same size, one per
    operator; allocated (pinned) when omitted.  Returns the list of outputs, same kind as
    ``fields``, after the last device-to-host copy has completed (the host buffers are safe to
    read on return).  All operators must be bound to the same grid.
    """
    torch = _lib.require_cuda()
    ops = list(ops)
    if not ops:
        return []
    u = _as_host_tensor(torch, fields)
    total = 1
    for n in ops[0].npts:
        total *= int(n)
    for op in ops:
        if tuple(op.npts) != tuple(ops[0].npts):
            raise ValueError("apply_host: all operators must be bound to the same grid")
    if total == 0 or u.numel() % total:
        raise ValueError("apply_host: fields are not a whole number of fields on the operators' grid")
    nfields = u.numel() // total
    as_numpy = not isinstance(fields, torch.Tensor)
    if outs is None:
        outs_t = [torch.empty(u.numel(), dtype=torch.float64, pin_memory=True) for _ in ops]
    else:
        outs_t = [_as_host_tensor(torch, o) for o in outs]
        if len(outs_t) != len(ops) or any(o.numel() != u.numel() for o in outs_t):
            raise ValueError("apply_host: need one output buffer of the fields' size per operator")
    u_flat = u.view(-1)
    outs_flat = [o.view(-1) for o in outs_t]

    # chunk schedule: a list of field counts, or one count for uniform chunks.  Default: a short ramp
    # (1, 1, 2, 4, ... fields) so that the device-to-host stream — the long pole: it carries one output per
    # operator — starts after ~0.1 ms instead of after a full-size chunk, then chunks of nfields/8
    # (few, large copies: every chunk boundary costs the copy engines an event round trip);
    # tools/e2e_sweep.py measures the alternatives
    if chunk_fields is None:
        big = max(1, -(-nfields // 8))
        sched, step, left = [], 1, nfields
        while left > 0:
            take = min(step, big, left)
            sched.append(take)
            left -= take
            if len(sched) >= 2:
                step *= 2
    elif isinstance(chunk_fields, (list, tuple)):
        sched = [int(c) for c in chunk_fields]
        if any(c < 1 for c in sched) or sum(sched) != nfields:
            raise ValueError("apply_host: the chunk schedule must be positive and sum to the number of fields")
    else:
        cf1 = max(1, int(chunk_fields))
        sched = [cf1] * (nfields // cf1) + ([nfields % cf1] if nfields % cf1 else [])
    nchunks = len(sched)
    cf = max(sched)
    starts = [0]
    for c in sched:
        starts.append(starts[-1] + c)
    cur = torch.cuda.current_stream()
    s_in, s_out = _copy_streams(torch)
    slots = min(2, nchunks)
    d_u = [torch.empty(cf * total, dtype=torch.float64, device="cuda") for _ in range(slots)]
    d_y = [[torch.empty(cf * total, dtype=torch.float64, device="cuda") for _ in range(slots)] for _ in ops]
    ev_in = [torch.cuda.Event() for _ in range(slots)]
    ev_comp = [torch.cuda.Event() for _ in range(slots)]
    ev_out = [torch.cuda.Event() for _ in range(slots)]
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    for c in range(nchunks):
        slot = c % slots
        lo, hi = starts[c], starts[c + 1]
        cnt = (hi - lo) * total
        with torch.cuda.stream(s_in):
            if c >= slots:
                s_in.wait_event(ev_comp[slot])       # kernels of chunk c-2 have consumed this slot
            d_u[slot][:cnt].copy_(u_flat[lo * total:hi * total], non_blocking=True)
            ev_in[slot].record(s_in)
        cur.wait_event(ev_in[slot])
        if c >= slots:
            cur.wait_event(ev_out[slot])             # D2H of chunk c-2 has drained this slot
        for k, op in enumerate(ops):
            after, n, before = op._split(cnt)
            op._apply_device(d_u[slot][:cnt], after, n, before, d_y[k][slot][:cnt], 1.0, 0.0, None, None)
        ev_comp[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_comp[slot])
            for k in range(len(ops)):
                outs_flat[k][lo * total:hi * total].copy_(d_y[k][slot][:cnt], non_blocking=True)
            ev_out[slot].record(s_out)
    cur.wait_stream(s_out)
    cur.wait_stream(s_in)
    # the outputs are HOST memory: they are valid only once the last device-to-host copy has
    # landed, so the call returns after waiting for it on the host (ndarray or tensor alike); the
    # asynchronous reads of `fields` have completed by then too (they precede the kernels)
    ev_out[(nchunks - 1) % slots].synchronize()
    if as_numpy:
        return [o.numpy().reshape(np.shape(fields)) for o in outs_t]
    return [o.view(u.shape) for o in outs_t]


def launches_per_apply(ops, after=1):
    """Number of kernel launches one pass of ``ops`` issues (for bench.py's gpu_launches)."""
    count = 0
    for op in ops:
        if hasattr(op, "launches"):
            count += op.launches
            continue
        band = getattr(op, "_band", None)
        if band is None:
            count += 1
        elif band["wuni"] is not None and not band["periodic"] and op.axis != 0:
            count += 2          # strided interior kernel + boundary-row table kernel
        else:
            count += 1
    return count
