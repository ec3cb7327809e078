"""Row-slab domain decomposition of the routing grid across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the
host-logic tests).  Rank k owns the contiguous rows [row0, row0 + ny) of the
global (NY, NX) grid; every per-cell tensor carries `halo` extra rows above and
below (1 for the kinematic wave, 2 for the diffusive wave whose parent depth is
re-derived from the parent's own neighbourhood).  Static halos (codes, channel
parameters) are filled once; per step only the discharge rows (and, for the
diffusive wave, the pre-step state rows) cross the link:

    kinematic : Q      , 1 row  per direction  (nx * 8 B)
    diffusive : Q      , 2 rows per direction  + vol (+ I, h_swe if fused), 1 row

plus ONE all-reduce(SUM) of TFB_NSUM doubles (volume integrals, bad-value counts,
outlet sample from its owner).  The reference has no counterpart (single
process); per-cell results are independent of the slab layout, so an N-GPU run
is bit-identical to a 1-GPU run in every grid (tests/test_slab_gpu.py), and the
global sums differ only by summation order.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import lib as L


def partition_rows(ny_global, world_size):
    """[(row0, ny)] per rank: contiguous, as even as possible, earlier ranks get
    the remainder."""
    base, rem = divmod(int(ny_global), int(world_size))
    out, r0 = [], 0
    for k in range(world_size):
        n = base + (1 if k < rem else 0)
        out.append((r0, n))
        r0 += n
    return out


class SlabLayout(object):
    def __init__(self, ny_global, nx, rank, world_size, halo):
        self.ny_global, self.nx = int(ny_global), int(nx)
        self.rank, self.world_size, self.halo = int(rank), int(world_size), int(halo)
        self.parts = partition_rows(ny_global, world_size)
        self.row0, self.ny = self.parts[rank]
        if self.world_size > 1 and min(n for _, n in self.parts) < max(1, halo):
            raise ValueError('slabs thinner than the halo: %r' % (self.parts,))
        self.up = rank - 1 if rank > 0 else None
        self.down = rank + 1 if rank < world_size - 1 else None

    def take(self, global_array):
        """Rows of a global (NY, NX) array this rank needs, halo included; rows
        outside the global grid are zero (code 0 / Q 0: nothing flows in)."""
        h = self.halo
        lo, hi = self.row0 - h, self.row0 + self.ny + h
        src = global_array[max(lo, 0):min(hi, self.ny_global)]
        if lo >= 0 and hi <= self.ny_global:
            return src
        if isinstance(global_array, torch.Tensor):
            out = torch.zeros((self.ny + 2 * h,) + tuple(global_array.shape[1:]),
                              dtype=global_array.dtype, device=global_array.device)
        else:
            import numpy as np
            out = np.zeros((self.ny + 2 * h,) + tuple(global_array.shape[1:]),
                           dtype=global_array.dtype)
        out[max(0, -lo):max(0, -lo) + src.shape[0]] = src
        return out


def exchange_halos(tensors, layout, rows=None, group=None):
    """Fill the halo rows of every (ny + 2*halo, nx) tensor in `tensors` from the
    neighbouring ranks' boundary rows.  `rows` (<= halo) limits how many rows
    are exchanged per tensor (default: all `halo` rows).  Works on CUDA tensors
    over NCCL and on CPU tensors over gloo."""
    if layout.world_size == 1 or not tensors:
        return
    h, ny = layout.halo, layout.ny
    if rows is None:
        rows = [h] * len(tensors)
    ops, keep = [], []
    for t, k in zip(tensors, rows):
        if k <= 0:
            continue
        if layout.up is not None:
            send = t[h:h + k]                       # my first k owned rows
            recv = t[h - k:h]                       # my upper halo (nearest k)
            ops.append(dist.P2POp(dist.isend, send, layout.up, group))
            ops.append(dist.P2POp(dist.irecv, recv, layout.up, group))
        if layout.down is not None:
            send = t[h + ny - k:h + ny]             # my last k owned rows
            recv = t[h + ny:h + ny + k]             # my lower halo
            ops.append(dist.P2POp(dist.isend, send, layout.down, group))
            ops.append(dist.P2POp(dist.irecv, recv, layout.down, group))
        keep.append(t)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class SlabChannels(object):
    """A ChannelsEngine per rank + the per-step exchange."""

    def __init__(self, layout, engine, group=None):
        self.layout, self.engine, self.group = layout, engine, group
        self.sums = torch.zeros(L.NSUM, dtype=torch.float64, device=engine.device)

    def exchange_state(self):
        e, lay = self.engine, self.layout
        ts, rows = [e.Q_buf[e.qcur]], [lay.halo]
        if not e.kinematic:
            ts.append(e._sbuf(e.vol_buf)); rows.append(1)
            for buf in (e.I_buf, e.h_swe_buf):
                if buf is not None:
                    ts.append(e._sbuf(buf)); rows.append(1)
        exchange_halos(ts, lay, rows, self.group)

    def step(self, time_min=0.0):
        e = self.engine
        e.step(time_min)                            # finalize = 0: sums stay per rank
        if self.layout.world_size > 1:
            self.sums.copy_(e.partials[:L.NSUM])
            dist.all_reduce(self.sums, op=dist.ReduceOp.SUM, group=self.group)
            # slot 16 (dt_cfl_min, a diagnostic) is a MIN, not a SUM: keep this
            # rank's own value; global_dt_min() reduces it on demand
            self.sums[16] = e.partials[16]
            src = self.sums
        else:
            src = e.partials[:L.NSUM]
        L.check(e.lib.tfb_channels_fold(C.byref(e.args), src.data_ptr(), L.stream_ptr()),
                'tfb_channels_fold')
        self.exchange_state()

    def global_dt_min(self):
        t = self.engine.partials[16:17].clone()
        if self.layout.world_size > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return float(t.item())
