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
import mmap
import os
import platform
import socket
import tempfile
import time

import numpy as np
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
        self.pitch = (self.nx + 3) // 4 * 4          # row pitch of the device arrays (routing.row_pitch)
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
        if t.dtype in (torch.int16, torch.bool):      # NCCL has no 16-bit integer type
            t = t.view(torch.uint8)
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


class _DevArray(object):
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw allocation."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr,
                                         'data': (int(ptr), False), 'version': 2, 'strides': None}


class PeerHalos(object):
    """Peer-memory halo exchange for the kinematic packed step (include/tfb.h,
    tfb_channels_step_packed_p2p): this rank's two discharge buffers and a 2-word flag
    block are allocated with tfb_peer_alloc, the CUDA-IPC handles go round with one
    all_gather_object, and the neighbours' buffers are mapped into this process.  After
    that a step needs no NCCL call: the kernel stores its boundary discharge rows into
    the neighbours' halo rows over NVLink and the streams order themselves with
    cuStreamWaitValue32 on the flag words."""

    def __init__(self, layout, device, group=None, _collective_guard=False, with_depth=False):
        self.layout, self.group = layout, group
        self.with_depth = bool(with_depth)       # diffusive wave: the depth grid crosses too
        self.lib = L.load()
        self.ok = True
        lay = layout
        rows = lay.ny + 2 * lay.halo
        nbytes = rows * lay.pitch * 8
        self._mine, handles = [], []
        try:
            sizes = [nbytes, nbytes, 256] + ([nbytes] if self.with_depth else [])
            for nb in sizes:                     # Q buffer 0, Q buffer 1, flags, (depth)
                p = C.c_void_p()
                h = C.create_string_buffer(64)
                L.check(self.lib.tfb_peer_alloc(nb, C.byref(p), h), 'tfb_peer_alloc')
                self._mine.append(p.value)
                handles.append(h.raw)
            self.q_buffers = [torch.as_tensor(_DevArray(self._mine[k], (rows, lay.pitch), '<f8'),
                                              device=device) for k in range(2)]
            self.d_buffer = (torch.as_tensor(_DevArray(self._mine[3], (rows, lay.pitch), '<f8'),
                                             device=device) if self.with_depth else None)
        except Exception:
            if not _collective_guard:
                raise
            self.ok, handles = False, None       # still take part in the collectives below
        info = [None] * lay.world_size
        dist.all_gather_object(info, dict(handles=handles, ny=lay.ny), group=group)
        self._open = {}
        self.nbr = {}
        if any(i['handles'] is None for i in info):
            self.ok = False
        try:
            for name, r in (('up', lay.up), ('down', lay.down)):
                if r is None or not self.ok:
                    continue
                ptrs = []
                for h in info[r]['handles']:
                    p = C.c_void_p()
                    L.check(self.lib.tfb_peer_open(h, C.byref(p)), 'tfb_peer_open')
                    ptrs.append(p.value)
                self._open[name] = ptrs
                self.nbr[name] = dict(q=ptrs[:2], flags=ptrs[2], ny=info[r]['ny'],
                                      d=ptrs[3] if len(ptrs) > 3 else None)
        except L.TfbError:
            if not _collective_guard:
                raise
            self.ok = False                       # make_peer_halos lets every rank fall back
        dist.barrier(group=group)

    def descriptor(self, q_out, step, depth=False):
        """tfb_peer_halo for one launch: the neighbours' halo rows of the discharge buffer
        `q_out` (or of the depth grid) and the flag words; `step` = step index for the
        kinematic kernel, phase counter (2n, 2n+1) for the two diffusive passes."""
        lay = self.layout
        row = lay.pitch * 8
        ph = L.PeerHalo()
        if 'up' in self.nbr:
            n = self.nbr['up']
            base = n['d'] if depth else n['q'][q_out]
            ph.up_Q = base + (lay.halo + n['ny']) * row              # its first LOWER halo row
            ph.up_flag = n['flags'] + 4                              # its "from_down" word
            ph.from_up = self._mine[2]
        if 'down' in self.nbr:
            n = self.nbr['down']
            base = n['d'] if depth else n['q'][q_out]
            ph.down_Q = base + (lay.halo - 1) * row                  # its last UPPER halo row
            ph.down_flag = n['flags']                                # its "from_up" word
            ph.from_down = self._mine[2] + 4
        ph.step = int(step)
        return ph

    def close(self):
        """Unmap the neighbours' buffers and free this rank's own peer allocations (the
        tensors wrapped around them must not be used afterwards)."""
        for ptrs in self._open.values():
            for p in ptrs:
                self.lib.tfb_peer_close(p)
        self._open = {}
        self.q_buffers, self.d_buffer = None, None
        for p in self._mine:
            self.lib.tfb_peer_free(p)
        self._mine = []


class ScalarBoard(object):
    """The ranks' scalar blocks side by side in ONE POSIX shared-memory segment (all ranks of
    a slab run sit on one box): how `SlabChannels.read_scalars` forms the global values
    without a collective.  A synchronous `update()` reads the scalars every step
    (CHECK_STABILITY, Q_outlet for the driver); an NCCL all-gather + device-to-host copy +
    stream synchronise there costs ~50 us per step on a 0.26 ms step, a 232-byte host copy
    and a poll of the neighbours' sequence words ~5 us.

    Layout: 2 banks x world slots of `SLOT` bytes; a slot = the block + a sequence word.
    The k-th collective read uses bank k & 1: a rank copies its block in, THEN stores k
    (x86-TSO: stores are seen in program order; other machines take the NCCL path), and
    polls until every slot of the bank holds k.  Bank k & 1 is written again at read
    k + 2, which a rank reaches only after EVERY rank has published k + 1, i.e. has finished
    reading bank k & 1 -- no rank can overwrite a block another one is still reading.
    No kernel writes here: each rank's kernel keeps writing its own pinned mirror."""
    SLOT = 256

    def __init__(self, layout, nbytes, group=None):
        self.world, self.rank, self.nbytes = layout.world_size, layout.rank, int(nbytes)
        assert self.nbytes + 8 <= self.SLOT and self.nbytes % 8 == 0
        self.nw = self.nbytes // 8
        size = 2 * self.world * self.SLOT
        box = [None]
        if self.rank == 0:
            try:
                fd, path = tempfile.mkstemp(prefix='tfb_scalars_', dir='/dev/shm')
                os.ftruncate(fd, size)
                os.close(fd)
                box[0] = path
            except OSError:
                box[0] = None             # the other ranks learn it through the broadcast
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast_object_list(box, src=src, group=group)
        self._path = box[0]
        # same host name and boot id do not prove a shared /dev/shm (separate mount namespaces):
        # every rank must SEE the segment, or all of them take the NCCL path
        seen = [None] * self.world
        dist.all_gather_object(seen, bool(self._path and os.path.exists(self._path)), group=group)
        if not all(seen):
            if self.rank == 0 and self._path and os.path.exists(self._path):
                os.unlink(self._path)
            raise OSError('ScalarBoard: /dev/shm is not shared by all ranks')
        fd = os.open(self._path, os.O_RDWR)
        try:
            self._mm = mmap.mmap(fd, size)
        finally:
            os.close(fd)
        dist.barrier(group=group)                 # everyone has it mapped: the name can go
        if self.rank == 0:
            os.unlink(self._path)
        words = np.frombuffer(self._mm, dtype=np.float64).reshape(2, self.world, self.SLOT // 8)
        self._blocks = words[:, :, :self.nw]                         # (2, world, nw) float64
        self._seq = words.view(np.int64)[:, :, self.SLOT // 8 - 1]   # (2, world) sequence words
        self.k = 0

    @staticmethod
    def usable(layout, group=None):
        """All ranks on one x86 host with /dev/shm (decided collectively)."""
        ok = (platform.machine() in ('x86_64', 'AMD64') and os.path.isdir('/dev/shm')
              and os.access('/dev/shm', os.W_OK) and os.environ.get('TFB_SCALAR_BOARD', '1') != '0')
        try:
            boot = open('/proc/sys/kernel/random/boot_id').read().strip()
        except OSError:
            boot = ''
        mine = (socket.gethostname(), boot, bool(ok))
        every = [None] * layout.world_size
        dist.all_gather_object(every, mine, group=group)
        return all(e[2] for e in every) and len(set(e[:2] for e in every)) == 1

    def exchange(self, block_words, timeout_s=120.0):
        """Collective: publish my block (float64 words), return all blocks (world, nw)."""
        self.k += 1
        k, b = self.k, self.k & 1
        self._blocks[b, self.rank, :] = block_words[:self.nw]
        self._seq[b, self.rank] = k
        seq = self._seq[b]
        t0 = None
        while not (seq == k).all():
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > timeout_s:
                raise L.TfbError('scalar board: rank(s) %s did not publish read %d within %g s'
                                 % (np.nonzero(seq != k)[0].tolist(), k, timeout_s))
        return self._blocks[b]

    def close(self):
        self._blocks = self._seq = None
        try:
            self._mm.close()
        except (BufferError, ValueError):
            pass


def make_peer_halos(layout, device, group=None, with_depth=False):
    """PeerHalos on every rank, or None on every rank: if CUDA IPC is not usable on any
    rank (containers without shared /dev/shm or IPC permission, no peer access) all
    ranks agree to fall back to the NCCL halo exchange."""
    ph, ok = None, 1.0
    try:
        ph = PeerHalos(layout, device, group, _collective_guard=True, with_depth=with_depth)
        ok = 1.0 if ph.ok else 0.0
    except Exception as exc:                      # pragma: no cover (depends on the host)
        import warnings
        warnings.warn('peer-memory halos unavailable, using NCCL send/recv: %s' % (exc,))
        ok = 0.0
    t = torch.tensor([ok], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if float(t.item()) < 1.0:
        if ph is not None:
            ph.close()
        return None
    return ph


class SlabChannels(object):
    """A ChannelsEngine per rank + the per-step halo exchange.

    Per step only the boundary discharge rows cross NVLink.  The volume
    integrals and bad-value counters are sums over cells and steps, so every
    rank folds ITS OWN share into its own scalar block each step (engine built
    with finalize=True and n_pixels_global = its own cell count) and the global
    values are formed on demand by ONE all-reduce in read_scalars(); the outlet
    sample, peaks and vol_Q live on the rank that owns the outlet cell and are
    zero elsewhere, so the same SUM delivers them.  No per-step collective."""

    _SUM_FIELDS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet',
                   'f_outlet', 'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak',
                   'vol_P', 'vol_PR', 'vol_PS', 'vol_SM', 'vol_ET', 'vol_IN')
    _CNT_FIELDS = ('n_neg_d', 'n_nan_d', 'n_inf_d', 'n_neg_u', 'n_nan_u', 'n_inf_u', 'n_nan_IN')

    def __init__(self, layout, engine, group=None, peer=None):
        self.layout, self.engine, self.group = layout, engine, group
        self.peer = peer if (peer is not None and layout.world_size > 1) else None
        if layout.world_size > 1 and not engine.args.finalize:
            raise L.TfbError('SlabChannels needs an engine built with finalize=True')
        self.device = engine.device
        self.n_launches_per_step = 1
        # the generic diffusive kernel re-derives the parent's depth from the HALO rows of
        # vol / I / h_swe: an initial depth d0 > 0 (or h0_swe, I0) is only written into the
        # owned rows, so the halo rows are filled from the neighbours before the first
        # generic step
        self._state_halos_stale = layout.world_size > 1 and not engine.kinematic
        self.n_peer_steps = 0         # steps (diffusive: phases) taken with the peer-memory halo
        if layout.world_size > 1 and not engine.kinematic and self.peer is None:
            # packed diffusive step: the new depths of the boundary rows cross the link
            # between pass A and pass B
            # (with FLOOD_OPTION the free surface is channel + flood depth: both rows)
            engine.mid_step_hook = lambda: exchange_halos(
                [engine.d] + ([engine.d_flood] if engine.flood else []), layout,
                [1] * (2 if engine.flood else 1), group)
        if self.peer is not None and not engine.kinematic and not self.peer.with_depth:
            raise L.TfbError('diffusive peer halos need PeerHalos(with_depth=True)')
        # global scalars without a collective where the ranks share a host (ScalarBoard)
        self.board = None
        if layout.world_size > 1 and dist.get_backend(group) == 'nccl' and ScalarBoard.usable(layout, group):
            try:
                self.board = ScalarBoard(layout, C.sizeof(L.ChannelsScalars), group)
            except OSError:               # raised on EVERY rank (collective decision): NCCL all-gather
                self.board = None

    def __getattr__(self, name):           # state access falls through to the engine
        return getattr(self.__dict__['engine'], name)

    def exchange_state(self, discharge_only=False):
        """Halo rows after a step.  The packed kernels read neighbour rows of the discharge
        only (and, between the diffusive passes, of the depth); the generic diffusive
        kernel re-derives the parent's new depth and needs the state rows as well."""
        e, lay = self.engine, self.layout
        ts, rows = [e.Q_buf[e.qcur]], [lay.halo]
        if not e.kinematic and not discharge_only:
            ts.append(e._sbuf(e.vol_buf)); rows.append(1)
            for buf in (e.I_buf, e.h_swe_buf, e.vol_stored_buf):
                if buf is not None:
                    ts.append(e._sbuf(buf)); rows.append(1)
        exchange_halos(ts, lay, rows, self.group)

    def step(self, time_min=0.0):
        e = self.engine
        packed = e._config()[0]
        if self.peer is not None and packed == 'lean':
            # halo rows travel inside the kernel(s); the flag words count PEER steps only,
            # so NCCL steps in between (below) leave the protocol in step
            e.step(time_min, peer=self.peer, peer_step=self.n_peer_steps)
            self.n_peer_steps += 1 if e.kinematic else 2
            self._state_halos_stale = not e.kinematic
            return
        # (with peer halos set up, a step the packed path cannot take -- grid forcing from a
        # provider, say -- uses the NCCL exchange: the peer buffers are ordinary device tensors)
        if not packed and self._state_halos_stale:
            # first generic step after packed ones: its state halo rows were not kept current
            self.exchange_state()
            self._state_halos_stale = False
        e.step(time_min)
        self.exchange_state(discharge_only=bool(packed))
        if packed and not e.kinematic:
            self._state_halos_stale = True

    def read_scalar_words(self):
        """Global scalar block (collective: every rank must call it) as the 8-byte words of
        lib.ChannelsScalars: (float64 view, int64 view) of one array.  The ranks' blocks are
        exchanged through the shared-memory ScalarBoard (one box) or ONE NCCL all-gather."""
        e = self.engine
        if self.layout.world_size == 1:
            return e.sync_scalar_mirror(), e._scalars_i64
        nf, nc = len(self._SUM_FIELDS), len(self._CNT_FIELDS)
        if self.board is not None:
            hf = self.board.exchange(e.sync_scalar_mirror())
        else:
            raw = e.scalars_dev.view(torch.float64)          # 20 doubles + 9 int64 bit patterns
            if getattr(self, '_allv', None) is None:
                self._allv = torch.empty((self.layout.world_size, raw.numel()), dtype=torch.float64,
                                         device=self.device)
                self._allh = torch.empty_like(self._allv, device='cpu').pin_memory()
            dist.all_gather_into_tensor(self._allv, raw, group=self.group)
            self._allh.copy_(self._allv, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            hf = self._allh.numpy()                          # (world, 29): reduced on the host
        hi = hf.view(np.int64)
        if getattr(self, '_red', None) is None:
            self._red = np.zeros(hf.shape[1], dtype=np.float64)
            self._red_i = self._red.view(np.int64)
        out, out_i = self._red, self._red_i
        out[:nf] = hf[:, :nf].sum(0)
        out[nf] = hf[:, nf].min()                            # dt_cfl_min: a MIN, not a SUM
        out_i[nf + 1:nf + 1 + nc] = hi[:, nf + 1:nf + 1 + nc].sum(0)
        out_i[nf + 1 + nc] = e.n_steps                       # step_count
        out_i[nf + 2 + nc] = hi[:, nf + 2 + nc].sum()        # n_step_failed
        return out, out_i

    def read_scalars(self):
        """The global scalar block as a lib.ChannelsScalars (collective)."""
        if self.layout.world_size == 1:
            return self.engine.read_scalars()
        return L.ChannelsScalars.from_buffer_copy(self.read_scalar_words()[0])

    def global_dt_min(self):
        return self.read_scalars().dt_cfl_min


def slab_d8(DEM, layout, dem_halo, link_flats=True, xres=30.0, yres=30.0, nodata=-9999.0,
            kinematic=True, device=None, group=None):
    """D8 codes and bed slope of one slab, computed on the device.

    `DEM`: float32 rows [row0 - dem_halo, row0 + ny + dem_halo) of the global grid
    (dem_halo >= layout.halo + 1 when world_size > 1).  Flat linking is the
    GLOBAL Jacobi iteration of d8_global.link_flats (:374-576): every rank sweeps
    its owned rows, the boundary code rows are exchanged and the flat / fixed
    counters all-reduced after each sweep, so the codes are bit-identical to a
    single-GPU run on the whole grid.  Returns code (uint8, halo rows included),
    S_bed (float64, owned rows; zero/negative slopes replaced by the GLOBAL
    minimum positive slope for the kinematic wave, channels_base.py:4164-4260)
    and the float32 pixel sizes."""
    import numpy as np
    from .d8 import D8Toolkit
    lib = L.load()
    dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    ny, nx, H, NY = layout.ny, layout.nx, layout.halo, layout.ny_global
    dx64, dy64 = np.float64(xres), np.float64(yres)
    dd64 = np.sqrt(dx64 ** 2 + dy64 ** 2)
    dx, dy, dd = np.float32(dx64), np.float32(dy64), np.float32(dd64)
    multi = layout.world_size > 1
    st = L.stream_ptr
    if not multi:
        tk = D8Toolkit(DEM, xres=xres, yres=yres, link_flats=link_flats, nodata=nodata, device=dev)
        tk.update_flow_grid()
        S = tk.slope().double()
        code = tk.code
        n_reps = tk.n_reps
    else:
        if dem_halo < H + 1:
            raise ValueError('dem_halo must be >= halo + 1')
        dem = (DEM.to(device=dev, dtype=torch.float32).contiguous() if isinstance(DEM, torch.Tensor)
               else torch.as_tensor(np.ascontiguousarray(DEM, dtype=np.float32)).to(dev))
        top = layout.row0 - dem_halo                      # global row of dem[0]
        lo, hi = max(0, layout.row0 - H), min(NY, layout.row0 + ny + H)
        code16 = torch.zeros((ny + 2 * H, nx), dtype=torch.int16, device=dev)
        reg = code16[lo - (layout.row0 - H):hi - (layout.row0 - H)]
        L.check(lib.tfb_d8_codes(dem[lo - top:].data_ptr(), hi - lo, nx, 1, lo, NY,
                                 float(dx), float(dy), float(dd), float(nodata),
                                 1 if link_flats else 0, reg.data_ptr(), st()), 'tfb_d8_codes')
        L.check(lib.tfb_d8_break_ties(reg.data_ptr(), reg.numel(), st()), 'tfb_d8_break_ties')
        n_reps = 0
        if link_flats:
            nxt = code16.clone()
            cnt = torch.zeros(2, dtype=torch.int32, device=dev)
            own = lambda t: t[H:H + ny]
            while True:
                n_reps += 1
                cnt.zero_()
                L.check(lib.tfb_d8_link_flats_sweep(
                    own(code16).data_ptr(), own(nxt).data_ptr(), ny, nx, H, layout.row0, NY,
                    cnt.data_ptr(), cnt[1:].data_ptr(), st()), 'tfb_d8_link_flats_sweep')
                tot = cnt.to(torch.int64)
                dist.all_reduce(tot, group=group)
                n_flats, n_fixed = (int(v) for v in tot.cpu())
                if n_flats == 0:
                    break
                code16, nxt = nxt, code16
                exchange_halos([code16], layout, [H], group)
                if n_fixed == n_flats or n_fixed == 0:
                    break
        code = torch.zeros((ny + 2 * H, nx), dtype=torch.uint8, device=dev)
        L.check(lib.tfb_d8_finalize_codes(code16.data_ptr(), code.data_ptr(), code.numel(), st()),
                'tfb_d8_finalize_codes')
        # slope over owned rows +- 1 (clipped): parents may sit in the neighbour slab
        s_lo, s_hi = max(0, layout.row0 - 1), min(NY, layout.row0 + ny + 1)
        n_s = s_hi - s_lo
        # keep the three row arrays alive across the launch (a temporary would be
        # recycled by the caching allocator before the kernel reads it)
        rdx, rdy, rdd = (torch.full((n_s,), float(v), dtype=torch.float32, device=dev)
                         for v in (dx, dy, dd))
        S32 = torch.empty((n_s, nx), dtype=torch.float32, device=dev)
        L.check(lib.tfb_d8_slope(dem[s_lo - top:].data_ptr(),
                                 code[s_lo - (layout.row0 - H):].data_ptr(), rdx.data_ptr(),
                                 rdy.data_ptr(), rdd.data_ptr(), n_s, nx,
                                 S32.data_ptr(), st()), 'tfb_d8_slope')
        # rows s_lo / s_hi-1 are treated as array borders by the kernel; they are
        # either true global borders (code 0) or outside the owned range
        S = S32[layout.row0 - s_lo:layout.row0 - s_lo + ny].double()
    if kinematic:
        good = S > 0
        big = torch.tensor(float('inf'), dtype=torch.float64, device=dev)
        smin = torch.where(good, S, big).min().reshape(1)
        if multi:
            dist.all_reduce(smin, op=dist.ReduceOp.MIN, group=group)
        bad = ~(good & torch.isfinite(S))
        S = torch.where(bad, smin.expand_as(S), S)
    return dict(code=code, S_bed=S.contiguous(), dx=np.array([dx]), dy=np.array([dy]),
                dd=np.array([dd]), n_reps=n_reps)


def slab_d8_area(code, layout, pixel_area, group=None, with_counts=True):
    """Total contributing area of a row-sharded grid (d8_global.update_area_grid
    :1029-1332), bit-identical to the single-GPU pass (tfb_d8_area).

    `code`: this rank's uint8 D8 codes INCLUDING layout.halo >= 1 halo rows (as slab_d8
    returns them); `pixel_area` in the units of the result (km^2: da / 1e6).  Each rank
    walks its own flow paths; where a path leaves the slab the {area, count} words of the
    boundary rows cross to the neighbour (two rows per exchange) and the walk resumes there
    -- one round per slab boundary the longest dependency chain crosses.  Returns
    (A float32 (ny, nx), count int64 (ny, nx) or None, rounds)."""
    lib = L.load()
    ny, nx, H = layout.ny, layout.nx, layout.halo
    dev = code.device
    st = L.stream_ptr
    if layout.world_size > 1 and H < 1:
        raise ValueError('slab_d8_area needs at least one halo row of codes')
    c1 = torch.zeros((ny + 2, nx), dtype=torch.uint8, device=dev)      # exactly one halo row
    if H >= 1:
        c1.copy_(code[H - 1:H + ny + 1])
    else:
        c1[1:ny + 1].copy_(code)
    cell = torch.zeros((ny + 2, nx), dtype=torch.int64, device=dev)
    word = torch.zeros((ny + 2, nx), dtype=torch.int32, device=dev)
    own = lambda t: t.data_ptr() + nx * t.element_size()
    L.check(lib.tfb_d8_area_slab_init(own(c1), ny, nx, layout.row0, layout.ny_global, own(cell),
                                      own(word), st()), 'tfb_d8_area_slab_init')
    L.check(lib.tfb_d8_area_slab_walk(ny, nx, float(pixel_area), None, own(cell), own(word), 0, st()),
            'tfb_d8_area_slab_walk')
    rounds = 0
    if layout.world_size > 1:
        lay1 = SlabLayout(layout.ny_global, nx, layout.rank, layout.world_size, 1)
        sent = None
        while True:
            edge = torch.stack((cell[1], cell[ny]))
            changed = 1 if (sent is None or not torch.equal(edge, sent)) else 0
            flag = torch.tensor([changed], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag.item()) == 0:
                break
            sent = edge
            exchange_halos([cell], lay1, [1], group)
            L.check(lib.tfb_d8_area_slab_walk(ny, nx, float(pixel_area), None, own(cell), own(word),
                                              1, st()), 'tfb_d8_area_slab_walk')
            rounds += 1
    A = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    cnt = torch.empty((ny, nx), dtype=torch.int64, device=dev) if with_counts else None
    n_done = C.c_int64(0)
    L.check(lib.tfb_d8_area_unpack(own(cell), ny * nx, A.data_ptr(),
                                   None if cnt is None else cnt.data_ptr(), C.byref(n_done), st()),
            'tfb_d8_area_unpack')
    return A, cnt, rounds


def slab_new_slope(DEM, code, layout, dem_halo, xres=30.0, yres=30.0, nodata=-9999.0,
                   device=None, group=None):
    """utils/new_slopes.get_new_slope_grid (:174-316) of a row-sharded grid: the drop to the
    first LOWER cell down the D8 flow path over the path length, fp64.

    That walk can cross any number of slabs, and the reference's global iteration count (the
    longest flow path of the WHOLE grid, which decides what walkers that have passed their
    noflow cell are compared with) couples all cells -- while the working set of even the
    65536^2 grid (DEM 17 GB, codes 4 GB, 2 x 17 GB scratch, slope 34 GB) fits the 180 GB of
    ONE B200.  So the owned rows are gathered on rank 0, the single-GPU kernels run there
    (tfb_d8_new_slope) and every rank receives its rows: bit-identical to one GPU by
    construction, one gather + one scatter over NVLink.

    `DEM`: float32 rows with `dem_halo` halo rows, `code`: uint8 with layout.halo halo rows
    (as slab_d8 takes / returns them).  Returns (slope float64 (ny, nx), longest path)."""
    from .d8 import D8Toolkit
    dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    ny, nx, H = layout.ny, layout.nx, layout.halo
    dem = DEM if isinstance(DEM, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(DEM))
    dem = dem.to(dev)[dem_halo:dem_halo + ny].contiguous()
    cod = code.to(dev)[H:H + ny].contiguous()
    if layout.world_size == 1:
        tk = D8Toolkit(dem, xres=xres, yres=yres, nodata=nodata, device=dev)
        tk.code = cod
        return tk.new_slope()
    root = dist.get_global_rank(group, 0) if group is not None else 0
    is_root = layout.rank == 0
    parts_z = [torch.empty((n, nx), dtype=torch.float32, device=dev) for _, n in layout.parts] \
        if is_root else None
    parts_c = [torch.empty((n, nx), dtype=torch.uint8, device=dev) for _, n in layout.parts] \
        if is_root else None
    dist.gather(dem, parts_z, dst=root, group=group)
    dist.gather(cod, parts_c, dst=root, group=group)
    out = torch.empty((ny, nx), dtype=torch.float64, device=dev)
    reps = torch.zeros(1, dtype=torch.int64, device=dev)
    pieces = None
    if is_root:
        tk = D8Toolkit(torch.cat(parts_z, 0), xres=xres, yres=yres, nodata=nodata, device=dev)
        del parts_z
        tk.code = torch.cat(parts_c, 0)
        S, n = tk.new_slope()
        reps[0] = n
        pieces = [S[r0:r0 + n_].contiguous() for r0, n_ in layout.parts]
    dist.scatter(out, pieces, src=root, group=group)
    dist.broadcast(reps, src=root, group=group)
    return out, int(reps.item())
