"""Device-resident routing state + one-call-per-step driver of the fused kernel.

``ChannelsEngine`` owns the torch CUDA tensors of one row slab (or the whole
grid), fills a ``tfb_channels_step_args`` once, and advances the slab with
``tfb_channels_step`` (libtfb_b200.so).  It is the layer between the BMI
component (``channels_base.py`` in this package) and the C ABI; it has no
numerical code of its own apart from the one-time initial volume
(channels_base.py:1331-1336 in the reference).

Layout: every per-cell tensor is allocated as (ny + 2*halo, pitch) and the kernels
get a pointer to the first OWNED row, so neighbour rows of an adjacent slab are
ordinary memory.  ``halo = 0`` for a single GPU.  ``pitch`` = nx rounded up to a
multiple of 4 elements (= nx for every grid whose width already is): the TMA unit
needs 16-byte row strides for the 8-byte and the 4-byte streams alike.  Columns
[nx, pitch) hold zeros (D8 code 0, no water), are never stepped and never show in
``own()`` views; the one-time set-up kernels simply see a grid ``pitch`` wide.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import lib as L

_DIAG_NAMES = ('Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
_SRC_SG = ('P_rain', 'SM', 'GW', 'MR', 'IN', 'ET', 'T_air', 'c0', 'T0', 'T_surf',
           'Qn_SW', 'Qn_LW', 'Ks', 'Ki', 'qs', 'qi', 'G', 'h_table', 'elev')


def row_pitch(nx):
    """Elements per row of every per-cell device array: nx rounded up to 4 (16-byte
    row strides for fp64, fp32 and 32-bit words alike; include/tfb.h `pitch`)."""
    return (int(nx) + 3) // 4 * 4


def _is_grid(x):
    return (isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim == 2)


class ChannelsEngine(object):

    def __init__(self, ny, nx, dt, code, dx, dy, dd, S_bed, width, angle,
                 rough, tan_angle=None, cos_angle=None, sinu=1.0, da=900.0,
                 d0=0.0, kinematic=True, manning=True, diag=False,
                 write_R=False, outlet=(0, 0), rho_H2O=1000.0, row0=0,
                 ny_global=None, halo=0, device=None, finalize=True,
                 n_pixels_global=None, packed=True, flood=False, d_bankfull=10.0,
                 flood_manning_n=0.10, width_ratio=5.0, attenuate=False, n_days=5.0,
                 q_buffers=None, d_buffer=None):
        if not torch.cuda.is_available():
            raise L.TfbError('ChannelsEngine needs a CUDA device (no CPU path)')
        self.lib = L.load()
        self.device = torch.device(device if device is not None
                                   else 'cuda:%d' % torch.cuda.current_device())
        self.ny, self.nx, self.halo = int(ny), int(nx), int(halo)
        self.pitch = row_pitch(self.nx)
        self.row0 = int(row0)
        self.ny_global = int(ny_global if ny_global is not None else ny)
        self.kinematic, self.manning, self.diag = kinematic, manning, diag
        self.dt = float(dt)
        self._keep = []           # tensors referenced only through raw pointers
        a = L.ChannelsStepArgs()
        self.args = a
        a.ny, a.nx, a.row0, a.ny_global = self.ny, self.nx, self.row0, self.ny_global
        a.halo = self.halo
        a.pitch = self.pitch
        a.outlet_row, a.outlet_col = int(outlet[0]), int(outlet[1])
        a.dt = self.dt
        a.rho_H2O = float(rho_H2O)
        a.n_pixels_global = int(n_pixels_global if n_pixels_global is not None
                                else self.ny_global * self.nx)
        a.finalize = 1 if finalize else 0
        flags = 0
        if not kinematic:
            flags |= L.CH_DIFFUSIVE
        if not manning:
            flags |= L.CH_LAW_OF_WALL
        if diag:
            flags |= L.CH_DIAG
        if write_R:
            flags |= L.CH_WRITE_R
        self.flood = bool(flood)
        if self.flood:
            flags |= L.CH_FLOOD
        self.attenuate = bool(attenuate)
        if self.attenuate:
            flags |= L.CH_ATTENUATE
            a.t_drain = 3600.0 * 24.0 * float(n_days)             # channels_base.py:2079
        self.flags = flags
        # ---- static inputs --------------------------------------------------
        self.code = self._grid(code, torch.uint8)
        # packed geo word per cell: D8 code + child mask (parent-side gather)
        self.geo = self._zeros(torch.int16)
        L.check(self.lib.tfb_channels_pack_geo(self._p(self.code), self.ny, self.pitch, self.halo,
                                               self._p(self.geo), L.stream_ptr()),
                'tfb_channels_pack_geo')
        a.geo = self._p(self.geo)
        self.dx = self._rows(dx, torch.float32)
        self.dy = self._rows(dy, torch.float32)
        self.dd = self._rows(dd, torch.float32)
        a.dx, a.dy, a.dd = self._prow(self.dx), self._prow(self.dy), self._prow(self.dd)
        self.S_bed = self._grid(S_bed, torch.float64)
        a.S_bed = self._p(self.S_bed)
        if kinematic and manning and not diag:
            # lean kinematic step only ever uses sqrt(|S_bed|) (channels_base
            # :3982-3983); evaluate it once -- by numpy when S_bed came as numpy,
            # so the value is the very one np.sqrt gives the reference
            if isinstance(S_bed, torch.Tensor):
                self.sqrtS = torch.sqrt(torch.abs(self.S_bed))
            else:
                self.sqrtS = self._grid(np.sqrt(np.abs(np.asarray(S_bed, dtype='float64'))),
                                        torch.float64)
            a.sqrtS = self._p(self.sqrtS)
        # tan / cos of the bank angle come from numpy on the host when the
        # angle is handed over as numpy (bit-identical to the reference's
        # np.tan / np.cos); torch inputs are evaluated on the device.
        if tan_angle is None:
            tan_angle = np.tan(angle) if not isinstance(angle, torch.Tensor) else torch.tan(angle)
        if cos_angle is None:
            cos_angle = np.cos(angle) if not isinstance(angle, torch.Tensor) else torch.cos(angle)
        for name, val in (('sinu', sinu), ('width', width),
                          ('tan_angle', tan_angle), ('cos_angle', cos_angle),
                          ('rough', rough)):
            setattr(a, name, self._sg(name, val))
        if isinstance(da, (np.ndarray, torch.Tensor)) and np.ndim(da) >= 1:
            da_rows = da[:, 0] if np.ndim(da) == 2 else da
            self.da_rows = self._rows(da_rows, torch.float64)
            a.da = L.SG(self._prow(self.da_rows), 0.0)
            self._da_is_grid = True
        else:
            a.da = L.SG(None, float(da))
            self._da_is_grid = False
        # ---- state ---------------------------------------------------------
        z = lambda: self._zeros(torch.float64)
        self.vol_buf = [z(), z() if not kinematic else None]
        # the two discharge buffers may come from peer-mappable allocations (slab.PeerHalos)
        self.Q_buf = list(q_buffers) if q_buffers is not None else [z(), z()]
        for t in self.Q_buf:
            if tuple(t.shape) != (self.ny + 2 * self.halo, self.pitch) or t.dtype != torch.float64:
                raise L.TfbError('q_buffers must be (ny + 2*halo, pitch) float64 CUDA tensors')
        self.qcur = 0             # Q buffer the NEXT step reads
        self.scur = 0             # state buffer (vol, I, h_swe) the next step reads
        self.d, self.u = (d_buffer if d_buffer is not None else z()), z()
        if tuple(self.d.shape) != (self.ny + 2 * self.halo, self.pitch) or self.d.dtype != torch.float64:
            raise L.TfbError('d_buffer must be a (ny + 2*halo, pitch) float64 CUDA tensor')
        a.d, a.u = self._p(self.d), self._p(self.u)
        if diag:
            for n in _DIAG_NAMES:
                t = z()
                setattr(self, n, t)
                setattr(a, n, self._p(t))
            if not kinematic:
                self.S_free = z()
                self.own(self.S_free).copy_(self.own(self.S_bed))
                a.S_free = self._p(self.S_free)
        if write_R:
            self.R = z()
            a.R = self._p(self.R)
        # ---- sources default to scalar zeros --------------------------------
        self.src = {}
        self._src_buf, self._stage_buf, self._registered, self._seen = {}, {}, {}, {}
        for n in _SRC_SG:
            self.set_input(n, 0.0)
        self.h_swe_buf = self.I_buf = None
        self._h_snow = None
        # ---- reductions ----------------------------------------------------
        self.scalars_dev = torch.zeros(C.sizeof(L.ChannelsScalars), dtype=torch.uint8,
                                       device=self.device)
        npart = self.lib.tfb_channels_partials_len(self.ny, self.nx)
        self.partials = torch.zeros(npart, dtype=torch.float64, device=self.device)
        self.ticket = torch.zeros(2, dtype=torch.int32, device=self.device)
        a.scalars = self.scalars_dev.data_ptr()
        a.partials = self.partials.data_ptr()
        a.ticket = self.ticket.data_ptr()
        # pinned host mirror of the scalar block: with unified addressing the device can
        # write it directly, so reading Q_outlet costs a stream synchronise, not a copy
        self._scalars_host = torch.zeros(C.sizeof(L.ChannelsScalars), dtype=torch.uint8).pin_memory()
        self._scalars_np = self._scalars_host.numpy()
        self._scalars_f64 = self._scalars_np.view(np.float64)
        self._scalars_i64 = self._scalars_np.view(np.int64)
        a.scalars_mirror = self._scalars_host.data_ptr()
        self.set_initial_depth(d0)
        self.n_steps = 0
        self.d_flood = self.vol_flood = self.vol_bankfull = None
        self.vol_stored_buf = None
        if self.attenuate:
            self.vol_stored_buf = [self._zeros(torch.float64),
                                   self._zeros(torch.float64) if not kinematic else None]
        if self.flood:
            self._init_flood(d_bankfull, flood_manning_n, width_ratio)
        self.packed = None
        self.lean_ok = self.ext_ok = False
        self.last_path = self.last_kernel = None
        # per-step host work is on the critical path of a synchronous update(): the kernel
        # choice and the buffer pointers of each double-buffer parity are computed once and
        # reused until an input / option changes
        self._dev_index = self.device.index if self.device.index is not None \
            else torch.cuda.current_device()
        self._cfg, self._cfg_key, self._ptr_cache = None, None, {}
        self.serpentine = os.environ.get('TFB_SERPENTINE', '1') != '0'
        self._scalars_dirty = False
        self.mid_step_hook = None
        self.use_packed = bool(packed)
        self._angle_in = angle
        self._rough_in = rough
        if self.use_packed:
            self._try_pack()

    # ------------------------------------------------------------ flood option
    def _ds32(self):
        """float32 D8 flow lengths of the haloed (pitch-wide) grid (d8_global.py:960-1027)."""
        ds32 = torch.empty((self.ny + 2 * self.halo, self.pitch), dtype=torch.float32,
                           device=self.device)
        L.check(self.lib.tfb_d8_ds_dw(self.code.data_ptr(), self.dx.data_ptr(),
                                      self.dy.data_ptr(), self.dd.data_ptr(),
                                      self.ny + 2 * self.halo, self.pitch,
                                      ds32.data_ptr(), None, L.stream_ptr()), 'tfb_d8_ds_dw')
        return ds32

    def _ds64(self):
        """Flow length ds = sinu * ds_d8 (float64 values of the float32 D8 grid)."""
        a = self.args
        sn = self.sinu_grid if self.sinu_grid is not None else a.sinu.val
        return sn * self._ds32().double()

    def _init_flood(self, d_bankfull, flood_manning_n, width_ratio):
        """FLOOD_OPTION state (channels_base.py:1346-1388): bankfull channel volume
        d_b (w + d_b tan(angle)) ds, flood depth and flood volume grids."""
        a = self.args
        g = lambda name: (getattr(self, name + '_grid') if getattr(self, name + '_grid') is not None
                          else getattr(a, name).val)
        db = (self._grid(d_bankfull, torch.float64) if _is_grid(d_bankfull) else float(d_bankfull))
        L3 = db * g('tan_angle')
        Ac = db * (g('width') + L3)
        vb = Ac * self._ds64()
        self.vol_bankfull = vb.contiguous()
        a.vol_bankfull = L.SG(self._p(self.vol_bankfull), 0.0)
        self.d_flood, self.vol_flood = self._zeros(torch.float64), self._zeros(torch.float64)
        a.d_flood, a.vol_flood = self._p(self.d_flood), self._p(self.vol_flood)
        a.flood_manning_n, a.width_ratio = float(flood_manning_n), float(width_ratio)

    # ------------------------------------------------------------ packed path
    def _try_pack(self):
        """Build the compact static inputs of the packed (TMA ring) kernels once, when the
        configuration allows it (at most 65536 distinct bank angles).  Everything is lossless and
        every per-cell array is (ny, pitch) like the state.  Two families share them:

          lean  tfb_channels_step_packed*: Manning, no per-step diagnostics, scalar forcing,
                optionally fused sources; <= 256 angles (table in shared memory); kinematic
                wave: coef = sqrt|S_bed| / n, diffusive: 1 / n and the float32 bed slope
          ext   tfb_channels_step_ext*: grid forcing, law of the wall, diagnostics, R grid,
                FLOOD_OPTION, ATTENUATE, a sinuosity grid; bed slope and width as float32
                where they are float32 values (RTG files), as fp64 grids otherwise (a scalar
                width like 3.3 m, slopes divided by a sinuosity != 1) + roughness (n or z0)
                + a {tan, cos} table in L2

        Otherwise the generic kernel (tfb_channels_step) runs."""
        a = self.args
        self.lean_ok = self.ext_ok = False
        self._cfg = None
        dev = self.device

        def padc(t):
            """(ny, nx) -> (ny, pitch), zero padding (no copy when pitch == nx)."""
            if t is None or self.pitch == self.nx:
                return None if t is None else t.contiguous()
            out = torch.zeros((self.ny, self.pitch), dtype=t.dtype, device=dev)
            out[:, :self.nx] = t
            return out
        full = lambda g, v: (self.own(g) if g is not None
                             else torch.full((self.ny, self.nx), float(v), dtype=torch.float64,
                                             device=dev))
        width = full(self.width_grid, a.width.val)
        w32 = width.float()
        w32_ok = bool((w32.double() == width).all())
        w64 = None if w32_ok else padc(width)
        del width
        if self.tan_angle_grid is None and self.cos_angle_grid is None:
            uniq = torch.tensor([[a.tan_angle.val, a.cos_angle.val]], dtype=torch.float64,
                                device=dev)
            inv = None
        else:
            tan = full(self.tan_angle_grid, a.tan_angle.val)
            cos = full(self.cos_angle_grid, a.cos_angle.val)
            key = torch.stack((tan.reshape(-1), cos.reshape(-1)), dim=1)
            uniq, inv = torch.unique(key, dim=0, return_inverse=True)
            del key, tan, cos
            if uniq.shape[0] > 65536:
                return
        n_angles = int(uniq.shape[0])
        t, c = uniq[:, 0], uniq[:, 1]
        lut = torch.zeros((max(n_angles, 1), 4), dtype=torch.float64, device=dev)
        lut[:, 0] = t
        lut[:, 1] = torch.where(t == 0, torch.zeros_like(t), 1.0 / (2.0 * t))
        lut[:, 2] = 1.0 / c
        lut[:, 3] = 2.0 * (2.0 * t)
        lut_tc = uniq.contiguous()                        # {tan, cos} as the reference uses them
        aidx = (None if inv is None
                else padc(inv.to(torch.int16).reshape(self.ny, self.nx)))   # 16-bit pattern
        rough = full(self.rough_grid, a.rough.val)
        # the bed slope as float32 (RTG precision): lossless unless sinuosity != 1 rescaled it
        S = self.own(self.S_bed)
        S32 = S.float()
        s32_ok = bool((S32.double() == S).all())
        S32 = padc(S32) if s32_ok else None
        coef = inv_rough = None
        S64 = None if s32_ok else padc(S)
        sinu64 = None if self.sinu_grid is None else padc(self.own(self.sinu_grid))
        lean_geom = w32_ok and sinu64 is None          # the lean kernels: float32 widths, row-wise ds
        if self.manning and lean_geom:
            if self.kinematic:
                sq = self.own(self.sqrtS) if getattr(self, 'sqrtS', None) is not None \
                    else torch.sqrt(torch.abs(S))
                coef = padc(sq / rough)
            elif s32_ok:
                inv_rough = padc(1.0 / rough)
        rough64 = padc(rough)
        geo32 = torch.empty((self.ny, self.pitch), dtype=torch.int32, device=dev)
        L.check(self.lib.tfb_channels_pack_geo32(self._p(self.code),
                                                 None if aidx is None else aidx.data_ptr(), self.ny,
                                                 self.pitch, self.halo, geo32.data_ptr(),
                                                 L.stream_ptr()), 'tfb_channels_pack_geo32')
        rowg = torch.empty((self.ny, 8), dtype=torch.float64, device=dev)
        L.check(self.lib.tfb_channels_pack_rows(self._prow(self.dx), self._prow(self.dy),
                                                self._prow(self.dd),
                                                1.0 if sinu64 is not None else float(a.sinu.val),
                                                self._prow(self.da_rows) if self._da_is_grid
                                                else None, float(a.da.val), self.ny,
                                                rowg.data_ptr(), L.stream_ptr()),
                'tfb_channels_pack_rows')
        orow = a.outlet_row - self.row0
        n_out = 0.0
        if 0 <= orow < self.ny and 0 <= a.outlet_col < self.nx:
            n_out = float(rough[orow, a.outlet_col])
        pk = L.ChannelsPacked()
        w32 = padc(w32) if w32_ok else None
        pk.geo32, pk.width32 = geo32.data_ptr(), (None if w32 is None else w32.data_ptr())
        pk.width64 = None if w64 is None else w64.data_ptr()
        pk.S64 = None if S64 is None else S64.data_ptr()
        pk.sinu64 = None if sinu64 is None else sinu64.data_ptr()
        pk.coef = None if coef is None else coef.data_ptr()
        pk.inv_rough = None if inv_rough is None else inv_rough.data_ptr()
        pk.S32 = None if S32 is None else S32.data_ptr()
        pk.rough64 = None if rough64 is None else rough64.data_ptr()
        pk.angle_lut, pk.row_geom, pk.lut_tc = lut.data_ptr(), rowg.data_ptr(), lut_tc.data_ptr()
        pk.outlet_rough, pk.n_angles = n_out, n_angles
        nf = None
        if not self.kinematic:
            # noflow cells of the owned rows, as flat indices relative to the first owned row
            # (whole pitched rows: the padding columns hold code 0 and depth 0 anyway)
            nf = torch.nonzero(self.code[self.halo:self.halo + self.ny].reshape(-1) == 0) \
                .reshape(-1).contiguous()
            pk.noflow_idx, pk.n_noflow = nf.data_ptr(), int(nf.numel())
        self._pack_keep = (geo32, w32, coef, inv_rough, S32, rough64, lut, lut_tc, rowg, nf, aidx,
                           w64, S64, sinu64)
        self.packed = pk
        self.lean_ok = (self.manning and n_angles <= 256 and not self.flood and not self.attenuate
                        and lean_geom
                        and (coef is not None if self.kinematic else inv_rough is not None))
        self.ext_ok = True

    def release_generic_inputs(self):
        """Free the fp64 static grids the packed path does not read (width, tan,
        cos, roughness, S_bed, sqrt|S_bed|, 16-bit geo): ~42 B/cell.  Afterwards
        only configurations the packed kernel accepts can be stepped."""
        if self.packed is None:
            raise L.TfbError('release_generic_inputs: no packed encoding was built')
        a = self.args
        for n in ('width', 'tan_angle', 'cos_angle', 'rough'):
            if getattr(self, n + '_grid') is not None:
                setattr(self, n + '_grid', None)
                setattr(a, n, L.SG(None, float('nan')))
        self.S_bed = self.sqrtS = self.geo = None
        a.S_bed = a.sqrtS = a.geo = None
        self.generic_released = True

    def _config(self):
        """(kernel family, R-is-a-grid) of the next step, cached until an option or input
        binding changes (`set_input`, `enable_fused_sources`, the flag word)."""
        key = (self.flags, self.use_packed, self.diag)
        if self._cfg is None or self._cfg_key != key:
            self._cfg = (self._packed_ok(), 1 if self.is_R_grid() else 0)
            self._cfg_key = key
        return self._cfg

    def _packed_ok(self):
        """Which packed family takes this step: 'lean' (the benchmarked production kernels),
        'ext' (grid forcing / law of the wall / diagnostics / R grid / flood option /
        ATTENUATE / sinuosity grid / non-float32 widths or slopes), or None (generic kernel:
        fused sources with grid parameters, more than 65536 bank angles)."""
        if self.packed is None or not self.use_packed:
            return None
        src = self.flags & (L.CH_SRC_MET | L.CH_SRC_SNOW | L.CH_SRC_EVAP | L.CH_SRC_INFIL)
        grid_forcing = any(isinstance(self.src[n], torch.Tensor)
                           for n in ('P_rain', 'SM', 'GW', 'MR', 'ET', 'IN'))
        plain = not (self.diag or (self.flags & L.CH_WRITE_R) or not self.manning)
        if self.lean_ok and plain and not grid_forcing:
            # fused sources with scalar parameters; only T_air may be a grid
            if any(isinstance(self.src[n], torch.Tensor) for n in _SRC_SG if n != 'T_air'):
                return None
            if (src & L.CH_SRC_INFIL) and self.args.h_table.val >= self.args.elev.val:
                return None
            return 'lean'
        if self.ext_ok and not src:
            return 'ext'
        return None

    # ------------------------------------------------------------ allocation
    def _zeros(self, dtype):
        return torch.zeros((self.ny + 2 * self.halo, self.pitch), dtype=dtype, device=self.device)

    def own(self, t):
        """(ny, nx) view of the rows this slab owns (without the row padding)."""
        t = t[self.halo:self.halo + self.ny]
        return t if t.shape[1] == self.nx else t[:, :self.nx]

    def _p(self, t):
        return t.data_ptr() + self.halo * t.stride(0) * t.element_size()

    def _prow(self, t):
        return t.data_ptr() + self.halo * t.element_size()

    def _grid(self, x, dtype):
        """Upload an (ny, nx) or (ny+2h, nx) array into a haloed (pitched) tensor."""
        full = self._zeros(dtype)
        src = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
        src = src.to(dtype=dtype, device=self.device)
        if src.shape[0] == self.ny + 2 * self.halo and self.halo:
            full[:, :src.shape[1]].copy_(src)
        else:
            if tuple(src.shape) != (self.ny, self.nx):
                raise L.TfbError('grid shape %s != (%d, %d)' % (tuple(src.shape), self.ny, self.nx))
            self.own(full).copy_(src)
        return full

    def _rows(self, x, dtype):
        src = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
        src = src.to(dtype=dtype, device=self.device).reshape(-1)
        full = torch.zeros(self.ny + 2 * self.halo, dtype=dtype, device=self.device)
        if src.numel() == 1:
            full.fill_(float(src))
        elif src.numel() == self.ny + 2 * self.halo:
            full.copy_(src)
        else:
            full[self.halo:self.halo + self.ny].copy_(src)
            if self.halo:
                full[:self.halo] = src[0]
                full[self.halo + self.ny:] = src[-1]
        return full

    def _sg(self, name, val):
        if _is_grid(val):
            t = self._grid(val, torch.float64)
            setattr(self, name + '_grid', t)
            return L.SG(self._p(t), 0.0)
        setattr(self, name + '_grid', None)
        return L.SG(None, float(val))

    # ---------------------------------------------------------------- inputs
    def set_input(self, name, val):
        """Bind a source-term input: python/0-D scalar, numpy grid (uploaded) or
        CUDA tensor (used in place when it already has the slab's layout)."""
        if name not in _SRC_SG:
            raise KeyError(name)
        self._cfg = None
        if _is_grid(val):
            if (isinstance(val, torch.Tensor) and val.is_cuda and val.dtype == torch.float64
                    and val.is_contiguous() and self.halo == 0 and self.pitch == self.nx
                    and tuple(val.shape) == (self.ny, self.nx)):
                t = val
                sgv = L.SG(t.data_ptr(), 0.0)
            else:
                t = self._src_buf.get(name)          # a device buffer of OURS, kept across steps
                if t is None:
                    t = self._src_buf[name] = self._zeros(torch.float64)
                haloed = bool(self.halo) and val.shape[0] == t.shape[0]
                dst = t[:, :self.nx] if haloed else self.own(t)
                if (isinstance(val, np.ndarray) and val.dtype == np.float64
                        and val.flags.c_contiguous and val.nbytes >= self.PIN_MIN_BYTES):
                    self._upload(name, val, dst)
                else:
                    src = torch.as_tensor(val) if not isinstance(val, torch.Tensor) else val
                    dst.copy_(src.to(dtype=torch.float64, device=self.device))
                sgv = L.SG(self._p(t), 0.0)
            self.src[name] = t
        else:
            v = float(val) if not isinstance(val, torch.Tensor) else float(val.item())
            self.src[name] = v
            sgv = L.SG(None, v)
        setattr(self.args, name, sgv)
        if name == 'ET':
            self.args.ET_rw = sgv.ptr

    # ---- host grids: pinned transfers ---------------------------------------------------
    # A coupled run hands the component (ny, nx) float64 numpy grids every step (EMELI's
    # set_value, time_interpolation.py:653-670).  A pageable cudaMemcpy moves them at ~10 GB/s
    # (measured on the B200 box: 13 ms per 4096^2 grid); here a grid goes through a pinned
    # staging buffer (multi-threaded host copy + DMA at ~54 GB/s: ~5.4 ms), and an array that
    # comes back at the same address -- providers that update their grids in place -- is
    # page-locked where it lies (cudaHostRegister, once) and moved by DMA alone (2.5 ms).
    PIN_MIN_BYTES = 1 << 20
    PIN_PROVIDER_ARRAYS = True

    def _stage(self, name):
        st = self._stage_buf.get(name)
        if st is None:
            st = self._stage_buf[name] = [torch.empty((self.ny + 2 * self.halo, self.nx),
                                                      dtype=torch.float64).pin_memory(), None]
        if st[1] is not None:
            st[1].synchronize()                    # the DMA that last used this buffer is done
            st[1] = None
        return st

    def _is_registered(self, arr):
        return self._registered.get(arr.ctypes.data) == arr.nbytes

    def _maybe_register(self, arr):
        """Page-lock a provider's array in place on its second sighting; unregistered again
        when numpy frees it."""
        ptr, nb = arr.ctypes.data, arr.nbytes
        if self._registered.get(ptr) == nb:
            return True
        if not self.PIN_PROVIDER_ARRAYS or self._registered.get(ptr) == -1:
            return False
        n = self._seen.get(ptr, 0) + 1
        if len(self._seen) > 64:
            self._seen.clear()
        self._seen[ptr] = n
        if n < 2:
            return False
        rt = torch.cuda.cudart()
        if int(rt.cudaHostRegister(ptr, nb, 0)) != 0:
            self._registered[ptr] = -1             # not page-lockable: staging buffer from now on
            return False
        self._registered[ptr] = nb
        import weakref
        reg = self._registered

        def _release(p=ptr):
            if reg.pop(p, None) not in (None, -1):
                rt.cudaHostUnregister(p)
        base = arr
        while isinstance(getattr(base, 'base', None), np.ndarray):
            base = base.base                       # the object that owns the memory
        try:
            weakref.finalize(base, _release)
        except TypeError:
            pass
        return True

    def _upload(self, name, arr, dst):
        """numpy (rows, nx) float64 -> device view `dst`, asynchronously on the current stream."""
        if self._maybe_register(arr):
            dst.copy_(torch.from_numpy(arr), non_blocking=True)
            return
        st = self._stage(name)
        h = st[0][:arr.shape[0]]
        h.copy_(torch.from_numpy(arr))             # host copy into pinned memory
        dst.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st[1] = ev

    def download(self, name, src, arr):
        """Device view `src` -> the provider's numpy array (in-place ET masking,
        channels_base.py:1625-1627).  Returns after the data has landed."""
        if (arr.dtype == np.float64 and arr.flags.c_contiguous and arr.nbytes >= self.PIN_MIN_BYTES):
            if self._is_registered(arr):
                torch.from_numpy(arr).copy_(src, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                return
            st = self._stage(name)
            h = st[0][:arr.shape[0]]
            h.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            torch.from_numpy(arr).copy_(h)
            return
        torch.from_numpy(arr).copy_(src)

    def enable_fused_sources(self, met=False, snow=False, evap=False, infil=False,
                             T_rain_snow=0.0, rho_snow=300.0, alpha=1.26,
                             K_soil=0.45, soil_x=0.05, T_soil_x=0.0,
                             h0_swe=0.0, I0=1e-6, use_table=False, materialize_h_snow=False):
        """Evaluate the met/snow/evap/infil source terms inside the routing
        kernel (TFB_CH_SRC_* flags) instead of taking them as inputs."""
        a = self.args
        self._cfg = None
        self._ptr_cache.clear()
        f = self.flags & ~(L.CH_SRC_MET | L.CH_SRC_SNOW | L.CH_SRC_EVAP | L.CH_SRC_INFIL)
        dbl = not self.kinematic
        if met:
            f |= L.CH_SRC_MET
            a.T_rain_snow = float(T_rain_snow)
        if snow:
            f |= L.CH_SRC_SNOW
            a.rho_snow = float(rho_snow)
            self.h_swe_buf = [self._grid_or_fill(h0_swe), self._zeros(torch.float64) if dbl else None]
            # snow depth = h_swe * rho_H2O / rho_snow (snow_base.py:631-703) is a pure
            # function of the state: materialised every step only on request
            if materialize_h_snow:
                self._h_snow = self._zeros(torch.float64)
                self.own(self._h_snow).copy_(self.own(self.h_swe_buf[0]) * (a.rho_H2O / a.rho_snow))
                a.h_snow = self._p(self._h_snow)
            else:
                self._h_snow = None
                a.h_snow = None
        if evap:
            f |= L.CH_SRC_EVAP
            a.alpha, a.K_soil = float(alpha), float(K_soil)
            a.soil_x, a.T_soil_x = float(soil_x), float(T_soil_x)
        if infil:
            f |= L.CH_SRC_INFIL
            self.I_buf = [self._grid_or_fill(I0), self._zeros(torch.float64) if dbl else None]
            if not use_table:
                a.h_table = L.SG(None, 0.0)
                a.elev = L.SG(None, 1.0)
        self.flags = f

    def _grid_or_fill(self, v):
        if _is_grid(v):
            return self._grid(v, torch.float64)
        t = self._zeros(torch.float64)
        self.own(t).fill_(float(v))
        return t

    def set_initial_depth(self, d0):
        """d = d0; vol = A_wet(d0) * ds   (channels_base.py:1262-1265, :1331-1336)."""
        a = self.args
        d = self.own(self.d)
        if _is_grid(d0):
            t = torch.as_tensor(d0).to(device=self.device, dtype=torch.float64)
            if self.halo and t.shape[0] == self.ny + 2 * self.halo:
                t = t[self.halo:self.halo + self.ny]       # a slab's rows came with their halo
            d.copy_(t)
        else:
            d.fill_(float(d0))
        vol = self.own(self._sbuf(self.vol_buf))
        def g(name):
            t = getattr(self, name + '_grid')
            return self.own(t) if t is not None else getattr(a, name).val
        def as_divisor(v):       # a python-scalar divisor would let torch multiply by 1/v
            return v if isinstance(v, torch.Tensor) else torch.full(
                (1, 1), float(v), dtype=torch.float64, device=self.device)
        if getattr(self, 'P_wet', None) is not None:
            # the diagnostics exist from initialize() on (channels_base.py:1331-1333)
            self.own(self.P_wet).copy_(g('width') + ((2.0 * d) / as_divisor(g('cos_angle'))))
            self.own(self.A_wet).zero_()
        if float(d.abs().max()) == 0.0:
            vol.zero_()
            return
        ds = g('sinu') * self.own(self._ds32()).double()
        L2 = d * g('tan_angle')
        A_wet = d * (g('width') + L2)
        vol.copy_(A_wet * ds)
        if getattr(self, 'A_wet', None) is not None:
            self.own(self.A_wet).copy_(A_wet)

    # ------------------------------------------------------------------ step
    # Q is always double-buffered (neighbours read the old Q while the new one
    # is written); vol / I / h_swe only for the diffusive wave, where the
    # parent's new depth is re-derived from its pre-step state.
    def _sbuf(self, buf):
        return buf[self.scur if buf[1] is not None else 0]

    @property
    def vol(self):
        return self.own(self._sbuf(self.vol_buf))

    @property
    def vol_stored(self):
        return self.own(self._sbuf(self.vol_stored_buf))

    @property
    def I(self):
        return self.own(self._sbuf(self.I_buf))

    @property
    def h_swe(self):
        return self.own(self._sbuf(self.h_swe_buf))

    @property
    def h_snow(self):
        if self._h_snow is not None:
            return self.own(self._h_snow)
        if self.h_swe_buf is None:
            return None
        return self.h_swe * (self.args.rho_H2O / self.args.rho_snow)

    @property
    def Q(self):
        """Discharge the LAST step used (= what the reference exposes as Q)."""
        return self.own(self.Q_buf[self.qcur ^ 1] if self.n_steps else self.Q_buf[self.qcur])

    @property
    def Q_next(self):
        """u * A_wet after the last step: the next step's discharge."""
        return self.own(self.Q_buf[self.qcur])

    def is_R_grid(self):
        if self._da_is_grid or (self.flags & 0xF00):
            return True
        for n in ('P_rain', 'SM', 'GW', 'MR', 'IN', 'ET'):
            if isinstance(self.src[n], torch.Tensor):
                return True
        return False

    def step(self, time_min=0.0, peer=None, peer_step=None):
        """Advance one time step.  `peer` (slab.PeerHalos) switches the packed step to the
        fused peer-memory halo exchange (kinematic: one launch; diffusive: both passes);
        `peer_step` = how many steps (diffusive: phases) used it before this one -- the
        value the flag words are compared with (default: every step so far did)."""
        a = self.args
        st = L.stream_ptr(self._dev_index)
        # odd kinematic steps walk the grid bottom-up (TFB_CH_REVERSE, include/tfb.h): they start
        # on the rows the previous step wrote last, which are still in L2
        a.flags = (self.flags | (L.CH_ET_INPLACE if a.ET.ptr else 0)
                   | (L.CH_REVERSE if (self.serpentine and
                                       ((self.n_steps & 1) or not self.kinematic)) else 0))
        a.time_min = float(time_min)
        ptrs = self._ptr_cache.get((self.qcur, self.scur))
        if ptrs is None:
            ptrs = [('Q_in', self._p(self.Q_buf[self.qcur])),
                    ('Q_out', self._p(self.Q_buf[self.qcur ^ 1]))]
            for name, buf in (('vol', self.vol_buf), ('h_swe', self.h_swe_buf),
                              ('I', self.I_buf), ('vol_stored', self.vol_stored_buf)):
                if buf is None:
                    continue
                dbl = buf[1] is not None
                ptrs.append((name, self._p(buf[self.scur if dbl else 0])))
                ptrs.append((name + '_out', self._p(buf[self.scur ^ 1 if dbl else 0])))
            self._ptr_cache[(self.qcur, self.scur)] = ptrs
        for name, p in ptrs:
            setattr(a, name, p)
        path, a.is_R_grid = self._config()
        if path == 'ext':
            if peer is not None:
                raise L.TfbError('the peer-memory halo needs the lean packed step')
            if self.kinematic:
                L.check(self.lib.tfb_channels_step_ext(C.byref(a), C.byref(self.packed),
                                                       st), 'tfb_channels_step_ext')
            else:
                L.check(self.lib.tfb_channels_step_ext_dw_a(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_ext_dw_a')
                if self.mid_step_hook is not None:
                    self.mid_step_hook()
                L.check(self.lib.tfb_channels_step_ext_dw_b(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_ext_dw_b')
                L.check(self.lib.tfb_channels_step_packed_dw_c(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_packed_dw_c')
            self.last_path = 'packed'
            self.last_kernel = 'ext'
        elif path == 'lean':
            self.last_kernel = 'lean'
            if self.kinematic and peer is not None:
                ps = self.n_steps if peer_step is None else int(peer_step)
                ph = peer.descriptor(self.qcur ^ 1, ps)
                L.check(self.lib.tfb_channels_step_packed_p2p(
                    C.byref(a), C.byref(self.packed), C.byref(ph), st),
                    'tfb_channels_step_packed_p2p')
            elif self.kinematic:
                L.check(self.lib.tfb_channels_step_packed(C.byref(a), C.byref(self.packed),
                                                          st), 'tfb_channels_step_packed')
            elif peer is not None:
                # diffusive wave, peer halo: pass A stores its boundary depths, pass B its
                # boundary discharges into the neighbours' halo rows; phases 2n and 2n + 1
                ps = 2 * self.n_steps if peer_step is None else int(peer_step)
                ph = peer.descriptor(0, ps, depth=True)
                L.check(self.lib.tfb_channels_step_packed_dw_a_p2p(
                    C.byref(a), C.byref(self.packed), C.byref(ph), st),
                    'tfb_channels_step_packed_dw_a_p2p')
                ph = peer.descriptor(self.qcur ^ 1, ps + 1)
                L.check(self.lib.tfb_channels_step_packed_dw_b_p2p(
                    C.byref(a), C.byref(self.packed), C.byref(ph), st),
                    'tfb_channels_step_packed_dw_b_p2p')
                L.check(self.lib.tfb_channels_step_packed_dw_c(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_packed_dw_c')
            else:
                # diffusive wave: new depths first, (slab: exchange one row of d), then
                # free-surface slope, velocity and discharge
                L.check(self.lib.tfb_channels_step_packed_dw_a(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_packed_dw_a')
                if self.mid_step_hook is not None:
                    self.mid_step_hook()
                L.check(self.lib.tfb_channels_step_packed_dw_b(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_packed_dw_b')
                L.check(self.lib.tfb_channels_step_packed_dw_c(
                    C.byref(a), C.byref(self.packed), st), 'tfb_channels_step_packed_dw_c')
            self.last_path = 'packed'
        else:
            if peer is not None:
                raise L.TfbError('the peer-memory halo needs the packed step')
            if getattr(self, 'generic_released', False):
                raise L.TfbError('this configuration needs the generic kernel, but its static '
                                 'inputs were released (release_generic_inputs)')
            L.check(self.lib.tfb_channels_step(C.byref(a), st), 'tfb_channels_step')
            self.last_path = 'generic'
            self.last_kernel = 'generic'
        self.qcur ^= 1
        self.scur ^= 1
        self.n_steps += 1

    def read_scalars(self):
        """The scalar block after the last step.  The kernel has already written it
        into the pinned host mirror; this only synchronises the stream.  (Before the
        first step, or after write_scalars, the block is copied explicitly.)"""
        self.sync_scalar_mirror()
        return L.ChannelsScalars.from_buffer_copy(self._scalars_np)

    def sync_scalar_mirror(self):
        """Wait until the pinned host mirror holds the scalar block of the last step;
        returns it as float64 words (the 20 fp64 fields first, lib.ChannelsScalars)."""
        if self.n_steps == 0 or self._scalars_dirty:
            self._scalars_host.copy_(self.scalars_dev, non_blocking=True)
            self._scalars_dirty = False
        L.check(self.lib.tfb_stream_synchronize(L.stream_ptr(self._dev_index)),
                'tfb_stream_synchronize')
        return self._scalars_f64

    def write_scalars(self, s):
        raw = torch.frombuffer(bytearray(bytes(s)), dtype=torch.uint8)
        self.scalars_dev.copy_(raw)
        self._scalars_dirty = True
