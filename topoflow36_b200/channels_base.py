"""GPU Channels component: drop-in for the reference's ``channels_component``.

Mirrors ``topoflow/components/channels_base.py`` (class ``channels_component``
:181): the same BMI surface (``initialize`` :633, ``update`` :803, ``finalize``
:964, the CSDMS long names / short names / units of :186-468), driven by the same
``<case>_channels_*.cfg`` + ``<case>_d8_global.cfg`` + RTI/RTG files -- but the
state lives in torch CUDA tensors and every ``update()`` is ONE launch of the
fused sm_100a routing kernel through the C ABI (``tfb_channels_step``,
include/tfb.h).  There is no CPU path: without the CUDA library/device
``initialize()`` raises ``TfbError``.

What the caller sees:
  * the 7 inputs arrive through ``set_value`` as 0-D numpy arrays, (ny, nx)
    numpy grids (copied host->device each update) or CUDA tensors (used in place);
  * outlet / peak / integral scalars are 0-D numpy arrays refreshed IN PLACE after
    every update by one small device->host copy (``tfb_channels_scalars``), so
    EMELI users holding references (e.g. topoflow_driver reading Q_outlet for its
    stop rule) see them change;
  * grids (Q, d, u, vol, ...) are numpy mirrors fetched from the device when the
    attribute is read; a grid handed out by ``get_value_ptr`` is re-filled in
    place after every update (EMELI's by-reference convention);
    ``get_value_tensor`` returns the device tensor itself;
  * the provider's ET grid is zeroed in place where ``vol < ET*da*dt``
    (channels_base.py:1625-1627), on the device and -- for a numpy provider --
    copied back into the provider's array;
  * ``CHECK_STABILITY`` raises the reference's RuntimeErrors (:3275, :3429) from
    the bad-value counters the kernel returns.
FLOOD_OPTION (rectangle-over-trapezoid overbank flow) is supported through the
generic kernel, as is ATTENUATE (per-cell linear reservoir).  Not handled (raise
NotImplementedError): dynamic wave,
CREATE_CHANNEL_FILES.
"""
import os

import numpy as np

from . import BMI_base, grid_io
from . import lib as L

_IN = [
    ('atmosphere_water__rainfall_volume_flux', 'P_rain', 'm s-1'),
    ('glacier_ice__melt_volume_flux', 'MR', 'm s-1'),
    ('land_surface_water__baseflow_volume_flux', 'GW', 'm s-1'),
    ('land_surface_water__evaporation_volume_flux', 'ET', 'm s-1'),
    ('soil_surface_water__infiltration_volume_flux', 'IN', 'm s-1'),
    ('snowpack__melt_volume_flux', 'SM', 'm s-1'),
    ('water-liquid__mass-per-volume_density', 'rho_H2O', 'kg m-3')]

_OUT = [
    ('basin_outlet_water_flow__half_of_fanning_friction_factor', 'f_outlet', '1'),
    ('basin_outlet_water_x-section__mean_depth', 'd_outlet', 'm'),
    ('basin_outlet_water_x-section__peak_time_of_depth', 'Td_peak', 'min'),
    ('basin_outlet_water_x-section__peak_time_of_volume_flow_rate', 'T_peak', 'min'),
    ('basin_outlet_water_x-section__peak_time_of_volume_flux', 'Tu_peak', 'min'),
    ('basin_outlet_water_x-section__time_integral_of_volume_flow_rate', 'vol_Q', 'm3'),
    ('basin_outlet_water_x-section__time_max_of_mean_depth', 'd_peak', 'm'),
    ('basin_outlet_water_x-section__time_max_of_volume_flow_rate', 'Q_peak', 'm3 s-1'),
    ('basin_outlet_water_x-section__time_max_of_volume_flux', 'u_peak', 'm s-1'),
    ('basin_outlet_water_x-section__volume_flow_rate', 'Q_outlet', 'm3 s-1'),
    ('basin_outlet_water_x-section__volume_flux', 'u_outlet', 'm s-1'),
    ('canals_entrance_water__volume_flow_rate', 'Q_canals_in', 'm3 s-1'),
    ('channel_bottom_surface__slope', 'S_bed', '1'),
    ('channel_bottom_water_flow__domain_max_of_log_law_roughness_length', 'z0val_max', 'm'),
    ('channel_bottom_water_flow__domain_min_of_log_law_roughness_length', 'z0val_min', 'm'),
    ('channel_bottom_water_flow__log_law_roughness_length', 'z0val', 'm'),
    ('channel_bottom_water_flow__magnitude_of_shear_stress', 'tau', 'kg m-1 s-2'),
    ('channel_bottom_water_flow__shear_speed', 'u_star', 'm s-1'),
    ('channel_centerline__sinuosity', 'sinu', '1'),
    ('channel_water__volume', 'vol', 'm3'),
    ('channel_water_flow__froude_number', 'froude', '1'),
    ('channel_water_flow__half_of_fanning_friction_factor', 'f', '1'),
    ('channel_water_flow__domain_max_of_manning_n_parameter', 'nval_max', 'm-1/3 s'),
    ('channel_water_flow__domain_min_of_manning_n_parameter', 'nval_min', 'm-1/3 s'),
    ('channel_water_flow__manning_n_parameter', 'nval', 'm-1/3 s'),
    ('channel_water_surface__slope', 'S_free', '1'),
    ('channel_water_x-section__domain_max_of_mean_depth', 'd_max', 'm'),
    ('channel_water_x-section__domain_min_of_mean_depth', 'd_min', 'm'),
    ('channel_water_x-section__domain_max_of_volume_flow_rate', 'Q_max', 'm3 s-1'),
    ('channel_water_x-section__domain_min_of_volume_flow_rate', 'Q_min', 'm3 s-1'),
    ('channel_water_x-section__domain_max_of_volume_flux', 'u_max', 'm s-1'),
    ('channel_water_x-section__domain_min_of_volume_flux', 'u_min', 'm s-1'),
    ('channel_water_x-section__hydraulic_radius', 'Rh', 'm'),
    ('channel_water_x-section__initial_mean_depth', 'd0', 'm'),
    ('channel_water_x-section__mean_depth', 'd', 'm'),
    ('channel_water_x-section__volume_flow_rate', 'Q', 'm3 s-1'),
    ('channel_water_x-section__volume_flux', 'u', 'm s-1'),
    ('channel_water_x-section__wetted_area', 'A_wet', 'm2'),
    ('channel_water_x-section__wetted_perimeter', 'P_wet', 'm'),
    ('channel_water_x-section_top__width', 'w_top', 'm'),
    ('channel_x-section_trapezoid_bottom__width', 'width', 'm'),
    ('channel_x-section_trapezoid_side__flare_angle', 'angle', 'rad'),
    ('land_surface_water__depth', 'df', 'm'),
    ('land_surface_water__runoff_volume_flux', 'R', 'm s-1'),
    ('land_surface_water__domain_time_integral_of_runoff_volume_flux', 'vol_R', 'm3'),
    ('model__time_step', 'dt', 's'),
    ('model_grid_cell__area', 'da', 'm2'),
    ('channel_water_x-section__boundary_time_integral_of_volume_flow_rate', 'vol_edge', 'm3'),
    ('river-network_channel_water__initial_volume', 'vol_chan_sum0', 'm3'),
    ('river-network_channel_water__volume', 'vol_chan_sum', 'm3'),
    ('land_surface_water__area_integral_of_depth', 'vol_flood_sum', 'm3')]

# names the reference also maps although they are neither inputs nor outputs
_EXTRA = [
    ('channel_water_x-section__bankfull_depth', 'd_bankfull', 'm'),
    ('canals__count', 'n_canals', '1'), ('sinks__count', 'n_sinks', '1'),
    ('sources__count', 'n_sources', '1')]

_SCALAR_BLOCK = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet',
                 'f_outlet', 'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak')
# word offsets inside lib.ChannelsScalars (all fields are 8 bytes wide)
_SC_FIELDS = [f[0] for f in L.ChannelsScalars._fields_]
assert tuple(_SC_FIELDS[:len(_SCALAR_BLOCK)]) == _SCALAR_BLOCK
_IX_DTMIN, _IX_FAIL = _SC_FIELDS.index('dt_cfl_min'), _SC_FIELDS.index('n_step_failed')
_IX_CNT = _SC_FIELDS.index('n_neg_d')
assert _SC_FIELDS[_IX_CNT:_IX_CNT + 6] == ['n_neg_d', 'n_nan_d', 'n_inf_d', 'n_neg_u', 'n_nan_u', 'n_inf_u']
_SRC_INPUTS = ('P_rain', 'SM', 'GW', 'MR', 'ET', 'IN')
_DIAG = ('Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
# attribute -> how to find the device tensor (state & diagnostics)
_DEVICE_GRIDS = ('Q', 'Qc', 'vol', 'vol_chan', 'd', 'u', 'R', 'S_free', 'd_flood', 'df',
                 'vol_flood', 'vol_stored') + _DIAG


class channels_component(BMI_base.BMI_component):

    _input_var_names = [t[0] for t in _IN]
    _output_var_names = [t[0] for t in _OUT]
    _var_name_map = {t[0]: t[1] for t in _IN + _OUT + _EXTRA}
    _var_units_map = {t[0]: t[2] for t in _IN + _OUT + _EXTRA}

    def __init__(self):
        BMI_base.BMI_component.__init__(self)
        self.engine = None
        self.device = None            # torch device; default = current CUDA device
        self._mirrors = {}            # grid name -> numpy array handed out by reference
        self._exported = set()        # grid names refreshed after every update
        self.SYNC_SCALARS = True      # D2H the scalar block after every update
        self.DIAGNOSTICS = False      # materialise Rh, A_wet, ... every step
        self._fused = None
        self._last_scalar_in = {}
        self._et_slab_ready = False
        self.layout = None            # SlabLayout in multi-GPU (row-slab) mode
        self.stepper = None           # SlabChannels wrapper in slab mode, else the engine
        self.group = None             # torch.distributed process group (None = world)
        self.P2P_HALO = False         # slab mode, kinematic, uniform forcing: halo rows by
                                      # NVLink peer stores inside the kernel instead of NCCL

    @property
    def input_device(self):
        """torch.device on which set_value() takes CUDA tensors as they are (None before
        initialize()); emeli_dropin.gpu_framework asks the time interpolator for tensors
        on it."""
        e = self.__dict__.get('engine')
        return None if e is None else e.device

    # ---------------------------------------------------------------- attributes
    def __getattr__(self, name):
        # only reached when normal lookup fails: device-backed grids
        if name in _DEVICE_GRIDS and self.__dict__.get('engine') is not None:
            return self._fetch(name)
        raise AttributeError(name)

    def _device_tensor(self, name):
        e = self.engine
        if name == 'vol_stored':
            if not e.attenuate:
                raise AttributeError('vol_stored exists only with ATTENUATE')
            return e.vol_stored
        if name in ('d_flood', 'df', 'vol_flood'):
            if not e.flood:
                raise AttributeError(name + ' exists only with FLOOD_OPTION')
            return e.own(e.vol_flood if name == 'vol_flood' else e.d_flood)
        if name == 'Q' or (name == 'Qc' and not e.flood):
            return e.Q
        if name == 'Qc':
            raise AttributeError('with FLOOD_OPTION only the total discharge Q is kept')
        if name == 'vol' or (name == 'vol_chan' and not e.flood):
            return e.vol
        if name == 'vol_chan':
            import torch
            return torch.minimum(e.vol, e.own(e.vol_bankfull))      # :2149
        if name in ('d', 'u'):
            return e.own(getattr(e, name))
        if name == 'R':
            if getattr(e, 'R', None) is None:
                raise AttributeError('R grid is not materialised; set WRITE_R = True '
                                     'before initialize()')
            return e.own(e.R)
        if name == 'S_free':
            if e.kinematic:
                return e.own(e.S_bed)
            self._need_diag()
            return e.own(e.S_free)
        if name in _DIAG:
            self._need_diag()
            return e.own(getattr(e, name))
        raise AttributeError(name)

    def _need_diag(self):
        if not self.engine.diag:
            raise AttributeError(
                'diagnostic grids (Rh, A_wet, P_wet, f, tau, u_star, froude, S_free) are '
                'only materialised when DIAGNOSTICS is True before initialize() or one of '
                'them was requested through get_value_ptr() before the first update')

    def _fetch(self, name):
        """numpy mirror of a device grid, refreshed now (same array object every
        time, so references stay valid)."""
        t = self._device_tensor(name)
        m = self._mirrors.get(name)
        if m is None or m.shape != tuple(t.shape):
            m = np.empty(tuple(t.shape), dtype='float64')
            self._mirrors[name] = m
        import torch
        torch.from_numpy(m).copy_(t)
        return m

    def gather_grid(self, name, dst=0):
        """Slab mode: the GLOBAL (ny, nx) numpy grid of a device variable on rank
        `dst` (None elsewhere); collective.  Single GPU: the grid itself."""
        import torch
        import torch.distributed as dist
        t = self._device_tensor(name).contiguous()
        if self.layout is None or self.layout.world_size == 1:
            return t.cpu().numpy()
        lay = self.layout
        parts = ([torch.empty((n, lay.nx), dtype=t.dtype, device=t.device) for _, n in lay.parts]
                 if lay.rank == dst else None)
        dist.gather(t, parts, dst=dst, group=self.group)
        if lay.rank != dst:
            return None
        return torch.cat(parts, 0).cpu().numpy()

    def get_value_tensor(self, long_var_name):
        """Device tensor behind a grid variable (zero copy; GPU-aware users)."""
        return self._device_tensor(self.get_var_name(long_var_name))

    def get_value_ptr(self, long_var_name):
        name = self.get_var_name(long_var_name)
        if name == 'R' and 'R' in self.__dict__:
            # uniform forcing: R is the 0-D array of update_R (:1649), kept on the host.
            # Should a provider switch to grids later, the R grid is materialised then.
            self._R_requested = True
            return self.__dict__['R']
        if name in _DEVICE_GRIDS and self.engine is not None:
            if name in _DIAG or (name == 'S_free' and not self.engine.kinematic):
                if not self.engine.diag and self.engine.n_steps == 0:
                    self._enable_diag()
            self._exported.add(name)
            return self._fetch(name)
        return getattr(self, name)

    # ---------------------------------------------------------------- constants
    def set_constants(self):
        self.g = np.float64(9.81)
        self.aval = np.float64(0.476)
        self.kappa = np.float64(0.408)
        self.law_const = np.sqrt(self.g) / self.kappa
        self.deg_to_rad = np.pi / np.float64(180)
        self.rad_to_deg = np.float64(180) / np.pi
        self.mmph_to_mps = 1.0 / (3600.0 * 1000.0)

    def set_missing_cfg_options(self):
        dflt = dict(CHECK_STABILITY=True, FLOOD_OPTION=False, ATTENUATE=False,
                    d_bankfull_type='Scalar', CREATE_CHANNEL_FILES=False,
                    slope_factor=1.0, nval_factor=1.0, z0val_factor=1.0, width_factor=1.0,
                    angle_factor=1.0, sinu_factor=1.0, d0_factor=1.0, d_bankfull_factor=1.0,
                    MANNING=1, LAW_OF_WALL=0, WRITE_R=False)
        for k, v in dflt.items():
            if not hasattr(self, k):
                setattr(self, k, v)
        if not hasattr(self, 'd_bankfull'):
            self.d_bankfull = self.initialize_scalar(10.0)
            self.d_bankfull_file = ''
        for k in ('Q', 'U', 'D', 'F'):
            for kind in ('GRIDS', 'PIXELS'):
                if not hasattr(self, 'SAVE_%s_%s' % (k, kind)):
                    setattr(self, 'SAVE_%s_%s' % (k, kind), False)
        for k in ('save_grid_dt', 'save_pixels_dt'):
            if not hasattr(self, k):
                setattr(self, k, self.initialize_scalar(60.0))
        self.MANNING = bool(int(self.MANNING))
        self.LAW_OF_WALL = bool(int(self.LAW_OF_WALL))
        self.FLOOD_OPTION = bool(int(self.FLOOD_OPTION))
        self.ATTENUATE = bool(int(self.ATTENUATE))
        if self.ATTENUATE and not hasattr(self, 'n_days'):
            self.n_days = 5                                          # :566-568
        if self.CREATE_CHANNEL_FILES not in (False, 0):
            raise NotImplementedError('CREATE_CHANNEL_FILES is input preparation '
                                      '(SURVEY.md 8f-3), not part of the GPU path')

    def set_computed_input_vars(self):
        ext = self.get_attribute('cfg_extension').lower()
        self.KINEMATIC_WAVE = 'kinematic' in ext
        self.DIFFUSIVE_WAVE = 'diffusive' in ext
        self.DYNAMIC_WAVE = 'dynamic' in ext
        if self.DYNAMIC_WAVE:
            raise NotImplementedError('dynamic wave is broken in the reference '
                                      '(channels_dynamic_wave.py:164) and not provided')
        self.code_type = 'Grid'
        self.slope_type = 'Grid'
        np.maximum(self.save_grid_dt, self.dt, out=self.save_grid_dt)
        np.maximum(self.save_pixels_dt, self.dt, out=self.save_pixels_dt)

    # ---------------------------------------------------------------- initialize
    def initialize(self, cfg_file=None, mode='nondriver', SILENT=False):
        self.SILENT = SILENT
        if not SILENT:
            print(' ')
            print('Channels component (B200): Initializing...')
        self.status = 'initializing'
        self.mode, self.cfg_file = mode, cfg_file
        L.load()                              # fail loudly if the CUDA library is absent
        import torch
        if not torch.cuda.is_available():
            raise L.TfbError('the GPU Channels component needs a CUDA device; there is '
                             'no CPU fallback')
        self.set_constants()
        self.initialize_config_vars()
        self.set_missing_cfg_options()
        self.initialize_basin_vars()
        self.initialize_time_vars()
        self._init_host_scalars()
        if self.comp_status.lower() == 'disabled':
            if not SILENT:
                print('Channels component: Disabled in CFG file.')
            self.DONE = True
            self.status = 'initialized'
            return
        self.set_computed_input_vars()
        self.initialize_slab_layout()
        self.initialize_d8_vars()
        self.read_input_files()
        self.initialize_computed_vars()
        self.open_output_files()
        self.status = 'initialized'

    def _init_host_scalars(self):
        z = lambda v=0.0: self.initialize_scalar(v, dtype='float64')
        # the 13 values every step refreshes are 0-D VIEWS of one float64 block (still
        # "mutable scalars" shared by reference, BMI_base.py:2230-2294): one vector copy
        # from the pinned mirror updates them all
        self._scalar_block = np.zeros(len(_SCALAR_BLOCK), dtype='float64')
        for i, n in enumerate(_SCALAR_BLOCK):
            setattr(self, n, self._scalar_block[i:i + 1].reshape(()))
        for n in ('vol_chan_sum', 'vol_chan_sum0', 'vol_flood_sum', 'Q_canals_in'):
            setattr(self, n, z())
        for n in ('Q_min', 'u_min', 'd_min'):
            setattr(self, n, z(1e6))
        for n in ('Q_max', 'u_max', 'd_max'):
            setattr(self, n, z(-1e6))
        for n in _SRC_INPUTS:
            if not hasattr(self, n):
                setattr(self, n, z())
        if not hasattr(self, 'rho_H2O'):
            self.rho_H2O = z(1000.0)
        self.dt_cfl_min = z()
        self.bad_counts = dict(n_neg_d=0, n_nan_d=0, n_inf_d=0, n_neg_u=0, n_nan_u=0,
                               n_inf_u=0)

    def initialize_slab_layout(self):
        """Row-slab sharding when the component runs under torch.distributed with
        more than one rank (one process per GPU): this rank owns rows [row0, row0+ny)
        of the global grid; `self.ny`, `self.nx` keep the GLOBAL shape (BMI grid
        info), `self.layout` describes the slab."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(self.group)
        if world == 1:
            return
        from .slab import SlabLayout
        halo = 1 if self.KINEMATIC_WAVE else 2
        self.layout = SlabLayout(self.ny, self.nx, dist.get_rank(self.group), world, halo)
        if np.ndim(self.da) != 0:
            raise NotImplementedError('slab mode needs fixed-length pixels (scalar da)')

    def _rows(self):
        """(row_lo, row_hi) of the rows this rank keeps of a per-cell input: its slab
        plus halo (the whole grid on a single GPU)."""
        if self.layout is None:
            return 0, self.ny
        lay = self.layout
        return lay.row0 - lay.halo, lay.row0 + lay.ny + lay.halo

    def initialize_d8_vars(self):
        """channels_base.py:1074-1135: an embedded D8 component on the same DEM."""
        from . import d8_global
        self.d8 = d8_global.d8_component()
        self.d8.device = self.device
        self.d8.layout, self.d8.group = self.layout, self.group
        cfg = self.cfg_directory + self.case_prefix + '_d8_global.cfg'
        self.d8.initialize(cfg_file=cfg, SILENT=self.SILENT, REPORT=self.REPORT)
        self.d8.update(self.time, REPORT=self.REPORT)

    def _read_var(self, name, factor):
        vtype = getattr(self, name + '_type', 'Scalar').lower()
        if vtype == 'scalar':
            v = getattr(self, name)
            v = np.float64(v) if np.ndim(v) == 0 else v
            if factor != 1.0:
                v = v * factor
            return np.array(v, dtype='float64')
        if vtype not in ('grid', 'grid_sequence'):
            raise RuntimeError('No match found for ' + vtype + '.')
        path = self.topo_directory + getattr(self, name + '_file')
        lo, hi = self._rows()
        g = np.float64(grid_io.read_rtg_rows(path, self.rti, lo, hi, 'FLOAT'))   # model_input.py:181
        if factor != 1.0:
            g *= factor
        return g

    def read_input_files(self):
        """channels_base.py:3498-3558 (float32 on disk -> float64)."""
        self.slope = self._read_var('slope', self.slope_factor)
        if self.MANNING:
            self.nval = self._read_var('nval', self.nval_factor)
        if self.LAW_OF_WALL:
            self.z0val = self._read_var('z0val', self.z0val_factor)
        self.width = self._read_var('width', self.width_factor)
        self.angle = self._read_var('angle', self.angle_factor)
        self.sinu = self._read_var('sinu', self.sinu_factor)
        self.d0 = self._read_var('d0', self.d0_factor)
        if getattr(self, 'd_bankfull_type', 'Scalar').lower() == 'scalar' \
                or os.path.exists(self.topo_directory + getattr(self, 'd_bankfull_file', '')):
            self.d_bankfull = self._read_var('d_bankfull', self.d_bankfull_factor)

    def initialize_computed_vars(self):
        """channels_base.py:1137-1405, with the grids created on the device."""
        import torch
        from .routing import ChannelsEngine
        z = lambda v=-1.0: self.initialize_scalar(v, dtype='float64')
        if self.MANNING:
            self.nval_min, self.nval_max = self.nval.min(), self.nval.max()
            self.z0val, self.z0val_min, self.z0val_max = z(), z(), z()
        elif self.LAW_OF_WALL:
            self.z0val_min, self.z0val_max = self.z0val.min(), self.z0val.max()
            self.nval, self.nval_min, self.nval_max = z(), z(), z()
        else:
            raise ValueError('In CFG file, MANNING=0 and LAW_OF_WALL=0.')
        self.angle = self.angle * self.deg_to_rad                      # :1196
        lay = self.layout
        if lay is None:
            tk = self.d8.toolkit
            code, dxh, dyh, ddh = tk.code, tk.dx_host, tk.dy_host, tk.dd_host
            dw = lambda: tk.dw.cpu().numpy()
            ny_loc, row0, halo = self.ny, 0, 0
        else:
            sd = self.d8.slab
            code, dxh, dyh, ddh = sd['code'], sd['dx'], sd['dy'], sd['dd']
            ny_loc, row0, halo = lay.ny, lay.row0, lay.halo

            def dw():
                t = torch.empty(code.shape, dtype=torch.float32, device=code.device)
                n = code.shape[0]
                r = [torch.full((n,), float(v[0]), dtype=torch.float32, device=code.device)
                     for v in (dxh, dyh, ddh)]
                L.check(L.load().tfb_d8_ds_dw(code.data_ptr(), r[0].data_ptr(), r[1].data_ptr(),
                                              r[2].data_ptr(), n, self.nx, None, t.data_ptr(),
                                              L.stream_ptr()), 'tfb_d8_ds_dw')
                return t.cpu().numpy()
        if np.ndim(self.width) == 2:
            w0 = (self.width == 0)                                     # :1206-1209
            if w0.any():
                self.width[w0] = dw()[w0]
        self.slope = self.slope / self.sinu                            # :1247
        if self.KINEMATIC_WAVE:
            self.remove_bad_slopes()                                   # :1316
        self.S_bed = self.slope
        rough = self.nval if self.MANNING else self.z0val
        peer = None
        if lay is not None and self.P2P_HALO:
            from .slab import make_peer_halos
            peer = make_peer_halos(lay, self.device or torch.device('cuda', torch.cuda.current_device()),
                                   self.group, with_depth=not self.KINEMATIC_WAVE)
        self.engine = ChannelsEngine(
            ny_loc, self.nx, dt=float(self.dt), code=code, dx=dxh, dy=dyh,
            dd=ddh, S_bed=self.S_bed, width=self.width, angle=self.angle,
            tan_angle=np.tan(self.angle), cos_angle=np.cos(self.angle), rough=rough,
            sinu=self.sinu if np.ndim(self.sinu) == 2 else float(self.sinu),
            da=self.da, d0=self.d0 if np.ndim(self.d0) == 2 else float(self.d0),
            kinematic=self.KINEMATIC_WAVE, manning=self.MANNING,
            diag=bool(self.DIAGNOSTICS or self.SAVE_F_GRIDS or self.SAVE_F_PIXELS),
            write_R=bool(self.WRITE_R), outlet=(int(self.outlet_ID[0]), int(self.outlet_ID[1])),
            rho_H2O=float(self.rho_H2O), device=self.device, row0=row0, ny_global=self.ny,
            halo=halo, finalize=True, n_pixels_global=ny_loc * self.nx,
            q_buffers=None if peer is None else peer.q_buffers,
            d_buffer=None if peer is None else peer.d_buffer,
            flood=self.FLOOD_OPTION, attenuate=self.ATTENUATE,
            n_days=float(getattr(self, 'n_days', 5.0)),
            d_bankfull=self.d_bankfull if np.ndim(self.d_bankfull) == 2 else float(self.d_bankfull))
        if self._fused:
            self.engine.enable_fused_sources(**self._fused)
        if lay is None:
            self.stepper = self.engine
        else:
            from .slab import SlabChannels
            self.stepper = SlabChannels(lay, self.engine, self.group, peer=peer)
        if getattr(self.engine, 'R', None) is None:
            # R as the reference has it with uniform forcing: a 0-D array (:1649); a
            # device grid takes its place when an input arrives as a grid (_update_R)
            self.__dict__['R'] = self.initialize_scalar(0.0, dtype='float64')
        self._R_requested = False
        self.update_total_channel_water_volume()
        self.vol_chan_sum0 = self.vol_chan_sum.copy()
        torch.cuda.current_stream().synchronize()

    def _enable_diag(self):
        """Switch the engine to the diagnostics kernel before the first step."""
        import torch
        e = self.engine
        if e.diag:
            return
        z = lambda: e._zeros(torch.float64)
        for n in _DIAG:
            t = z()
            setattr(e, n, t)
            setattr(e.args, n, e._p(t))
        if not e.kinematic:
            e.S_free = z()
            e.own(e.S_free).copy_(e.own(e.S_bed))
            e.args.S_free = e._p(e.S_free)
        e.diag = True
        e.flags |= L.CH_DIAG
        # initial A_wet / P_wet / Rh of the initial depth (channels_base :1331-1333)
        d = e.own(e.d)
        g = lambda name: (e.own(getattr(e, name + '_grid')) if getattr(e, name + '_grid')
                          is not None else getattr(e.args, name).val)
        e.own(e.A_wet).copy_(d * (g('width') + d * g('tan_angle')))
        e.own(e.P_wet).copy_(g('width') + (2.0 * d) / g('cos_angle'))

    def remove_bad_slopes(self):
        """NaN / zero / negative / infinite slopes -> smallest good slope
        (channels_base.py:4164-4260)."""
        S = self.slope
        bad = np.isnan(S) | (S == 0.0) | (S < 0.0) | np.isinf(S)
        if not bad.any() and self.layout is None:
            return
        good = ~bad
        S_min = S[good].min() if good.any() else np.inf
        if self.layout is not None:
            import torch
            import torch.distributed as dist
            t = torch.tensor([float(S_min)], dtype=torch.float64, device=self.device or 'cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
            S_min = float(t.item())
        if not np.isfinite(S_min):
            raise ValueError('remove_bad_slopes: no positive finite slope in the grid')
        S[bad] = S_min
        if not self.SILENT:
            print('WARNING: %d zero, negative or NaN slopes replaced with min(S) = %s'
                  % (int(bad.sum()), S_min))
        self.slope = np.float64(S)

    def enable_fused_sources(self, **kw):
        """Evaluate met / snow / evap / infil inside the routing kernel (north-star
        fusion) instead of receiving their outputs.  Call before initialize() or
        before the first update(); keyword arguments as
        ChannelsEngine.enable_fused_sources; parameters are then bound with
        set_source_input()."""
        self._fused = dict(kw)
        if self.engine is not None:
            self.engine.enable_fused_sources(**self._fused)

    def set_source_input(self, name, value):
        """Bind a fused-source parameter/forcing (T_air, c0, T0, T_surf, Qn_SW,
        Qn_LW, Ks, Ki, qs, qi, G, h_table, elev): scalar, numpy grid or CUDA tensor."""
        self.engine.set_input(name, value)

    # -------------------------------------------------------------------- update
    def _bind_inputs(self):
        """Hand the 6 source inputs to the engine.  Scalars are re-bound only when
        their value changed (EMELI re-sets every input every step, mostly unchanged
        0-D arrays); numpy grids are uploaded every step (the provider may have
        mutated them in place); CUDA tensors are used where they are."""
        import torch
        e = self.engine
        self._et_host = None
        last = self._last_scalar_in
        for n in _SRC_INPUTS:
            v = getattr(self, n)
            if type(v) is np.ndarray and v.ndim == 0 or isinstance(v, (float, int, np.floating)):
                f = float(v)
                if last.get(n) != f:
                    e.set_input(n, f)
                    last[n] = f
                continue
            last[n] = None
            if isinstance(v, torch.Tensor):
                e.set_input(n, v if v.dim() == 2 else float(v))
            elif np.ndim(v) == 2:
                lay = self.layout
                if lay is not None and v.shape[0] == self.ny:
                    # a GLOBAL grid: this rank takes its rows (+ halo); the in-place ET
                    # masking goes back into the owned rows of the provider's array.
                    # The masked ET is STATE (zeros persist, channels_base.py:1625-1627):
                    # after the first upload only the owned rows are refreshed from the
                    # host, the halo rows come from the neighbours (update()).
                    own = v[lay.row0:lay.row0 + lay.ny]
                    if n == 'ET' and self._et_slab_ready:
                        e.set_input(n, own)
                    else:
                        e.set_input(n, lay.take(v))
                    if n == 'ET':
                        self._et_host = own
                        self._et_slab_ready = True
                else:
                    e.set_input(n, v)
                    if n == 'ET':
                        self._et_host = v
            else:
                e.set_input(n, float(v))
        rho = float(self.rho_H2O)
        if rho != e.args.rho_H2O:
            e.args.rho_H2O = rho

    def _update_R(self):
        """Host side of update_R (:1548-1655) for uniform forcing: R = (P_rain + SM + GW +
        MR) - (ET + IN) as a 0-D array, a scalar ET being dropped (:1629).  With any grid
        input (or fused sources) R is a grid: the 0-D stand-in is retired, and if some
        user holds a reference to R the kernel materialises the grid from now on."""
        if 'R' not in self.__dict__:
            return
        e = self.engine
        if not (e.flags & 0xF00) and not any(not isinstance(e.src[n], float) for n in _SRC_INPUTS):
            v = e.src
            self.__dict__['R'].fill((((v['P_rain'] + v['SM']) + v['GW']) + v['MR']) - (0.0 + v['IN']))
            return
        del self.__dict__['R']
        if self._R_requested:
            import torch
            e.R = e._zeros(torch.float64)
            e.args.R = e._p(e.R)
            e.flags |= L.CH_WRITE_R
            self._exported.add('R')

    def update(self, dt=-1.0):
        if self.comp_status.lower() == 'disabled':
            return
        self.status = 'updating'
        if self.mode == 'driver' and not self.SILENT:
            self.print_time_and_value(self.Q_outlet, 'Q_out', '[m^3/s]')
        self._bind_inputs()
        self._update_R()
        self.stepper.step(float(self.time_min))
        if self._et_host is not None:
            # provider's ET array is masked in place (channels_base.py:1625-1627)
            import torch
            self.engine.download('ET', self.engine.own(self.engine.src['ET']), self._et_host)
            if self.layout is not None and not self.KINEMATIC_WAVE:
                from .slab import exchange_halos
                exchange_halos([self.engine.src['ET']], self.layout, [1], self.group)
        OK = True
        if self.SYNC_SCALARS or self.CHECK_STABILITY:
            self.sync_scalars()
            if self.CHECK_STABILITY:
                OK = self.check_flow_depth() and self.check_flow_velocity()
        for name in self._exported:
            self._fetch(name)
        self.write_output_files()
        self.update_time(dt)
        if OK:
            self.status = 'updated'
        else:
            self.status = 'failed'
            self.DONE = True

    def sync_scalars(self):
        """Refresh the 0-D arrays in place from the scalar block the kernel wrote
        into pinned host memory (one stream synchronise, no copy; in slab mode the
        ranks' blocks meet in shared memory, slab.ScalarBoard, or in one all-gather, and
        are reduced on the host -- collective)."""
        b = self.bad_counts
        # the 8-byte words of lib.ChannelsScalars: straight from the pinned mirror on one GPU,
        # reduced over the ranks' blocks in slab mode
        if self.layout is None:
            e = self.engine
            f, i = e.sync_scalar_mirror(), e._scalars_i64
        else:
            f, i = self.stepper.read_scalar_words()
        if i[_IX_FAIL]:
            # a bounded wait inside the kernel ran out (TMA pipeline, or a slab neighbour that
            # never delivered its halo rows): the grids of that step are not valid
            self._step_failed(int(i[_IX_FAIL]))
        self._scalar_block[:] = f[:len(_SCALAR_BLOCK)]
        self.dt_cfl_min.fill(f[_IX_DTMIN])
        (b['n_neg_d'], b['n_nan_d'], b['n_inf_d'], b['n_neg_u'], b['n_nan_u'],
         b['n_inf_u']) = i[_IX_CNT:_IX_CNT + 6].tolist()

    def _step_failed(self, n):
        self.status = 'failed'
        self.DONE = True
        raise L.TfbError('routing step failed on the device: %d launch(es) gave up waiting for '
                         'the tile pipeline or for a neighbour slab\'s halo rows' % n)

    def _abort_banner(self, lines):
        star = '*******************************************'
        print(star)
        print('ERROR: Simulation aborted.')
        print(' ')
        for ln in lines:
            print(ln)
        print('Time step may be too large for stability.')
        print('Time step:      ' + str(self.dt) + ' [s]')
        print('Try reducing timestep in channels CFG file.')
        print(star)
        print(' ')

    def check_flow_depth(self):
        b = self.bad_counts
        if b['n_neg_d'] == 0 and b['n_inf_d'] == 0:        # :3218 (NaN alone passes)
            return True
        msg = []
        if b['n_neg_d']:
            msg.append('Found %d negative depths.' % b['n_neg_d'])
        if b['n_nan_d']:
            msg.append('Found %d NaN depths.' % b['n_nan_d'])
        if b['n_inf_d']:
            msg.append('Found %d infinite depths.' % b['n_inf_d'])
        self._abort_banner(msg)
        self.status = 'failed'
        self.DONE = True
        raise RuntimeError('Negative or NaN depth found.')

    def check_flow_velocity(self):
        b = self.bad_counts
        if b['n_neg_u'] == 0 and b['n_nan_u'] == 0 and b['n_inf_u'] == 0:
            return True
        msg = []
        if b['n_neg_u']:
            msg.append('Found %d negative velocities.' % b['n_neg_u'])
        if b['n_nan_u']:
            msg.append('Found %d NaN velocities.' % b['n_nan_u'])
        if b['n_inf_u']:
            msg.append('Found %d infinite velocities.' % b['n_inf_u'])
        self._abort_banner(msg)
        self.status = 'failed'
        self.DONE = True
        raise RuntimeError('Found bad velocities.')

    # ------------------------------------------------------------------ finalize
    def update_total_channel_water_volume(self):
        self._minmax()

    def update_mins_and_maxes(self, REPORT=False):
        self._minmax()

    def _minmax(self):
        """Q/u/d interior extrema and total channel volume (channels_base.py
        :2986-3110): one reduction kernel on a single GPU; in slab mode local torch
        reductions over the owned interior cells + one all-reduce each."""
        import torch
        e = self.engine
        if self.layout is not None:
            import torch.distributed as dist
            lay = self.layout
            r0 = 1 if lay.row0 == 0 else 0
            r1 = lay.ny - 1 if lay.row0 + lay.ny == self.ny else lay.ny
            inner = lambda t: t[r0:r1, 1:-1]
            big = 1.0e308
            mn = torch.stack([inner(t).min() if r1 > r0 else torch.tensor(big, device=e.device,
                              dtype=torch.float64) for t in (e.Q, e.own(e.u), e.own(e.d))])
            mx = torch.stack([inner(t).max() if r1 > r0 else torch.tensor(-big, device=e.device,
                              dtype=torch.float64) for t in (e.Q, e.own(e.u), e.own(e.d))])
            sm = e.vol.sum().reshape(1)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=self.group)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=self.group)
            mn, mx = mn.cpu().numpy(), mx.cpu().numpy()
            if e.n_steps > 0:
                for k, n in enumerate(('Q', 'u', 'd')):
                    getattr(self, n + '_min').fill(mn[k])
                    getattr(self, n + '_max').fill(mx[k])
            self.vol_chan_sum.fill(float(sm))
            return
        out = torch.zeros(7, dtype=torch.float64, device=e.device)
        npart = e.lib.tfb_sources_partials_len()
        part = torch.zeros(npart, dtype=torch.float64, device=e.device)
        tick = torch.zeros(1, dtype=torch.int32, device=e.device)
        # dense (ny, nx) views for the reduction kernel (copies only when rows are padded)
        Q, u_, d_, vol_ = (t.contiguous() for t in (e.Q, e.own(e.u), e.own(e.d), e.vol))
        L.check(e.lib.tfb_channels_minmax(
            Q.data_ptr(), u_.data_ptr(), d_.data_ptr(), vol_.data_ptr(),
            e.ny, e.nx, out.data_ptr(), out[6:].data_ptr(), part.data_ptr(), tick.data_ptr(),
            L.stream_ptr()), 'tfb_channels_minmax')
        h = out.cpu().numpy()
        if e.n_steps > 0:
            for k, n in enumerate(('Q_min', 'Q_max', 'u_min', 'u_max', 'd_min', 'd_max')):
                getattr(self, n).fill(h[k])
        self.vol_chan_sum.fill(h[6])

    def finalize(self):
        if self.comp_status.lower() == 'disabled':
            return
        self.status = 'finalizing'
        self.sync_scalars()
        self.update_total_channel_water_volume()
        self.update_mins_and_maxes()
        if self.FLOOD_OPTION and self.layout is None:
            # update_total_channel/flood_water_volume :3083-3135
            self.vol_chan_sum.fill(float(self._device_tensor('vol_chan').sum()))
            self.vol_flood_sum.fill(float(self._device_tensor('vol_flood').sum()))
        if not self.SILENT:
            self.print_final_report()
        self.close_output_files()
        self.status = 'finalized'

    def print_final_report(self):
        print('Channels component (B200): final report')
        for n in ('Q_peak', 'T_peak', 'u_peak', 'd_peak', 'vol_Q', 'vol_R', 'vol_edge',
                  'vol_chan_sum'):
            print('   %-14s = %r' % (n, float(getattr(self, n))))

    # -------------------------------------------------------------------- output
    # Grid stacks as RTS (flat float32 frames) and outlet series as text -- the
    # two writers of the reference that need no netCDF (utils/rts_files.py,
    # utils/text_ts_files.py); frames are fetched from the device only when due.
    def open_output_files(self):
        """Grid stacks: ONE file per variable.  In slab mode rank 0 creates it (and its RTI
        side-car), every rank opens it and later writes ITS rows of each frame at their
        offset -- no gather, the ranks write in parallel.  Outlet series: rank 0."""
        self._out = {}
        self._frame_no = {}
        slab = self.layout is not None
        rank0 = not slab or self.layout.rank == 0
        self._out_any = any(getattr(self, 'SAVE_%s_%s' % (K, kind))
                            for K in 'QUDF' for kind in ('GRIDS', 'PIXELS'))
        kinds = {'Q': 'Q', 'U': 'u', 'D': 'd', 'F': 'f'}
        gs_paths = {}
        for K, v in kinds.items():
            if getattr(self, 'SAVE_%s_GRIDS' % K) and rank0:
                f = getattr(self, v + '_gs_file', self.case_prefix + '_2D-' + v + '.nc')
                path = self.out_directory + os.path.splitext(os.path.basename(f))[0] + '.rts'
                os.makedirs(self.out_directory, exist_ok=True)
                path = self._unique(path)
                self._out[('gs', v)] = open(path, 'w+b')
                gs_paths[v] = path
                # side-car header like rts_files.open_new_file (:255-268, MAKE_RTI): the
                # data set's own geometry and BYTE ORDER, FLOAT frames, min/max unknown
                grid_io.write_output_rti(path, self.rti, 'FLOAT')
        if slab and any(getattr(self, 'SAVE_%s_GRIDS' % K) for K in kinds):
            import torch.distributed as dist
            box = [gs_paths]
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            dist.broadcast_object_list(box, src=src, group=self.group)     # (also orders the create)
            if not rank0:
                for v, path in box[0].items():
                    self._out[('gs', v)] = open(path, 'r+b')
        for v in [k[1] for k in self._out if k[0] == 'gs']:
            self._frame_no[v] = 0
        if not rank0:
            return
        for K, v in kinds.items():
            if getattr(self, 'SAVE_%s_PIXELS' % K):
                f = getattr(self, v + '_ts_file', self.case_prefix + '_0D-' + v + '.txt')
                path = self.out_directory + os.path.splitext(os.path.basename(f))[0] + '.txt'
                os.makedirs(self.out_directory, exist_ok=True)
                fh = open(self._unique(path), 'w')
                rows, cols = self.outlet_IDs
                hdr = ['Time [min]'] + ['%s_%d_%d' % (v, r, c) for r, c in zip(rows, cols)]
                fh.write(''.join(h.rjust(15) for h in hdr) + '\n')
                fh.write('-' * (15 * len(hdr)) + '\n')
                self._out[('ts', v)] = fh

    def _unique(self, path):
        if self.OVERWRITE_OK or not os.path.exists(path):
            return path
        stem, ext = os.path.splitext(path)
        k = 1
        while os.path.exists('%s_%d%s' % (stem, k, ext)):
            k += 1
        return '%s_%d%s' % (stem, k, ext)

    def write_output_files(self, time_seconds=None):
        """Save grids / outlet series when due (channels_base.py:3848-3958).  Single
        GPU: the frame is snapshotted on the device (cast to float32, ~50 us for 4096^2),
        copied to pinned host memory on a side stream and written by a background
        thread -- the step loop never waits for PCIe or the disk.  Outlet series are a
        device-side gather of the monitored cells.  Slab mode: every rank writes its own
        rows of a frame into the shared file the same way (positional writes, no gather);
        the monitored cells are summed over the ranks (each lives on one)."""
        if not getattr(self, '_out', None) and not (self.layout is not None and
                                                    getattr(self, '_out_any', False)):
            return
        t = int(self.time_sec if time_seconds is None else time_seconds)
        slab = self.layout is not None
        names = {'Q': 'SAVE_Q', 'u': 'SAVE_U', 'd': 'SAVE_D', 'f': 'SAVE_F'}
        for kind, dt_save in (('gs', self.save_grid_dt), ('ts', self.save_pixels_dt)):
            if t % int(dt_save) != 0:
                continue
            for v, flag in names.items():
                if not getattr(self, flag + ('_GRIDS' if kind == 'gs' else '_PIXELS')):
                    continue
                fh = self._out.get((kind, v))
                if kind == 'gs':
                    # this rank's rows of the frame, at their place in the file
                    self._async_frame(v, fh)
                    continue
                import torch
                if getattr(self, '_outlet_idx', None) is None:
                    rows, cols = (np.asarray(x) for x in self.outlet_IDs)
                    r0 = self.layout.row0 if slab else 0
                    mine = (rows >= r0) & (rows < r0 + self.engine.ny)
                    dev = self.engine.device
                    self._outlet_idx = (torch.as_tensor(np.where(mine, rows - r0, 0), device=dev),
                                        torch.as_tensor(np.where(mine, cols, 0), device=dev),
                                        torch.as_tensor(mine, device=dev))
                ri, ci, mine = self._outlet_idx
                vals = torch.where(mine, self._device_tensor(v)[ri, ci], 0.0)
                if slab:
                    # a monitored cell lives on one rank: the sum over ranks is its value
                    import torch.distributed as dist
                    dist.all_reduce(vals, op=dist.ReduceOp.SUM, group=self.group)
                if fh is not None:
                    self._write_ts_line(fh, vals.cpu().numpy())

    def _write_ts_line(self, fh, vals):
        # text_ts_files.add_values :149-161 (20-wide columns), values_at_IDs :170-198 (float32)
        fh.write(('%20.7f' % float(self.time_min))
                 + ''.join('%20.7f' % x for x in np.float32(vals)) + '\n')

    def _async_frame(self, v, fh):
        import queue
        import threading
        import torch
        w = getattr(self, '_writer', None)
        ny_own = self.engine.ny                       # all rows on one GPU, this slab's otherwise
        row0 = self.layout.row0 if self.layout is not None else 0
        if w is None:
            dev = self.engine.device
            n_slots = 3
            w = self._writer = dict(
                stream=torch.cuda.Stream(device=dev), q=queue.Queue(), free=queue.Queue(),
                stage=[torch.empty((ny_own, self.nx), dtype=torch.float32, device=dev)
                       for _ in range(n_slots)],
                host=[torch.empty((ny_own, self.nx), dtype=torch.float32).pin_memory()
                      for _ in range(n_slots)], error=[])
            for k in range(n_slots):
                w['free'].put(k)

            def run():
                while True:
                    item = w['q'].get()
                    if item is None:
                        return
                    k, ev, out, offset = item
                    try:
                        ev.synchronize()
                        buf = memoryview(w['host'][k].numpy()).cast('B')
                        done = 0
                        while done < len(buf):        # positional: ranks share the file
                            done += os.pwrite(out.fileno(), buf[done:], offset + done)
                    except Exception as exc:            # surfaced at close_output_files()
                        w['error'].append(exc)
                    w['free'].put(k)
            w['thread'] = threading.Thread(target=run, daemon=True)
            w['thread'].start()
        k = w['free'].get()                           # blocks only if 3 frames are in flight
        cur = torch.cuda.current_stream()
        w['stage'][k].copy_(self._device_tensor(v))    # device snapshot + fp32 cast, compute stream
        if self.rti.SWAP_ENDIAN:
            # frames go out in the data set's byte order (rts_files.add_grid :387-389):
            # the four bytes of every float32 are reversed on the device, so the host
            # side still writes the pinned buffer as it is
            b = w['stage'][k].view(torch.uint8).view(ny_own, self.nx, 4)
            b.copy_(b.flip(2))
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(w['stream']):
            w['stream'].wait_event(ready)
            w['host'][k].copy_(w['stage'][k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(w['stream'])
        # frame n of an (ny, nx) float32 stack (rts_files.add_grid :375-389), rows from row0
        offset = (self._frame_no[v] * self.ny + row0) * self.nx * 4
        self._frame_no[v] += 1
        w['q'].put((k, done, fh, offset))

    def close_output_files(self):
        w = getattr(self, '_writer', None)
        if w is not None:
            w['q'].put(None)                          # drain the background writer
            w['thread'].join()
            self._writer = None
            if w['error']:
                raise w['error'][0]
        for fh in getattr(self, '_out', {}).values():
            fh.close()
        self._out = {}
