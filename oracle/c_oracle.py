"""ctypes wrapper of oracle/tf_oracle.c (CPU restatement, TEST INFRASTRUCTURE).

Mirrors oracle/np_oracle.Channels closely enough that tests can drive either;
used where numpy would be too slow/large (>= 1024^2) and as bench.py's
`cpu_baseline` (kind "port").  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libtf_oracle.so')
_lib = None


class SG(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('val', C.c_double)]


vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double


class ChannelsC(C.Structure):
    _fields_ = [
        ('ny', i32), ('nx', i32), ('diffusive', i32), ('law_of_wall', i32), ('fused', i32),
        ('outlet_row', i32), ('outlet_col', i32), ('is_R_grid', i32),
        ('dt', f64), ('time_min', f64), ('rho_H2O', f64),
        ('code', vp), ('dx', vp), ('dy', vp), ('dd', vp),
        ('sinu', SG), ('da_row', SG), ('width', SG), ('angle', SG), ('rough', SG),
        ('S_bed', vp), ('vol', vp), ('Q_in', vp), ('Q_out', vp), ('d', vp), ('u', vp),
        ('d_pre', vp),
        ('Rh', vp), ('A_wet', vp), ('P_wet', vp), ('f', vp), ('tau', vp), ('u_star', vp),
        ('froude', vp), ('S_free', vp), ('R', vp),
        ('P_rain', SG), ('SM', SG), ('GW', SG), ('MR', SG), ('IN', SG), ('ET', SG), ('ET_rw', vp),
        ('T_air', SG), ('T_rain_snow', f64),
        ('c0', SG), ('T0', SG), ('h_swe', vp), ('h_snow', vp), ('rho_snow', f64),
        ('T_surf', SG), ('Qn_SW', SG), ('Qn_LW', SG),
        ('alpha', f64), ('K_soil', f64), ('soil_x', f64), ('T_soil_x', f64),
        ('Ks', SG), ('Ki', SG), ('qs', SG), ('qi', SG), ('G', SG), ('I', vp),
    ] + [(n, f64) for n in (
        'vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
        'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak',
        'vol_P', 'vol_PR', 'vol_PS', 'vol_SM', 'vol_ET', 'vol_IN')] + [
        (n, C.c_int64) for n in ('n_neg_d', 'n_nan_d', 'n_inf_d', 'n_neg_u', 'n_nan_u', 'n_inf_u')]


def build(force=False):
    src = os.path.join(_HERE, 'tf_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE])
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(_SO)
        lib.tfo_sizeof_channels.restype = C.c_int64
        assert lib.tfo_sizeof_channels() == C.sizeof(ChannelsC), 'tfo_channels layout mismatch'
        lib.tfo_channels_step.argtypes = [C.POINTER(ChannelsC), C.c_int]
        lib.tfo_d8_flow_grid.argtypes = [vp, i32, i32, C.c_float, C.c_float, C.c_float, f64, i32,
                                         vp, vp, C.POINTER(C.c_int)]
        lib.tfo_d8_area.argtypes = [vp, i32, i32, f64, vp, vp, vp]
        lib.tfo_fill_pits.argtypes = [vp, i32, i32, C.c_float, C.c_float, i32]
        lib.tfo_fill_pits.restype = C.c_int64
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def d8_flow_grid(DEM, dx, dy, dd, nodata=-9999.0, link_flats=True):
    from . import np_oracle
    DEM = np.ascontiguousarray(DEM, dtype=np.float32)
    ny, nx = DEM.shape
    res = np_oracle.resolve_table()
    code = np.zeros((ny, nx), dtype=np.uint8)
    n = C.c_int(0)
    rc = load().tfo_d8_flow_grid(_p(DEM), ny, nx, float(dx[0]), float(dy[0]), float(dd[0]),
                                 float(nodata), int(link_flats), _p(res), _p(code), C.byref(n))
    assert rc == 0
    return code, n.value


def fill_pits(DEM, nodata=-9999.0, closed_basin_code=-32000.0, close_nodata=False):
    """In place on a C-contiguous float32 array; returns the number of raised cells."""
    assert DEM.dtype == np.float32 and DEM.flags.c_contiguous and DEM.ndim == 2
    n = load().tfo_fill_pits(_p(DEM), DEM.shape[0], DEM.shape[1], nodata, closed_basin_code,
                             int(bool(close_nodata)))
    if n < 0:
        raise MemoryError('tfo_fill_pits')
    return int(n)


def d8_area(code, pixel_area, with_counts=False):
    code = np.ascontiguousarray(code, dtype=np.uint8)
    ny, nx = code.shape
    A = np.zeros((ny, nx), dtype=np.float32)
    N = np.zeros((ny, nx), dtype=np.int64) if with_counts else None
    pa = np.asarray(pixel_area, dtype=np.float64)
    rows = None if pa.ndim == 0 else np.ascontiguousarray(pa[:, 0] if pa.ndim == 2 else pa)
    rc = load().tfo_d8_area(_p(code), ny, nx, float(pa) if pa.ndim == 0 else 0.0, _p(rows),
                            _p(A), _p(N))
    assert rc == 0
    return (A, N) if with_counts else A


class Channels(object):
    """Same constructor conventions as np_oracle.Channels (angle in DEGREES)."""

    GRIDS = ('Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')

    def __init__(self, code, dx, dy, dd, slope, width, angle_deg, rough, sinu=1.0, da=900.0,
                 dt=6.0, kinematic=True, manning=True, outlet=(0, 0), rho_H2O=1000.0,
                 diag=True, d0=0.0):
        from . import np_oracle as o
        self.lib = load()
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.code = np.ascontiguousarray(code, dtype=np.uint8)
        ny, nx = self.code.shape
        self.ny, self.nx = ny, nx
        c = self.c = ChannelsC()
        c.ny, c.nx = ny, nx
        c.diffusive, c.law_of_wall = int(not kinematic), int(not manning)
        c.outlet_row, c.outlet_col = int(outlet[0]), int(outlet[1])
        c.dt, c.rho_H2O = float(dt), float(rho_H2O)
        self._keep = {}

        def sg(name, v):
            if np.ndim(v) == 2:
                self._keep[name] = f(v)
                return SG(self._keep[name].ctypes.data, 0.0)
            return SG(None, float(v))
        self.dx, self.dy, self.dd = (np.ascontiguousarray(x, dtype=np.float32) for x in (dx, dy, dd))
        c.code, c.dx, c.dy, c.dd = _p(self.code), _p(self.dx), _p(self.dy), _p(self.dd)
        ds32, dw32 = o.d8_ds_dw(self.code, self.dx, self.dy, self.dd)
        width = np.array(width, dtype=np.float64)
        if width.ndim == 2:
            w0 = width == 0
            width[w0] = dw32[w0]
        angle = np.array(angle_deg, dtype=np.float64) * o.DEG_TO_RAD
        S = f(np.array(slope, dtype=np.float64) / sinu)
        if kinematic:
            o.remove_bad_slopes(S)
        self.S_bed = S
        c.S_bed = _p(S)
        c.sinu, c.width, c.angle, c.rough = (sg('sinu', sinu), sg('width', width),
                                             sg('angle', angle), sg('rough', rough))
        if np.ndim(da) >= 1:
            self._keep['da'] = f(da[:, 0] if np.ndim(da) == 2 else da)
            c.da_row = SG(self._keep['da'].ctypes.data, 0.0)
        else:
            c.da_row = SG(None, float(da))
        z = lambda: np.zeros((ny, nx), dtype=np.float64)
        self.vol, self.d, self.u, self.d_pre = z(), z(), z(), z()
        self.Qbuf = [z(), z()]
        self.qcur = 0
        c.vol, c.d, c.u, c.d_pre = _p(self.vol), _p(self.d), _p(self.u), _p(self.d_pre)
        if diag:
            for n in self.GRIDS:
                setattr(self, n, z())
                setattr(c, n, _p(getattr(self, n)))
            if not kinematic:
                self.S_free = S.copy()
                c.S_free = _p(self.S_free)
        if np.any(np.asarray(d0) != 0):
            self.d += d0
            ds = np.asarray(sinu) * ds32
            self.vol[:] = (self.d * (width + self.d * np.tan(angle))) * ds
        self.src = {}
        for n in ('P_rain', 'SM', 'GW', 'MR', 'IN', 'ET', 'T_air', 'c0', 'T0', 'T_surf',
                  'Qn_SW', 'Qn_LW', 'Ks', 'Ki', 'qs', 'qi', 'G'):
            self.set_input(n, 0.0)
        self.n_steps = 0
        self.time_min = 0.0
        self.nthreads = 1

    def set_input(self, name, v):
        if np.ndim(v) == 2:
            a = np.ascontiguousarray(v, dtype=np.float64)
            self.src[name] = a
            setattr(self.c, name, SG(a.ctypes.data, 0.0))
            if name == 'ET':
                self.c.ET_rw = a.ctypes.data
        else:
            self.src[name] = float(v)
            setattr(self.c, name, SG(None, float(v)))
            if name == 'ET':
                self.c.ET_rw = None

    def enable_fused(self, met=False, snow=False, evap=False, infil=False, T_rain_snow=0.0,
                     rho_snow=300.0, alpha=1.26, K_soil=0.45, soil_x=0.05, T_soil_x=0.0,
                     h0_swe=0.0, I0=1e-6):
        c = self.c
        c.fused = (1 if met else 0) | (2 if snow else 0) | (4 if evap else 0) | (8 if infil else 0)
        c.T_rain_snow, c.rho_snow = float(T_rain_snow), float(rho_snow)
        c.alpha, c.K_soil, c.soil_x, c.T_soil_x = map(float, (alpha, K_soil, soil_x, T_soil_x))
        full = lambda v: np.zeros((self.ny, self.nx)) + v
        if snow:
            self.h_swe = np.ascontiguousarray(full(h0_swe))
            self.h_snow = self.h_swe * (c.rho_H2O / c.rho_snow)
            c.h_swe, c.h_snow = _p(self.h_swe), _p(self.h_snow)
        if infil:
            self.I = np.ascontiguousarray(full(I0))
            c.I = _p(self.I)

    @property
    def Q(self):
        return self.Qbuf[self.qcur ^ 1] if self.n_steps else self.Qbuf[self.qcur]

    def step(self):
        c = self.c
        c.time_min = float(self.time_min)
        grid_R = bool(c.fused) or c.da_row.ptr is not None or any(
            isinstance(self.src[n], np.ndarray) for n in ('P_rain', 'SM', 'GW', 'MR', 'IN', 'ET'))
        c.is_R_grid = int(grid_R)
        c.Q_in, c.Q_out = _p(self.Qbuf[self.qcur]), _p(self.Qbuf[self.qcur ^ 1])
        assert self.lib.tfo_channels_step(C.byref(c), int(self.nthreads)) == 0
        self.qcur ^= 1
        self.n_steps += 1
        self.time_min += c.dt / 60.0

    def __getattr__(self, name):
        c = self.__dict__.get('c')
        if c is not None and name in dict(ChannelsC._fields_) and not name.startswith('_'):
            return getattr(c, name)
        raise AttributeError(name)
