"""Device-resident source-term steppers: thin drivers of the stand-alone
``tfb_met_precip`` / ``tfb_infil_green_ampt`` / ``tfb_snow_degree_day`` /
``tfb_evap_priestley_taylor`` kernels (one per reference component ``update()``),
for couplings where the source components keep their own time step instead of
being fused into the routing kernel (``ChannelsEngine.enable_fused_sources``).

Reference formulas: met_base.py:1300-1388 (+ integrals :1247-1446);
infil_green_ampt.py:511-578 with infil_base.py:654-811, :954-995;
snow_degree_day.py:293-360 with snow_base.py:490-703;
evap_priestley_taylor.py:308-413 with evap_base.py:326-377.
Outputs are CUDA float64 tensors that can be handed to the Channels component
through ``set_value`` without leaving the device.
"""
import numpy as np
import torch

from . import lib as L


class _Base(object):
    def __init__(self, ny, nx, dt, da=900.0, device=None):
        if not torch.cuda.is_available():
            raise L.TfbError('source kernels need a CUDA device (no CPU path)')
        self.lib = L.load()
        self.device = torch.device(device if device is not None
                                   else 'cuda:%d' % torch.cuda.current_device())
        self.ny, self.nx, self.dt = int(ny), int(nx), float(dt)
        self._keep = {}
        self.da = self._sg('da', da if np.ndim(da) == 0 else np.asarray(da)[:, 0])
        n = self.lib.tfb_sources_partials_len()
        self._partials = torch.zeros(n, dtype=torch.float64, device=self.device)
        self._ticket = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.sums = torch.zeros(4, dtype=torch.float64, device=self.device)

    def _zeros(self):
        return torch.zeros((self.ny, self.nx), dtype=torch.float64, device=self.device)

    def _sg(self, name, v):
        """scalar-or-grid binding; grids are uploaded once and kept alive."""
        if isinstance(v, torch.Tensor) and v.dim() >= 1:
            t = v.to(device=self.device, dtype=torch.float64).contiguous()
        elif np.ndim(v) >= 1:
            t = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)).to(self.device)
        else:
            self._keep.pop(name, None)
            return L.SG(None, float(v))
        self._keep[name] = t
        return L.SG(t.data_ptr(), 0.0)

    def _tail(self):
        return (self._partials.data_ptr(), self._ticket.data_ptr(), L.stream_ptr())


class MetPrecip(_Base):
    """P -> P_rain, P_snow and the integrals vol_P, vol_PR, vol_PS."""

    def __init__(self, ny, nx, dt, T_rain_snow=0.0, **kw):
        _Base.__init__(self, ny, nx, dt, **kw)
        self.T_rain_snow = float(T_rain_snow)
        self.P_rain, self.P_snow = self._zeros(), self._zeros()

    def update(self, P, T_air):
        L.check(self.lib.tfb_met_precip(self._sg('P', P), self._sg('T_air', T_air),
                                        self.T_rain_snow, self.da, self.dt, self.ny, self.nx,
                                        self.P_rain.data_ptr(), self.P_snow.data_ptr(),
                                        self.sums.data_ptr(), *self._tail()), 'tfb_met_precip')

    vol_P = property(lambda s: float(s.sums[0]))
    vol_PR = property(lambda s: float(s.sums[1]))
    vol_PS = property(lambda s: float(s.sums[2]))


class GreenAmpt(_Base):
    """Green-Ampt v1 with state I (cumulative infiltrated depth)."""

    def __init__(self, ny, nx, dt, Ks, Ki, qs, qi, G, I0=1e-6, h_table=None, elev=None, **kw):
        _Base.__init__(self, ny, nx, dt, **kw)
        self.par = [self._sg(n, v) for n, v in (('Ks', Ks), ('Ki', Ki), ('qs', qs), ('qi', qi),
                                                ('G', G))]
        self.use_table = int(h_table is not None and elev is not None)
        self.h_table = self._sg('h_table', 0.0 if h_table is None else h_table)
        self.elev = self._sg('elev', 1.0 if elev is None else elev)
        self.IN = self._zeros()
        self.I = self._zeros()
        self.I += torch.as_tensor(I0, dtype=torch.float64, device=self.device)

    def update(self, P_rain, SM=0.0):
        L.check(self.lib.tfb_infil_green_ampt(
            self._sg('P_rain', P_rain), self._sg('SM', SM), *self.par, self.h_table, self.elev,
            self.use_table, self.da, self.dt, self.ny, self.nx, self.IN.data_ptr(),
            self.I.data_ptr(), self.sums.data_ptr(), *self._tail()), 'tfb_infil_green_ampt')

    vol_IN = property(lambda s: float(s.sums[0]))
    n_bad = property(lambda s: int(s.sums[1]))


class DegreeDay(_Base):
    """Degree-day snowmelt with state h_swe / h_snow."""

    def __init__(self, ny, nx, dt, c0, T0, h0_swe=0.0, rho_H2O=1000.0, rho_snow=300.0, **kw):
        _Base.__init__(self, ny, nx, dt, **kw)
        self.c0, self.T0 = self._sg('c0', c0), self._sg('T0', T0)
        self.rho_H2O, self.rho_snow = float(rho_H2O), float(rho_snow)
        self.SM, self.h_swe, self.h_snow = self._zeros(), self._zeros(), self._zeros()
        self.h_swe += torch.as_tensor(h0_swe, dtype=torch.float64, device=self.device)
        self.h_snow.copy_(self.h_swe * (self.rho_H2O / self.rho_snow))

    def update(self, P_snow, T_air):
        L.check(self.lib.tfb_snow_degree_day(
            self._sg('P_snow', P_snow), self._sg('T_air', T_air), self.c0, self.T0, self.rho_H2O,
            self.rho_snow, self.da, self.dt, self.ny, self.nx, self.SM.data_ptr(),
            self.h_swe.data_ptr(), self.h_snow.data_ptr(), self.sums.data_ptr(), *self._tail()),
            'tfb_snow_degree_day')

    vol_SM = property(lambda s: float(s.sums[0]))


class PriestleyTaylor(_Base):
    """Qc and ET by the Priestley-Taylor method."""

    def __init__(self, ny, nx, dt, alpha=1.26, K_soil=0.45, soil_x=0.05, T_soil_x=0.0, **kw):
        _Base.__init__(self, ny, nx, dt, **kw)
        self.alpha, self.K_soil = float(alpha), float(K_soil)
        self.soil_x, self.T_soil_x = float(soil_x), float(T_soil_x)
        self.Qc, self.ET = self._zeros(), self._zeros()

    def update(self, T_air, T_surf, Qn_SW, Qn_LW):
        L.check(self.lib.tfb_evap_priestley_taylor(
            self._sg('T_air', T_air), self._sg('T_surf', T_surf), self._sg('Qn_SW', Qn_SW),
            self._sg('Qn_LW', Qn_LW), self.alpha, self.K_soil, self.soil_x, self.T_soil_x, self.da,
            self.dt, self.ny, self.nx, self.Qc.data_ptr(), self.ET.data_ptr(),
            self.sums.data_ptr(), *self._tail()), 'tfb_evap_priestley_taylor')

    vol_ET = property(lambda s: float(s.sums[0]))


class MetEnergy(_Base):
    """The meteorology energy chain of met_base.update() (:790-809, PRECIP_ONLY = No)
    in one point-wise kernel: e_sat/e (air and surface), T_dew, T_surf, Ri, Dn, Dh, Qh,
    W_p, Qe, clear-sky Qn_SW on sloping cells (solar_funcs.Clear_Sky_Radiation), em_air,
    Qn_LW, Q_sum.  Inputs are scalars or (ny, nx) grids; `julian_day` and `TSN_offset`
    come from solar_host.julian_day / tsn_offset (the date arithmetic stays on the
    host).  Albedo is an input (see Albedo below)."""

    def __init__(self, ny, nx, satterlund=False, dust_atten=0.08, **kw):
        _Base.__init__(self, ny, nx, 1.0, **kw)
        self.satterlund, self.dust_atten = bool(satterlund), float(dust_atten)
        for n in L.MetEnergyArgs._OUT:
            setattr(self, n, self._zeros())

    def update(self, julian_day, **inputs):
        a = L.MetEnergyArgs()
        a.ny, a.nx, a.satterlund = self.ny, self.nx, int(self.satterlund)
        a.julian_day, a.dust_atten = float(julian_day), self.dust_atten
        for n in L.MetEnergyArgs._IN:
            setattr(a, n, self._sg(n, inputs.get(n, 0.0)))
        for n in L.MetEnergyArgs._OUT:
            setattr(a, n, getattr(self, n).data_ptr())
        import ctypes as C
        L.check(self.lib.tfb_met_energy(C.byref(a), L.stream_ptr()), 'tfb_met_energy')


class Albedo(_Base):
    """met_base.update_albedo (:1995-2045): 'aging' snow albedo (Rohrer & Braun 1994) with
    the 3-day snowfall ring of met_base.py:1121-1133 kept on the device, or the 'simple'
    method.  `albedo` is updated in place every update()."""

    def __init__(self, ny, nx, dt, albedo=0.3, method='aging', rho_H2O=1000.0, rho_snow=300.0, **kw):
        _Base.__init__(self, ny, nx, dt, **kw)
        self.method = method
        self.rho_H2O, self.rho_snow = float(rho_H2O), float(rho_snow)
        self.albedo = self._zeros()
        self.albedo += torch.as_tensor(np.float64(albedo)).to(self.device)
        self.slots = int(np.int64(259200 / self.dt))
        self.head = self.slots - 1 if self.slots > 0 else 0      # roll(-1): the new value goes last
        if method == 'aging':
            if self.slots < 1:
                raise ValueError('albedo aging needs dt <= 3 days')
            self.ring = torch.zeros((self.slots, self.ny, self.nx), dtype=torch.float64,
                                    device=self.device)
            self.n = self._zeros()
        else:
            self.ring = self.n = None

    def update(self, P_snow, T_air, h_snow, h_ice):
        aging = self.method == 'aging'
        L.check(self.lib.tfb_met_albedo(
            self._sg('P_snow', P_snow), self._sg('T_air', T_air), self._sg('h_snow', h_snow),
            self._sg('h_ice', h_ice), self.ny, self.nx, self.dt, self.rho_H2O, self.rho_snow,
            int(aging), self.ring.data_ptr() if aging else None, max(self.slots, 1), self.head,
            self.n.data_ptr() if aging else None, self.albedo.data_ptr(), L.stream_ptr()),
            'tfb_met_albedo')
        if aging:
            self.head = (self.head + 1) % self.slots
