"""Stand-alone source-term kernels (C ABI) against values produced by the
REFERENCE's own component methods (tests/golden/sources.npz, made by
oracle/gen_golden.gen_sources).  Only + - * / min max are involved and the TU is
compiled without FMA contraction, so equality is exact."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_green_ampt_three_updates(sources_gold):
    from topoflow36_b200.sources import GreenAmpt
    g = sources_gold
    ny, nx = g['ga_Ks'].shape
    ga = GreenAmpt(ny, nx, 5.0, g['ga_Ks'], g['ga_Ki'], g['ga_qs'], g['ga_qi'], g['ga_G'],
                   I0=1e-6, h_table=g['ga_h_table'], elev=g['ga_elev'])
    vol = 0.0
    for it in range(3):
        ga.update(g['ga_P_rain'], g['ga_SM'])
        assert np.array_equal(ga.IN.cpu().numpy(), g['ga_IN%d' % it])
        assert np.array_equal(ga.I.cpu().numpy(), g['ga_I%d' % it])
        vol += np.sum(g['ga_IN%d' % it] * 900.0 * 5.0)
    assert abs(ga.vol_IN - vol) <= 1e-12 * vol and ga.n_bad == 0


def test_degree_day_three_updates(sources_gold):
    from topoflow36_b200.sources import DegreeDay
    g = sources_gold
    ny, nx = g['dd_T_air'].shape
    dd = DegreeDay(ny, nx, 3600.0, 2.7, -0.2, h0_swe=g['dd_h_swe_init'])
    for it in range(3):
        dd.update(g['dd_P_snow'], g['dd_T_air'])
        assert np.array_equal(dd.SM.cpu().numpy(), g['dd_SM%d' % it])
        assert np.array_equal(dd.h_swe.cpu().numpy(), g['dd_h_swe%d' % it])
        assert np.array_equal(dd.h_snow.cpu().numpy(), g['dd_h_snow%d' % it])


def test_priestley_taylor(sources_gold):
    from topoflow36_b200.sources import PriestleyTaylor
    g = sources_gold
    ny, nx = g['pt_T_air'].shape
    pt = PriestleyTaylor(ny, nx, 60.0, alpha=1.26, K_soil=0.45, soil_x=0.05, T_soil_x=12.0)
    pt.update(g['pt_T_air'], g['pt_T_surf'], g['pt_Qn_SW'], g['pt_Qn_LW'])
    assert np.array_equal(pt.Qc.cpu().numpy(), g['pt_Qc'])
    assert np.array_equal(pt.ET.cpu().numpy(), g['pt_ET'])
    want = np.sum(g['pt_ET'] * 900.0 * 60.0)
    assert abs(pt.vol_ET - want) <= 1e-12 * want


def test_met_rain_snow_split(sources_gold):
    from topoflow36_b200.sources import MetPrecip
    g = sources_gold
    ny, nx = g['met_P'].shape
    m = MetPrecip(ny, nx, 60.0, T_rain_snow=0.5)
    m.update(g['met_P'], g['met_T_air'])
    assert np.array_equal(m.P_rain.cpu().numpy(), g['met_P_rain'])
    assert np.array_equal(m.P_snow.cpu().numpy(), g['met_P_snow'])
    assert abs(m.vol_P - np.sum(g['met_P'] * 900.0 * 60.0)) <= 1e-12 * m.vol_P
    assert abs(m.vol_PR + m.vol_PS - m.vol_P) <= 1e-12 * m.vol_P
    # scalar forcing and a scalar air temperature
    m.update(2e-5, 3.0)
    assert float(m.P_rain.min()) == 2e-5 and float(m.P_snow.max()) == 0.0


@pytest.mark.parametrize('case', ['day', 'dusk', 'night'])
def test_met_energy_chain(met_gold, case):
    """tfb_met_energy against the REFERENCE's met_base / solar_funcs methods.  The
    chain is ~25 transcendental calls deep (exp, log, pow, sin, cos, tan, acos, asin,
    atan): libdevice and numpy agree to ~1-2 ulp per call, asserted rel <= 1e-11 of
    each field's range."""
    from conftest import rel_err
    from topoflow36_b200.sources import MetEnergy
    g = met_gold
    I = {k[len(case) + 4:]: g[k] for k in g if k.startswith(case + '_in_')}
    ny, nx = I['T_air'].shape
    m = MetEnergy(ny, nx, dust_atten=0.08)
    m.update(float(I['julian_day']), z=10.0,
             **{k: I[k] for k in ('T_air', 'RH', 'uz', 'z0_air', 'h_snow', 'h_ice', 'p0', 'albedo',
                                  'em_surf', 'lat_deg', 'alpha', 'beta', 'TSN_offset',
                                  'cloud_factor', 'canopy_factor')})
    for k in ('e_sat_air', 'e_air', 'T_dew', 'T_surf', 'e_sat_surf', 'Ri', 'Dn', 'Dh', 'Qh', 'W_p',
              'e_surf', 'Qe', 'Qn_SW', 'em_air', 'Qn_LW', 'Q_sum'):
        e = rel_err(getattr(m, k).cpu().numpy(), g[case + '_' + k])
        assert e <= 1e-11, '%s rel err %.3e' % (k, e)
    dark_ref = (g[case + '_Qn_SW'] == 0)
    assert np.array_equal(m.Qn_SW.cpu().numpy() == 0, dark_ref)      # same sunrise/sunset cells


def test_unfused_coupling_on_device_equals_fused(sources_gold):
    """EMELI-style coupling without leaving the GPU: the stand-alone met / snow / evap /
    infil steppers hand CUDA tensors to the Channels engine through its inputs (zero
    copy), in provider order met, snow, evap, infil, channels -- the result equals the
    run with the same source terms fused into the routing kernel."""
    import torch
    from conftest import rel_err
    from oracle import np_oracle as o
    from topoflow36_b200 import cases, synth
    from topoflow36_b200.routing import ChannelsEngine
    from topoflow36_b200.sources import MetPrecip, DegreeDay, PriestleyTaylor, GreenAmpt
    ny, nx = 64, 96
    DEM = synth.fractal_dem(ny, nx, seed=12, sigma=1.0)
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    ds, _ = o.d8_ds_dw(code, dx, dy, dd)
    slope = np.array(o.d8_slope(DEM, code, ds), dtype='float64')
    o.remove_bad_slopes(slope)
    width, angle, nval = synth.channel_params(ny, nx, seed=12)
    rng = np.random.default_rng(8)
    T_air = 2.0 + 3.0 * (2.0 * rng.random((ny, nx)) - 1.0)
    P, dt = 4e-5, 2.0
    src = dict(cases.FULL_SOURCES, h0_swe=0.004, I0=0.05)
    sc = cases.FULL_SCALARS

    def engine():
        return ChannelsEngine(ny, nx, dt=dt, code=code, dx=dx, dy=dy, dd=dd, S_bed=slope,
                              width=np.float64(width), angle=np.float64(angle) * o.DEG_TO_RAD,
                              rough=np.float64(nval), outlet=(ny - 2, nx // 2))
    fused = engine()
    fused.enable_fused_sources(**src)
    fused.set_input('P_rain', P)
    fused.set_input('T_air', T_air)
    for k, v in list(sc.items()) + list(cases.GA_TREYNOR.items()):
        fused.set_input(k, v)
    eng = engine()
    met = MetPrecip(ny, nx, dt, T_rain_snow=src['T_rain_snow'])
    snow = DegreeDay(ny, nx, dt, sc['c0'], sc['T0'], h0_swe=src['h0_swe'], rho_snow=src['rho_snow'])
    evap = PriestleyTaylor(ny, nx, dt, alpha=src['alpha'], K_soil=src['K_soil'],
                           soil_x=src['soil_x'], T_soil_x=src['T_soil_x'])
    infil = GreenAmpt(ny, nx, dt, I0=src['I0'], **cases.GA_TREYNOR)
    Ta = torch.as_tensor(T_air).cuda()
    for i in range(60):
        met.update(P, Ta)
        snow.update(met.P_snow, Ta)
        evap.update(Ta, sc['T_surf'], sc['Qn_SW'], sc['Qn_LW'])
        infil.update(met.P_rain, snow.SM)
        for n, t in (('P_rain', met.P_rain), ('SM', snow.SM), ('ET', evap.ET), ('IN', infil.IN)):
            eng.set_input(n, t)                      # CUDA tensors: bound in place
        eng.step(time_min=i * dt / 60.0)
        fused.step(time_min=i * dt / 60.0)
    # fused: lean packed kernel; un-fused (four forcing grids from device providers): extended one
    assert fused.last_kernel == 'lean' and eng.last_kernel == 'ext'
    for a, b in ((eng.vol, fused.vol), (eng.own(eng.d), fused.own(fused.d)),
                 (eng.own(eng.u), fused.own(fused.u)), (eng.Q, fused.Q),
                 (infil.I, fused.I), (snow.h_swe, fused.h_swe)):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-12
    s1, s2 = eng.read_scalars(), fused.read_scalars()
    assert abs(s1.vol_R - s2.vol_R) <= 1e-12 * abs(s2.vol_R)
    assert abs(infil.vol_IN - s2.vol_IN) <= 1e-12 * s2.vol_IN
    assert abs(snow.vol_SM - s2.vol_SM) <= 1e-12 * max(s2.vol_SM, 1e-300)
    assert float(fused.Q.max()) > 0


@pytest.mark.parametrize('method', ['aging', 'simple'])
def test_albedo_vs_reference_fixture(method):
    """met_base.update_albedo over 30 six-hour steps with two snowfalls: the 3-day ring,
    the day counter and the albedo after every step (exp() differs from numpy's by <= 1 ulp)."""
    from conftest import load_golden
    from topoflow36_b200 import sources as S
    d = load_golden('albedo.npz')
    alb = S.Albedo(9, 13, 21600.0, albedo=d[method + '_albedo0'], method=method,
                   rho_H2O=1000.0, rho_snow=250.0)
    assert method == 'simple' or alb.slots == 12
    for k in range(30):
        alb.update(d[method + '_P_snow'][k], d[method + '_T_air'][k], d[method + '_h_snow'][k],
                   d[method + '_h_ice'][k])
        got = alb.albedo.cpu().numpy()
        assert rel_err(got, d[method + '_albedo'][k]) <= 1e-15, k
        if method == 'aging':
            assert np.array_equal(alb.n.cpu().numpy(), d['aging_n'][k]), k
