"""Stand-alone source-term kernels (C ABI) against values produced by the
REFERENCE's own component methods (tests/golden/sources.npz, made by
oracle/gen_golden.gen_sources).  Only + - * / min max are involved and the TU is
compiled without FMA contraction, so equality is exact."""
import numpy as np
import pytest

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
