"""The C restatement (oracle/tf_oracle.c) against the numpy oracle, which is
itself bit-exact against the reference fixtures: D8 is bit-exact, routing agrees
to ~1e-15 (glibc pow/log vs numpy's)."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import c_oracle, np_oracle as o

GRIDS = ('vol', 'd', 'u', 'Q', 'Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
SCALARS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
           'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak')


def test_d8_bit_exact_vs_fixtures(treynor, d8_synth):
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    code, _ = c_oracle.d8_flow_grid(DEM, dx, dy, dd)
    assert np.array_equal(code, treynor['flow'])
    A, N = c_oracle.d8_area(code, np.float64(900.0) / 1e6, with_counts=True)
    assert np.array_equal(A, treynor['area']) and N.max() == 554
    for lf in (1, 0):
        DEM = d8_synth['DEM']
        dx, dy, dd = o.pixel_dims(DEM.shape[0], 30.0, 30.0)
        code, _ = c_oracle.d8_flow_grid(DEM, dx, dy, dd, link_flats=bool(lf))
        assert np.array_equal(code, d8_synth['lf%d_code' % lf])
        assert np.array_equal(c_oracle.d8_area(code, np.float64(900.0) / 1e6), d8_synth['lf%d_area' % lf])


@pytest.mark.parametrize('tag', ['kw_man', 'kw_low', 'dw_man', 'dw_low'])
def test_channels_vs_reference_fixture(synth_channels, tag):
    g = synth_channels
    kin, man = tag.startswith('kw'), tag.endswith('man')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    ch = c_oracle.Channels(code, dx, dy, dd, g['slope'], g['width'], g['angle'],
                           g['nval'] if man else g['z0val'], dt=2.0, kinematic=kin, manning=man,
                           outlet=(ny - 2, nx // 2))
    ch.set_input('P_rain', 80 / 3.6e6)
    ch.set_input('GW', 1e-7)
    ch.set_input('SM', 2e-7)
    ch.set_input('ET', g['ET0'].copy())
    ch.set_input('IN', g['IN0'].copy())
    for i in range(150):
        if i == 100:
            ch.set_input('P_rain', 0.0)
        ch.step()
    for n in GRIDS:
        assert rel_err(getattr(ch, n), g[tag + '_' + n]) <= 1e-12, n
    for n in SCALARS:
        assert abs(getattr(ch, n) - float(g[tag + '_' + n])) <= 1e-12 * abs(float(g[tag + '_' + n])), n
    assert np.array_equal(ch.src['ET'], g[tag + '_ET_after'])


def test_treynor_storm(treynor, treynor_kw):
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    ch = c_oracle.Channels(treynor['flow'], dx, dy, dd, treynor['slope'], treynor['width'],
                           treynor['angle'], treynor['nval'], dt=6.0, outlet=(34, 16))
    ch.set_input('ET', np.zeros(DEM.shape))
    ch.set_input('IN', np.zeros(DEM.shape))
    rain = treynor['rain_mmph']
    series = treynor_kw['outlet_series']
    for i in range(1200):
        m = i // 10
        ch.set_input('P_rain', np.float64(rain[m]) / 3.6e6 if m < rain.size else 0.0)
        ch.step()
        got = np.array([ch.Q_outlet, ch.u_outlet, ch.d_outlet, ch.f_outlet])
        assert rel_err(got, series[i]) <= 1e-12
    for n in GRIDS:
        assert rel_err(getattr(ch, n), treynor_kw['s1200_' + n]) <= 1e-12, n


def test_treynor_dw_storm(treynor, treynor_dw):
    """Diffusive wave, dt = 2 s, 3600 steps of the same storm (rise + recession)."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    ch = c_oracle.Channels(treynor['flow'], dx, dy, dd, treynor['slope'], treynor['width'],
                           treynor['angle'], treynor['nval'], dt=2.0, kinematic=False, outlet=(34, 16))
    ch.set_input('ET', np.zeros(DEM.shape))
    ch.set_input('IN', np.zeros(DEM.shape))
    rain = treynor['rain_mmph']
    series = treynor_dw['outlet_series']
    for i in range(3600):
        m = i // 30
        ch.set_input('P_rain', np.float64(rain[m]) / 3.6e6 if m < rain.size else 0.0)
        ch.step()
        got = np.array([ch.Q_outlet, ch.u_outlet, ch.d_outlet, ch.f_outlet])
        assert rel_err(got, series[i]) <= 1e-12, i
        if (i + 1) in (600, 1300, 3600):
            for n in GRIDS:
                assert rel_err(getattr(ch, n), treynor_dw['s%d_%s' % (i + 1, n)]) <= 1e-12, n


def test_c_fill_pits_vs_reference_fixture_and_numpy_oracle():
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'prep.npz')
    d = np.load(gold)
    from oracle import c_oracle as co, np_oracle as o
    for t in 'ab':
        D = d['fill_%s_in' % t].copy()
        n = co.fill_pits(D)
        assert np.array_equal(D, d['fill_%s_out' % t]) and n > 1000
        D1, D2 = d['fill_%s_in' % t].copy(), d['fill_%s_in' % t].copy()
        assert co.fill_pits(D1, close_nodata=True) == o.fill_pits(D2, close_nodata=True)
        assert np.array_equal(D1, D2)
