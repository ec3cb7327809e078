"""The numpy oracle against the committed reference fixtures (CPU only).

These are the checks that PIN the oracle: every fixture under tests/golden was
produced by the unmodified reference (oracle/gen_golden.py) or ships with it
(Treynor_flow.rtg / Treynor_area.rtg, RiverTools output)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import np_oracle as o

GRIDS = ('vol', 'd', 'u', 'Q', 'Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
SCALARS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
           'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak')


def test_resolve_table_golden(d8_synth):
    t = o.resolve_table()
    assert t.sum() == 10161 and (t != 0).sum() == 255
    assert list(t[:16]) == [0, 1, 2, 2, 4, 1, 2, 2, 8, 8, 2, 2, 8, 8, 2, 2]
    assert hashlib.sha256(t.tobytes()).hexdigest().startswith('b0cd03c82342c315')
    assert np.array_equal(t, d8_synth['resolve'])


def test_treynor_d8_matches_rivertools_fixtures(treynor):
    """Treynor_flow.rtg / Treynor_area.rtg (shipped with the reference): bit-exact."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd, link_flats=True)
    assert np.array_equal(code, treynor['flow'])
    pa = np.float64(treynor['xres'] * treynor['yres']) / 1e6
    A_lv, _ = o.d8_area_levels(code, pa)
    A_kahn, N = o.d8_area_kahn(code, pa, with_counts=True)
    assert np.array_equal(A_lv, treynor['area'])
    assert np.array_equal(A_kahn, treynor['area'])
    assert np.array_equal(A_kahn, treynor['d8_area'])
    hist = {int(k): int((code == k).sum()) for k in np.unique(code)}
    assert hist == {0: 142, 1: 102, 2: 291, 4: 180, 8: 132, 16: 163, 32: 157, 64: 99, 128: 10}
    assert N.max() == 554


@pytest.mark.parametrize('lf', [1, 0])
def test_d8_synth_bit_exact(d8_synth, lf):
    g = d8_synth
    DEM = g['DEM']
    ny, nx = DEM.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    res = o.resolve_table()
    start = o.d8_start_codes(DEM, dx, dy, dd, link_flats=bool(lf))
    assert np.array_equal(start, g['lf%d_start' % lf])
    ties = o.d8_break_ties(start.copy(), res)
    assert np.array_equal(ties, g['lf%d_ties' % lf])
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd, link_flats=bool(lf))
    assert np.array_equal(code, g['lf%d_code' % lf])
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    assert np.array_equal(ds, g['lf%d_ds' % lf]) and np.array_equal(dw, g['lf%d_dw' % lf])
    A = o.d8_area_kahn(code, np.float64(900.0) / 1e6)
    assert np.array_equal(A, g['lf%d_area' % lf])
    pr, pc = o.d8_parent_ids(code)
    assert np.array_equal(pr * nx + pc, g['lf%d_pid' % lf])
    S = o.d8_slope(DEM, code, ds)
    assert np.array_equal(S, g['lf%d_slope' % lf], equal_nan=True)


def _treynor_channels(treynor):
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    code = treynor['flow']
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    ch = o.Channels(code, ds, dw, treynor['slope'], treynor['width'], treynor['angle'],
                    nval=treynor['nval'], da=900.0, dt=6.0, outlet=(34, 16))
    ch.ET = np.zeros(code.shape)
    ch.IN = np.zeros(code.shape)
    return ch


def test_treynor_kw_storm_bit_exact(treynor, treynor_kw):
    """1200 steps of the June-20-1967 storm, stand-alone reference run."""
    ch = _treynor_channels(treynor)
    rain = treynor['rain_mmph']
    series = treynor_kw['outlet_series']
    for i in range(1200):
        m = i // 10
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        ch.P_rain = np.array(P, dtype='float64')
        ch.step()
        assert (float(ch.Q_outlet), float(ch.u_outlet), float(ch.d_outlet),
                float(ch.f_outlet)) == tuple(series[i])
        if (i + 1) in (100, 400, 1200):
            for n in GRIDS:
                assert np.array_equal(getattr(ch, n), treynor_kw['s%d_%s' % (i + 1, n)]), n
            for n in SCALARS:
                assert float(getattr(ch, n)) == float(treynor_kw['s%d_%s' % (i + 1, n)]), n


def test_treynor_dw_storm_bit_exact(treynor, treynor_dw):
    """The June-20-1967 storm on the DIFFUSIVE wave (dt = 2 s, 3600 steps, rise and two hours
    of recession): the numpy oracle against the stand-alone reference run, outlet values every
    step and every grid (incl. S_free) at three checkpoints, bit for bit."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    code = treynor['flow']
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    ch = o.Channels(code, ds, dw, treynor['slope'], treynor['width'], treynor['angle'],
                    nval=treynor['nval'], da=900.0, dt=2.0, kinematic=False, outlet=(34, 16))
    ch.ET = np.zeros(code.shape)
    ch.IN = np.zeros(code.shape)
    rain = treynor['rain_mmph']
    series = treynor_dw['outlet_series']
    assert 12.0 < series[:, 0].max() < 12.6                  # a smooth hydrograph, not the dt = 6 s saw-tooth
    for i in range(3600):
        m = i // 30
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        ch.P_rain = np.array(P, dtype='float64')
        ch.step()
        assert (float(ch.Q_outlet), float(ch.u_outlet), float(ch.d_outlet),
                float(ch.f_outlet)) == tuple(series[i]), i
        if (i + 1) in (600, 1300, 3600):
            for n in GRIDS + ('S_free',):
                assert np.array_equal(getattr(ch, n), treynor_dw['s%d_%s' % (i + 1, n)]), n
            for n in SCALARS:
                assert float(getattr(ch, n)) == float(treynor_dw['s%d_%s' % (i + 1, n)]), n


@pytest.mark.parametrize('tag', ['kw_man', 'kw_low', 'dw_man', 'dw_low'])
def test_synth_channels_bit_exact(synth_channels, tag):
    g = synth_channels
    kin, man = tag.startswith('kw'), tag.endswith('man')
    code = g['code']
    ch = o.Channels(code, g['ds32'], g['dw32'], g['slope'], g['width'], g['angle'],
                    nval=g['nval'] if man else None, z0val=None if man else g['z0val'],
                    da=900.0, dt=2.0, kinematic=kin, manning=man,
                    outlet=(code.shape[0] - 2, code.shape[1] // 2))
    z = lambda v: np.array(v, dtype='float64')
    ch.P_rain, ch.GW, ch.SM = z(80 / 3.6e6), z(1e-7), z(2e-7)
    ch.ET, ch.IN = g['ET0'].copy(), g['IN0'].copy()
    assert np.array_equal(ch.S_bed, g[tag + '_S_bed'])
    for i in range(150):
        if i == 100:
            ch.P_rain = z(0)
        ch.step()
    for n in GRIDS + (() if kin else ('S_free',)):
        assert np.array_equal(getattr(ch, n), g[tag + '_' + n], equal_nan=True), n
    for n in SCALARS:
        assert float(getattr(ch, n)) == float(g[tag + '_' + n]), n
    assert np.array_equal(ch.ET, g[tag + '_ET_after'])


def test_sources_bit_exact(sources_gold):
    g = sources_gold
    I = np.zeros_like(g['ga_P_rain']) + 1e-6
    for it in range(3):
        IN, I = o.green_ampt_step(I, g['ga_P_rain'], g['ga_SM'], g['ga_Ks'], g['ga_Ki'],
                                  g['ga_qs'], g['ga_qi'], g['ga_G'], 5.0,
                                  h_table=g['ga_h_table'], elev=g['ga_elev'])
        assert np.array_equal(IN, g['ga_IN%d' % it])
        assert np.array_equal(I, g['ga_I%d' % it])
    h = g['dd_h_swe_init']
    for it in range(3):
        SM, h, hs = o.degree_day_step(h, g['dd_P_snow'], g['dd_T_air'], 2.7, -0.2, 3600.0)
        assert np.array_equal(SM, g['dd_SM%d' % it])
        assert np.array_equal(h, g['dd_h_swe%d' % it])
        assert np.array_equal(hs, g['dd_h_snow%d' % it])
    ET = o.priestley_taylor_step(g['pt_T_air'], g['pt_T_surf'], g['pt_Qn_SW'], g['pt_Qn_LW'],
                                 1.26, 0.45, 0.05, 12.0)
    assert np.array_equal(ET, g['pt_ET'])
    pr, ps = o.met_split_precip(g['met_P'], g['met_T_air'], 0.5)
    assert np.array_equal(pr, g['met_P_rain']) and np.array_equal(ps, g['met_P_snow'])


@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_flood_option_bit_exact(flood_gold, tag):
    """FLOOD_OPTION (channels_base.py:1701-1862, :2149-2185, :2393-2448): the numpy
    restatement against a stand-alone run of the reference with overbank flow."""
    g = flood_gold
    kin = (tag == 'kw')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    ds32, dw32 = o.d8_ds_dw(code, dx, dy, dd)
    ch = o.Channels(code, ds32, dw32, g['slope'], g['width'], g['angle'], nval=g['nval'],
                    da=900.0, dt=2.0, kinematic=kin, outlet=(ny - 2, nx // 2), flood=True,
                    d_bankfull=g['d_bankfull'])
    z = lambda v: np.array(v, dtype='float64')
    ch.P_rain = z(150 / 3.6e6)
    for i in range(150):
        if i == 100:
            ch.P_rain = z(0)
        ch.step()
    assert g[tag + '_d_flood'].max() > 0.1          # the banks really overflow
    for n in GRIDS + ('d_flood', 'vol_flood', 'vol_chan', 'Qc', 'Qf') + (() if kin else ('S_free',)):
        assert np.array_equal(getattr(ch, n), g[tag + '_' + n], equal_nan=True), n
    for n in SCALARS:
        assert float(getattr(ch, n)) == float(g[tag + '_' + n]), n


@pytest.mark.parametrize('case', ['day', 'dusk', 'night'])
def test_met_energy_chain_bit_exact(met_gold, case):
    """The numpy restatement of met_base.update()'s energy chain and of
    solar_funcs.Clear_Sky_Radiation equals the reference's own methods bit for bit
    (full sun, partial darkness on sloping cells, night)."""
    g = met_gold
    I = {k[len(case) + 4:]: g[k] for k in g if k.startswith(case + '_in_')}
    r = o.met_energy_chain(I['T_air'], I['RH'], I['uz'], np.float64(10.0), I['z0_air'], I['h_snow'],
                           I['h_ice'], I['p0'], I['albedo'], I['em_surf'], I['lat_deg'], I['alpha'],
                           I['beta'], I['julian_day'], I['TSN_offset'], dust_atten=np.float64(0.08),
                           cloud_factor=I['cloud_factor'], canopy_factor=I['canopy_factor'])
    for k, v in r.items():
        assert np.array_equal(v, g[case + '_' + k], equal_nan=True), k
    if case == 'dusk':
        sw = g['dusk_Qn_SW']
        assert (sw == 0).any() and (sw > 0).any()


@pytest.mark.parametrize('case', ['day', 'dusk', 'night'])
def test_solar_host_date_arithmetic(met_gold, case):
    """solar_host.day_and_tsn == met_base.update_julian_day (Julian day with integer
    hour, equation of time, true solar noon) for a longitude grid."""
    from topoflow36_b200 import solar_host as sh
    g = met_gold
    h = float(g[case + '_in_hour'])
    JD, tsn = sh.day_and_tsn('1967-06-20 00:00:00', h * 60.0, g[case + '_in_lon_deg'], -6,
                             current_year=int(g['current_year']))
    assert JD == float(g[case + '_ujd_julian_day'])
    assert np.array_equal(tsn, g[case + '_ujd_TSN_offset'])


@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_attenuate_bit_exact(flood_gold, attenuate_gold, tag):
    """ATTENUATE (channels_base.py:2077-2084): numpy restatement == reference run."""
    g0, g = flood_gold, attenuate_gold
    code = g0['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    ds32, dw32 = o.d8_ds_dw(code, dx, dy, dd)
    ch = o.Channels(code, ds32, dw32, g0['slope'], g0['width'], g0['angle'], nval=g0['nval'],
                    da=900.0, dt=2.0, kinematic=(tag == 'kw'), outlet=(ny - 2, nx // 2),
                    attenuate=True, n_days=0.002)
    ch.P_rain = np.array(150 / 3.6e6)
    for i in range(150):
        if i == 100:
            ch.P_rain = np.array(0.0)
        ch.step()
    for n in GRIDS + ('vol_stored',):
        assert np.array_equal(getattr(ch, n), g[tag + '_' + n], equal_nan=True), n
    for n in SCALARS:
        assert float(getattr(ch, n)) == float(g[tag + '_' + n]), n


def _geo_info(tmp_path, ny, nx):
    # through the RTI file, like the reference run that produced the fixture (the header
    # keeps 12 decimals of the box edges)
    from topoflow36_b200 import grid_io
    info = grid_io.make_rti(nx, ny, 1.0, 1.0, pixel_geom=0, y_south_edge=41.0, x_west_edge=-95.7)
    grid_io.write_rti(str(tmp_path / 'geo.rti'), info)
    return grid_io.read_rti(str(tmp_path / 'geo.rti'))


def test_geographic_pixels_bit_exact(geo_gold, tmp_path):
    """Fixed-angle pixels: per-row dx/dd, per-cell da (utils/pixels.py:71-233) through
    the D8 toolkit and both wave types -- numpy restatement == reference run."""
    from topoflow36_b200 import grid_io
    g = geo_gold
    ny, nx = g['DEM'].shape
    info = _geo_info(tmp_path, ny, nx)
    dx, dy, dd = (np.float32(v) for v in grid_io.pixel_sizes_by_row(info))
    assert np.array_equal(dx, g['dx']) and np.array_equal(dy, g['dy']) and np.array_equal(dd, g['dd'])
    da = grid_io.pixel_area(info)
    assert np.array_equal(da, g['da'])
    code, _ = o.d8_flow_grid(g['DEM'], dx, dy, dd)
    assert np.array_equal(code, g['code'])
    ds32, dw32 = o.d8_ds_dw(code, dx, dy, dd)
    assert np.array_equal(ds32, g['ds32']) and np.array_equal(dw32, g['dw32'])
    with np.errstate(all='ignore'):
        assert np.array_equal(np.float32(o.d8_slope(g['DEM'], code, ds32)), g['slope'])
    for tag in ('kw', 'dw'):
        ch = o.Channels(code, ds32, dw32, g['slope'], g['width'], g['angle'], nval=g['nval'],
                        da=da, dt=2.0, kinematic=(tag == 'kw'), outlet=(ny - 2, nx // 2))
        ch.P_rain = np.array(100 / 3.6e6)
        for i in range(120):
            if i == 80:
                ch.P_rain = np.array(0.0)
            ch.step()
        for n in GRIDS:
            assert np.array_equal(getattr(ch, n), g[tag + '_' + n], equal_nan=True), n
        for n in SCALARS:
            assert float(getattr(ch, n)) == float(g[tag + '_' + n]), n


def test_prep_oracle_vs_reference_fixture():
    """fill_pits / get_new_slope_grid / get_grid_from_TCA restatements against the
    reference's outputs (tests/golden/prep.npz, gen_prep)."""
    from conftest import load_golden
    d = load_golden('prep.npz')
    for t in 'ab':
        D = d['fill_%s_in' % t].copy()
        o.fill_pits(D)
        assert np.array_equal(D, d['fill_%s_out' % t]), t
    DEM, code = d['slope_DEM'], d['slope_code']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], 30.0, 30.0)
    ds, _ = o.d8_ds_dw(code, dx, dy, dd)
    S, reps = o.new_slope_grid(DEM, code, ds)
    assert reps > 3 and np.array_equal(np.float32(S), d['slope_new'])
    assert np.array_equal(o.d8_aspect(code), d['aspect'])
    A = d['tca_area']
    for tag, kw in {'w': dict(g1=25.0, g2=1.5), 'n': dict(g1=0.03, p=-0.2),
                    'd': dict(A1=50.0, g1=3.0, A2=0.01, g2=0.2)}.items():
        g, c, p = o.grid_from_TCA(A, **kw)
        assert np.array_equal(np.float32(g), d['tca_' + tag]), tag


def test_albedo_oracle_vs_reference_fixture():
    """met_base.update_albedo ('aging' with the 3-day snowfall ring, 'simple')."""
    from conftest import load_golden
    d = load_golden('albedo.npz')
    for method in ('aging', 'simple'):
        slots = int(np.int64(259200 / 21600.0))
        st = dict(albedo=d[method + '_albedo0'].copy(), n=0, ring=np.zeros((slots, 9, 13)))
        for k in range(30):
            a = o.albedo_update(st, d[method + '_P_snow'][k], d[method + '_T_air'][k],
                                d[method + '_h_snow'][k], d[method + '_h_ice'][k], 21600.0,
                                1000.0, 250.0, method)
            assert np.array_equal(a, d[method + '_albedo'][k]), (method, k)
            if method == 'aging':
                assert np.array_equal(np.zeros((9, 13)) + st['n'], d['aging_n'][k])
    assert d['aging_n'].max() > 3.0


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
def test_committed_fixtures_are_what_the_reference_produces(tmp_path):
    """Re-run oracle/gen_golden.py (the UNMODIFIED reference, a few seconds) into a scratch
    directory: every array of every committed fixture must come out identical."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TFB_GOLDEN_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(root, 'oracle', 'gen_golden.py')], env=env,
                       cwd=root, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = os.path.join(root, 'tests', 'golden')
    names = sorted(f for f in os.listdir(gold) if f.endswith('.npz'))
    assert len(names) == 14 and names == sorted(f for f in os.listdir(str(tmp_path)) if f.endswith('.npz'))
    for f in names:
        a, b = np.load(os.path.join(gold, f)), np.load(os.path.join(str(tmp_path), f))
        assert sorted(a.files) == sorted(b.files), f
        for k in a.files:
            assert np.array_equal(a[k], b[k], equal_nan=True), (f, k)
    assert json.load(open(os.path.join(gold, 'meta.json'))) == \
        json.load(open(os.path.join(str(tmp_path), 'meta.json')))


@pytest.mark.parametrize('kin', [True, False])
def test_sinuosity_and_initial_depth_bit_exact(kin):
    """sinu = 1.3 (flow lengths * sinu, slope / sinu) and d0 = 0.05 m (initial A_wet,
    P_wet, vol): initialize_computed_vars :1137-1405 and 120 steps, KW and DW."""
    from conftest import load_golden
    g = load_golden('sinu_d0_channels.npz')
    tag = 'kw' if kin else 'dw'
    code = g['code']
    ch = o.Channels(code, g['ds32'], g['dw32'], g['slope'], g['width'], g['angle'], nval=g['nval'],
                    da=900.0, dt=2.0, kinematic=kin, outlet=(code.shape[0] - 2, code.shape[1] // 2),
                    sinu=float(g['sinu']), d0=float(g['d0']))
    assert np.array_equal(ch.S_bed, g[tag + '_S_bed'])
    for n in GRIDS:
        assert np.array_equal(getattr(ch, n), g[tag + '_init_' + n], equal_nan=True), n
    z = lambda v: np.array(v, dtype='float64')
    ch.P_rain = z(60 / 3.6e6)
    for i in range(120):
        if i == 80:
            ch.P_rain = z(0)
        ch.step()
    for n in GRIDS + (() if kin else ('S_free',)):
        assert np.array_equal(getattr(ch, n), g[tag + '_' + n], equal_nan=True), n
    for n in SCALARS:
        assert float(getattr(ch, n)) == float(g[tag + '_' + n]), n
    assert float(g[tag + '_Q_outlet']) > 0.1
