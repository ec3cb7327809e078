"""The GPU Channels component driven the way EMELI drives the reference's
(initialize(cfg_file, mode) -> set_value() x7 -> update() ... -> finalize()),
from CFG + RTI + RTG files, against the stand-alone REFERENCE run stored in
tests/golden/treynor_kw_standalone.npz.  Tolerance as in test_channels_gpu.py:
RTOL = 1e-11 (north-star: 1e-9)."""
import os

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-11

LONG = {'P_rain': 'atmosphere_water__rainfall_volume_flux',
        'MR': 'glacier_ice__melt_volume_flux',
        'GW': 'land_surface_water__baseflow_volume_flux',
        'ET': 'land_surface_water__evaporation_volume_flux',
        'IN': 'soil_surface_water__infiltration_volume_flux',
        'SM': 'snowpack__melt_volume_flux',
        'rho_H2O': 'water-liquid__mass-per-volume_density'}
z = lambda v: np.array(v, dtype='float64')


def _stage_treynor(tmp_path, treynor, kinematic=True, byte_order='MSB', **kw):
    from topoflow36_b200 import synth
    return synth.write_case(str(tmp_path), treynor['DEM'], treynor['slope'], treynor['width'],
                            treynor['angle'], nval=treynor['nval'], site='Treynor',
                            case='June_20_67', dx=float(treynor['xres']), dy=float(treynor['yres']),
                            dt=6.0, kinematic=kinematic, outlet=(34, 16), n_steps=1200,
                            byte_order=byte_order, **kw)


@pytest.mark.parametrize('diag', [True, False])
def test_treynor_storm_through_the_bmi_component(tmp_path, treynor, treynor_kw, diag):
    """diag=True: every diagnostic grid materialised, ET / IN arrive as grids (what the
    stand-alone reference run was given).  diag=False: what the EMELI run of this case
    delivers -- all six inputs 0-D -- which must ride the benchmarked packed kernel
    although nx = 29 is odd (pitched rows)."""
    from topoflow36_b200 import channels_kinematic_wave as ckw
    cfg = _stage_treynor(tmp_path, treynor)
    c = ckw.channels_component()
    c.DIAGNOSTICS = diag
    c.initialize(cfg_file=cfg, mode='driver', SILENT=True)
    assert c.get_status() == 'initialized'
    assert np.array_equal(c.d8.d8_grid, treynor['flow'])          # D8 codes, bit-exact
    assert np.array_equal(c.d8.A, treynor['area'])                # contributing area, bit-exact
    ny, nx = treynor['DEM'].shape
    ET, IN = (np.zeros((ny, nx)), np.zeros((ny, nx))) if diag else (z(0), z(0))
    for k, v in dict(MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0), ET=ET, IN=IN).items():
        c.set_value(LONG[k], v)
    # EMELI takes references ONCE; they must track the model in place
    q_ref = c.get_value_ptr('basin_outlet_water_x-section__volume_flow_rate')
    Q_ref = c.get_value_ptr('channel_water_x-section__volume_flow_rate')
    rain = treynor['rain_mmph']
    series = treynor_kw['outlet_series']
    worst = 0.0
    n = 0
    while not c.DONE:
        m = n // 10
        c.set_value(LONG['P_rain'], z(np.float64(rain[m]) / 3.6e6 if m < rain.size else 0.0))
        c.update()
        got = np.array([float(q_ref), float(c.u_outlet), float(c.d_outlet), float(c.f_outlet)])
        worst = max(worst, rel_err(got, series[n]))
        n += 1
        if n in (100, 400):
            assert rel_err(Q_ref, treynor_kw['s%d_Q' % n]) <= RTOL
    assert n == 1200 and c.time_index == 1200
    assert c.engine.last_path == 'packed'
    assert c.engine.last_kernel == ('ext' if diag else 'lean')
    assert abs(float(c.time_min) - 120.0) < 1e-9
    assert worst <= RTOL, worst
    for g in ('vol', 'd', 'u', 'Q') + (('Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
                                       if diag else ()):
        assert rel_err(getattr(c, g), treynor_kw['s1200_' + g]) <= RTOL, g
    for s in ('vol_R', 'vol_Q', 'vol_edge', 'Q_peak', 'T_peak', 'u_peak', 'd_peak', 'Td_peak'):
        want = float(treynor_kw['s1200_' + s])
        assert abs(float(getattr(c, s)) - want) <= RTOL * abs(want), s
    c.finalize()
    assert c.get_status() == 'finalized'
    want_vol = treynor_kw['s1200_vol'].sum()
    assert abs(float(c.vol_chan_sum) - want_vol) <= 1e-12 * want_vol
    assert float(c.Q_max) == treynor_kw['s1200_Q'][1:-1, 1:-1].max()


def test_bmi_surface_and_outputs(tmp_path, treynor):
    """get_value / get_value_at_indices / units / grid info, RTS + text output,
    little-endian files, stability abort."""
    from topoflow36_b200 import channels_kinematic_wave as ckw
    cfg = _stage_treynor(tmp_path, treynor, byte_order='LSB')
    txt = open(cfg).read().replace('SAVE_Q_GRIDS        | No', 'SAVE_Q_GRIDS        | Yes') \
                          .replace('SAVE_Q_PIXELS       | No', 'SAVE_Q_PIXELS       | Yes')
    open(cfg, 'w').write(txt)
    c = ckw.channels_component()
    c.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    ny, nx = treynor['DEM'].shape
    assert c.get_var_units('channel_water_x-section__volume_flow_rate') == 'm3 s-1'
    assert 'atmosphere_water__rainfall_volume_flux' in c.get_input_var_names()
    assert c.get_attribute('cfg_extension') == '_channels_kinematic_wave.cfg'
    assert c.get_time_units() == 'seconds' and float(c.get_time_step()) == 6.0
    shape = np.zeros(2, dtype='int64')
    assert tuple(c.get_grid_shape(0, shape)) == (ny, nx)
    c.set_value(LONG['P_rain'], z(100 / 3.6e6))
    Q_at_60s = None
    for i in range(20):
        c.update(-1.0)
        if i == 10:                      # the update that starts at t = 60 s saves a frame
            Q_at_60s = c.Q.copy()
    dest = np.zeros((ny, nx))
    c.get_value('channel_water_x-section__mean_depth', dest)
    assert dest.max() > 0 and np.array_equal(dest, c.d)
    ids = np.array([34 * nx + 16, 30 * nx + 15])
    got = c.get_value_at_indices('channel_water_x-section__mean_depth', np.zeros(2), ids)
    assert np.array_equal(got, dest.flat[ids])
    t = c.get_value_tensor('channel_water_x-section__mean_depth')
    assert t.is_cuda and tuple(t.shape) == (ny, nx)
    c.finalize()
    rts = os.path.join(str(tmp_path), 'June_20_67_2D-Q.rts')
    assert os.path.getsize(rts) == 2 * ny * nx * 4          # frames at t = 0 s and 60 s
    frames = np.fromfile(rts, dtype='float32').reshape(2, ny, nx)   # written asynchronously
    assert np.array_equal(frames[1], np.float32(Q_at_60s)) and frames[1].max() > 0
    last = np.loadtxt(os.path.join(str(tmp_path), 'June_20_67_0D-Q.txt'), skiprows=2)[-1]
    assert abs(last[0] - 1.0) < 1e-9 and abs(last[1] - Q_at_60s[34, 16]) <= 5e-8
    lines = open(os.path.join(str(tmp_path), 'June_20_67_0D-Q.txt')).read().splitlines()
    assert len(lines) == 2 + 2
    # an absurd time step must abort like the reference (RuntimeError + 'failed')
    c2 = ckw.channels_component()
    cfg2 = open(cfg).read().replace('CHECK_STABILITY     | Yes', 'CHECK_STABILITY     | Yes')
    open(cfg, 'w').write(cfg2)
    c2.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    c2.set_value(LONG['P_rain'], z(np.float64('nan')))
    with pytest.raises(RuntimeError):
        for _ in range(3):
            c2.update()
    assert c2.get_status() == 'failed' and c2.DONE


def test_diffusive_component_vs_engine(tmp_path, synth_channels):
    """DW component from files == reference fixture (S_free included)."""
    from topoflow36_b200 import channels_diffusive_wave as cdw, synth
    g = synth_channels
    ny, nx = g['code'].shape
    cfg = synth.write_case(str(tmp_path), g['DEM'], g['slope'], g['width'], g['angle'],
                           nval=g['nval'], dt=2.0, kinematic=False, outlet=(ny - 2, nx // 2))
    c = cdw.channels_component()
    c.DIAGNOSTICS = True
    c.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    assert np.array_equal(c.d8.d8_grid, g['code'])
    ET = g['ET0'].copy()
    for k, v in dict(P_rain=z(80 / 3.6e6), MR=z(0), GW=z(1e-7), SM=z(2e-7), rho_H2O=z(1000.0),
                     ET=ET, IN=g['IN0'].copy()).items():
        c.set_value(LONG[k], v)
    for i in range(150):
        if i == 100:
            c.set_value(LONG['P_rain'], z(0))
        c.update()
    for n in ('vol', 'd', 'u', 'Q', 'Rh', 'f', 'froude', 'S_free'):
        assert rel_err(getattr(c, n), g['dw_man_' + n]) <= RTOL, n
    # the PROVIDER's ET array was masked in place, bit-exactly
    assert np.array_equal(ET, g['dw_man_ET_after'])
    for s in ('vol_R', 'vol_Q', 'vol_edge', 'Q_peak'):
        want = float(g['dw_man_' + s])
        assert abs(float(getattr(c, s)) - want) <= RTOL * abs(want), s


@pytest.mark.parametrize('kin', [True, False])
def test_sinuosity_and_initial_depth_from_cfg(tmp_path, kin):
    """sinu = 1.3 and d0 = 0.05 m read from the CFG: flow lengths * sinu, slope / sinu
    (channels_base.py:1242-1247) and the initial A_wet / P_wet / vol (:1331-1336), then
    120 steps, against the reference fixture (tests/golden/sinu_d0_channels.npz)."""
    from conftest import load_golden
    from topoflow36_b200 import channels_kinematic_wave as ckw, channels_diffusive_wave as cdw, synth
    g = load_golden('sinu_d0_channels.npz')
    tag = 'kw' if kin else 'dw'
    ny, nx = g['code'].shape
    cfg = synth.write_case(str(tmp_path), g['DEM'], g['slope'], g['width'], g['angle'],
                           nval=g['nval'], dt=2.0, kinematic=kin, outlet=(ny - 2, nx // 2),
                           sinu=float(g['sinu']), d0=float(g['d0']))
    c = (ckw if kin else cdw).channels_component()
    c.DIAGNOSTICS = True
    c.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    assert np.array_equal(c.d8.d8_grid, g['code'])
    for n in ('vol', 'd', 'A_wet', 'P_wet'):                 # state right after initialize()
        assert rel_err(getattr(c, n), g[tag + '_init_' + n]) <= RTOL, n
    for k, v in dict(P_rain=z(60 / 3.6e6), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0),
                     ET=z(0), IN=z(0)).items():
        c.set_value(LONG[k], v)
    for i in range(120):
        if i == 80:
            c.set_value(LONG['P_rain'], z(0))
        c.update()
    # slope / 1.3 is no float32 value any more: the fp64 slope stream of the extended kernels
    assert (c.engine.last_path, c.engine.last_kernel) == ('packed', 'ext')
    for n in ('vol', 'd', 'u', 'Q', 'Rh', 'f', 'froude') + (() if kin else ('S_free',)):
        assert rel_err(getattr(c, n), g[tag + '_' + n]) <= RTOL, n
    for s in ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'Q_peak'):
        want = float(g[tag + '_' + s])
        assert abs(float(getattr(c, s)) - want) <= RTOL * abs(want), s
