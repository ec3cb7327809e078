"""Generate the committed fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference) in this container.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.gen_golden            # writes tests/golden/*.npz / *.json

The GPU box has no reference tree; the `-m gpu` tests and the oracle tests read
only these files.  Everything here is deterministic (fixed seeds, numpy 2.3.5).
"""
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from topoflow36_b200 import grid_io, synth    # noqa: E402

GOLD = os.environ.get('TFB_GOLDEN_DIR') or os.path.join(ROOT, 'tests', 'golden')

LONG = {'P_rain': 'atmosphere_water__rainfall_volume_flux',
        'MR': 'glacier_ice__melt_volume_flux',
        'GW': 'land_surface_water__baseflow_volume_flux',
        'ET': 'land_surface_water__evaporation_volume_flux',
        'IN': 'soil_surface_water__infiltration_volume_flux',
        'SM': 'snowpack__melt_volume_flux',
        'rho_H2O': 'water-liquid__mass-per-volume_density'}
GRIDS = ('vol', 'd', 'u', 'Q', 'Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star',
         'froude')
SCALARS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet',
           'f_outlet', 'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak',
           'Td_peak')


def z(v):
    return np.array(v, dtype='float64')


def standalone_channels(cfg_file, kinematic, inputs):
    """Reference channels component driven without EMELI."""
    from topoflow.components import channels_kinematic_wave as ckw
    from topoflow.components import channels_diffusive_wave as cdw
    c = (ckw if kinematic else cdw).channels_component()
    rh.quiet_call(c.initialize, cfg_file=cfg_file, mode='driver', SILENT=True)
    for k, v in inputs.items():
        c.set_value(LONG[k], v)
    c.n_steps = 10 ** 9
    return c


def snapshot(c, diag=True):
    out = {n: np.array(getattr(c, n), dtype='float64') for n in GRIDS}
    if not c.KINEMATIC_WAVE:
        out['S_free'] = np.array(c.S_free)
    for n in SCALARS:
        out[n] = np.float64(getattr(c, n))
    return out


def gen_treynor(work):
    cfg_dir, out_dir = rh.stage_treynor(work)
    topo = os.path.join(os.path.dirname(cfg_dir.rstrip(os.sep)), '__topo')
    info = grid_io.read_rti(os.path.join(topo, 'Treynor.rti'))
    rd = lambda n, t=None: grid_io.read_rtg(os.path.join(topo, 'Treynor_' + n + '.rtg'), info, t)
    rain = np.loadtxt(os.path.join(cfg_dir, 'June_20_67_rain_rates.txt'), dtype='float32')
    inputs = dict(DEM=rd('DEM'), slope=rd('slope'), width=rd('chan-w'), angle=rd('chan-a'),
                  nval=rd('chan-n'), z0val=rd('chan-z0'), flow=rd('flow', 'BYTE'),
                  area=rd('area'), d8_area=rd('d8-area'), rain_mmph=rain,
                  xres=np.float64(info.xres), yres=np.float64(info.yres),
                  outlets_rc=np.array([[34, 16], [32, 16], [37, 19], [26, 14], [30, 15]]))
    np.savez_compressed(os.path.join(GOLD, 'treynor_inputs.npz'), **inputs)

    # ---- (1) full EMELI run: the reference's own integration "test" -----------
    from topoflow.framework import emeli
    f = emeli.framework()
    rh.quiet_call(f.run_model, driver_comp_name='topoflow_driver', cfg_prefix='June_20_67',
                  cfg_directory=cfg_dir, SILENT=True)
    ch = f.comp_set['tf_channels_kin_wave']
    emeli_out = {n: float(getattr(ch, n)) for n in SCALARS}
    emeli_out['n_steps'] = int(ch.time_index)
    emeli_out['vol_chan_sum'] = float(ch.vol_chan_sum)
    final = {n: np.array(getattr(ch, n)) for n in ('vol', 'd', 'u', 'Q')}
    # the files that run wrote (utils/rts_files.py, text_ts_files.py, rti_files.py): RTS
    # stacks in the data set's byte order (Treynor: MSB), outlet series, RTI side-car
    raw = lambda fn: np.frombuffer(open(os.path.join(out_dir, fn), 'rb').read(), dtype=np.uint8)
    final['file_2D_Q_rts'] = raw('June_20_67_2D-Q.rts')
    final['file_2D_d_rts'] = raw('June_20_67_2D-d.rts')
    final['file_0D_Q_txt'] = raw('June_20_67_0D-Q.txt')
    rti = open(os.path.join(out_dir, 'June_20_67_2D-Q.rti')).read().replace(out_dir, '<out_directory>/')
    final['file_2D_Q_rti'] = np.frombuffer(rti.encode(), dtype=np.uint8)
    # ---- (1b) the same run without time interpolation (emeli.py:893): the coupling the
    # GPU drop-in recommends
    f = emeli.framework()
    rh.quiet_call(f.run_model, driver_comp_name='topoflow_driver', cfg_prefix='June_20_67',
                  cfg_directory=cfg_dir, SILENT=True, time_interp_method='None')
    ch = f.comp_set['tf_channels_kin_wave']
    for n in ('vol', 'd', 'u', 'Q'):
        final['none_' + n] = np.array(getattr(ch, n))
    for n in SCALARS + ('vol_chan_sum',):
        final['none_' + n] = np.float64(getattr(ch, n))
    final['none_n_steps'] = np.int64(ch.time_index)
    np.savez_compressed(os.path.join(GOLD, 'treynor_emeli_final.npz'), **final)

    # ---- (2) stand-alone KW run forced by the 1-minute rain series -----------
    # P is held for 10 channel steps (60 s / 6 s) like the met component does
    # with time interpolation off; all fp64 grids captured in memory.
    ny, nx = info.nrows, info.ncols
    c = standalone_channels(cfg_dir + 'June_20_67_channels_kinematic_wave.cfg', True,
                            dict(P_rain=z(0), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0),
                                 ET=np.zeros((ny, nx)), IN=np.zeros((ny, nx))))
    n_steps = 1200
    series = np.zeros((n_steps, 4))
    snaps = {}
    for i in range(n_steps):
        m = i // 10
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        c.set_value(LONG['P_rain'], z(P))
        rh.quiet_call(c.update)
        series[i] = (c.Q_outlet, c.u_outlet, c.d_outlet, c.f_outlet)
        if (i + 1) in (100, 400, 1200):
            for k, v in snapshot(c).items():
                snaps['s%d_%s' % (i + 1, k)] = v
    np.savez_compressed(os.path.join(GOLD, 'treynor_kw_standalone.npz'),
                        outlet_series=series, **snaps)
    return emeli_out


def gen_treynor_dw(work):
    """The June-20-1967 storm on the reference's DIFFUSIVE-wave component (its Treynor data set
    ships a kinematic CFG only: the same file under the diffusive name, with dt = 2 s -- at the
    CFG's 6 s the explicit scheme of channels_diffusive_wave oscillates, outlet peaks of
    1000 m^3/s; at 2 s the hydrograph is smooth, Q_peak 12.33 against 12.44 kinematic, and a
    one-ulp perturbation of Manning's n stays at 1e-15 over the whole run).  3600 steps:
    outlet series every step, all grids at three checkpoints."""
    import re
    cfg_dir, out_dir = rh.stage_treynor(os.path.join(work, 'dw'))
    src = cfg_dir + 'June_20_67_channels_kinematic_wave.cfg'
    dst = cfg_dir + 'June_20_67_channels_diffusive_wave.cfg'
    txt = open(src).read()
    txt2 = re.sub(r'^(dt\s*\|\s*)[0-9.]+', lambda m: m.group(1) + '2.0', txt, flags=re.M)
    assert txt2 != txt
    with open(dst, 'w') as fh:
        fh.write(txt2)
    rain = np.loadtxt(os.path.join(cfg_dir, 'June_20_67_rain_rates.txt'), dtype='float32')
    ny, nx = 44, 29
    c = standalone_channels(dst, False,
                            dict(P_rain=z(0), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0),
                                 ET=np.zeros((ny, nx)), IN=np.zeros((ny, nx))))
    assert c.DIFFUSIVE_WAVE and float(c.dt) == 2.0
    n_steps = 3600
    series = np.zeros((n_steps, 4))
    snaps = {}
    for i in range(n_steps):
        m = i // 30                                # the 1-minute rain value is held for 30 steps
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        c.set_value(LONG['P_rain'], z(P))
        rh.quiet_call(c.update)
        series[i] = (c.Q_outlet, c.u_outlet, c.d_outlet, c.f_outlet)
        if (i + 1) in (600, 1300, 3600):
            for k, v in snapshot(c).items():
                snaps['s%d_%s' % (i + 1, k)] = v
    np.savez_compressed(os.path.join(GOLD, 'treynor_dw_standalone.npz'),
                        outlet_series=series, **snaps)


def gen_synth_channels(work):
    """KW/DW x Manning/law-of-wall on a 61 x 83 synthetic case, 150 steps, rain
    for the first 100; ET and IN given as grids (exercises the ET masking)."""
    import topoflow.components.d8_global as d8g
    ny, nx = 61, 83
    DEM = synth.fractal_dem(ny, nx, seed=7, sigma=1.0)
    width, angle, nval = synth.channel_params(ny, nx, seed=7)
    rng = np.random.default_rng(99)
    angle = np.float32(np.where(rng.random((ny, nx)) < 0.15, 0.0, 20.0 + 40.0 * rng.random((ny, nx))))
    z0val = np.float32(0.005 + 0.02 * rng.random((ny, nx)))
    out = dict(DEM=DEM, width=width, angle=angle, nval=nval, z0val=z0val)
    ETg = 2e-6 * rng.random((ny, nx))
    INg = 3e-6 * rng.random((ny, nx))
    out['ET0'] = ETg.copy()
    out['IN0'] = INg.copy()
    for kin in (True, False):
        for man in (True, False):
            tag = ('kw' if kin else 'dw') + ('_man' if man else '_low')
            cdir = os.path.join(work, 'synth_' + tag)
            # slope grid: the reference's own D8 slope of this DEM
            cfg = synth.write_case(cdir, DEM, np.zeros_like(DEM), width, angle,
                                   nval=nval if man else None,
                                   z0val=None if man else z0val, dt=2.0,
                                   kinematic=kin, manning=man, outlet=(ny - 2, nx // 2))
            d8 = d8g.d8_component()
            rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                          SILENT=True, REPORT=False)
            rh.quiet_call(d8.update, 0)
            d8.update_noflow_IDs()
            d8.update_slope_grid()
            info = grid_io.read_rti(os.path.join(cdir, 'Synth.rti'))
            grid_io.write_rtg(os.path.join(cdir, 'Synth_slope.rtg'), d8.S, info)
            if kin and man:
                out['code'] = np.array(d8.d8_grid)
                out['slope'] = np.float32(d8.S)
                out['ds32'] = np.float32(d8.ds)
                out['dw32'] = np.float32(d8.dw)
                out['area'] = np.float32(d8.A)
            ET = ETg.copy()
            c = standalone_channels(cfg, kin, dict(
                P_rain=z(80 / 3.6e6), MR=z(0), GW=z(1e-7), SM=z(2e-7), rho_H2O=z(1000.0),
                ET=ET, IN=INg.copy()))
            for i in range(150):
                if i == 100:
                    c.set_value(LONG['P_rain'], z(0))
                rh.quiet_call(c.update)
            for k, v in snapshot(c).items():
                out[tag + '_' + k] = v
            out[tag + '_ET_after'] = ET
            out[tag + '_S_bed'] = np.array(c.S_bed)
    np.savez_compressed(os.path.join(GOLD, 'synth_channels.npz'), **out)


def gen_sinu_d0(work):
    """Sinuosity 1.3 (flow lengths * sinu, slope / sinu: channels_base.py:1242-1247) and an
    initial depth of 5 cm (:1331-1336) on the 61 x 83 synthetic case, KW and DW, scalar
    forcing, 120 steps with rain for the first 80."""
    import topoflow.components.d8_global as d8g
    ny, nx = 61, 83
    DEM = synth.fractal_dem(ny, nx, seed=7, sigma=1.0)
    width, angle, nval = synth.channel_params(ny, nx, seed=7)
    out = dict(DEM=DEM, width=width, angle=angle, nval=nval, sinu=np.float64(1.3),
               d0=np.float64(0.05))
    for kin in (True, False):
        tag = 'kw' if kin else 'dw'
        cdir = os.path.join(work, 'sinu_' + tag)
        cfg = synth.write_case(cdir, DEM, np.zeros_like(DEM), width, angle, nval=nval, dt=2.0,
                               kinematic=kin, outlet=(ny - 2, nx // 2), sinu=1.3, d0=0.05)
        d8 = d8g.d8_component()
        rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                      SILENT=True, REPORT=False)
        rh.quiet_call(d8.update, 0)
        d8.update_noflow_IDs()
        d8.update_slope_grid()
        info = grid_io.read_rti(os.path.join(cdir, 'Synth.rti'))
        grid_io.write_rtg(os.path.join(cdir, 'Synth_slope.rtg'), d8.S, info)
        if kin:
            out['code'] = np.array(d8.d8_grid)
            out['slope'] = np.float32(d8.S)
            out['ds32'] = np.float32(d8.ds)
            out['dw32'] = np.float32(d8.dw)
        c = standalone_channels(cfg, kin, dict(
            P_rain=z(60 / 3.6e6), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0), ET=z(0), IN=z(0)))
        for k, v in snapshot(c).items():
            out[tag + '_init_' + k] = v
        for i in range(120):
            if i == 80:
                c.set_value(LONG['P_rain'], z(0))
            rh.quiet_call(c.update)
        for k, v in snapshot(c).items():
            out[tag + '_' + k] = v
        out[tag + '_S_bed'] = np.array(c.S_bed)
    np.savez_compressed(os.path.join(GOLD, 'sinu_d0_channels.npz'), **out)


def gen_flood(work):
    """FLOOD_OPTION (rectangle-over-trapezoid overbank flow) on the 61 x 83 synthetic
    case with shallow banks, KW and DW, 150 steps of heavy rain then recession."""
    import topoflow.components.d8_global as d8g
    ny, nx = 61, 83
    DEM = synth.fractal_dem(ny, nx, seed=7, sigma=1.0)
    width, angle, nval = synth.channel_params(ny, nx, seed=7)
    rng = np.random.default_rng(99)
    angle = np.float32(np.where(rng.random((ny, nx)) < 0.15, 0.0, 20.0 + 40.0 * rng.random((ny, nx))))
    d_bank = np.float32(0.02 + 0.05 * rng.random((ny, nx)))
    out = dict(DEM=DEM, width=width, angle=angle, nval=nval, d_bankfull=d_bank)
    for kin in (True, False):
        tag = 'kw' if kin else 'dw'
        cdir = os.path.join(work, 'flood_' + tag)
        cfg = synth.write_case(cdir, DEM, np.zeros_like(DEM), width, angle, nval=nval, dt=2.0,
                               kinematic=kin, outlet=(ny - 2, nx // 2), flood=True,
                               d_bankfull=d_bank)
        d8 = d8g.d8_component()
        rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                      SILENT=True, REPORT=False)
        rh.quiet_call(d8.update, 0)
        d8.update_noflow_IDs()
        d8.update_slope_grid()
        info = grid_io.read_rti(os.path.join(cdir, 'Synth.rti'))
        grid_io.write_rtg(os.path.join(cdir, 'Synth_slope.rtg'), d8.S, info)
        if kin:
            out['code'] = np.array(d8.d8_grid)
            out['slope'] = np.float32(d8.S)
        c = standalone_channels(cfg, kin, dict(
            P_rain=z(150 / 3.6e6), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0),
            ET=z(0), IN=z(0)))
        assert c.FLOOD_OPTION
        for i in range(150):
            if i == 100:
                c.set_value(LONG['P_rain'], z(0))
            rh.quiet_call(c.update)
            if i == 99:
                out[tag + '_dflood_max_at_100'] = np.float64(c.d_flood.max())
        for k, v in snapshot(c).items():
            out[tag + '_' + k] = v
        for k in ('d_flood', 'vol_flood', 'vol_chan', 'Qc', 'Qf'):
            out[tag + '_' + k] = np.array(getattr(c, k), dtype='float64')
    np.savez_compressed(os.path.join(GOLD, 'flood_channels.npz'), **out)


def gen_attenuate(work):
    """ATTENUATE (per-cell linear reservoir in front of the channel, :2077-2084) on the
    61 x 83 synthetic case, KW and DW, rain for 100 of 150 steps."""
    import topoflow.components.d8_global as d8g
    g0 = dict(np.load(os.path.join(GOLD, 'flood_channels.npz')))
    DEM, width, angle, nval = g0['DEM'], g0['width'], g0['angle'], g0['nval']
    ny, nx = DEM.shape
    out = {}
    for kin in (True, False):
        tag = 'kw' if kin else 'dw'
        cdir = os.path.join(work, 'atten_' + tag)
        cfg = synth.write_case(cdir, DEM, g0['slope'], width, angle, nval=nval, dt=2.0,
                               kinematic=kin, outlet=(ny - 2, nx // 2), attenuate_days=0.002)
        c = standalone_channels(cfg, kin, dict(
            P_rain=z(150 / 3.6e6), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0),
            ET=z(0), IN=z(0)))
        assert c.ATTENUATE
        for i in range(150):
            if i == 100:
                c.set_value(LONG['P_rain'], z(0))
            rh.quiet_call(c.update)
        for k, v in snapshot(c).items():
            out[tag + '_' + k] = v
        out[tag + '_vol_stored'] = np.array(c.vol_stored, dtype='float64')
    np.savez_compressed(os.path.join(GOLD, 'attenuate_channels.npz'), **out)


def gen_geographic(work):
    """Fixed-angle (geographic) pixels, 1 arcsec at 41-42 N: per-row dx / dd / ds / dw and
    a per-cell area grid da (utils/pixels.py:71-233).  D8 + KW and DW routing."""
    import topoflow.components.d8_global as d8g
    ny, nx = 75, 64
    DEM = synth.fractal_dem(ny, nx, seed=31, sigma=1.0, dx=23.0, dy=31.0)
    width, angle, nval = synth.channel_params(ny, nx, seed=31)
    out = dict(DEM=DEM, width=width, angle=angle, nval=nval)
    for kin in (True, False):
        tag = 'kw' if kin else 'dw'
        cdir = os.path.join(work, 'geo_' + tag)
        cfg = synth.write_case(cdir, DEM, np.zeros_like(DEM), width, angle, nval=nval, dx=1.0, dy=1.0,
                               dt=2.0, kinematic=kin, outlet=(ny - 2, nx // 2),
                               geographic=(41.0, -95.7))
        d8 = d8g.d8_component()
        rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                      SILENT=True, REPORT=False)
        rh.quiet_call(d8.update, 0)
        d8.update_noflow_IDs()
        d8.update_slope_grid()
        info = grid_io.read_rti(os.path.join(cdir, 'Synth.rti'))
        grid_io.write_rtg(os.path.join(cdir, 'Synth_slope.rtg'), d8.S, info)
        if kin:
            out.update(code=np.array(d8.d8_grid), slope=np.float32(d8.S), ds32=np.float32(d8.ds),
                       dw32=np.float32(d8.dw), area=np.float32(d8.A), dx=np.float32(d8.dx),
                       dy=np.float32(d8.dy), dd=np.float32(d8.dd), da=np.float64(d8.da))
        c = standalone_channels(cfg, kin, dict(
            P_rain=z(100 / 3.6e6), MR=z(0), GW=z(0), SM=z(0), rho_H2O=z(1000.0), ET=z(0), IN=z(0)))
        assert np.ndim(c.da) == 2
        for i in range(120):
            if i == 80:
                c.set_value(LONG['P_rain'], z(0))
            rh.quiet_call(c.update)
        for k, v in snapshot(c).items():
            out[tag + '_' + k] = v
    np.savez_compressed(os.path.join(GOLD, 'geo_channels.npz'), **out)


def gen_d8(work):
    """D8 toolkit on a DEM with plenty of ties and flats (elevations quantised
    to 0.25 m) and a few nodata cells; LINK_FLATS on and off."""
    import topoflow.components.d8_global as d8g
    ny, nx = 150, 131
    DEM = synth.fractal_dem(ny, nx, seed=21, sigma=1.5, sy=0.004, sx=0.002)
    DEM = np.float32(np.round(DEM * 4.0) / 4.0)
    DEM[40:43, 50:53] = -9999.0
    DEM[100, 20] = np.nan
    out = dict(DEM=DEM)
    w, a, n = synth.channel_params(ny, nx, seed=21)
    for lf in (1, 0):
        cdir = os.path.join(work, 'd8_lf%d' % lf)
        synth.write_case(cdir, DEM, np.zeros_like(DEM), w, a, nval=n, link_flats=lf)
        d8 = d8g.d8_component()
        rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                      SILENT=True, REPORT=False)
        # intermediate stages of update_flow_grid (d8_global.py:99-146)
        rh.quiet_call(d8.start_new_d8_codes, None)
        out['lf%d_start' % lf] = np.array(d8.d8_grid, dtype='int16', copy=True)
        d8.break_flow_grid_ties()
        out['lf%d_ties' % lf] = np.array(d8.d8_grid, dtype='int16', copy=True)
        rh.quiet_call(d8.update, 0)
        d8.update_noflow_IDs()
        with np.errstate(all='ignore'):
            d8.update_slope_grid()
        out['lf%d_code' % lf] = np.uint8(d8.d8_grid)
        out['lf%d_area' % lf] = np.float32(d8.A)
        out['lf%d_ds' % lf] = np.float32(d8.ds)
        out['lf%d_dw' % lf] = np.float32(d8.dw)
        out['lf%d_slope' % lf] = np.float32(d8.S)
        out['lf%d_pid' % lf] = np.int32(d8.parent_ID_grid)
    d8.get_resolve_array()
    out['resolve'] = np.uint8(d8.resolve)
    np.savez_compressed(os.path.join(GOLD, 'd8_synth.npz'), **out)


def gen_sources():
    """Pin the per-cell source formulas by calling the reference METHODS on bare
    component objects whose attributes are set by hand (no CFG machinery)."""
    from topoflow.components import infil_green_ampt, snow_degree_day, \
        evap_priestley_taylor, met_base
    rng = np.random.default_rng(5)
    ny, nx = 37, 41
    g = lambda lo, hi: lo + (hi - lo) * rng.random((ny, nx))
    out = {}
    # ---- Green-Ampt, 3 consecutive updates
    c = infil_green_ampt.infil_component()
    c.DEBUG = False
    c.RICHARDS = False
    c.dt = z(5.0)
    c.P_rain = g(0, 2e-5)
    c.SM = g(0, 2e-6)
    c.Ks = [g(1e-6, 9e-6)]
    c.Ki = [g(1e-8, 9e-7)]
    c.qs = [g(0.4, 0.5)]
    c.qi = [g(0.2, 0.39)]
    c.G = [g(0.1, 0.9)]
    c.I = np.zeros((ny, nx)) + 1e-6
    c.h_table = g(0, 10)
    c.elev = g(5, 20)
    for k in ('P_rain', 'SM', 'h_table', 'elev'):
        out['ga_' + k] = np.array(getattr(c, k))
    for k in ('Ks', 'Ki', 'qs', 'qi', 'G'):
        out['ga_' + k] = np.array(getattr(c, k)[0])
    for it in range(3):
        c.update_surface_influx()
        c.update_infil_rate()
        c.adjust_infil_rate()
        c.update_I()
        out['ga_IN%d' % it] = np.array(c.IN)
        out['ga_I%d' % it] = np.array(c.I)
    # ---- degree-day, 3 updates
    s = snow_degree_day.snow_component()
    s.dt = z(3600.0)
    s.T_air = g(-3, 6)
    s.c0 = z(2.7)
    s.T0 = z(-0.2)
    s.P_snow = g(0, 1e-6)
    s.h_swe = g(0, 0.002)
    s.h_snow = s.h_swe * (1000.0 / 300.0)
    s.SM = np.zeros((ny, nx))
    s.rho_H2O = z(1000.0)
    s.rho_snow = z(300.0)
    s.rho_snow_type = 'Scalar'
    s.density_ratio = (s.rho_H2O / s.rho_snow)
    out['dd_T_air'] = np.array(s.T_air)
    out['dd_P_snow'] = np.array(s.P_snow)
    out['dd_h_swe_init'] = np.array(s.h_swe)
    for it in range(3):
        s.update_meltrate()
        s.enforce_max_meltrate()
        s.update_swe()
        s.update_depth()
        out['dd_SM%d' % it] = np.array(s.SM)
        out['dd_h_swe%d' % it] = np.array(s.h_swe)
        out['dd_h_snow%d' % it] = np.array(s.h_snow)
    # ---- Priestley-Taylor
    e = evap_priestley_taylor.evap_component()
    e.T_air = g(-5, 30)
    e.T_surf = g(-5, 35)
    e.Qn_SW = g(0, 800)
    e.Qn_LW = g(-120, 50)
    e.alpha = z(1.26)
    e.K_soil = z(0.45)
    e.soil_x = z(0.05)
    e.T_soil_x = z(12.0)
    e.Qc = np.zeros((ny, nx))
    e.ET = np.zeros((ny, nx))
    e.update_Qc()
    e.update_ET_rate()
    for k in ('T_air', 'T_surf', 'Qn_SW', 'Qn_LW', 'Qc', 'ET'):
        out['pt_' + k] = np.array(getattr(e, k))
    # ---- met rain/snow split
    m = met_base.met_component()
    m.DEBUG = False
    m.P = g(0, 3e-5)
    m.T_air = g(-4, 4)
    m.T_air_type = 'Grid'
    m.T_rain_snow = z(0.5)
    m.P_rain = np.zeros((ny, nx))
    m.P_snow = np.zeros((ny, nx))
    m.update_P_rain()
    m.update_P_snow()
    for k in ('P', 'T_air', 'P_rain', 'P_snow'):
        out['met_' + k] = np.array(getattr(m, k))
    np.savez_compressed(os.path.join(GOLD, 'sources.npz'), **out)


def gen_met_energy():
    """Pin the meteorology energy chain (met_base.update :790-809, PRECIP_ONLY = No)
    by calling the reference METHODS, in the reference's order, on a bare component
    whose attributes are set by hand.  Albedo is held fixed (albedo_type that skips
    the 'aging' branch); julian_day / TSN_offset come from the reference's own
    solar functions for 1967-06-20 14:30 local time."""
    from topoflow.components import met_base, solar_funcs as solar
    rng = np.random.default_rng(17)
    ny, nx = 33, 47
    g = lambda lo, hi: lo + (hi - lo) * rng.random((ny, nx))
    out = {}
    for case, hour in (('day', 14.5), ('dusk', 19.6), ('night', 2.0)):
        m = met_base.met_component()
        m.DEBUG = False
        m.set_constants()
        m.SATTERLUND = False
        m.T_air, m.T_surf = g(-6, 28), g(-6, 30)
        m.RH = g(0.2, 0.98)
        m.uz = g(0.5, 9.0)
        m.z, m.z0_air = z(10.0), g(0.005, 0.04)
        m.h_snow = np.where(rng.random((ny, nx)) < 0.3, g(0, 0.6), 0.0)
        m.h_ice = np.where(rng.random((ny, nx)) < 0.1, 0.2, 0.0)
        m.p0 = g(880, 1020)
        m.albedo, m.albedo_type = g(0.1, 0.8), 'Grid_Sequence'
        m.T_surf_type = 'Grid'
        m.em_surf = g(0.9, 0.99)
        m.dust_atten, m.cloud_factor, m.canopy_factor = z(0.08), g(0, 1), g(0, 0.7)
        m.lat_deg = g(41.0, 41.6)
        m.lon_deg = g(-95.8, -95.2)
        m.alpha = g(0, 2 * np.pi)            # aspect [rad]
        m.beta = g(0, 0.6)                   # slope angle [rad]
        m.GMT_offset = np.int16(-6)
        m.Qn_SW, m.Qn_LW = np.zeros((ny, nx)), np.zeros((ny, nx))
        m.Q_sum = np.zeros((ny, nx))
        m.Qa = m.Qc = z(0)
        m.julian_day = solar.Julian_Day(6, 20, hour_num=hour, year=1967)
        dec_part = m.julian_day - np.int16(m.julian_day)
        clock_hour = dec_part * m.hours_per_day
        solar_noon = solar.True_Solar_Noon(m.julian_day, m.lon_deg, m.GMT_offset,
                                           DST_offset=None, year=1967)
        m.TSN_offset = (clock_hour - solar_noon)
        # the component's own date arithmetic (update_julian_day :1938-1993), kept
        # separately: it drops the minutes of the hour
        m.start_datetime, m.time_min = '1967-06-20 00:00:00', np.float64(hour * 60.0)
        jd_keep, tsn_keep = m.julian_day, m.TSN_offset
        m.update_julian_day()
        out[case + '_ujd_julian_day'] = np.float64(m.julian_day)
        out[case + '_ujd_TSN_offset'] = np.array(m.TSN_offset, dtype='float64')
        m.julian_day, m.TSN_offset = jd_keep, tsn_keep
        for k in ('T_air', 'RH', 'uz', 'z0_air', 'h_snow', 'h_ice', 'p0', 'albedo', 'em_surf',
                  'cloud_factor', 'canopy_factor', 'lat_deg', 'lon_deg', 'alpha', 'beta',
                  'TSN_offset'):
            out[case + '_in_' + k] = np.array(getattr(m, k), dtype='float64')
        out[case + '_in_julian_day'] = np.float64(m.julian_day)
        out[case + '_in_hour'] = np.float64(hour)
        out['current_year'] = np.int64(solar.Current_Year())
        m.update_saturation_vapor_pressure(MBAR=True)
        m.update_vapor_pressure()
        m.update_dew_point()
        m.update_T_surf()
        m.update_saturation_vapor_pressure(MBAR=True, SURFACE=True)
        m.update_bulk_richardson_number()
        m.update_bulk_aero_conductance()
        m.update_sensible_heat_flux()
        m.update_precipitable_water_content()
        m.update_vapor_pressure(SURFACE=True)
        m.update_latent_heat_flux()
        m.update_albedo(method='aging')
        m.update_net_shortwave_radiation()
        m.update_em_air()
        m.update_net_longwave_radiation()
        m.update_net_energy_flux()
        for k in ('e_sat_air', 'e_air', 'T_dew', 'T_surf', 'e_sat_surf', 'Ri', 'Dn', 'Dh', 'Qh',
                  'W_p', 'e_surf', 'Qe', 'Qn_SW', 'em_air', 'Qn_LW', 'Q_sum'):
            out[case + '_' + k] = np.array(getattr(m, k), dtype='float64')
    np.savez_compressed(os.path.join(GOLD, 'met_energy.npz'), **out)


def gen_albedo():
    """met_base.update_albedo (:1995-2045), 'aging' and 'simple', called on a bare
    meteorology component over 30 six-hour steps with two snowfalls; the ring-buffer
    state is set up as initialize() does (:1121-1133)."""
    from topoflow.components import met_base
    ny, nx, dt = 9, 13, 21600.0
    rng = np.random.default_rng(77)
    out = {}
    for method in ('aging', 'simple'):
        m = met_base.met_component()
        m.dt = dt
        m.albedo_type = 'Grid'
        m.albedo = rng.random((ny, nx)) * 0.5 + 0.2
        m.rho_H2O, m.rho_snow = np.float64(1000.0), np.float64(250.0)
        m.n = 0
        m.days_per_dt = m.dt / 86400
        m.P_snow_3day_grid = np.zeros((int(np.int64(259200 / m.dt)), ny, nx))
        out[method + '_albedo0'] = m.albedo.copy()
        series = {k: [] for k in ('P_snow', 'T_air', 'h_snow', 'h_ice', 'albedo', 'n')}
        for step in range(30):
            snowing = step in (2, 3, 27)
            m.P_snow = (rng.random((ny, nx)) * (4e-6 if snowing else 2e-8))
            m.T_air = rng.random((ny, nx)) * 8.0 - 5.0
            m.h_snow = np.where(rng.random((ny, nx)) < 0.7, rng.random((ny, nx)) * 0.4, 0.0)
            m.h_ice = np.where(rng.random((ny, nx)) < 0.3, 0.5, 0.0)
            m.update_albedo(method=method)
            for k in ('P_snow', 'T_air', 'h_snow', 'h_ice', 'albedo'):
                series[k].append(np.array(getattr(m, k), dtype='float64'))
            series['n'].append(np.zeros((ny, nx)) + m.n)
        for k, v in series.items():
            out[method + '_' + k] = np.stack(v)
    np.savez_compressed(os.path.join(GOLD, 'albedo.npz'), **out)


def gen_prep(work):
    """DEM / geometry preparation: the reference's fill_pits() on DEMs
    with many depressions, nodata holes, NaNs and a closed-basin code; its
    get_new_slope_grid() and get_grid_from_TCA() run from files."""
    from topoflow.utils import fill_pits as fp, new_slopes as ns, parameterize as pz
    from topoflow36_b200 import grid_io
    import time as _time
    import types
    # fill_pits.py:511 calls time.sleep(float32(0.001)), a TypeError on Python 3.12:
    # give the module a `time` whose sleep() takes anything (nothing else is touched)
    fp.time = types.SimpleNamespace(time=_time.time, sleep=lambda s: None)
    out = {}
    # 1. depression filling: rough DEM without regional tilt -> depressions of all sizes
    for tag, (ny, nx), seed in (('a', (97, 131), 5), ('b', (160, 75), 6)):
        DEM = synth.fractal_dem(ny, nx, seed=seed, sigma=3.0, sy=0.0005, sx=0.0002)
        DEM = np.float32(np.round(DEM * 8.0) / 8.0)
        DEM[30:34, 40:47] = -9999.0
        DEM[0, 5:9] = -9999.0                    # nodata on the border
        DEM[1, 20] = -9999.0                     # on the interior's rim (wrapped rolls)
        DEM[60, 10] = np.nan
        DEM[70, 50] = np.inf
        DEM[50, 60] = -32000.0                   # closed-basin code
        if tag == 'b':
            DEM[ny - 2, 30:33] = -9999.0
            DEM[100:103, nx - 2] = -9999.0
        out['fill_%s_in' % tag] = DEM.copy()
        D2 = DEM.copy()
        rh.quiet_call(fp.fill_pits, D2, 'FLOAT', nx, ny, SILENT=True)
        out['fill_%s_out' % tag] = D2
    # 2. path-length slopes and power-law grids, from files
    ny, nx = 150, 131
    DEM = synth.fractal_dem(ny, nx, seed=21, sigma=1.5, sy=0.004, sx=0.002)
    DEM = np.float32(np.round(DEM * 4.0) / 4.0)
    DEM[40:43, 50:53] = -9999.0
    w, a, n = synth.channel_params(ny, nx, seed=21)
    cdir = os.path.join(work, 'prep')
    synth.write_case(cdir, DEM, np.zeros_like(DEM), w, a, nval=n, link_flats=1)
    cfg_dir = cdir + os.sep
    rh.quiet_call(ns.get_new_slope_grid, site_prefix='Synth', case_prefix='Case1',
                  cfg_dir=cfg_dir, slope_file='Synth_newslope.rtg')
    info = grid_io.read_rti(os.path.join(cdir, 'Synth.rti'))
    out['slope_DEM'] = DEM
    out['slope_new'] = grid_io.read_rtg(os.path.join(cdir, 'Synth_newslope.rtg'), info)
    import topoflow.components.d8_global as d8g
    d8 = d8g.d8_component()
    rh.quiet_call(d8.initialize, cfg_file=os.path.join(cdir, 'Case1_d8_global.cfg'),
                  SILENT=True, REPORT=False)
    rh.quiet_call(d8.update, 0)
    out['slope_code'] = np.uint8(d8.d8_grid)
    d8.update_aspect_grid()
    out['aspect'] = np.float32(d8.aspect)
    area = np.float32(d8.A)
    grid_io.write_rtg(os.path.join(cdir, 'Synth_area.rtg'), area, info)
    out['tca_area'] = area
    cases = {'w': dict(g1=25.0, g2=1.5),                        # min/max of A (float64 p)
             'n': dict(g1=0.03, p=-0.2),                        # p given (float32 branch)
             'd': dict(A1=50.0, g1=3.0, A2=0.01, g2=0.2)}       # two points
    for tag, kw in cases.items():
        rh.quiet_call(pz.get_grid_from_TCA, site_prefix='Synth', topo_dir=cfg_dir,
                      area_file='Synth_area.rtg', out_file='Synth_tca_%s.rtg' % tag,
                      REPORT=False, **kw)
        out['tca_' + tag] = grid_io.read_rtg(os.path.join(cdir, 'Synth_tca_%s.rtg' % tag), info)
    np.savez_compressed(os.path.join(GOLD, 'prep.npz'), **out)


def main():
    rh.import_reference()
    os.makedirs(GOLD, exist_ok=True)
    work = tempfile.mkdtemp(prefix='tf_golden_')
    try:
        meta = {'numpy': np.__version__, 'reference': 'peckhams/topoflow36 @ /root/reference'}
        meta['treynor_emeli'] = gen_treynor(work)
        gen_treynor_dw(work)
        gen_synth_channels(work)
        gen_sinu_d0(work)
        gen_flood(work)
        gen_attenuate(work)
        gen_geographic(work)
        gen_d8(work)
        gen_prep(work)
        gen_sources()
        gen_met_energy()
        gen_albedo()
        with open(os.path.join(GOLD, 'meta.json'), 'w') as fh:
            json.dump(meta, fh, indent=1, sort_keys=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    for f in sorted(os.listdir(GOLD)):
        print('%10d  %s' % (os.path.getsize(os.path.join(GOLD, f)), f))


if __name__ == '__main__':
    main()
