"""GPU parity of the fused routing kernel (through the C ABI) against the numpy
oracle and the committed reference fixtures.

Tolerance: the north-star asks for rel 1e-9 on Q / depth / velocity over a
storm.  Everything except x**(2/3), x**(1/3) and log() is IEEE-exact and in the
reference's order, so the observed error is ~1e-15; the bound asserted here is
RTOL = 1e-11 (two orders inside the spec) on every fp64 grid, and exact
equality where no libm call is on the data path."""
import numpy as np
import pytest

from conftest import cell_rel_err, rel_err
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-11
GRIDS = ('vol', 'd', 'u', 'Q', 'Rh', 'A_wet', 'P_wet', 'f', 'tau', 'u_star', 'froude')
SCALARS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
           'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak')


def _engine(code, dx, dy, dd, slope, width, angle_deg, rough, **kw):
    import torch  # noqa: F401
    from topoflow36_b200.routing import ChannelsEngine
    ny, nx = code.shape
    angle = np.float64(angle_deg) * o.DEG_TO_RAD
    S = np.array(slope, dtype='float64')
    if kw.get('kinematic', True):
        o.remove_bad_slopes(S)
    width = np.array(width, dtype='float64')
    if width.ndim == 2:
        _, dw = o.d8_ds_dw(code, dx, dy, dd)
        w0 = width == 0
        width[w0] = dw[w0]
    return ChannelsEngine(ny, nx, code=code, dx=dx, dy=dy, dd=dd, S_bed=S, width=width,
                          angle=angle, rough=np.float64(rough), **kw)


def _cmp_grids(eng, ch, names, rtol=RTOL):
    for n in names:
        got = (eng.Q if n == 'Q' else eng.vol if n == 'vol' else eng.own(getattr(eng, n))).cpu().numpy()
        want = getattr(ch, n) if not isinstance(ch, dict) else ch[n]
        e = rel_err(got, want)
        assert e <= rtol, '%s rel err %.3e' % (n, e)


def _cmp_scalars(eng, ch, rtol=RTOL):
    s = eng.read_scalars()
    for n in SCALARS:
        want = float(getattr(ch, n)) if not isinstance(ch, dict) else float(ch[n])
        got = getattr(s, n)
        assert abs(got - want) <= rtol * max(abs(want), 1e-300) + 0.0, \
            '%s: got %r want %r' % (n, got, want)
    return s


def test_treynor_storm_vs_reference_fixture(treynor, treynor_kw):
    """June-20-1967 storm, 1200 steps: hydrograph at the outlet every step and
    all grids at three checkpoints against the stand-alone REFERENCE run."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    eng = _engine(treynor['flow'], dx, dy, dd, treynor['slope'], treynor['width'],
                  treynor['angle'], treynor['nval'], dt=6.0, da=900.0, outlet=(34, 16),
                  diag=True)
    ET = np.zeros(DEM.shape)
    eng.set_input('ET', ET)
    eng.set_input('IN', np.zeros(DEM.shape))
    rain = treynor['rain_mmph']
    series = treynor_kw['outlet_series']
    worst = 0.0
    for i in range(1200):
        m = i // 10
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        eng.set_input('P_rain', P)
        eng.step(time_min=(i * 6.0) / 60.0)
        if i % 25 == 24 or (i + 1) in (100, 400, 1200):
            s = eng.read_scalars()
            got = np.array([s.Q_outlet, s.u_outlet, s.d_outlet, s.f_outlet])
            worst = max(worst, rel_err(got, series[i]))
        if (i + 1) in (100, 400, 1200):
            snap = {n: treynor_kw['s%d_%s' % (i + 1, n)] for n in GRIDS + SCALARS}
            _cmp_grids(eng, snap, GRIDS)
            _cmp_scalars(eng, snap)
    assert worst <= RTOL, worst
    # grid ET / IN, all diagnostics, nx = 29: the extended TMA-ring kernel, not the generic one
    assert eng.last_path == 'packed' and eng.last_kernel == 'ext'
    s = eng.read_scalars()
    assert s.n_neg_d == 0 and s.n_inf_d == 0 and s.n_neg_u == 0 and s.n_nan_u == 0


def test_treynor_storm_on_the_packed_kernel(treynor, treynor_kw):
    """The same 1200-step storm on the BENCHMARKED kernel (channels_packed_kernel, TMA
    ring): nx = 29 rides on pitched rows (include/tfb.h `pitch`), forcing is scalar as
    in the EMELI run.  Hydrograph every step, state grids at three checkpoints --
    also per cell (cell_rel_err <= 1e-9, the north-star tolerance)."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    eng = _engine(treynor['flow'], dx, dy, dd, treynor['slope'], treynor['width'],
                  treynor['angle'], treynor['nval'], dt=6.0, da=900.0, outlet=(34, 16))
    rain = treynor['rain_mmph']
    series = treynor_kw['outlet_series']
    worst = worst_cell = 0.0
    for i in range(1200):
        m = i // 10
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        eng.set_input('P_rain', P)
        eng.step(time_min=(i * 6.0) / 60.0)
        assert eng.last_path == 'packed'
        s = eng.read_scalars()
        got = np.array([s.Q_outlet, s.u_outlet, s.d_outlet, s.f_outlet])
        worst = max(worst, rel_err(got, series[i]))
        if (i + 1) in (100, 400, 1200):
            snap = {n: treynor_kw['s%d_%s' % (i + 1, n)] for n in GRIDS + SCALARS}
            _cmp_grids(eng, snap, ('vol', 'd', 'u', 'Q'))
            _cmp_scalars(eng, snap)
            for n in ('vol', 'd', 'u', 'Q'):
                got = (eng.Q if n == 'Q' else eng.vol if n == 'vol' else eng.own(getattr(eng, n)))
                worst_cell = max(worst_cell, cell_rel_err(got.cpu().numpy(), snap[n]))
    assert worst <= RTOL, worst
    assert worst_cell <= 1e-9, worst_cell
    s = eng.read_scalars()
    assert s.n_neg_d == 0 and s.n_inf_d == 0 and s.n_neg_u == 0 and s.n_nan_u == 0


@pytest.mark.parametrize('kernels', ['lean', 'ext'])
def test_treynor_dw_storm_vs_reference_fixture(treynor, treynor_dw, kernels):
    """The June-20-1967 storm on the DIFFUSIVE wave against the stand-alone run of the
    reference's channels_diffusive_wave (dt = 2 s, 3600 steps: rise, peak 12.33 m^3/s, two hours
    of recession; tests/golden/treynor_dw_standalone.npz): outlet values every 30 steps, all
    grids at three checkpoints, also per cell.  'lean': scalar forcing -> pass A / pass B of the
    production kernels on pitched rows (nx = 29); 'ext': grid ET / IN and all diagnostics
    (incl. S_free) -> the extended kernels."""
    DEM = treynor['DEM']
    dx, dy, dd = o.pixel_dims(DEM.shape[0], treynor['xres'], treynor['yres'])
    ext = (kernels == 'ext')
    eng = _engine(treynor['flow'], dx, dy, dd, treynor['slope'], treynor['width'],
                  treynor['angle'], treynor['nval'], dt=2.0, da=900.0, outlet=(34, 16),
                  kinematic=False, diag=ext)
    if ext:
        eng.set_input('ET', np.zeros(DEM.shape))
        eng.set_input('IN', np.zeros(DEM.shape))
    rain = treynor['rain_mmph']
    series = treynor_dw['outlet_series']
    worst = worst_cell = 0.0
    for i in range(3600):
        m = i // 30
        P = np.float64(rain[m]) / (1000.0 * 3600.0) if m < rain.size else 0.0
        eng.set_input('P_rain', P)
        eng.step(time_min=(i * 2.0) / 60.0)
        if i % 30 == 29 or (i + 1) in (600, 1300, 3600):
            s = eng.read_scalars()
            got = np.array([s.Q_outlet, s.u_outlet, s.d_outlet, s.f_outlet])
            worst = max(worst, rel_err(got, series[i]))
        if (i + 1) in (600, 1300, 3600):
            snap = {n: treynor_dw['s%d_%s' % (i + 1, n)] for n in GRIDS + SCALARS + ('S_free',)}
            names = GRIDS if ext else ('vol', 'd', 'u', 'Q')
            _cmp_grids(eng, snap, names)
            if ext:
                assert rel_err(eng.own(eng.S_free).cpu().numpy(), snap['S_free']) <= RTOL
            _cmp_scalars(eng, snap)
            for n in ('vol', 'd', 'u', 'Q'):
                got = (eng.Q if n == 'Q' else eng.vol if n == 'vol' else eng.own(getattr(eng, n)))
                worst_cell = max(worst_cell, cell_rel_err(got.cpu().numpy(), snap[n]))
    assert worst <= RTOL, worst
    assert worst_cell <= 1e-9, worst_cell
    assert (eng.last_path, eng.last_kernel) == ('packed', 'ext' if ext else 'lean')
    s = eng.read_scalars()
    assert s.n_neg_d == 0 and s.n_inf_d == 0 and s.n_neg_u == 0 and s.n_nan_u == 0
    assert 12.0 < s.Q_peak < 12.6


@pytest.mark.parametrize('variant', ['kw_ga', 'dw_full'])
def test_packed_storm_rise_and_recession(variant):
    """A whole storm on the benchmarked kernels at 512^2: 500 steps of 50 mm/h rain on
    pre-wetted soil, then 1000 steps of recession (1500 steps, dt = 2 s).  kw_ga = the
    bench default (kinematic wave + fused Green-Ampt) against the numpy restatement of
    the reference path; dw_full = diffusive wave + all fused sources (configs[3]) against
    the C restatement.  Grids are compared at four checkpoints, by grid maximum
    (<= 1e-11) and per cell (<= 1e-9 with the floor of conftest.cell_rel_err).
    Per cell, the diffusive wave gets 1e-6 (measured: 2e-8 in Q, 1.3e-8 in vol after 1500
    steps): u ~ sqrt|S_free| with S_free = S_bed + (d - d_parent)/ds (:2606), so where the
    water surface is almost level the last-bit differences of the packed path's depths
    (reciprocal multiplications, include/tfb.h) are amplified by S_bed/|S_free| and fed back
    into the volumes of nearly balanced cells -- ill-conditioned in the reference's own
    formula, invisible in the grid-maximum norm and in the outlet hydrograph (<= 1e-11)."""
    import os
    from oracle import c_oracle
    from topoflow36_b200 import cases
    n = 512
    kin = variant == 'kw_ga'
    tile = cases.noise_tile(n, n, seed=7)
    DEM = cases.slab_dem(n, n, 0, n, 0, tile)
    dx, dy, dd = o.pixel_dims(n, 30.0, 30.0)
    code, _ = c_oracle.d8_flow_grid(DEM, dx, dy, dd)
    ds32, dw32 = o.d8_ds_dw(code, dx, dy, dd)
    with np.errstate(all='ignore'):
        slope = np.float32(o.d8_slope(DEM, code, ds32))
    width, angle, nval = cases.slab_params(n, n, 0, 0, seed=7)
    outlet = tuple(int(v) for v in np.unravel_index(
        np.argmax(c_oracle.d8_area(code, 900.0 / 1e6, with_counts=True)[1]), code.shape))
    eng = _engine(code, dx, dy, dd, slope, width, angle, nval, dt=2.0, da=900.0,
                  kinematic=kin, outlet=outlet)
    eng.set_input('P_rain', cases.RAIN_50MMPH)
    if kin:
        ch = o.Channels(code, ds32, dw32, slope, width, angle, nval=nval, da=900.0, dt=2.0,
                        outlet=outlet)
        ch.P_rain = np.array(cases.RAIN_50MMPH)
        eng.enable_fused_sources(infil=True, I0=cases.I0_WET)
        I = np.zeros((n, n)) + cases.I0_WET
    else:
        ch = c_oracle.Channels(code, dx, dy, dd, slope, width, angle, nval, dt=2.0,
                               kinematic=False, outlet=outlet, diag=False)
        ch.nthreads = os.cpu_count() or 1
        ch.set_input('P_rain', cases.RAIN_50MMPH)
        src = dict(FULL_SRC, I0=0.5)            # soil wet enough for 50 mm/h to run off
        eng.enable_fused_sources(**src)
        ch.enable_fused(**src)
        rng = np.random.default_rng(3)
        Ta = 2.0 + 3.0 * (2.0 * rng.random((n, n)) - 1.0)
        for nm, v in (('T_air', Ta), ('c0', 2.7), ('T0', -0.2), ('T_surf', 5.0), ('Qn_SW', 300.0),
                      ('Qn_LW', -40.0)):
            eng.set_input(nm, v)
            ch.set_input(nm, v)
    for k, v in cases.GA_TREYNOR.items():
        eng.set_input(k, v)
        if not kin:
            ch.set_input(k, v)
    worst = 0.0
    worst_cell = dict(vol=0.0, d=0.0, u=0.0, Q=0.0)
    q_series = []
    with np.errstate(all='ignore'):
        for i in range(1500):
            if i == 500:
                eng.set_input('P_rain', 0.0)
                if kin:
                    ch.P_rain = np.array(0.0)
                else:
                    ch.set_input('P_rain', 0.0)
            if kin:
                ch.IN, I = o.green_ampt_step(I, ch.P_rain, 0.0, dt=2.0, **cases.GA_TREYNOR)
            else:
                ch.time_min = i * 2.0 / 60.0
            ch.step()
            eng.step(time_min=i * 2.0 / 60.0)
            if (i + 1) % 100 == 0:
                s = eng.read_scalars()
                q_series.append(s.Q_outlet)
                want = float(ch.Q_outlet)
                assert abs(s.Q_outlet - want) <= RTOL * max(abs(want), 1e-300), (i, s.Q_outlet, want)
            if (i + 1) in (250, 500, 900, 1500):
                for nm in ('vol', 'd', 'u', 'Q'):
                    got = (eng.Q if nm == 'Q' else eng.vol if nm == 'vol'
                           else eng.own(getattr(eng, nm))).cpu().numpy()
                    worst = max(worst, rel_err(got, getattr(ch, nm)))
                    worst_cell[nm] = max(worst_cell[nm], cell_rel_err(got, getattr(ch, nm)))
    assert eng.last_path == 'packed'
    assert worst <= RTOL, worst
    for nm, e in worst_cell.items():
        assert e <= (1e-9 if kin else 1e-6), worst_cell
    # a real hydrograph at the largest-area cell: it rises, peaks and recedes
    q = np.array(q_series)
    assert q.max() > 0 and q.argmax() < len(q) - 2 and q[-1] < 0.8 * q.max(), q
    s = eng.read_scalars()
    for nm in ('vol_R', 'vol_edge', 'vol_Q', 'Q_peak', 'T_peak', 'vol_IN'):
        want = float(getattr(ch, nm)) if hasattr(ch, nm) else None
        if want is not None:
            assert abs(getattr(s, nm) - want) <= RTOL * max(abs(want), 1e-300), nm
    assert s.n_neg_d == s.n_inf_d == s.n_neg_u == s.n_nan_u == s.n_inf_u == 0


@pytest.mark.parametrize('tag', ['kw_man', 'kw_low', 'dw_man', 'dw_low'])
def test_synth_variants_vs_reference_fixture(synth_channels, tag):
    """KW/DW x Manning/law-of-the-wall, grid ET (masked in place) and IN, zero
    bank angles mixed in, 150 steps -- against the stand-alone REFERENCE run."""
    g = synth_channels
    kin, man = tag.startswith('kw'), tag.endswith('man')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    eng = _engine(code, dx, dy, dd, g['slope'], g['width'], g['angle'],
                  g['nval'] if man else g['z0val'], dt=2.0, da=900.0, kinematic=kin,
                  manning=man, diag=True, outlet=(ny - 2, nx // 2))
    eng.set_input('P_rain', 80 / 3.6e6)
    eng.set_input('GW', 1e-7)
    eng.set_input('SM', 2e-7)
    eng.set_input('ET', g['ET0'].copy())
    eng.set_input('IN', g['IN0'].copy())
    for i in range(150):
        if i == 100:
            eng.set_input('P_rain', 0.0)
        eng.step(time_min=(i * 2.0) / 60.0)
    # grid forcing, law of the wall, diagnostics, 4286 distinct bank angles, nx = 83: all on
    # the extended packed (TMA ring) kernels
    assert eng.last_path == 'packed' and eng.last_kernel == 'ext'
    want = {n: g[tag + '_' + n] for n in GRIDS + SCALARS}
    _cmp_grids(eng, want, GRIDS)
    if not kin:
        assert rel_err(eng.own(eng.S_free).cpu().numpy(), g[tag + '_S_free']) <= RTOL
    _cmp_scalars(eng, want)
    # provider's ET grid is zeroed in place where vol < ET*da*dt (bit-exact)
    assert np.array_equal(eng.own(eng.src['ET']).cpu().numpy(), g[tag + '_ET_after'])


@pytest.mark.parametrize('tag', ['kw_man', 'kw_low', 'dw_man', 'dw_low'])
def test_extended_packed_equals_generic_kernel(synth_channels, tag):
    """The extended packed kernels keep the generic kernel's operation order: every grid
    within 1e-13 of it after 150 steps (only x^(2/3) is evaluated differently), the in-place
    ET mask identical."""
    g = synth_channels
    kin, man = tag.startswith('kw'), tag.endswith('man')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    res = {}
    for packed in (True, False):
        eng = _engine(code, dx, dy, dd, g['slope'], g['width'], g['angle'],
                      g['nval'] if man else g['z0val'], dt=2.0, da=900.0, kinematic=kin,
                      manning=man, diag=True, write_R=True, outlet=(ny - 2, nx // 2), packed=packed)
        eng.set_input('P_rain', 80 / 3.6e6)
        eng.set_input('GW', 1e-7)
        eng.set_input('SM', 2e-7)
        eng.set_input('ET', g['ET0'].copy())
        eng.set_input('IN', g['IN0'].copy())
        for i in range(150):
            eng.step(time_min=(i * 2.0) / 60.0)
        assert eng.last_kernel == ('ext' if packed else 'generic')
        res[packed] = {n: (eng.Q if n == 'Q' else eng.vol if n == 'vol'
                           else eng.own(getattr(eng, n))).cpu().numpy().copy()
                       for n in GRIDS + ('R',)}
        res[packed]['ET'] = eng.own(eng.src['ET']).cpu().numpy().copy()
        res[packed]['s'] = eng.read_scalars()
    for n in GRIDS + ('R',):
        assert rel_err(res[True][n], res[False][n]) <= 1e-13, n
    assert np.array_equal(res[True]['ET'], res[False]['ET'])
    for n in SCALARS:
        a_, b_ = getattr(res[True]['s'], n), getattr(res[False]['s'], n)
        assert abs(a_ - b_) <= 1e-13 * max(abs(b_), 1e-300), n


def test_lean_mode_equals_diag_mode(synth_channels):
    """Without TFB_CH_DIAG the generic kernel's state grids are bit-identical to
    the diagnostics kernel's; the packed fast path (reciprocal multiplications,
    include/tfb.h) agrees to a few ulp per step (asserted: 1e-12 after 60 steps)."""
    g = synth_channels
    code = g['code']
    ny, nx = code.shape
    nx2 = nx - (nx % 4)                      # the packed path needs nx % 4 == 0
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    sl = (slice(None), slice(0, nx2))
    code2, _ = o.d8_flow_grid(g['DEM'][sl], dx, dy, dd)     # D8 of the cropped DEM
    res = {}
    for name, kw in (('diag', dict(diag=True, packed=False)), ('lean', dict(packed=False)), ('packed', dict())):
        # bank angles rounded to whole degrees: <= 256 distinct values (packed path)
        eng = _engine(code2, dx, dy, dd, g['slope'][sl], g['width'][sl],
                      np.round(g['angle'][sl]), g['nval'][sl], dt=2.0, da=900.0,
                      outlet=(ny - 2, nx2 // 2), **kw)
        eng.set_input('P_rain', 80 / 3.6e6)
        for i in range(60):
            eng.step(time_min=i / 30.0)
        assert eng.last_path == ('packed' if name == 'packed' else 'generic')
        res[name] = [t.cpu().numpy().copy() for t in (eng.vol, eng.own(eng.d), eng.own(eng.u), eng.Q)]
        res[name + '_s'] = eng.read_scalars()
    for a, b in zip(res['lean'], res['diag']):
        assert np.array_equal(a, b)
    for a, b in zip(res['packed'], res['diag']):
        assert rel_err(a, b) <= 1e-12
    sp, sd = res['packed_s'], res['diag_s']
    for n in SCALARS:
        assert abs(getattr(sp, n) - getattr(sd, n)) <= 1e-12 * max(abs(getattr(sd, n)), 1e-300), n
    assert abs(sp.dt_cfl_min - sd.dt_cfl_min) <= 1e-5 * sd.dt_cfl_min


@pytest.mark.parametrize('shape', [(1, 5), (3, 3), (2, 70), (70, 2), (33, 257)])
def test_ragged_and_tiny_grids(shape):
    """Edge shapes: smaller than a tile, single row/column, non-multiples of the
    tile; all-border grids have code 0 everywhere (pure edge bookkeeping)."""
    ny, nx = shape
    rng = np.random.default_rng(3)
    DEM = np.float32(100 - 0.3 * np.arange(ny)[:, None] - 0.1 * np.arange(nx)[None, :]
                     + 0.05 * rng.random((ny, nx)))
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    with np.errstate(all='ignore'):
        slope = np.float32(o.d8_slope(DEM, code, ds))
    if not (slope > 0).any():
        slope[:] = 0.01
    width = np.float32(1 + rng.random((ny, nx)))
    angle = np.float32(30 + 20 * rng.random((ny, nx)))
    nval = np.float32(0.03 + 0.01 * rng.random((ny, nx)))
    ch = o.Channels(code, ds, dw, slope, width, angle, nval=nval, da=900.0, dt=3.0,
                    outlet=(ny // 2, nx // 2))
    ch.P_rain = np.array(1e-5)
    eng = _engine(code, dx, dy, dd, slope, width, angle, nval, dt=3.0, da=900.0,
                  diag=True, outlet=(ny // 2, nx // 2))
    eng.set_input('P_rain', 1e-5)
    for i in range(40):
        ch.step()
        eng.step(time_min=i * 3.0 / 60.0)
    _cmp_grids(eng, ch, GRIDS)
    _cmp_scalars(eng, ch)


def test_bad_value_counts_reported():
    """A time step far beyond stability makes depths/velocities blow up; the
    kernel must report the counts the reference's checks raise on."""
    ny, nx = 40, 40
    DEM = np.float32(100 - 0.5 * np.arange(ny)[:, None] + 0 * np.arange(nx)[None, :])
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    slope = np.full((ny, nx), 0.05, dtype=np.float32)
    eng = _engine(code, dx, dy, dd, slope, np.float32(np.ones((ny, nx))),
                  np.float32(np.full((ny, nx), 45.0)), np.float32(np.full((ny, nx), 0.03)),
                  dt=6.0, da=900.0, outlet=(ny - 2, 5))
    eng.set_input('P_rain', float('inf'))
    eng.step()
    s = eng.read_scalars()
    assert s.n_inf_d > 0 or s.n_nan_d > 0
    assert s.n_nan_u + s.n_inf_u > 0


def test_fused_sources_match_oracle_composition(sources_gold):
    """met split + degree-day + Priestley-Taylor + Green-Ampt evaluated inside
    the routing kernel == the oracle's per-component functions feeding the
    oracle's channel step, in provider order met, snow, evap, infil, channels."""
    rng = np.random.default_rng(11)
    ny, nx = 45, 52
    from topoflow36_b200 import synth
    DEM = synth.fractal_dem(ny, nx, seed=5, sigma=1.0)
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    slope = np.float32(o.d8_slope(DEM, code, ds))
    width, angle, nval = synth.channel_params(ny, nx, seed=5)
    g = lambda lo, hi: lo + (hi - lo) * rng.random((ny, nx))
    T_air, T_surf = g(-2, 6), g(-2, 8)
    Qn_SW, Qn_LW = g(0, 600), g(-100, 20)
    P = g(0, 4e-5)
    h_swe = g(0, 0.01)
    I = np.zeros((ny, nx)) + 1e-6
    dt = 2.0
    ga = dict(Ks=7.2e-6, Ki=1e-7, qs=0.485, qi=0.375808, G=0.724)
    for kin in (True, False):
        ch = o.Channels(code, ds, dw, slope, width, angle, nval=nval, da=900.0, dt=dt,
                        kinematic=kin, outlet=(ny - 2, nx // 2))
        eng = _engine(code, dx, dy, dd, slope, width, angle, nval, dt=dt, da=900.0,
                      kinematic=kin, diag=True, outlet=(ny - 2, nx // 2))
        eng.enable_fused_sources(met=True, snow=True, evap=True, infil=True, T_rain_snow=1.0,
                                 rho_snow=300.0, alpha=0.95, K_soil=0.45, soil_x=0.05,
                                 T_soil_x=20.0, h0_swe=h_swe.copy(), I0=1e-6)
        eng.set_input('P_rain', P)
        for n, v in (('T_air', T_air), ('T_surf', T_surf), ('Qn_SW', Qn_SW), ('Qn_LW', Qn_LW)):
            eng.set_input(n, v)
        eng.set_input('c0', 2.7)
        eng.set_input('T0', -0.2)
        for n, v in ga.items():
            eng.set_input(n, v)
        hs, Iv = h_swe.copy(), I.copy()
        vol_SM = vol_IN = vol_ET = 0.0
        for i in range(50):
            P_rain, P_snow = o.met_split_precip(P, T_air, 1.0)
            SM, hs, _ = o.degree_day_step(hs, P_snow, T_air, 2.7, -0.2, dt)
            ET = o.priestley_taylor_step(T_air, T_surf, Qn_SW, Qn_LW, 0.95, 0.45, 0.05, 20.0)
            IN, Iv = o.green_ampt_step(Iv, P_rain, SM, ga['Ks'], ga['Ki'], ga['qs'], ga['qi'],
                                       ga['G'], dt)
            vol_SM += np.sum(SM * 900.0 * dt)
            vol_IN += np.sum(IN * 900.0 * dt)
            vol_ET += np.sum(ET * 900.0 * dt)
            ch.P_rain, ch.SM, ch.ET, ch.IN = P_rain, SM, ET.copy(), IN
            ch.step()
            eng.step(time_min=i * dt / 60.0)
        _cmp_grids(eng, ch, GRIDS)
        s = _cmp_scalars(eng, ch)
        assert rel_err(eng.I.cpu().numpy(), Iv) <= RTOL
        assert rel_err(eng.h_swe.cpu().numpy(), hs) <= RTOL
        assert abs(s.vol_SM - vol_SM) <= 1e-11 * abs(vol_SM)
        assert abs(s.vol_IN - vol_IN) <= 1e-11 * abs(vol_IN)
        assert abs(s.vol_ET - vol_ET) <= 1e-11 * abs(vol_ET)


FULL_SRC = dict(met=True, snow=True, evap=True, infil=True, T_rain_snow=1.0, rho_snow=300.0,
                alpha=0.95, K_soil=0.45, soil_x=0.05, T_soil_x=20.0, h0_swe=0.004, I0=0.05)


@pytest.mark.parametrize('kinematic,src', [(True, 'none'), (True, 'ga'), (True, 'full'),
                                           (False, 'none'), (False, 'ga'), (False, 'full')])
def test_packed_path_vs_c_oracle(kinematic, src):
    """The benchmarked kernels (TMA ring, packed inputs; kinematic = one launch,
    diffusive = two passes) at a size the numpy oracle cannot handle -- 2048^2
    kinematic, 1024^2 diffusive -- for 60 steps against the multithreaded C
    restatement (pinned to the reference in test_c_oracle.py), with uniform
    forcing, fused Green-Ampt, or all four fused sources with a T_air grid; plus
    exact mass conservation."""
    import os
    import torch
    from oracle import c_oracle
    from topoflow36_b200 import cases
    from topoflow36_b200.d8 import D8Toolkit
    ny = nx = 2048 if kinematic else 1024
    tile = cases.noise_tile(ny, nx, seed=5)
    DEM = cases.slab_dem(ny, nx, 0, ny, 0, tile)
    width, angle, nval = cases.slab_params(ny, nx, 0, 0, seed=5)
    angle[::3, ::5] = 30.0                         # three distinct bank angles
    angle[::7, 1::4] = 0.0                         # including vertical banks (tan == 0)
    tk = D8Toolkit(DEM)
    tk.update_flow_grid()
    code = tk.code.cpu().numpy()
    slope = tk.slope().cpu().numpy()
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    eng = _engine(code, dx, dy, dd, slope, width, angle, nval, dt=1.0, da=900.0,
                  kinematic=kinematic, outlet=(ny - 2, nx // 2))
    ch = c_oracle.Channels(code, dx, dy, dd, slope, width, angle, nval, dt=1.0,
                           kinematic=kinematic, outlet=(ny - 2, nx // 2), diag=False)
    ch.nthreads = os.cpu_count() or 1
    both = lambda n, v: (eng.set_input(n, v), ch.set_input(n, v))
    both('P_rain', cases.RAIN_50MMPH)
    if src == 'ga':
        eng.enable_fused_sources(infil=True, I0=cases.I0_WET)
        ch.enable_fused(infil=True, I0=cases.I0_WET)
    elif src == 'full':
        eng.enable_fused_sources(**FULL_SRC)
        ch.enable_fused(**FULL_SRC)
        rng = np.random.default_rng(3)
        both('T_air', 2.0 + 3.0 * (2.0 * rng.random((ny, nx)) - 1.0))
        for n, v in (('c0', 2.7), ('T0', -0.2), ('T_surf', 5.0), ('Qn_SW', 300.0),
                     ('Qn_LW', -40.0), ('GW', 1e-7), ('MR', 2e-8)):
            both(n, v)
    if src != 'none':
        for n, v in cases.GA_TREYNOR.items():
            both(n, v)
    for i in range(60):
        eng.step(time_min=i / 60.0)
        ch.time_min = i / 60.0
        ch.step()
    assert eng.last_path == 'packed'
    names = ('vol', 'd', 'u', 'Q') + (('I',) if src != 'none' else ()) + \
        (('h_swe',) if src == 'full' else ())
    for n in names:
        got = (eng.Q if n == 'Q' else getattr(eng, n) if n in ('vol', 'I', 'h_swe')
               else eng.own(getattr(eng, n)))
        e = rel_err(got.cpu().numpy(), getattr(ch, n))
        assert e <= RTOL, '%s rel err %.3e' % (n, e)
    s = eng.read_scalars()
    for n in ('vol_R', 'vol_edge', 'vol_Q', 'Q_outlet', 'Q_peak', 'vol_IN', 'vol_SM', 'vol_ET',
              'vol_P', 'vol_PR', 'vol_PS'):
        want = float(getattr(ch, n))
        assert abs(getattr(s, n) - want) <= 1e-11 * max(abs(want), 1e-300), n
    # water is conserved: runoff in = stored in channels + left through the edges
    stored = float(eng.vol.sum())
    assert abs(s.vol_R - (stored + s.vol_edge)) <= 1e-10 * abs(s.vol_R)
    assert s.n_neg_d == s.n_nan_d == s.n_inf_d == s.n_neg_u == s.n_nan_u == s.n_inf_u == 0
    assert float(eng.Q.max()) > 0


@pytest.mark.parametrize('packed', [True, False])
@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_flood_option_vs_reference_fixture(flood_gold, tag, packed):
    """FLOOD_OPTION (rectangle-over-trapezoid overbank flow, channels_base.py
    :1701-1862, :2149-2185, :2393-2448) against a stand-alone REFERENCE run: on the
    extended packed kernels (the path a run takes) and on the generic kernel."""
    g = flood_gold
    kin = (tag == 'kw')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    eng = _engine(code, dx, dy, dd, g['slope'], g['width'], g['angle'], g['nval'], dt=2.0,
                  da=900.0, kinematic=kin, diag=True, outlet=(ny - 2, nx // 2), flood=True,
                  d_bankfull=np.float64(g['d_bankfull']), packed=packed)
    eng.set_input('P_rain', 150 / 3.6e6)
    for i in range(150):
        if i == 100:
            eng.set_input('P_rain', 0.0)
        eng.step(time_min=(i * 2.0) / 60.0)
    assert (eng.last_path, eng.last_kernel) == (('packed', 'ext') if packed else ('generic', 'generic'))
    want = {n: g[tag + '_' + n] for n in GRIDS + SCALARS}
    _cmp_grids(eng, want, GRIDS)
    assert rel_err(eng.own(eng.d_flood).cpu().numpy(), g[tag + '_d_flood']) <= RTOL
    assert rel_err(eng.own(eng.vol_flood).cpu().numpy(), g[tag + '_vol_flood']) <= RTOL
    assert g[tag + '_d_flood'].max() > 0.1
    if not kin:
        assert rel_err(eng.own(eng.S_free).cpu().numpy(), g[tag + '_S_free']) <= RTOL
    _cmp_scalars(eng, want)


@pytest.mark.parametrize('packed', [True, False])
@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_attenuate_vs_reference_fixture(flood_gold, attenuate_gold, tag, packed):
    """ATTENUATE (per-cell linear reservoir in front of the channel,
    channels_base.py:2077-2084) against a stand-alone REFERENCE run (extended packed
    kernels and generic kernel)."""
    g0, g = flood_gold, attenuate_gold
    kin = (tag == 'kw')
    code = g0['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    eng = _engine(code, dx, dy, dd, g0['slope'], g0['width'], g0['angle'], g0['nval'], dt=2.0,
                  da=900.0, kinematic=kin, diag=True, outlet=(ny - 2, nx // 2), attenuate=True,
                  n_days=0.002, packed=packed)
    eng.set_input('P_rain', 150 / 3.6e6)
    for i in range(150):
        if i == 100:
            eng.set_input('P_rain', 0.0)
        eng.step(time_min=(i * 2.0) / 60.0)
    assert (eng.last_path, eng.last_kernel) == (('packed', 'ext') if packed else ('generic', 'generic'))
    want = {n: g[tag + '_' + n] for n in GRIDS + SCALARS}
    _cmp_grids(eng, want, GRIDS)
    assert rel_err(eng.vol_stored.cpu().numpy(), g[tag + '_vol_stored']) <= RTOL
    assert g[tag + '_vol_stored'].max() > 1.0
    _cmp_scalars(eng, want)


@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_flood_with_attenuate_ext_equals_generic(flood_gold, tag):
    """FLOOD_OPTION and ATTENUATE together with a grid infiltration rate and the law of the
    wall (no reference fixture for the combination): the extended packed kernels against the
    generic kernel, which keeps the reference's operation order."""
    g = flood_gold
    kin = (tag == 'kw')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    rng = np.random.default_rng(5)
    IN = 2e-6 * rng.random((ny, nx))
    engs = []
    for packed in (True, False):
        eng = _engine(code, dx, dy, dd, g['slope'], g['width'], g['angle'], 0.01 + 0 * g['nval'],
                      dt=2.0, da=900.0, kinematic=kin, manning=False, diag=True,
                      outlet=(ny - 2, nx // 2), flood=True, d_bankfull=np.float64(g['d_bankfull']),
                      attenuate=True, n_days=0.002, packed=packed)
        eng.set_input('P_rain', 150 / 3.6e6)
        eng.set_input('IN', IN)
        for i in range(120):
            if i == 90:
                eng.set_input('P_rain', 0.0)
            eng.step(time_min=(i * 2.0) / 60.0)
        engs.append(eng)
    a, b = engs
    assert (a.last_kernel, b.last_kernel) == ('ext', 'generic')
    for n in ('vol', 'Q'):
        assert rel_err(getattr(a, n).cpu().numpy(), getattr(b, n).cpu().numpy()) <= 1e-12, n
    for n in ('d', 'u', 'd_flood', 'vol_flood', 'Rh', 'f', 'tau', 'froude'):
        assert rel_err(a.own(getattr(a, n)).cpu().numpy(), b.own(getattr(b, n)).cpu().numpy()) <= 1e-12, n
    assert rel_err(a.vol_stored.cpu().numpy(), b.vol_stored.cpu().numpy()) <= 1e-12
    assert float(a.own(a.d_flood).max()) > 0.01 and float(a.vol_stored.max()) > 0.5
    sa, sb = a.read_scalars(), b.read_scalars()
    for n in ('vol_R', 'vol_edge', 'vol_Q', 'Q_peak', 'd_peak', 'u_peak'):
        assert abs(getattr(sa, n) - getattr(sb, n)) <= 1e-12 * max(abs(getattr(sb, n)), 1e-300), n


@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_sinuosity_grid_and_fp64_inputs_ext_equals_generic(flood_gold, tag):
    """A sinuosity GRID (ds = sinu * ds_d8 per cell, channels_base.py:1242), a scalar width
    that is no float32 value (3.3 m) and slopes that are no float32 values: the extended
    packed kernels take them as fp64 streams; compared with the generic kernel."""
    g = flood_gold
    kin = (tag == 'kw')
    code = g['code']
    ny, nx = code.shape
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    rng = np.random.default_rng(11)
    sinu = 1.0 + 0.4 * rng.random((ny, nx))
    slope = np.float64(g['slope']) / sinu
    engs = []
    for packed in (True, False):
        eng = _engine(code, dx, dy, dd, slope, 3.3, g['angle'], g['nval'], dt=2.0, da=900.0,
                      kinematic=kin, diag=True, outlet=(ny - 2, nx // 2), sinu=sinu, packed=packed)
        eng.set_input('P_rain', 120 / 3.6e6)
        for i in range(100):
            if i == 70:
                eng.set_input('P_rain', 0.0)
            eng.step(time_min=(i * 2.0) / 60.0)
        engs.append(eng)
    a, b = engs
    assert (a.last_kernel, b.last_kernel) == ('ext', 'generic')
    assert a.packed.width64 and a.packed.S64 and a.packed.sinu64 and not a.packed.width32
    for n in ('vol', 'Q'):
        assert rel_err(getattr(a, n).cpu().numpy(), getattr(b, n).cpu().numpy()) <= 1e-12, n
    for n in ('d', 'u', 'Rh', 'f', 'tau', 'froude'):
        assert rel_err(a.own(getattr(a, n)).cpu().numpy(), b.own(getattr(b, n)).cpu().numpy()) <= 1e-12, n
    assert float(a.Q.max()) > 0
    sa, sb = a.read_scalars(), b.read_scalars()
    for n in ('vol_R', 'vol_edge', 'vol_Q', 'Q_peak', 'd_peak', 'u_peak'):
        assert abs(getattr(sa, n) - getattr(sb, n)) <= 1e-12 * max(abs(getattr(sb, n)), 1e-300), n


def test_flood_component_from_cfg(tmp_path, flood_gold):
    """FLOOD_OPTION read from the CFG file by the BMI component."""
    from topoflow36_b200 import channels_kinematic_wave as ckw, synth
    g = flood_gold
    ny, nx = g['code'].shape
    cfg = synth.write_case(str(tmp_path), g['DEM'], g['slope'], g['width'], g['angle'],
                           nval=g['nval'], dt=2.0, kinematic=True, outlet=(ny - 2, nx // 2),
                           flood=True, d_bankfull=g['d_bankfull'])
    c = ckw.channels_component()
    c.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    assert c.FLOOD_OPTION and np.array_equal(c.d8.d8_grid, g['code'])
    c.set_value('atmosphere_water__rainfall_volume_flux', np.array(150 / 3.6e6))
    for i in range(150):
        if i == 100:
            c.set_value('atmosphere_water__rainfall_volume_flux', np.array(0.0))
        c.update()
    for n in ('vol', 'd', 'u', 'Q', 'd_flood', 'vol_flood'):
        assert rel_err(getattr(c, n), g['kw_' + n]) <= RTOL, n
    assert rel_err(c.get_value_ptr('land_surface_water__depth'), g['kw_d_flood']) <= RTOL
    c.finalize()
    assert abs(float(c.vol_flood_sum) - g['kw_vol_flood'].sum()) <= 1e-11 * g['kw_vol_flood'].sum()


@pytest.mark.parametrize('shape', [(1, 4), (2, 8), (15, 64), (16, 68), (31, 132), (47, 260), (100, 60),
                                   (44, 29), (5, 7), (3, 1), (20, 33), (17, 67), (33, 130),
                                   (24, 64), (25, 66), (20, 64), (21, 70), (49, 129), (73, 65)])
@pytest.mark.parametrize('kinematic', [True, False])
def test_packed_path_ragged_shapes(shape, kinematic):
    """TMA tiles that hang over the grid edge (zero-filled by the TMA unit), grids
    smaller than one 64 x 15 tile, a single row or column, widths that are not a
    multiple of 4 or odd (rows padded to the pitch; the lane holding the last column
    of an odd grid has no second cell): packed kinematic (1 launch) and diffusive
    (2 launches) steps against the numpy oracle, with fused Green-Ampt."""
    from topoflow36_b200 import cases
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx)
    DEM = np.float32(100 - 0.3 * np.arange(ny)[:, None] - 0.1 * np.arange(nx)[None, :]
                     + 0.2 * rng.random((ny, nx)))
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    ds, dw = o.d8_ds_dw(code, dx, dy, dd)
    with np.errstate(all='ignore'):
        slope = np.float32(o.d8_slope(DEM, code, ds))
    if not (slope > 0).any():
        slope[:] = 0.01
    width = np.float32(1 + rng.random((ny, nx)))
    angle = np.float32(rng.choice([0.0, 30.0, 45.0], size=(ny, nx)))
    nval = np.float32(0.03 + 0.01 * rng.random((ny, nx)))
    out = (ny // 2, nx // 2)
    ch = o.Channels(code, ds, dw, slope, width, angle, nval=nval, da=900.0, dt=3.0,
                    kinematic=kinematic, outlet=out)
    ch.P_rain = np.array(3e-5)
    eng = _engine(code, dx, dy, dd, slope, width, angle, nval, dt=3.0, da=900.0,
                  kinematic=kinematic, outlet=out)
    eng.set_input('P_rain', 3e-5)
    eng.enable_fused_sources(infil=True, I0=0.05)
    for k, v in cases.GA_TREYNOR.items():
        eng.set_input(k, v)
    I = np.zeros((ny, nx)) + 0.05
    for i in range(40):
        ch.IN, I = o.green_ampt_step(I, ch.P_rain, 0.0, dt=3.0, **cases.GA_TREYNOR)
        ch.step()
        eng.step(time_min=i * 3.0 / 60.0)
    assert eng.last_path == 'packed'
    _cmp_grids(eng, ch, ('vol', 'd', 'u', 'Q'))
    assert rel_err(eng.I.cpu().numpy(), I) <= RTOL
    s = eng.read_scalars()
    for n in ('vol_R', 'vol_edge', 'vol_Q', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
              'Q_peak', 'T_peak'):
        want = float(getattr(ch, n))
        assert abs(getattr(s, n) - want) <= RTOL * max(abs(want), 1e-300), n


@pytest.mark.parametrize('tag', ['kw', 'dw'])
def test_geographic_pixels_vs_reference_fixture(geo_gold, tag):
    """Fixed-angle (geographic) pixels: per-row dx / dd and a per-cell area grid, D8
    toolkit and routing against a stand-alone REFERENCE run."""
    import torch
    from topoflow36_b200.d8 import D8Toolkit
    from topoflow36_b200.routing import ChannelsEngine
    g = geo_gold
    ny, nx = g['DEM'].shape
    tk = D8Toolkit(g['DEM'], dx=g['dx'], dy=g['dy'], dd=g['dd'], da=g['da'])
    tk.update()
    assert np.array_equal(tk.code.cpu().numpy(), g['code'])
    assert np.array_equal(tk.ds.cpu().numpy(), g['ds32'])
    assert np.array_equal(tk.dw.cpu().numpy(), g['dw32'])
    assert np.array_equal(tk.A.cpu().numpy(), g['area'])
    assert np.array_equal(tk.slope().cpu().numpy(), g['slope'])
    kin = (tag == 'kw')
    for diag in (True, False):
        eng = _engine(g['code'], g['dx'], g['dy'], g['dd'], g['slope'], g['width'], g['angle'],
                      g['nval'], dt=2.0, da=g['da'], kinematic=kin, diag=diag,
                      outlet=(ny - 2, nx // 2))
        eng.set_input('P_rain', 100 / 3.6e6)
        for i in range(120):
            if i == 80:
                eng.set_input('P_rain', 0.0)
            eng.step(time_min=(i * 2.0) / 60.0)
        want = {n: g[tag + '_' + n] for n in GRIDS + SCALARS}
        _cmp_grids(eng, want, GRIDS if diag else ('vol', 'd', 'u', 'Q'))
        _cmp_scalars(eng, want)
        if not diag:
            assert eng.last_path == 'packed'       # per-row pixel areas ride the fast path too


@pytest.mark.parametrize('workload', ['kw_ga', 'dw'])
def test_bench_size_parity_vs_c_oracle(workload):
    """BASELINE configs[1] at ITS size (4096^2, the grid bench.py times) on the lean packed
    kernels against the C restatement of the reference path, 40 steps from the initial state --
    the same leg bench.py runs inside its JSON line (sanity.parity_rel_err)."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from topoflow36_b200 import cases
    n = 4096
    kin = workload.startswith('kw')
    eng, parts = cases.build_engine(n, n, kinematic=kin, sources='ga' if kin else 'none',
                                    return_parts=True)
    res = bench.parity_vs_c_oracle(eng, parts, n, n_steps=40)
    assert res['parity_path'] == 'packed' and eng.last_kernel == 'lean'
    assert max(res['parity_rel_err'].values()) <= 1e-11, res['parity_rel_err']
    s = eng.read_scalars()
    stored = float(eng.vol.sum())
    infil = s.vol_IN if kin else 0.0
    # water balance at full size: rain in = channels + edges + infiltrated
    assert abs((s.vol_R) - (stored + s.vol_edge)) <= 1e-12 * abs(s.vol_R + infil)
    assert s.n_neg_d == s.n_nan_d == s.n_inf_d == s.n_neg_u == s.n_nan_u == s.n_inf_u == 0
