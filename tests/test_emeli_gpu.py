"""The real thing: the reference's UNMODIFIED EMELI coupler (framework/emeli.py:890-1093)
runs the Treynor June-20-1967 storm with the GPU Channels component plugged in under the
reference's own component name (emeli_dropin.install()), next to the reference's own
meteorology / snow / evap / infil / ice / diversions / satzone / driver components.

The reference tree comes from /root/reference in the build container and from the archive
`make -C oracle _ref` staged (oracle/_ref/, shipped to the GPU box) elsewhere.  Expected
values are what the all-reference run produced here (oracle/gen_golden.py ->
tests/golden/treynor_emeli_final.npz, meta.json): Q_peak = 12.443873381178541 after 1592
steps with Linear time interpolation, 12.457869898593854 after 1585 steps without.

Tolerance: the north-star bar is rel 1e-9 on hydrographs; asserted here: 1e-10 on the
outlet scalars and the final fp64 grids, and the float32 output files value for value."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, load_golden, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-10
SCALARS = ('vol_R', 'vol_Q', 'vol_edge', 'Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet',
           'Q_peak', 'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak', 'vol_chan_sum')


def _reference():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip('reference not staged (make -C oracle _ref)')
    rh.import_reference()
    return rh


def _run(rh, work, method, device_interpolator):
    from topoflow36_b200 import emeli_dropin, channels_base as gpu_base
    from topoflow.framework import emeli
    cfg_dir, out_dir = rh.stage_treynor(str(work))
    swapped = emeli_dropin.install()
    try:
        assert 'channels_kinematic_wave' in swapped
        F = emeli_dropin.gpu_framework() if device_interpolator else emeli.framework
        f = F()
        rh.quiet_call(f.run_model, driver_comp_name='topoflow_driver', cfg_prefix='June_20_67',
                      cfg_directory=cfg_dir, SILENT=True, time_interp_method=method)
    finally:
        emeli_dropin.uninstall()
    ch = f.comp_set['tf_channels_kin_wave']
    assert isinstance(ch, gpu_base.channels_component)
    assert type(f.comp_set['tf_meteorology']).__module__.startswith('topoflow.components')
    return ch, out_dir


def _check_state(ch, want_scalars, want_grids, n_steps):
    assert int(ch.time_index) == int(n_steps)
    for n in SCALARS:
        got, want = float(getattr(ch, n)), float(want_scalars[n])
        assert abs(got - want) <= RTOL * max(abs(want), 1e-300), (n, got, want)
    for n, want in want_grids.items():
        e = rel_err(getattr(ch, n), want)
        assert e <= RTOL, (n, e)
    assert ch.engine.last_path == 'packed'          # nx = 29: pitched rows, TMA ring kernel


def test_emeli_runs_treynor_storm_on_the_gpu_component(tmp_path):
    """time_interp_method='None' (the coupling INTEGRATION.md recommends)."""
    rh = _reference()
    g = load_golden('treynor_emeli_final.npz')
    ch, _ = _run(rh, tmp_path, 'None', False)
    _check_state(ch, {n: g['none_' + n] for n in SCALARS},
                 {n: g['none_' + n] for n in ('vol', 'd', 'u', 'Q')}, g['none_n_steps'])


@pytest.mark.parametrize('device_interpolator', [False, True])
def test_emeli_linear_interpolation_and_output_files(tmp_path, device_interpolator):
    """EMELI's default Linear interpolation, with the reference's interpolator and with the
    device-aware one of emeli_dropin.gpu_framework(); the RTS stacks (MSB, like the data
    set), their RTI side-cars and the outlet text series equal the files the reference
    wrote."""
    rh = _reference()
    from topoflow36_b200 import grid_io
    g = load_golden('treynor_emeli_final.npz')
    meta = json.load(open(os.path.join(GOLD, 'meta.json')))['treynor_emeli']
    ch, out = _run(rh, tmp_path, 'Linear', device_interpolator)
    _check_state(ch, meta, {n: g[n] for n in ('vol', 'd', 'u', 'Q')}, meta['n_steps'])
    # ---- grid stacks: float32 frames in the byte order of Treynor.rti (MSB) ------------
    for v in ('Q', 'd'):
        want = np.frombuffer(g['file_2D_%s_rts' % v].tobytes(), dtype='>f4')
        raw = open(os.path.join(out, 'June_20_67_2D-%s.rts' % v), 'rb').read()
        got = np.frombuffer(raw, dtype='>f4')
        assert got.size == want.size == 160 * 44 * 29
        differ = np.flatnonzero(got != want)
        # fp64 values within 1e-12 of each other almost always round to the same float32
        assert differ.size <= 8, differ.size
        assert rel_err(got, want) <= 2e-7
    # ---- side-car: same geometry / byte order / type as the reference's ----------------
    ref_rti = tmp_path / 'ref_2D-Q.rti'
    ref_rti.write_bytes(g['file_2D_Q_rti'].tobytes())
    a = grid_io.read_rti(os.path.join(out, 'June_20_67_2D-Q.rti'))
    b = grid_io.read_rti(str(ref_rti))
    assert a is not None and a.byte_order == 'MSB' and a.data_type == 'FLOAT'
    for k in vars(b):
        if k != 'grid_file':
            assert getattr(a, k) == getattr(b, k), k
    # ---- outlet series: header verbatim, values to the 7 decimals that are written -----
    want = g['file_0D_Q_txt'].tobytes().decode().splitlines()
    got = open(os.path.join(out, 'June_20_67_0D-Q.txt')).read().splitlines()
    assert got[:2] == want[:2] and len(got) == len(want)
    A = np.array([[float(x) for x in ln.split()] for ln in got[2:]])
    B = np.array([[float(x) for x in ln.split()] for ln in want[2:]])
    assert np.abs(A - B).max() <= 1.5e-7
    assert sum(x != y for x, y in zip(got, want)) <= 3      # last-digit rounding at most
