"""Host-side logic that needs no GPU: the C ABI exports what include/tfb.h
declares, RTI/RTG IO, the CFG machinery of the BMI base, loud failure without a
device, and -- when the read-only reference tree is present -- that the GPU
component satisfies what the reference's EMELI asks of a component."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_library_exports_every_declared_symbol():
    from topoflow36_b200 import lib as L
    hdr = open(os.path.join(ROOT, 'include', 'tfb.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(tfb_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 29
    assert os.path.exists(L.LIB_PATH), 'build libtfb_b200.so first (__graft_entry__.build())'
    so = ctypes.CDLL(L.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(so, n)]
    assert not missing, missing
    # the binding covers the header, name for name
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.load()                       # binds all symbols, checks struct layouts
    assert lib.tfb_abi_version() == 1
    tbl = (ctypes.c_uint8 * 256)()
    assert lib.tfb_d8_resolve_table(tbl) == 0          # host-only entry point
    assert sum(tbl) == 10161 and list(tbl[:8]) == [0, 1, 2, 2, 4, 1, 2, 2]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from topoflow36_b200 import lib as L
    from topoflow36_b200.routing import ChannelsEngine
    from topoflow36_b200.d8 import D8Toolkit
    with pytest.raises(L.TfbError):
        D8Toolkit(np.zeros((4, 4), dtype=np.float32))
    with pytest.raises(L.TfbError):
        ChannelsEngine(4, 4, dt=1.0, code=np.zeros((4, 4), np.uint8), dx=[30.], dy=[30.], dd=[42.],
                       S_bed=np.ones((4, 4)), width=np.ones((4, 4)), angle=np.ones((4, 4)),
                       rough=np.ones((4, 4)))


def test_rti_rtg_roundtrip_both_byte_orders(tmp_path):
    from topoflow36_b200 import grid_io
    g = np.float32(np.random.default_rng(0).random((7, 5)) * 100)
    for order in ('LSB', 'MSB'):
        info = grid_io.make_rti(5, 7, 30.0, 30.0, byte_order=order, gmin=1.0, gmax=2.0)
        p = str(tmp_path / ('g_%s' % order))
        grid_io.write_rti(p + '.rti', info)
        grid_io.write_rtg(p + '.rtg', g, info)
        back = grid_io.read_rti(p + '.rti')
        assert (back.ncols, back.nrows, back.byte_order) == (5, 7, order)
        assert back.xres == 30.0 and back.pixel_geom == 1
        assert np.array_equal(grid_io.read_rtg(p + '.rtg', back), g)
    raw = np.fromfile(str(tmp_path / 'g_MSB.rtg'), dtype='>f4').reshape(7, 5)
    assert np.array_equal(raw, g)
    assert float(grid_io.pixel_area(back)) == 900.0
    assert grid_io.read_rti(str(tmp_path / 'nope.rti')) is None


def _write_case(tmp_path, **kw):
    from topoflow36_b200 import synth
    ny, nx = 12, 16
    DEM = synth.fractal_dem(ny, nx, seed=1, sigma=0.5)
    w, a, n = synth.channel_params(ny, nx, seed=1)
    return synth.write_case(str(tmp_path), DEM, np.full((ny, nx), 0.01, np.float32), w, a, nval=n,
                            dt=3.0, n_steps=7, **kw), (ny, nx)


def test_cfg_machinery_of_the_component(tmp_path):
    """read_path_info / read_time_info / read_grid_info / read_config_file /
    set_directories / initialize_basin_vars / time stepping -- the inherited
    host logic, exercised without touching the device."""
    from topoflow36_b200 import channels_kinematic_wave as ckw
    cfg, (ny, nx) = _write_case(tmp_path, scalar_angle=30.0)
    os.makedirs(str(tmp_path / '__topo'), exist_ok=True)      # empty: files stay beside the CFG
    c = ckw.channels_component()
    c.mode, c.cfg_file = 'driver', cfg
    c.set_constants()
    for f in os.listdir(str(tmp_path)):                       # move inputs into __topo like Treynor
        if f.startswith('Synth'):
            os.rename(str(tmp_path / f), str(tmp_path / '__topo' / f))
    c.initialize_config_vars()
    c.set_missing_cfg_options()
    c.initialize_basin_vars()
    c.initialize_time_vars()
    c.set_computed_input_vars()
    assert c.topo_directory.endswith('__topo' + os.sep) and c.in_directory == c.out_directory
    assert (c.ny, c.nx) == (ny, nx) and float(c.da) == 900.0
    assert c.site_prefix == 'Synth' and c.case_prefix == 'Case1'
    assert c.dt.ndim == 0 and float(c.dt) == 3.0 and c.dt.dtype == np.float64
    assert c.n_steps.dtype == np.int32 and int(c.n_steps) == 7
    assert c.MANNING is True and c.LAW_OF_WALL is False and c.KINEMATIC_WAVE
    assert c.nval_type == 'Grid' and c.nval_file == 'Synth_chan-n.rtg'
    assert c.angle_type == 'Scalar' and c.angle_file == '' and float(c.angle) == 30.0
    assert c.width_file == 'Synth_chan-w.rtg' and c.slope_file == 'Synth_slope.rtg'
    assert c.CHECK_STABILITY is True and c.SAVE_Q_GRIDS is False
    assert c.Q_gs_file == 'Case1_2D-Q.nc'
    assert c.outlet_ID == (ny - 2, nx // 2) and c.n_outlets == 1
    assert float(c.save_grid_dt) == 60.0
    assert c.time_info.duration == 1440.0
    # inputs are read from the float32 RTG files into float64 (model_input.py:181)
    c.read_input_files()
    assert c.width.dtype == np.float64 and c.width.shape == (ny, nx)
    assert np.array_equal(np.float32(c.width), c.width)
    assert c.sinu.ndim == 0 and float(c.sinu) == 1.0
    # clock: 0-D arrays updated IN PLACE, driver stop rule after n_steps
    t_ref, tmin_ref = c.time, c.time_min
    for _ in range(7):
        assert not c.DONE
        c.update_time()
    assert c.DONE and c.time is t_ref and float(t_ref) == 21.0
    assert tmin_ref is c.time_min and abs(float(tmin_ref) - 0.35) < 1e-15
    # BMI surface
    assert c.get_var_name('channel_water_x-section__volume_flow_rate') == 'Q'
    assert c.get_var_units('basin_outlet_water_x-section__volume_flow_rate') == 'm3 s-1'
    assert len(c.get_input_var_names()) == 7 and len(c.get_output_var_names()) == 51
    v = np.array(5e-6)
    c.set_value('atmosphere_water__rainfall_volume_flux', v)
    assert c.get_value_ptr('atmosphere_water__rainfall_volume_flux') is v      # by reference
    c.remove_bad_slopes()                                       # all 0.01: untouched
    c.slope[0, 0], c.slope[1, 1], c.slope[2, 2] = 0.0, -1.0, np.nan
    c.slope[3, 3] = 0.004
    c.remove_bad_slopes()
    assert c.slope[0, 0] == c.slope[1, 1] == c.slope[2, 2] == 0.004


def test_initialize_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from topoflow36_b200 import channels_kinematic_wave as ckw, lib as L
    cfg, _ = _write_case(tmp_path)
    with pytest.raises(L.TfbError):
        ckw.channels_component().initialize(cfg_file=cfg, mode='driver', SILENT=True)


def test_unsupported_options_raise(tmp_path):
    from topoflow36_b200 import channels_kinematic_wave as ckw
    cfg, _ = _write_case(tmp_path)
    open(cfg, 'a').write('CREATE_CHANNEL_FILES | 1     | int        | make width / n grids from area\n')
    c = ckw.channels_component()
    c.mode, c.cfg_file = 'driver', cfg
    c.set_constants()
    c.initialize_config_vars()
    with pytest.raises(NotImplementedError):
        c.set_missing_cfg_options()


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_emeli_resolves_to_the_gpu_component():
    """With emeli_dropin.install(), the UNMODIFIED EMELI repository / lookup code
    instantiates the GPU class under the reference's component names, and the
    class offers every method and attribute EMELI calls."""
    from oracle import ref_harness as rh
    rh.import_reference()
    from topoflow36_b200 import emeli_dropin, channels_base as gpu_base
    from topoflow.framework import emeli
    swapped = emeli_dropin.install()
    try:
        assert set(swapped) == {'channels_kinematic_wave', 'channels_diffusive_wave'}
        f = emeli.framework(SILENT=True)
        rh.quiet_call(f.read_repository)
        for name, ext in (('tf_channels_kin_wave', '_channels_kinematic_wave.cfg'),
                          ('tf_channels_diff_wave', '_channels_diffusive_wave.cfg')):
            rh.quiet_call(f.instantiate_comp, name)
            bmi = f.comp_set[name]
            assert isinstance(bmi, gpu_base.channels_component)
            # ... and it INHERITS the reference's plug-in base (north star), with the GPU
            # class's methods in front of the reference's in the MRO
            from topoflow.utils import BMI_base as ref_bmi
            assert isinstance(bmi, ref_bmi.BMI_component)
            mro = type(bmi).__mro__
            assert mro.index(gpu_base.channels_component) < mro.index(ref_bmi.BMI_component)
            assert type(bmi).update is gpu_base.channels_component.update
            assert type(bmi).read_grid_info.__module__.startswith('topoflow36_b200')
            assert type(bmi).get_bmi_version(bmi) == '2.1'
            assert bmi.get_attribute('cfg_extension') == ext
            for m in ('initialize', 'update', 'finalize', 'get_status', 'get_time_units',
                      'get_current_time', 'get_time_step', 'get_input_var_names',
                      'get_output_var_names', 'get_var_units', 'get_value_ptr', 'set_value',
                      'get_attribute'):
                assert callable(getattr(bmi, m)), m
            assert bmi.comp_status == 'Enabled' and bmi.DONE is False
        # same public names as the reference class
        from topoflow.components import channels_base as ref_base
        sys.modules.pop('topoflow.components.channels_base', None)
        R, G = ref_base.channels_component, gpu_base.channels_component
        assert R._input_var_names == G._input_var_names
        assert R._output_var_names == G._output_var_names
        for k, v in R._var_name_map.items():
            if k in G._var_name_map:
                assert G._var_name_map[k] == v, k
        for k in R._input_var_names + R._output_var_names:
            if k in R._var_units_map:          # the reference lacks units for w_top
                assert G._var_units_map[k] == R._var_units_map[k], k
    finally:
        emeli_dropin.uninstall()


def _run_bench(args, env_extra=None):
    import json
    import subprocess
    env = dict(os.environ)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    return [json.loads(l) for l in lines]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm):
    one JSON line with the contract's keys; under torchrun only rank 0 works."""
    out = _run_bench(['--impl', 'reference', '--steps', '2', '--warmup', '1',
                      '--rows-per-gpu', '192', '--nx', '160'])
    assert len(out) == 1
    d = out[0]
    assert d['impl'] == 'reference' and d['metric'] == 'routing Gcell-updates/sec'
    assert d['unit'] == 'Gcell-updates/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 2 and d['n_gpus'] == 1
    # the unmodified reference when it is staged (here: /root/reference), else the C port
    from oracle import ref_harness as rh
    if rh.reference_available():
        assert d['cpu_baseline']['kind'] == 'reference' and d['cpu_baseline']['cores'] == 1
        assert d['cpu_baseline']['port_all_threads']['kind'] == 'port'
    else:
        assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    for k in ('workload', 'grid', 'rows_per_gpu', 'partition', 'l2'):   # same keys as the GPU arm
        assert k in d['config'], k
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and d['vs_baseline'] is None
    # a non-zero rank of a torchrun launch exits 0 without printing or working
    assert _run_bench(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                      {'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}) == []


def test_tile_walkers_on_the_host(tmp_path):
    """The division-free tile walkers of the routing kernels (tfb_pk.cuh TileWalk, tile_of:
    __host__ __device__) compiled as HOST code and checked against each other: every tile
    exactly once for any grid shape / CTA count, forward, reversed (serpentine) and in the
    boundary-first order of the peer-memory halo."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    exe = str(tmp_path / 'tilewalk_test')
    src = os.path.join(ROOT, 'tests', 'csrc', 'tilewalk_host_test.cu')
    r = subprocess.run([nvcc, '-std=c++17', '-O1', '-gencode', 'arch=compute_100a,code=sm_100a',
                        '-I', os.path.join(ROOT, 'include'), '-x', 'cu', src, '-o', exe],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith('ok'), r.stdout[-500:]
