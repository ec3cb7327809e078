"""The device-aware time interpolator (topoflow36_b200/time_interpolation.py, SURVEY 8f-4:
subclasses of the reference's framework/time_interpolation.py classes) against the
reference's own, driven with identical mock providers (CPU), and its CUDA path against
its numpy path."""
import os
import sys

import numpy as np
import pytest

REF = '/root/reference'


class _Provider(object):
    """A minimal provider: P is REBOUND to a new array every update (as e.g.
    met_base.update_P_rain does) and changes everywhere -> interpolated; S has cells that
    do not change -> step function; Z is updated IN PLACE -> the interpolator's v2 is the
    provider's live array, so v1 == v2 and it degenerates to a step function (reference
    behaviour, :40-52 and :88-100); T is a 0-D scalar; clock in `units`."""
    comp_status = 'Enabled'

    def __init__(self, dt=1800.0, units='seconds', shape=(5, 7), seed=3):
        self.dt, self.units, self.shape = dt, units, shape
        self.rng = np.random.default_rng(seed)
        self.time = np.float64(0)
        self.status = 'initialized'
        self.P = self.rng.random(shape)
        self.S = np.zeros(shape)
        self.Z = self.rng.random(shape)
        self.T = np.array(2.5)

    def get_time_units(self):
        return self.units

    def get_current_time(self):
        return self.time

    def get_status(self):
        return self.status

    def get_value_ptr(self, name):
        return getattr(self, name)

    def update(self, dt=-1.0):
        self.status = 'updated'
        self.time = self.time + np.float64(self.dt if self.units == 'seconds' else self.dt / 60.0)
        self.P = self.P + self.rng.random(self.shape) + 0.01         # new array, every cell changes
        self.S = self.S + (self.rng.random(self.shape) > 0.5)        # some cells do not
        self.Z[:] = self.Z + self.rng.random(self.shape) + 0.01      # in place
        self.T.fill(float(self.T) + 0.3)


def _drive(module, method, units):
    prov = _Provider(units=units)
    ti = module.time_interpolator({'met': prov}, ['met'], {'met': ['P', 'S', 'Z', 'T']}, method)
    ti.initialize(SILENT=True)
    out = []
    t = 0.0
    for k in range(40):                       # a user with dt = 200 s; the provider has 1800 s
        t += 200.0
        ti.update_all(t)
        out.append([np.array(ti.get_values(v, 'met', t), dtype='float64', copy=True)
                    for v in ('P', 'S', 'Z', 'T')])
    return out


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
@pytest.mark.parametrize('method', ['Linear', 'None'])
@pytest.mark.parametrize('units', ['seconds', 'minutes'])
def test_interpolator_equals_the_reference(method, units):
    from oracle import ref_harness as rh
    rh.import_reference()
    from topoflow.framework import time_interpolation as ref
    from topoflow36_b200 import time_interpolation as ours
    want = rh.quiet_call(_drive, ref, method, units)
    got = rh.quiet_call(_drive, ours, method, units)
    for k, (w, g) in enumerate(zip(want, got)):
        for a, b in zip(w, g):
            assert np.array_equal(a, b), (method, units, k)
    if method == 'Linear':                     # the values really are interpolated
        assert not np.array_equal(got[0][0], got[1][0])      # P moves inside a provider interval
        assert np.array_equal(got[0][1], got[1][1])          # S: step function
        assert np.array_equal(got[0][2], got[1][2])          # Z: in-place provider -> step


def _ref_or_skip():
    """The device-aware classes subclass the reference's: it must be importable (the live
    tree here, the staged oracle/_ref archive on the GPU box)."""
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip('reference not available (make -C oracle _ref)')
    rh.import_reference()


def test_interpolation_hits_the_end_points():
    _ref_or_skip()
    from topoflow36_b200 import time_interpolation as ours
    d = ours.time_interp_data(v1=np.array([[1.0, 2.0]]), t1=np.float64(0), long_var_name='x')
    d.update(np.array([[3.0, 6.0]]), np.float64(10.0))
    assert np.array_equal(d.value(10.0), [[3.0, 6.0]])
    assert np.allclose(d.value(5.0), [[2.0, 4.0]])
    d.update(np.array([[3.0, 7.0]]), np.float64(20.0))       # first cell unchanged -> step
    assert np.array_equal(d.value(12.0), [[3.0, 7.0]]) and float(d.a) == 0.0
    ti = ours.time_interpolator({}, [], {}, 'linear')
    assert ti.interpolation_method == 'Linear'
    assert ti.convert_time_units(np.float64(2), 'hours') == 7200.0
    assert ti.convert_time_units(np.float64(3), 'minutes') == 180.0


@pytest.mark.gpu
def test_device_values_equal_host_values():
    import torch
    _ref_or_skip()
    from topoflow36_b200 import time_interpolation as ours
    prov = _Provider(shape=(33, 41))
    ti = ours.time_interpolator({'met': prov}, ['met'], {'met': ['P', 'S', 'Z', 'T']}, 'Linear')
    ti.initialize(SILENT=True)
    dev = torch.device('cuda', 0)
    t = 0.0
    uploads = 0
    for k in range(30):
        t += 200.0
        ti.update_all(t)
        for v in ('P', 'S', 'Z'):
            before = ti.time_interp_vars[v]._dev
            got = ti.get_values(v, 'met', t, device=dev)
            uploads += ti.time_interp_vars[v]._dev is not before
            assert got.is_cuda and got.dtype == torch.float64
            assert np.array_equal(got.cpu().numpy(), ti.get_values(v, 'met', t)), (v, k)
        assert isinstance(ti.get_values('T', 'met', t, device=dev), (np.ndarray, np.floating, float))
    assert uploads <= 3 * 5                      # once per provider step, not per user step


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present')
def test_gpu_framework_asks_for_device_values():
    """emeli_dropin.gpu_framework(): an emeli.framework subclass whose get_required_vars
    passes device= for users that expose input_device, and nothing else changes."""
    from oracle import ref_harness as rh
    rh.import_reference()
    from topoflow.framework import emeli
    from topoflow36_b200 import emeli_dropin

    calls = []

    class _TI(object):
        device_aware = True

        def get_values(self, name, provider, time, device=None):
            calls.append((name, provider, time, device))
            return np.float64(1.5)

    class _User(object):
        def __init__(self, dev):
            self.input_device = dev
            self.got = {}

        def get_input_var_names(self):
            return ['a__x', 'b__y']

        def set_value(self, name, v):
            self.got[name] = v

    F = emeli_dropin.gpu_framework()
    assert issubclass(F, emeli.framework)
    f = F(SILENT=True)
    f.time_interpolator = _TI()
    f.var_providers = {'a__x': ['met'], 'b__y': ['snow']}
    f.comp_set = {'gpu_user': _User('cuda:0'), 'cpu_user': _User(None)}
    f.get_required_vars('gpu_user', 120.0)
    f.get_required_vars('cpu_user', 120.0)
    assert calls == [('a__x', 'met', 120.0, 'cuda:0'), ('b__y', 'snow', 120.0, 'cuda:0'),
                     ('a__x', 'met', 120.0, None), ('b__y', 'snow', 120.0, None)]
    assert f.comp_set['gpu_user'].got == {'a__x': 1.5, 'b__y': 1.5}
