"""Device-aware time interpolation for EMELI: SUBCLASSES of the reference's
own ``framework/time_interpolation.py`` classes -- nothing of the reference is restated.

The reference interpolates every provided variable linearly between two provider updates
(``time_interp_data`` :27-141 keeps ``a = (v2 - v1) / (t2 - t1)``, ``b = v2 - a t2``;
``time_interpolator.get_values`` :591-690 returns ``a * time + b`` as a fresh host array).
For a GPU user that runs many steps per provider step this means a host->device copy of
every grid every step.  Added here, and only this:

  * ``time_interp_data.value_on(device, time)``: the same expression evaluated on the GPU
    from device copies of ``a`` and ``b`` that are uploaded once per PROVIDER update
    (``update()`` invalidates them); fp64 multiply and add are the same on both sides, so
    the tensors equal the numpy values bit for bit (tests/test_time_interp.py);
  * ``time_interpolator.get_values(..., device=dev)``: grids come back as CUDA tensors;
    ``emeli_dropin.gpu_framework`` makes EMELI ask that way for users with an
    ``input_device``;
  * ``convert_time_units``: the reference's 'hours' branch (:707) misses a ``self.``.

The reference must be importable (``import topoflow``); the classes are built on first use,
so this module can be handed to EMELI in place of the reference module
(``emeli.time_interpolation``, looked up at emeli.py:1004).
"""
import numpy as np

_classes = None


def _build():
    global _classes
    if _classes is not None:
        return _classes
    from topoflow.framework import time_interpolation as ref

    class time_interp_data(ref.time_interp_data):
        _dev = None                    # (device, a, b) mirrors, rebuilt after update()

        def update(self, v2=None, t2=None):
            ref.time_interp_data.update(self, v2=v2, t2=t2)
            self._dev = None

        def value(self, time):
            return (self.a * time) + self.b

        def value_on(self, device, time):
            """(a * time) + b as a float64 CUDA tensor; scalars (0-D) stay on the host."""
            if np.ndim(self.b) == 0:
                return self.value(time)
            import torch
            if self._dev is None or self._dev[0] != device:
                up = lambda x: (float(x) if np.ndim(x) == 0 else
                                torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(device))
                self._dev = (device, up(self.a), up(self.b))
            _, a, b = self._dev
            return (a * float(time)) + b

    class time_interpolator(ref.time_interpolator):
        device_aware = True

        def initialize(self, SILENT=False):
            ref.time_interpolator.initialize(self, SILENT=SILENT)
            # the reference's initialize() instantiates ITS data class by name (:199-208):
            # adopt the instances (same attributes, plus the device mirror)
            for v in (getattr(self, 'time_interp_vars', None) or {}).values():
                v.__class__ = time_interp_data

        def get_values(self, long_var_name, comp_name, time, device=None):
            if device is not None and self.interpolation_method == 'Linear':
                bmi = self.comp_set[comp_name]
                if bmi.get_status() not in ('initialized', 'finalized'):
                    return self.time_interp_vars[long_var_name].value_on(device, time)
            return ref.time_interpolator.get_values(self, long_var_name, comp_name, time)

        def convert_time_units(self, in_time, in_units):
            if in_units in ('hours', 'h'):
                return in_time * self.secs_per_hour
            return ref.time_interpolator.convert_time_units(self, in_time, in_units)

    _classes = dict(time_interp_data=time_interp_data, time_interpolator=time_interpolator)
    return _classes


def __getattr__(name):                 # PEP 562: the two class names resolve lazily
    if name in ('time_interp_data', 'time_interpolator'):
        return _build()[name]
    raise AttributeError(name)
