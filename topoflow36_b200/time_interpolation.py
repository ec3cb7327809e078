"""EMELI's time interpolation of provided variables, aware of device consumers
(SURVEY.md 8f-4).

Same classes, methods and numerics as the reference's
``framework/time_interpolation.py`` (``time_interp_data`` :27-141,
``time_interpolator`` :145-716): between two updates of a provider the framework
hands its users ``a * time + b`` with ``a = (v2 - v1) / (t2 - t1)``,
``b = v2 - a * t2`` -- and the step function ``b = v2`` when any cell did not
change (``dv.min() == 0``, :118-132) or with ``method='None'``.

What is added: ``get_values(..., device=dev)`` evaluates the same expression on
the GPU from device copies of ``a`` and ``b`` that are uploaded once per PROVIDER
update, so a device user (the GPU Channels component) that runs many steps per
provider step gets a CUDA tensor every step without a host->device copy of the
grid (``emeli_dropin.gpu_framework`` makes EMELI ask that way).  fp64 multiply and
add are the same on both sides, so the tensors equal the numpy values bit for
bit (tests/test_time_interp.py).
"""
import numpy as np


class time_interp_data(object):
    """(v1, t1), (v2, t2) and the line through them for ONE shared variable."""

    def __init__(self, v1=None, t1=None, long_var_name=None):
        # update() first moves (v2, t2) to (v1, t1): start with both equal (:40-52)
        self.v2 = v1
        self.t2 = t1
        self.long_var_name = long_var_name
        self.v1 = v1.copy()
        self.t1 = t1.copy()
        self._dev = None               # (device, a, b) mirrors, rebuilt after update()

    def update(self, v2=None, t2=None):
        """New (v2, t2) from the provider; values may be 0-D .. 3-D (:63-141)."""
        self.t1 = self.t2.copy()
        try:
            self.v1[:] = self.v2.copy()
        except Exception:              # 0-D arrays cannot be sliced
            self.v1 = self.v2.copy()
        self.t2 = t2
        try:
            self.v2[:] = v2.copy()
        except Exception:
            self.v2 = v2.copy()
        dv_min = np.abs(v2 - self.v1).min()
        if (dv_min != 0) and (t2 != self.t1):
            self.a = (v2 - self.v1) / (t2 - self.t1)
            self.b = v2 - (self.a * t2)
        else:
            # some cell did not change (or no time passed): step function
            self.a = np.float64(0)
            self.b = v2
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


class time_interpolator(object):

    secs_per_min = 60
    secs_per_hour = 60 * secs_per_min
    secs_per_day = 24 * secs_per_hour
    secs_per_year = 365 * secs_per_day
    secs_per_month = secs_per_year / 12
    device_aware = True

    def __init__(self, comp_set, comp_names, vars_provided, method='Linear'):
        self.comp_set = comp_set
        self.provider_comp_list = comp_names
        self.vars_provided = vars_provided
        self.interpolation_method = method.title()

    def _time_of(self, bmi):
        return self.convert_time_units(bmi.get_current_time(), bmi.get_time_units())

    def initialize(self, SILENT=False):
        """:163-227.  'None': every provider is advanced once.  'Linear': (v1, t1) is
        taken, the provider advanced, (v2, t2) taken."""
        self.SILENT = SILENT
        method = self.interpolation_method
        if not SILENT:
            print('Time interpolation method = ' + method)
            print()
        if method == 'None':
            self.time_interp_vars = None
            for comp_name in self.provider_comp_list:
                self.comp_set[comp_name].update(-1.0)
            return
        if method == 'Linear':
            self.time_interp_vars = dict()
            for comp_name in self.provider_comp_list:
                bmi = self.comp_set[comp_name]
                t1 = self._time_of(bmi)
                for long_var_name in self.vars_provided[comp_name]:
                    v1 = bmi.get_value_ptr(long_var_name)
                    self.time_interp_vars[long_var_name] = time_interp_data(
                        v1=v1, t1=t1, long_var_name=long_var_name)
                bmi.update(-1.0)
                t2 = self._time_of(bmi)
                for long_var_name in self.vars_provided[comp_name]:
                    v2 = bmi.get_value_ptr(long_var_name)
                    self.time_interp_vars[long_var_name].update(v2, t2)

    def _refresh(self, bmi, comp_name):
        t2 = self._time_of(bmi)
        for long_var_name in self.vars_provided[comp_name]:
            v2 = bmi.get_value_ptr(long_var_name)
            self.time_interp_vars[long_var_name].update(v2, t2)

    def update(self, comp_name, time):
        """:329-388.  Advance a provider whose clock is behind the framework time."""
        bmi = self.comp_set[comp_name]
        if time <= self._time_of(bmi):
            return
        if hasattr(bmi, 'comp_status') and bmi.comp_status.lower() == 'disabled':
            return
        if self.interpolation_method == 'None':
            bmi.update(-1.0)
        elif self.interpolation_method == 'Linear':
            bmi.update(-1.0)
            self._refresh(bmi, comp_name)

    def update2(self, comp_name):
        """:390-438.  The caller has already advanced the component."""
        if self.interpolation_method == 'None':
            return
        bmi = self.comp_set[comp_name]
        if bmi.get_status() == 'initialized':
            return
        if self.interpolation_method == 'Linear':
            self._refresh(bmi, comp_name)

    def update_all(self, time):
        for comp_name in self.provider_comp_list:
            self.update(comp_name, time)

    def get_values(self, long_var_name, comp_name, time, device=None):
        """:591-690.  `device` (a torch.device): grids come back as CUDA tensors."""
        bmi = self.comp_set[comp_name]
        if bmi.get_status() in ('initialized', 'finalized'):
            return bmi.get_value_ptr(long_var_name)
        if self.interpolation_method == 'None':
            bmi_time = bmi.get_current_time()
            if time > bmi_time:
                print('--------------------------------------------')
                print(' WARNING: (in time_interpolation.py)')
                print('     time > bmi_time in get_values().')
                print('     time, bmi_time =', time, bmi_time)
                print('     comp_name =', comp_name)
                print('--------------------------------------------')
                print(' ')
            return bmi.get_value_ptr(long_var_name)
        if self.interpolation_method == 'Linear':
            i_vars = self.time_interp_vars[long_var_name]
            if time > i_vars.t2:
                print('#######################################')
                print(' ERROR: time > t2 in get_values().')
                print('        time, t2 =', time, i_vars.t2)
                print('        comp_name =', comp_name)
                print('#######################################')
                print(' ')
            if device is not None:
                return i_vars.value_on(device, time)
            return i_vars.value(time)

    def convert_time_units(self, in_time, in_units):
        """To seconds (:692-716; the reference's 'hours' branch misses a `self.`)."""
        if in_units in ('years', 'y'):
            return in_time * self.secs_per_year
        if in_units == 'months':
            return in_time * self.secs_per_month
        if in_units in ('days', 'd'):
            return in_time * self.secs_per_day
        if in_units in ('hours', 'h'):
            return in_time * self.secs_per_hour
        if in_units in ('minutes', 'm'):
            return in_time * self.secs_per_min
        return in_time.copy()
