"""Hand the GPU Channels / D8 components to the reference's EMELI coupler.

EMELI discovers components BY NAME: ``component_repository.xml``
maps ``tf_channels_kin_wave -> module channels_kinematic_wave, class
channels_component`` (framework/component_repository.xml:6-42) and
``instantiate_comp`` runs ``from topoflow.components import <module>``
(framework/emeli.py:536-547).  ``install()`` pre-seeds ``sys.modules`` and the
``topoflow.components`` package attribute with the modules of this package, so
an UNMODIFIED providers file and repository resolve to the GPU classes.  Nothing
in the reference tree is edited; ``uninstall()`` restores the originals.

Needs the reference importable (``import topoflow``); it is only the coupler
that comes from there -- the components run on the GPU through libtfb_b200.so.
Run EMELI with ``time_interp_method='None'`` (emeli.py:893): the reference's Linear
interpolator copies every provided grid each step (time_interpolation.py:81-118).
For Linear interpolation use ``gpu_framework()``: an ``emeli.framework`` subclass
that builds the device-aware interpolator of ``topoflow36_b200.time_interpolation``
and asks it for CUDA tensors when the user component takes them.
"""
import importlib
import sys
import types

_SWAPS = {'channels_kinematic_wave': 'topoflow36_b200.channels_kinematic_wave',
          'channels_diffusive_wave': 'topoflow36_b200.channels_diffusive_wave'}
_saved = {}


def _inheriting(mod):
    """A copy of one of this package's component modules whose ``channels_component``
    also derives from the reference's ``utils.BMI_base.BMI_component`` (north star: the
    BMI API "inherited from BMI_base").  The GPU class comes first in the MRO, so every
    method it defines is the one that runs; whatever else a caller may reach for on a
    TopoFlow component (the remaining ~100 helpers of utils/BMI_base.py:189-2435) resolves
    to the reference's own code, and ``isinstance(c, BMI_base.BMI_component)`` holds."""
    from topoflow.utils import BMI_base as ref_bmi
    shim = types.ModuleType(mod.__name__ + '_ref_bmi', mod.__doc__)
    shim.__dict__.update({k: v for k, v in vars(mod).items() if not k.startswith('__')})
    gpu_cls = mod.channels_component
    shim.channels_component = type('channels_component', (gpu_cls, ref_bmi.BMI_component),
                                   {'__module__': shim.__name__, '__doc__': gpu_cls.__doc__})
    return shim


def install(modules=None, inherit=True):
    """Route EMELI's component lookup to the GPU components.  Returns the list
    of swapped reference module names.  ``inherit``: derive the installed classes from
    the reference's ``BMI_component`` as well (see ``_inheriting``)."""
    import topoflow.components as pkg          # the reference package
    done = []
    for ref_name, ours in _SWAPS.items():
        if modules is not None and ref_name not in modules:
            continue
        full = 'topoflow.components.' + ref_name
        mod = importlib.import_module(ours)
        if inherit:
            mod = _inheriting(mod)
        _saved[full] = (sys.modules.get(full), getattr(pkg, ref_name, None))
        sys.modules[full] = mod
        setattr(pkg, ref_name, mod)
        done.append(ref_name)
    return done


def uninstall():
    import topoflow.components as pkg
    for full, (old_mod, old_attr) in list(_saved.items()):
        name = full.rsplit('.', 1)[1]
        if old_mod is None:
            sys.modules.pop(full, None)
        else:
            sys.modules[full] = old_mod
        if old_attr is None:
            if hasattr(pkg, name):
                delattr(pkg, name)
        else:
            setattr(pkg, name, old_attr)
        del _saved[full]


def gpu_framework():
    """``emeli.framework`` subclass (option ii) for Linear time interpolation
    with device users: identical coupling logic, but

    * the ``time_interpolation`` module EMELI instantiates (emeli.py:1004) is the
      device-aware one of this package, and
    * ``get_required_vars`` (emeli.py:1500-1515) passes ``device=`` for users that
      expose ``input_device`` (the GPU Channels component), so their grids arrive as
      CUDA tensors computed on the device instead of fresh host arrays."""
    from topoflow.framework import emeli
    from . import time_interpolation as ti

    class framework(emeli.framework):

        def run_model(self, *args, **kwargs):
            saved = emeli.time_interpolation
            emeli.time_interpolation = ti            # the name emeli.py resolves at :1004
            try:
                return emeli.framework.run_model(self, *args, **kwargs)
            finally:
                emeli.time_interpolation = saved

        def get_required_vars(self, user_name, bmi_time):
            u_bmi = self.comp_set[user_name]
            dev = getattr(u_bmi, 'input_device', None)
            aware = getattr(self.time_interpolator, 'device_aware', False)
            for long_var_name in u_bmi.get_input_var_names():
                provider_name = self.var_providers[long_var_name][0]
                if dev is not None and aware:
                    values = self.time_interpolator.get_values(long_var_name, provider_name,
                                                               bmi_time, device=dev)
                else:
                    values = self.time_interpolator.get_values(long_var_name, provider_name,
                                                               bmi_time)
                u_bmi.set_value(long_var_name, values)

    return framework
