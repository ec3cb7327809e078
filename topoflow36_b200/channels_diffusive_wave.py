"""Diffusive-wave GPU Channels component (drop-in for the reference's
``components/channels_diffusive_wave.py``, class ``channels_component`` :38).

Velocity uses the FREE-SURFACE slope ``S_bed + (d - d[parent]) / ds``
(channels_base.py:2583-2618, channels_diffusive_wave.py:85-139); the kernel
re-derives the parent's new depth from the parent's pre-step state, so the step
stays one launch.
"""
from . import channels_base


class channels_component(channels_base.channels_component):

    _att_map = {
        'model_name': 'Channels_Diffusive_Wave', 'version': '3.6-b200',
        'author_name': 'b200-topoflow-routing', 'grid_type': 'uniform',
        'time_step_type': 'fixed', 'step_method': 'explicit',
        'comp_name': 'ChannelsDiffWave', 'model_family': 'TopoFlow',
        'cfg_template_file': 'Channels_Diffusive_Wave.cfg.in',
        'cfg_extension': '_channels_diffusive_wave.cfg',
        'cmt_var_prefix': '/ChannelsDiffWave/Input/Var/',
        'dialog_title': 'Channels: Diffusive Wave Parameters',
        'time_units': 'seconds'}

    def get_component_name(self):
        return 'TopoFlow_Channels_Diffusive_Wave'
