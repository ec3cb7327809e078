"""Kinematic-wave GPU Channels component (drop-in for the reference's
``components/channels_kinematic_wave.py``, class ``channels_component`` :38).

The reference subclass only supplies ``update_velocity`` (Manning :3960 or law
of the wall :4005 with the BED slope, channels_kinematic_wave.py:84-127); here
that choice is a kernel template flag, selected from the ``cfg_extension``
attribute exactly as ``set_computed_input_vars`` does (channels_base.py:1000-1004).
"""
from . import channels_base


class channels_component(channels_base.channels_component):

    _att_map = {
        'model_name': 'Channels_Kinematic_Wave', 'version': '3.6-b200',
        'author_name': 'b200-topoflow-routing', 'grid_type': 'uniform',
        'time_step_type': 'fixed', 'step_method': 'explicit',
        'comp_name': 'ChannelsKinWave', 'model_family': 'TopoFlow',
        'cfg_template_file': 'Channels_Kinematic_Wave.cfg.in',
        'cfg_extension': '_channels_kinematic_wave.cfg',
        'cmt_var_prefix': '/ChannelsKinWave/Input/Var/',
        'dialog_title': 'Channels: Kinematic Wave Parameters',
        'time_units': 'seconds'}

    def get_component_name(self):
        return 'TopoFlow_Channels_Kinematic_Wave'
