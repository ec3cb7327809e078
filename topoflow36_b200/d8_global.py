"""GPU ``d8_component``: the D8 toolkit behind the reference's component API.

Mirrors the interface of the reference's ``components/d8_base.py`` (class
``d8_component`` :98, ``initialize`` :251-356, ``update`` :358-416) and
``components/d8_global.py`` (the whole-grid method: ``update_flow_grid`` :99,
``update_parent_ID_grid`` :578, ``update_noflow_IDs`` :843,
``update_flow_width_grid`` :873, ``update_flow_length_grid`` :960,
``update_area_grid`` :1029, ``update_slope_grid`` :1334) and reads the same
``<case>_d8_global.cfg``.  All arithmetic runs in the tfb_d8_* CUDA kernels
(``D8Toolkit``); host-facing attributes (``d8_grid``, ``A``, ``ds``, ``dw``,
``S``, ``parent_ID_grid``, ``noflow_IDs``) are numpy copies fetched on access,
the device tensors are ``.toolkit.<name>``.
"""
import os

import numpy as np

from . import BMI_base, grid_io
from . import lib as L


class d8_component(BMI_base.BMI_component):

    _att_map = {
        'model_name': 'D8_Global_B200', 'version': '3.6-b200',
        'author_name': 'b200-topoflow-routing', 'grid_type': 'uniform',
        'time_step_type': 'fixed', 'step_method': 'explicit',
        'comp_name': 'D8Global', 'model_family': 'TopoFlow',
        'cfg_extension': '_d8_global.cfg', 'time_units': 'seconds'}

    _DEVICE_ATTRS = {'d8_grid': 'code', 'flow_grid': 'code', 'A': 'A', 'ds': 'ds',
                     'dw': 'dw', 'area_count': 'count'}

    def __init__(self):
        BMI_base.BMI_component.__init__(self)
        self.toolkit = None
        self.device = None
        self.layout = None            # SlabLayout when the grid is row-sharded over ranks
        self.group = None
        self._slope_dev = None
        self._pid_dev = None

    # device results are fetched lazily; everything else is a normal attribute
    def __getattr__(self, name):
        dev = type(self)._DEVICE_ATTRS.get(name)
        tk = self.__dict__.get('toolkit')
        if dev is not None and tk is not None:
            return getattr(tk, dev).cpu().numpy()
        raise AttributeError(name)

    def initialize(self, cfg_file=None, mode='nondriver', SILENT=True, REPORT=False):
        self.SILENT, self.REPORT = SILENT, REPORT
        self.status = 'initializing'
        self.mode, self.cfg_file = mode, cfg_file
        self.initialize_config_vars()
        self.initialize_time_vars()
        for k, v in (('LINK_FLATS', 1), ('A_units', 'km^2'), ('FILL_PITS_IN_Z0', 0),
                     ('LR_PERIODIC', 0), ('TB_PERIODIC', 0)):
            if not hasattr(self, k):
                setattr(self, k, v)
        if int(self.LR_PERIODIC) or int(self.TB_PERIODIC):
            raise NotImplementedError('periodic D8 boundaries are not part of the GPU path')
        if not hasattr(self, 'DEM_file'):
            self.DEM_file = self.site_prefix + '_DEM.rtg'
        self.DEM_file = self.topo_directory + os.path.basename(self.DEM_file) \
            if not os.path.isabs(self.DEM_file) else self.DEM_file
        if self.layout is not None and self.layout.world_size > 1:
            self._initialize_slab()
            self.status = 'initialized'
            return
        self.read_DEM()
        # FILL_PITS_IN_Z0 is read but never acted on by the reference's d8_global (only
        # d8_local.py:218 honours it); same here.  Fill explicitly with prep.fill_pits().
        dx, dy, dd = grid_io.pixel_sizes_by_row(self.rti)
        # cast to float32 as the reference does (d8_base.py:515-518)
        self.dx, self.dy, self.dd = np.float32(dx), np.float32(dy), np.float32(dd)
        from .d8 import D8Toolkit
        self.toolkit = D8Toolkit(self.DEM, link_flats=bool(int(self.LINK_FLATS)),
                                 nodata=getattr(self, 'nodata', -9999.0),
                                 dx=self.dx, dy=self.dy, dd=self.dd, da=self.da,
                                 A_units=str(self.A_units), device=self.device)
        self.status = 'initialized'

    def _initialize_slab(self):
        """Row-sharded D8: this rank reads its rows (+ halo) of the DEM and runs the
        distributed flow-grid construction of slab.slab_d8 (global Jacobi flat
        linking) and the cross-slab contributing-area pass of slab.slab_d8_area."""
        if int(self.rti.pixel_geom) != 1:
            raise NotImplementedError('slab mode needs fixed-length pixels')
        if self.rti.data_type != 'FLOAT':
            raise L.TfbError('GPU D8 kernels take a FLOAT (float32) DEM')
        from .slab import slab_d8
        lay = self.layout
        hd = lay.halo + 1
        DEM = grid_io.read_rtg_rows(self.DEM_file, self.rti, lay.row0 - hd, lay.row0 + lay.ny + hd)
        self.slab = slab_d8(DEM, lay, hd, link_flats=bool(int(self.LINK_FLATS)),
                            xres=float(self.rti.xres), yres=float(self.rti.yres),
                            nodata=getattr(self, 'nodata', -9999.0), kinematic=False,
                            device=self.device, group=self.group)
        self.n_flat_link_reps = self.slab['n_reps']
        # contributing area (and exact cell counts) of the owned rows: the topological walk
        # continues across the slab boundaries (slab.slab_d8_area), bit-identical to one GPU
        from .slab import slab_d8_area
        unit = 1.0e6 if str(self.A_units).lower().startswith('km') else 1.0
        A, cnt, rounds = slab_d8_area(self.slab['code'], lay, float(self.da) / unit, self.group)
        self.slab['A'], self.slab['count'], self.slab['area_rounds'] = A, cnt, rounds

    def read_DEM(self):
        if self.rti.data_type != 'FLOAT':
            raise L.TfbError('GPU D8 kernels take a FLOAT (float32) DEM; %s is %s'
                             % (self.DEM_file, self.rti.data_type))
        self.DEM = grid_io.read_rtg(self.DEM_file, self.rti)

    def update(self, time=None, REPORT=False, SILENT=True):
        """d8_base.update (:358-416): flow grid, parent IDs, ds/dw, area."""
        self.status = 'updating'
        if self.toolkit is None:              # slab mode: everything was done in initialize()
            self.status = 'updated'
            return
        self.toolkit.update(with_counts=True)
        self._slope_dev = self._pid_dev = None
        self.n_flat_link_reps = self.toolkit.n_reps
        self.status = 'updated'

    # ---- individual stages, named as in d8_global.py ---------------------
    def update_flow_grid(self, SILENT=True, REPORT=False):
        self.toolkit.update_flow_grid()

    def update_flow_width_grid(self, SILENT=True, REPORT=False):
        self.toolkit.update_ds_dw()

    update_flow_length_grid = update_flow_width_grid

    def update_area_grid(self, SILENT=True, REPORT=False):
        self.toolkit.update_area(with_counts=True)

    def update_parent_ID_grid(self, SILENT=True):
        self._pid_dev = self.toolkit.parent_ids()

    @property
    def parent_ID_grid(self):
        if self._pid_dev is None:
            self.update_parent_ID_grid()
        return self._pid_dev.cpu().numpy()

    @property
    def parent_IDs(self):
        return divmod(self.parent_ID_grid, self.nx)

    def update_noflow_IDs(self, REPORT=False):
        code = self.toolkit.code.cpu().numpy()
        self.noflow_IDs = np.where(code <= 0)
        self.noflow_ID_count = self.noflow_IDs[0].size

    def update_slope_grid(self, SILENT=True, REPORT=False):
        self._slope_dev = self.toolkit.slope()

    @property
    def S(self):
        if self._slope_dev is None:
            self.update_slope_grid()
        return self._slope_dev.cpu().numpy()

    def update_aspect_grid(self, REPORT=False):
        self.aspect = self.toolkit.aspect().cpu().numpy()

    def finalize(self):
        self.status = 'finalized'
