"""BMI component base for the B200 components -- the plug-in API the EMELI coupler
drives.

A from-scratch mirror of the INTERFACE of the reference's
``topoflow/utils/BMI_base.py`` class ``BMI_component`` (:189-2435): same method
names, argument meaning, status strings and CFG-file conventions, so that a
class derived from it satisfies every call ``framework/emeli.py`` makes on a
component and reads the reference's own CFG files
unchanged.  It deliberately does not import the reference (which is absent on
the GPU box); ``emeli_dropin.py`` shows how a class derived from THIS base is
handed to the reference's EMELI.

Conventions kept from the reference:
  * scalars are 0-D numpy arrays ("mutable scalars", BMI_base.py:2230-2294) that
    are updated in place so other components holding a reference see changes;
  * ``set_value`` REBINDS the attribute (:634-654); ``get_value_ptr`` returns the
    attribute itself (:596-601); ``get_value`` copies into ``dest`` (:564-594);
  * CFG files are ``name | value | type | help`` lines (:1670-1938):
    ``<var>_type`` = Scalar | Grid | Time_Series | Grid_Sequence decides whether
    the following ``<var>`` line is a value or a file name (stored as
    ``<var>_file``); ``[case_prefix]`` / ``[site_prefix]`` are expanded; yes/no
    style strings become booleans;
  * ``<prefix>_path_info.cfg`` / ``<prefix>_time_info.cfg`` sit next to the
    component CFG (:1468-1668) and ``__topo/__soil/__met/__misc`` sub-folders of
    the input directory are used when they exist (:2069-2228).
"""
import os
import time

import numpy as np

from . import grid_io

_TRUE = ('yes', 'true', 'on', '1')
_FALSE = ('no', 'false', 'off', '0')
_SEC_PER = {'seconds': 1.0, 'minutes': 60.0, 'hours': 3600.0, 'days': 86400.0,
            'weeks': 604800.0}


def parse_cfg_lines(path):
    """Yield (name, value, type) of every 4-field, non-comment line."""
    with open(path, 'r') as fh:
        for line in fh:
            if line[:1] == '#':
                continue
            words = line.split('|')
            if len(words) == 4:
                yield words[0].strip(), words[1].strip(), words[2].strip()


class BMI_component(object):

    _att_map = {'model_name': 'BMI_component', 'version': '0', 'time_units': 'seconds',
                'cfg_extension': '_bmi.cfg', 'comp_name': 'BMI_component'}
    _input_var_names = []
    _output_var_names = []
    _var_name_map = {}
    _var_units_map = {}

    def __init__(self):
        self.SILENT = True
        self.DEBUG = False
        self.REPORT = False
        self.DONE = False
        self.SKIP_ERRORS = False
        self.status = 'created'            # OpenMI 2.0 status strings
        self.in_directory = None
        self.out_directory = None
        self.site_prefix = None
        self.case_prefix = None
        self.cfg_prefix = None
        self.comp_status = 'Enabled'
        self.OVERWRITE_OK = False
        self.mode = 'nondriver'
        self.cfg_file = None

    # ------------------------------------------------------------- BMI: info
    def get_bmi_version(self):
        return '2.1'

    def get_component_name(self):
        return self.get_attribute('comp_name')

    def get_attribute(self, att_name):
        try:
            return self._att_map[att_name.lower()]
        except KeyError:
            print('###################################################')
            print(' ERROR: Could not find attribute: ' + att_name)
            print('###################################################')
            return None

    def get_grid_attribute(self, long_var_name, att_name):
        amap = {'x_units': 'meters', 'y_units': 'meters', 'z_units': 'meters',
                'ellipsoid': 'None', 'datum': 'None', 'projection': 'None'}
        return amap.get(att_name.lower())

    def get_input_item_count(self):
        return len(self._input_var_names)

    def get_output_item_count(self):
        return len(self._output_var_names)

    def get_input_var_names(self):
        return self._input_var_names

    def get_output_var_names(self):
        return self._output_var_names

    def get_var_name(self, long_var_name):
        return self._var_name_map[long_var_name]

    def get_var_units(self, long_var_name):
        return self._var_units_map[long_var_name]

    def get_var_grid(self, long_var_name):
        return 0

    def get_var_type(self, long_var_name):
        v = self.get_value_ptr(long_var_name)
        return str(np.asarray(v).dtype)

    def get_var_itemsize(self, long_var_name):
        return np.dtype(self.get_var_type(long_var_name)).itemsize

    def get_var_nbytes(self, long_var_name):
        return int(np.asarray(self.get_value_ptr(long_var_name)).nbytes)

    def get_var_location(self, long_var_name=None):
        return 'node'

    def get_status(self):
        return self.status

    # ------------------------------------------------------------- BMI: time
    def get_current_time(self):
        return self.time

    def get_start_time(self):
        return np.float64(0)

    def get_end_time(self):
        return np.float64(getattr(self, 'n_steps', 1)) * np.float64(self.dt)

    def get_time_units(self):
        return self.get_attribute('time_units')

    def get_time_step(self):
        return np.float64(self.dt)

    # ----------------------------------------------------------- BMI: values
    def get_value_ptr(self, long_var_name):
        return getattr(self, self.get_var_name(long_var_name))

    def get_value(self, long_var_name, dest):
        dest[:] = np.asarray(self.get_value_ptr(long_var_name)).reshape(np.shape(dest))
        return dest

    def get_value_at_indices(self, long_var_name, dest, IDs):
        dest[:] = np.asarray(self.get_value_ptr(long_var_name)).take(IDs)
        return dest

    def set_value(self, long_var_name, value):
        setattr(self, self.get_var_name(long_var_name), value)

    def set_value_at_indices(self, long_var_name, IDs, value):
        np.asarray(self.get_value_ptr(long_var_name)).flat[IDs] = value

    # ------------------------------------------------------------ BMI: grids
    def get_grid_rank(self, grid_id=0):
        return 2

    def get_grid_size(self, grid_id=0):
        return int(self.nx) * int(self.ny)

    def get_grid_type(self, grid_id=0):
        return 'uniform_rectilinear'

    def get_grid_shape(self, grid_id, shape):
        shape[:] = (self.ny, self.nx)
        return shape

    def get_grid_spacing(self, grid_id, spacing):
        spacing[:] = (self.rti.yres, self.rti.xres)
        return spacing

    def get_grid_origin(self, grid_id, origin):
        origin[:] = (self.rti.y_south_edge, self.rti.x_west_edge)
        return origin

    # ------------------------------------------------------------ life cycle
    def initialize(self, cfg_file=None, mode='nondriver'):
        self.status = 'initializing'
        self.mode = mode
        self.cfg_file = cfg_file
        self.initialize_config_vars()
        self.initialize_basin_vars()
        self.initialize_time_vars()
        self.status = 'initialized'

    def update(self, dt=-1.0):
        self.status = 'updating'
        self.update_time(dt)
        self.status = 'updated'

    def update_frac(self, time_frac):
        dt0 = self.get_time_step()
        self.update(time_frac * dt0)

    def update_until(self, later_time):
        n = (later_time - self.get_current_time()) / self.get_time_step()
        for _ in range(int(n)):
            self.update()
        if n - int(n) > 0:
            self.update_frac(n - int(n))

    def finalize(self):
        self.status = 'finalizing'
        self.status = 'finalized'

    def run_model(self, cfg_directory=None, cfg_prefix=None, n_steps=5):
        """Stand-alone driving of one component (reference :874-948)."""
        if cfg_directory is not None and cfg_prefix is not None:
            self.cfg_directory = os.path.join(cfg_directory, '')
            self.cfg_prefix = cfg_prefix
            self.cfg_file = (self.cfg_directory + cfg_prefix +
                             self.get_attribute('cfg_extension'))
        self.initialize(cfg_file=self.cfg_file, mode='driver')
        while not self.DONE:
            self.update(-1.0)
        self.finalize()

    # ------------------------------------------------------------------ time
    def initialize_time_vars(self, units='seconds'):
        self.start_time = time.time()
        self.time_units = units.lower()
        self.time_index = np.int32(0)
        self.time = self.initialize_scalar(0, dtype='float64')
        self.DONE = False
        if not hasattr(self, 'stop_code'):
            self.stop_code = 0
        if not hasattr(self, 'n_steps'):
            self.n_steps = 1
        self.time_sec = self.initialize_scalar(0, dtype='float64')
        self.time_min = self.initialize_scalar(0, dtype='float64')
        self.last_print_time = time.time() - 100.0

    def update_time(self, dt=-1):
        """Advance the clock by self.dt (or `dt`), reference :1029-1090.  The
        0-D time arrays are updated IN PLACE."""
        self.time_index += 1
        step = self.dt if (dt == -1) else dt
        self.time += step
        fac = _SEC_PER.get(self.time_units)
        if fac is None:
            if self.time_units == 'months':
                fac = 365.0 * 86400.0 / 12.0
            elif self.time_units == 'years':
                fac = 365.0 * 86400.0
            else:
                fac = 1.0
        self.time_sec.fill(float(self.time) * fac)
        self.time_min.fill(float(self.time_sec) / 60.0)
        self.check_finished()

    def check_finished(self):
        if self.mode != 'driver':
            return
        if self.stop_code == 0:
            up = (self.time_index >= self.n_steps)
        elif self.stop_code == 1:
            up = (self.time >= self.stop_time)
        else:
            up = False
        self.DONE = bool(self.DONE or up)

    def print_time_and_value(self, value, var_name='Q_out', units_name='[m^3/s]',
                             interval=2.0):
        now = time.time()
        if now - self.last_print_time > interval:
            print('Time = %.2f [min],  %s = %.6g %s'
                  % (float(self.time_min), var_name, float(value), units_name))
            self.last_print_time = now

    # ------------------------------------------------------------ CFG files
    def _cfg_stem(self):
        return self.cfg_file.replace(self.get_attribute('cfg_extension'), '')

    def read_path_info(self):
        path = self._cfg_stem() + '_path_info.cfg'
        if not os.path.exists(path):
            print('WARNING: path info CFG file not found:')
            print('         ' + path)
            return
        for name, value, _ in parse_cfg_lines(path):
            setattr(self, name, value)
        if isinstance(self.OVERWRITE_OK, str):
            self.OVERWRITE_OK = self.OVERWRITE_OK.lower() in _TRUE
        self.set_directories()

    def read_time_info(self):
        from types import SimpleNamespace
        ti = SimpleNamespace(start_date='2020-01-01', start_time='00:00:00',
                             end_date='2020-01-02', end_time='00:00:00')
        path = self._cfg_stem() + '_time_info.cfg'
        if os.path.exists(path):
            for name, value, _ in parse_cfg_lines(path):
                setattr(ti, name, value)
        try:
            from datetime import datetime
            f = '%Y-%m-%d %H:%M:%S'
            t0 = datetime.strptime(ti.start_date + ' ' + ti.start_time, f)
            t1 = datetime.strptime(ti.end_date + ' ' + ti.end_time, f)
            ti.duration = (t1 - t0).total_seconds() / 60.0
        except ValueError:
            ti.duration = 1440.0
        ti.duration_units = 'minutes'
        ti.start_datetime = ti.start_date + ' ' + ti.start_time
        ti.end_datetime = ti.end_date + ' ' + ti.end_time
        self.time_info = ti

    def read_grid_info(self):
        rti_file = self.site_prefix + '.rti'
        self.grid_info_file = self.topo_directory + rti_file
        info = grid_io.read_rti(self.grid_info_file)
        if info is None:
            self.grid_info_file = self.out_directory + self.case_prefix + '.rti'
            info = grid_io.read_rti(self.grid_info_file)
            if info is None:
                raise IOError('RTI file %s not found in %s (nor %s.rti in %s)'
                              % (rti_file, self.topo_directory, self.case_prefix,
                                 self.out_directory))
        self.rti = info
        self.grid_info = info
        self.nx = info.ncols
        self.ny = info.nrows
        self.da = grid_io.pixel_area(info)

    def read_config_file(self):
        if not os.path.exists(self.cfg_file):
            print('WARNING: cfg_file not found:')
            print('         ' + self.cfg_file)
            return
        last_name, last_value = '', ''
        for name, value, vtype in parse_cfg_lines(self.cfg_file):
            read_scalar = read_filename = False
            p1, p2 = name.rfind('['), name.rfind(']')
            if p1 > 0 and p2 > p1:
                base, file_attr = name[:p1], name[:p1] + '_file' + name[p1:p2 + 1]
            else:
                base, file_attr = name, name + '_file'
            if last_name.startswith(base + '_type'):
                if str(last_value).lower() == 'scalar':
                    setattr(self, file_attr, '')
                    read_scalar = True
                else:
                    setattr(self, name, 0.0)            # placeholder
                    read_filename = True
            if vtype in ('double', 'float'):
                setattr(self, name, self.initialize_scalar(np.float64(value)))
            elif vtype in ('long', 'int'):
                setattr(self, name, self.initialize_scalar(np.int32(value), dtype='int32'))
            elif vtype == 'string':
                s = value
                if s[:13] == '[case_prefix]':
                    vs = self.case_prefix + s[13:]
                elif s[:13] == '[site_prefix]':
                    vs = self.site_prefix + s[13:]
                else:
                    vs = s
                if s.lower() in _TRUE:
                    setattr(self, name, True)
                elif s.lower() in _FALSE:
                    setattr(self, name, False)
                elif read_filename:
                    setattr(self, file_attr, vs)
                elif read_scalar:
                    # the shipped CFGs sometimes type a Scalar value "string"
                    # (e.g. z0val in the Treynor KW CFG); keep numbers numeric
                    try:
                        setattr(self, name, self.initialize_scalar(np.float64(vs)))
                    except ValueError:
                        setattr(self, name, vs)
                else:
                    setattr(self, name, vs)
            else:
                print('ERROR in BMI_base.read_config_file().')
                print('   Unsupported data type = ' + vtype + '.')
            last_name, last_value = name, value

    def initialize_config_vars(self, READ_GRID_INFO=True):
        self.read_path_info()
        self.read_time_info()
        if READ_GRID_INFO:
            self.read_grid_info()
        self.read_config_file()
        self.set_directories()

    def set_computed_input_vars(self):
        pass

    def set_directories(self, REPORT=False):
        if not hasattr(self, 'cfg_directory') or self.cfg_directory is None:
            d = os.path.dirname(self.cfg_file) if self.cfg_file else ''
            cfg_dir = (d or os.getcwd()) + os.sep
        else:
            cfg_dir = self.cfg_directory
        in_dir = os.path.expanduser(self.in_directory or '.')
        out_dir = os.path.expanduser(self.out_directory or '.')
        if in_dir[0] == '.':
            in_dir = os.path.abspath(cfg_dir + in_dir)
        if out_dir[0] == '.':
            out_dir = os.path.abspath(cfg_dir + out_dir)
        in_dir = os.path.join(in_dir, '')
        out_dir = os.path.join(out_dir, '')
        sub = {}
        for k in ('topo', 'soil', 'met', 'misc'):
            d = in_dir + '__' + k + os.sep
            sub[k] = d if os.path.exists(d) else in_dir
        self.cfg_directory = cfg_dir
        self.in_directory, self.out_directory = in_dir, out_dir
        self.topo_directory, self.soil_directory = sub['topo'], sub['soil']
        self.met_directory, self.misc_directory = sub['met'], sub['misc']

    def initialize_basin_vars(self):
        """Monitored-cell file (reference utils/outlets.py:59-190): 5 header
        lines, then `col row area relief` per outlet."""
        outlet_file = getattr(self, 'pixel_file', None) or (self.case_prefix + '_outlets.txt')
        self.outlet_file = self.cfg_directory + outlet_file
        if not os.path.exists(self.outlet_file):
            raise IOError('Could not find outlet_file: ' + self.outlet_file)
        cols, rows, areas, reliefs = [], [], [], []
        with open(self.outlet_file, 'r') as fh:
            for line in fh.readlines()[5:]:
                w = line.split()
                if len(w) >= 4:
                    cols.append(np.int64(np.float64(w[0])))
                    rows.append(np.int64(np.float64(w[1])))
                    areas.append(np.float64(w[2]))
                    reliefs.append(np.float64(w[3]))
        if not rows:
            raise IOError('No outlets listed in ' + self.outlet_file)
        rows, cols = np.array(rows, dtype='int64'), np.array(cols, dtype='int64')
        ids = rows * self.nx + cols
        if (ids < 0).any() or (ids >= self.rti.n_pixels).any() or (cols >= self.nx).any():
            raise ValueError('Some outlet_IDs lie outside of the DEM: ' + self.outlet_file)
        self.basin_area, self.basin_relief = areas[0], reliefs[0]
        self.n_outlets = len(rows)
        self.outlet_flat_ID = ids[0]
        self.outlet_IDs = (rows, cols)
        self.outlet_ID = (rows[0], cols[0])

    # ------------------------------------------------------------ var helpers
    def initialize_scalar(self, value=0.0, dtype='float64'):
        return np.array(value, dtype=dtype)

    def initialize_grid(self, value=0.0, dtype='float64'):
        g = np.zeros((self.ny, self.nx), dtype=dtype)
        if value != 0:
            g += value
        return g

    def initialize_var(self, var_type, value=0.0, dtype='float64'):
        if var_type.lower() in ('scalar', 'time_series'):
            return self.initialize_scalar(value, dtype)
        return self.initialize_grid(value, dtype)

    def is_scalar(self, var_name):
        return np.ndim(getattr(self, var_name)) == 0

    def is_grid(self, var_name):
        return np.ndim(getattr(self, var_name)) == 2

    def update_var(self, var_name, value, factor=1.0):
        if value is None:
            return
        if factor != 1.0:
            value = value * factor
        cur = getattr(self, var_name, None)
        if isinstance(cur, np.ndarray) and cur.ndim == 0:
            cur.fill(value)
        elif isinstance(cur, np.ndarray) and cur.shape == np.shape(value):
            cur[:] = value
        else:
            setattr(self, var_name, np.array(value, dtype='float64'))
