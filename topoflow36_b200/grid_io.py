"""RiverTools grid files: RTI (text header) + RTG (flat binary grid).

Own implementation of the two on-disk formats the routing path reads at
``initialize()`` time (never per step).  Behaviour mirrors the reference's
``utils/rti_files.py:read_info`` (:291-385), ``make_info`` (:405-470),
``write_info`` (:524-600) and ``utils/rtg_files.py:read_grid`` (:478-534),
``write_grid`` (:536-575): values are read in file order from the
``label: value`` lines (';' lines are comments), grids are row-major with row
0 = north, and are byte-swapped when the header's byte order differs from the
machine's.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np

_NUMPY_TYPE = {'BYTE': 'uint8', 'INTEGER': 'int16', 'LONG': 'int32',
               'LONG64': 'int64', 'FLOAT': 'float32', 'DOUBLE': 'float64'}
_BPE = {'BYTE': 1, 'INTEGER': 2, 'LONG': 4, 'LONG64': 8, 'FLOAT': 4, 'DOUBLE': 8}

# (attribute, kind) in the order the values appear in an RTI file
_RTI_FIELDS = (
    ('grid_file', 'str'), ('data_source', 'str'),
    ('ncols', 'int'), ('nrows', 'int'),
    ('data_type', 'STR'), ('byte_order', 'STR'),
    ('pixel_geom', 'int'), ('xres', 'f64'), ('yres', 'f64'), ('zres', 'f32'),
    ('z_units', 'STR'),
    ('y_south_edge', 'f64'), ('y_north_edge', 'f64'),
    ('x_east_edge', 'f64'), ('x_west_edge', 'f64'), ('box_units', 'STR'),
    ('gmin', 'f32'), ('gmax', 'f32'), ('UTM_zone', 'STR'))


def machine_byte_order():
    return 'MSB' if sys.byteorder == 'big' else 'LSB'


def _finish_info(info):
    info.n_pixels = int(info.ncols) * int(info.nrows)
    info.bpe = _BPE[info.data_type]
    info.grid_size = info.bpe * info.n_pixels
    info.SWAP_ENDIAN = (machine_byte_order() != info.byte_order)
    info.nx = info.ncols
    info.ny = info.nrows
    return info


def read_rti(path):
    """Parse an RTI header.  Returns None if the file is missing/invalid
    (the reference's convention, rti_files.py:291-305)."""
    if not path.lower().endswith('.rti'):
        path = os.path.splitext(path)[0] + '.rti'
    if not os.path.exists(path):
        return None
    with open(path, 'r') as fh:
        first = fh.readline().upper().strip()
        if first != 'RIVERTOOLS INFO FILE':
            return None
        values = []
        for line in fh:
            if line[:1] == ';':
                continue
            pos = line.find(':')
            if pos == -1:
                continue
            values.append(line[pos + 1:].strip())
            if len(values) == len(_RTI_FIELDS):
                break
    if len(values) < len(_RTI_FIELDS):
        return None
    info = SimpleNamespace()
    for (name, kind), s in zip(_RTI_FIELDS, values):
        if kind == 'str':
            v = s
        elif kind == 'STR':
            v = s.upper()
        elif kind == 'int':
            v = int(float(s))
        elif kind == 'f32':
            v = np.float32(float(s))
        else:
            v = np.float64(float(s))
        setattr(info, name, v)
    if info.UTM_zone == '-1':
        info.UTM_zone = 'unknown'
    return _finish_info(info)


def make_rti(ncols, nrows, xres, yres, data_type='FLOAT', byte_order=None,
             pixel_geom=1, grid_file='grid.rtg', y_south_edge=0.0,
             x_west_edge=0.0, gmin=-9999.0, gmax=-9999.0):
    """Header for a new grid: fixed-length pixels (pixel_geom 1: xres, yres [m]) or
    fixed-angle pixels (pixel_geom 0: xres, yres [arcsec], edges in degrees)."""
    unit = 1.0 if int(pixel_geom) == 1 else 1.0 / 3600.0
    info = SimpleNamespace(
        grid_file=grid_file, data_source='Unknown', ncols=int(ncols),
        nrows=int(nrows), data_type=data_type.upper(),
        byte_order=(byte_order or machine_byte_order()),
        pixel_geom=int(pixel_geom), xres=np.float64(xres), yres=np.float64(yres),
        zres=np.float32(1.0), z_units='METERS',
        y_south_edge=np.float64(y_south_edge),
        y_north_edge=np.float64(y_south_edge + nrows * yres * unit),
        x_east_edge=np.float64(x_west_edge + ncols * xres * unit),
        x_west_edge=np.float64(x_west_edge),
        box_units='METERS' if int(pixel_geom) == 1 else 'DEGREES',
        gmin=np.float32(gmin), gmax=np.float32(gmax), UTM_zone='UNKNOWN')
    return _finish_info(info)


def write_rti(path, info):
    lines = [
        'RiverTools Info File', '',
        ';--------------------', ';Description of Grid', ';--------------------',
        'DEM filename:  %s' % info.grid_file,
        'DEM source:    %s' % info.data_source, '',
        ';----------------------', ';Grid dimensions, etc.',
        ';----------------------',
        'Number of columns:  %d' % info.ncols,
        'Number of rows:     %d' % info.nrows,
        'Data type:          %s' % info.data_type,
        'Byte order:         %s' % info.byte_order, '',
        ';--------------------', ';Pixel geometry info', ';--------------------',
        'Pixel geometry code:  %d' % info.pixel_geom,
        'Pixel x-resolution:   %.8f' % info.xres,
        'Pixel y-resolution:   %.8f' % info.yres,
        'Pixel z-resolution:   %.8f' % info.zres,
        'Z-resolution units:   %s' % info.z_units, '',
        ';--------------------------------------------------',
        ';Bounding box coordinates (degrees lat/lon or UTM)',
        ';--------------------------------------------------',
        'South edge y-coordinate:  %.12f' % info.y_south_edge,
        'North edge y-coordinate:  %.12f' % info.y_north_edge,
        'East  edge x-coordinate:  %.12f' % info.x_east_edge,
        'West  edge x-coordinate:  %.12f' % info.x_west_edge,
        'Measurement units:        %s' % info.box_units, '',
        ';----------------------', ';Min and max elevation',
        ';----------------------',
        'Minimum elevation:  %s' % repr(float(info.gmin)),
        'Maximum elevation:  %s' % repr(float(info.gmax)), '',
        ';---------', ';UTM zone', ';---------',
        'UTM zone: %s' % info.UTM_zone, '']
    with open(path, 'w') as fh:
        fh.write('\n'.join(lines))


def write_output_rti(grid_path, info, data_type='FLOAT'):
    """Side-car header of an OUTPUT grid stack: what rts_files.open_new_file (:255-268)
    derives from the data set's RTI -- same geometry, same byte order, the stack's data
    type, `TopoFlow 3.0` as source, min / max unknown (-9999) -- laid out with the labels
    and number formats of rti_files.write_info (:430-540), which both the reference and
    read_rti() parse."""
    f = lambda v: ('%22.12f' % float(v)).strip()
    lines = [
        'RiverTools Info File', '',
        ';--------------------', ';Description of Grid', ';--------------------',
        'Grid filename:  ' + str(grid_path),
        'Grid source:    TopoFlow 3.0', ' ',
        ';----------------------', ';Grid dimensions, etc.', ';----------------------',
        'Number of columns:  %d' % info.ncols,
        'Number of rows:     %d' % info.nrows,
        'Data type:          ' + data_type.upper(),
        'Byte order:         ' + info.byte_order, '',
        ';--------------------', ';Pixel geometry info', ';--------------------',
        'Pixel geometry code:  %d' % int(info.pixel_geom),
        'Pixel x-resolution:   ' + f(info.xres),
        'Pixel y-resolution:   ' + f(info.yres),
        'Pixel z-resolution:   ' + f(info.zres),
        'Z-resolution units:   ' + info.z_units, '',
        ';--------------------------------------------------',
        ';Bounding box coordinates (degrees lat/lon or UTM)',
        ';--------------------------------------------------',
        'South edge y-coordinate:  ' + f(info.y_south_edge),
        'North edge y-coordinate:  ' + f(info.y_north_edge),
        'East  edge x-coordinate:  ' + f(info.x_east_edge),
        'West  edge x-coordinate:  ' + f(info.x_west_edge),
        'Measurement units:        ' + info.box_units, '',
        ';------------------', ';Min and max value', ';------------------',
        'Minimum value:  ' + f(-9999.0),
        'Maximum value:  ' + f(-9999.0), '',
        ';---------', ';UTM zone', ';---------',
        'UTM zone: ' + str(info.UTM_zone), '', '']
    with open(os.path.splitext(grid_path)[0] + '.rti', 'w') as fh:
        fh.write('\n'.join(lines))


def read_rtg(path, info, rtg_type=None):
    """Read a whole grid as (nrows, ncols) in the file's own dtype."""
    dtype = _NUMPY_TYPE[(rtg_type or info.data_type).upper()]
    grid = np.fromfile(path, count=info.n_pixels, dtype=dtype)
    if grid.size != info.n_pixels:
        raise IOError('RTG file %s holds %d values, expected %d'
                      % (path, grid.size, info.n_pixels))
    grid = grid.reshape(info.nrows, info.ncols)
    if info.SWAP_ENDIAN:
        grid = grid.byteswap()
    return grid


def read_rtg_rows(path, info, row_lo, row_hi, rtg_type=None):
    """Rows [row_lo, row_hi) of a grid file (clipped to the grid); rows outside the
    grid are returned as copies of the nearest valid row.  Used by slab-sharded
    components so that every rank touches only its own part of the file."""
    dtype = np.dtype(_NUMPY_TYPE[(rtg_type or info.data_type).upper()])
    lo, hi = max(0, int(row_lo)), min(int(info.nrows), int(row_hi))
    if os.path.getsize(path) < info.n_pixels * dtype.itemsize:
        raise IOError('RTG file %s is shorter than %d values' % (path, info.n_pixels))
    mm = np.memmap(path, dtype=dtype, mode='r', shape=(info.nrows, info.ncols))
    part = np.array(mm[lo:hi])
    del mm
    if info.SWAP_ENDIAN:
        part = part.byteswap()
    top, bot = lo - int(row_lo), int(row_hi) - hi
    if top or bot:
        part = np.concatenate([np.repeat(part[:1], top, 0), part, np.repeat(part[-1:], bot, 0)], 0)
    return part


def write_rtg(path, grid, info, rtg_type='FLOAT'):
    rtg_type = rtg_type.upper()
    if rtg_type == 'BYTE':
        out = np.uint8(np.int16(grid))
    else:
        out = np.asarray(grid).astype(_NUMPY_TYPE[rtg_type])
    if info.SWAP_ENDIAN:
        out = out.byteswap()
    out.tofile(path)


def pixel_sizes_by_row(info):
    """(dx, dy, dd) per row in metres, float64 -- the values the reference's
    ``utils/pixels.py:get_sizes_by_row`` (:71-171, METERS=True) produces.
    Fixed-length pixels (geom 1) are constant; fixed-angle pixels (geom 0,
    arcseconds) depend on the latitude of each row."""
    ny = int(info.nrows)
    if int(info.pixel_geom) == 0:
        yresdeg = info.yres / np.float64(3600)
        ycoords = np.arange(ny, dtype='int16')
        lats = info.y_north_edge - (yresdeg * (ycoords + np.float64(1)))
        mpd_lon, mpd_lat = _meters_per_degree(lats)
        dx = info.xres / np.float64(3600) * mpd_lon
        dy = info.yres / np.float64(3600) * mpd_lat
        dd = np.sqrt(dx ** 2 + dy ** 2)
    else:
        dx0 = np.float64(info.xres)
        dy0 = np.float64(info.yres)
        dd0 = np.sqrt(dx0 ** 2 + dy0 ** 2)
        z = np.zeros(ny, dtype='float64')
        dx, dy, dd = z + dx0, z + dy0, z + dd0
    return dx, dy, dd


def _meters_per_degree(lat_deg, a_radius=6378137.0, b_radius=6356752.3):
    """Ellipsoid arc lengths per degree of longitude / latitude
    (reference: utils/pixels.py:236-330, WGS-84 radii)."""
    a = np.float64(a_radius)
    b = np.float64(b_radius)
    flat = (a - b) / a
    ecc = np.sqrt((np.float64(2) * flat) - flat ** np.float64(2))
    dtor = np.pi / np.float64(180)
    lat = np.asarray(lat_deg, dtype='float64') * dtor
    es2 = (ecc * np.sin(lat)) ** np.float64(2)
    # Snyder (1987) p.25: prime-vertical radius N and meridional radius M
    N = a * np.cos(lat) / np.sqrt(np.float64(1) - es2)
    M = a * (np.float64(1) - ecc ** np.float64(2)) / (np.float64(1) - es2) ** np.float64(1.5)
    return dtor * N, dtor * M


def pixel_area(info):
    """``da``: 0-D-like float64 for fixed-length pixels, (ny, nx) grid for
    geographic ones (utils/pixels.py:173-233)."""
    dx, dy, _ = pixel_sizes_by_row(info)
    if int(info.pixel_geom) == 1:
        return np.float64(dx[0] * dy[0])
    return np.repeat(dx * dy, info.ncols).reshape(info.nrows, info.ncols)
