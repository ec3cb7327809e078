"""DEM and channel-geometry preparation on the GPU: the steps the
reference runs before D8 (depression filling) and between D8 and routing (power-law
width / roughness / bankfull-depth grids from contributing area, path-length slopes).

Function names, arguments and file conventions follow the reference's
``utils/fill_pits.py`` (``fill_pits`` :358), ``utils/parameterize.py``
(``get_grid_from_TCA`` :93) and ``utils/new_slopes.py`` (``get_new_slope_grid`` :174);
the arithmetic runs in ``tfb_fill_pits``, ``tfb_power_law_grid`` and
``tfb_d8_new_slope``.  There is no CPU path.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import grid_io
from . import lib as L


def _need_gpu(what):
    if not torch.cuda.is_available():
        raise L.TfbError('%s needs a CUDA device (no CPU path)' % what)


def _st():
    return L.stream_ptr()


# ---------------------------------------------------------------- fill_pits
def fill_pits_tensor(dem, nodata=-9999.0, closed_basin_code=-32000.0, close_nodata=False):
    """In place on a float32 CUDA tensor.  Returns (n_raised, n_rounds).
    close_nodata=False is the reference's actual behaviour (nodata cells are flooded up
    to their rim, see include/tfb.h); True keeps nodata cells closed."""
    _need_gpu('fill_pits')
    if dem.dtype != torch.float32 or not dem.is_cuda or not dem.is_contiguous() or dem.dim() != 2:
        raise L.TfbError('fill_pits_tensor takes a contiguous 2-D float32 CUDA tensor')
    lib = L.load()
    ny, nx = dem.shape
    scratch = torch.empty(int(lib.tfb_fill_pits_scratch_bytes(ny, nx)), dtype=torch.uint8,
                          device=dem.device)
    raised, rounds = C.c_int64(0), C.c_int32(0)
    with torch.cuda.device(dem.device):
        L.check(lib.tfb_fill_pits(dem.data_ptr(), ny, nx, C.c_float(nodata),
                                  C.c_float(closed_basin_code), int(bool(close_nodata)),
                                  scratch.data_ptr(),
                                  C.byref(raised), C.byref(rounds), _st()), 'tfb_fill_pits')
    return int(raised.value), int(rounds.value)


def fill_pits(DEM, DEM_type='FLOAT', nx=None, ny=None, nodata=np.float32(-9999),
              USE_64_BITS=False, SILENT=True, close_nodata=False):
    """utils/fill_pits.fill_pits (:358): fills the depressions of ``DEM`` IN PLACE
    (non-finite values become ``nodata``, as in get_start_pixels :232-235).
    ``DEM``: (ny, nx) float32 numpy array or CUDA tensor."""
    if str(DEM_type).upper() != 'FLOAT':
        raise L.TfbError('GPU fill_pits takes a FLOAT (float32) DEM, got %s' % DEM_type)
    if isinstance(DEM, torch.Tensor):
        n_raised, n_rounds = fill_pits_tensor(DEM, float(nodata), close_nodata=close_nodata)
    else:
        _need_gpu('fill_pits')
        if DEM.dtype != np.float32 or DEM.ndim != 2:
            raise L.TfbError('GPU fill_pits takes a 2-D float32 DEM, got %s' % DEM.dtype)
        dev = torch.as_tensor(np.ascontiguousarray(DEM)).cuda()
        n_raised, n_rounds = fill_pits_tensor(dev, float(nodata), close_nodata=close_nodata)
        DEM[...] = dev.cpu().numpy()
    if not SILENT:
        print('Raised  pixels = ' + str(n_raised))
        print('Relaxation rounds = ' + str(n_rounds))
        print('Finished with fill_pits().')
    return n_raised


# ---------------------------------------------------------------- power-law grids
def power_law_params(A, A1=None, g1=None, A2=None, g2=None, p=None):
    """(c, p, single) as utils/parameterize.get_grid_from_TCA (:121-150) derives them.
    ``single`` tells that numpy evaluates c * A**p in float32 there (p given as a Python
    float keeps the float32 of the area grid; a p computed with np.log is a float64)."""
    lib = L.load()
    rng = None

    def area_range():
        mn, mx = C.c_float(0), C.c_float(0)
        with torch.cuda.device(A.device):
            L.check(lib.tfb_area_range(A.data_ptr(), A.numel(), C.byref(mn), C.byref(mx), _st()),
                    'tfb_area_range')
        return np.float32(mn.value), np.float32(mx.value)

    if g1 is not None and g2 is not None and A1 is not None and A2 is not None:
        p = np.log(g1 / g2) / np.log(A1 / A2)
        c = g1 / (A1 ** p)
        return float(c), float(p), False
    if p is not None and g1 is not None:
        _, A1 = area_range()
        c = g1 / (A1 ** p)              # float32 ** python float -> float32
        return float(c), float(p), isinstance(c, np.float32)
    if g1 is not None and g2 is not None:
        A2, A1 = area_range()
        p = np.log(g1 / g2) / np.log(A1 / A2)
        c = g1 / (A1 ** p)
        return float(c), float(p), False
    raise ValueError('Sorry, not enough input parameters to compute c & p.')


def grid_from_TCA(A, A1=None, g1=None, A2=None, g2=None, p=None, out_dtype=None):
    """c * A**p (1.0 where A <= 0) for a float32 area grid (numpy or CUDA tensor).
    Returns a CUDA tensor: float32 when numpy's result is float32, else float64."""
    _need_gpu('grid_from_TCA')
    At = A if isinstance(A, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(A))
    if At.dtype != torch.float32:
        raise L.TfbError('area grid must be float32 (RTG FLOAT)')
    At = At.cuda().contiguous()
    c, p, single = power_law_params(At, A1, g1, A2, g2, p)
    dt = torch.float32 if (single if out_dtype is None else out_dtype == torch.float32) else torch.float64
    out = torch.empty(At.shape, dtype=dt, device=At.device)
    lib = L.load()
    with torch.cuda.device(At.device):
        L.check(lib.tfb_power_law_grid(At.data_ptr(), At.numel(), c, p, int(single),
                                       out.data_ptr() if dt == torch.float64 else None,
                                       out.data_ptr() if dt == torch.float32 else None, _st()),
                'tfb_power_law_grid')
    return out, c, p


def get_grid_from_TCA(site_prefix=None, topo_dir=None, area_file=None, out_file=None,
                      A1=None, g1=None, A2=None, g2=None, p=None, REPORT=True):
    """File-level mirror of utils/parameterize.get_grid_from_TCA (:93-215)."""
    if site_prefix is None:
        site_prefix = 'Treynor'
    if area_file is None:
        area_file = site_prefix + '_area.rtg'
    if out_file is None:
        out_file = site_prefix + 'chan-w.rtg'
    info = grid_io.read_rti(topo_dir + site_prefix + '.rti')
    area_file2 = area_file if os.sep in area_file else topo_dir + area_file
    A = grid_io.read_rtg(area_file2, info)
    grid, c, pp = grid_from_TCA(A, A1=A1, g1=g1, A2=A2, g2=g2, p=p)
    if REPORT:
        print('Power-law parameters are:')
        print('c = ' + str(c))
        print('p = ' + str(pp))
    out_file2 = out_file if os.sep in area_file else topo_dir + out_file
    grid_io.write_rtg(out_file2, grid.float().cpu().numpy(), info)
    return out_file2


# ---------------------------------------------------------------- new slopes
def get_new_slope_grid(site_prefix=None, case_prefix=None, cfg_dir=None,
                       slope_file='TEST_newslope.rtg', device=None):
    """File-level mirror of utils/new_slopes.get_new_slope_grid (:174-316): D8 from
    <case>_d8_global.cfg, then the path-length slope, written as an RTG FLOAT grid."""
    _need_gpu('get_new_slope_grid')
    from . import d8_global
    d8 = d8_global.d8_component()
    d8.site_prefix, d8.case_prefix, d8.in_directory = site_prefix, case_prefix, cfg_dir
    d8.device = device
    d8.initialize(cfg_file=cfg_dir + case_prefix + '_d8_global.cfg', SILENT=True)
    d8.update(0)
    slope, n_reps = d8.toolkit.new_slope()
    out = slope_file if os.sep in slope_file else d8.topo_directory + slope_file
    grid_io.write_rtg(out, np.float32(slope.cpu().numpy()), d8.rti)
    return out, n_reps
