"""Device-resident D8 toolkit: drives the tfb_d8_* kernels of libtfb_b200.so.

Stage order follows the reference's ``d8_component.update()``
(components/d8_base.py:358-416) and ``update_flow_grid``
(components/d8_global.py:99-146): start codes -> break ties -> link flats ->
finalize to uint8 -> ds/dw -> contributing area; parent IDs and slope on
request.  Everything lives in torch CUDA tensors; results are bit-identical to
the reference's numpy arrays (tests/test_d8_gpu.py).
"""
import ctypes as C

import numpy as np
import torch

from . import lib as L


class D8Toolkit(object):

    def __init__(self, DEM, xres=30.0, yres=30.0, link_flats=True, nodata=-9999.0,
                 dx=None, dy=None, dd=None, da=None, A_units='km^2', device=None):
        if not torch.cuda.is_available():
            raise L.TfbError('D8Toolkit needs a CUDA device (no CPU path)')
        self.lib = L.load()
        self.device = torch.device(device if device is not None
                                   else 'cuda:%d' % torch.cuda.current_device())
        dem = DEM if isinstance(DEM, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(DEM))
        if dem.dtype != torch.float32:
            raise L.TfbError('D8 kernels take a float32 DEM (RTG FLOAT), got %s' % dem.dtype)
        self.DEM = dem.to(self.device).contiguous()
        self.ny, self.nx = self.DEM.shape
        self.link_flats = bool(link_flats)
        self.nodata = float(nodata)
        if dx is None:
            # per-row float32 pixel sizes (d8_base.get_pixel_dimensions :505-518)
            dx64, dy64 = np.float64(xres), np.float64(yres)
            dd64 = np.sqrt(dx64 ** 2 + dy64 ** 2)
            one = np.zeros(self.ny, dtype='float64')
            dx, dy, dd = np.float32(one + dx64), np.float32(one + dy64), np.float32(one + dd64)
        self.dx_host, self.dy_host, self.dd_host = (np.float32(dx), np.float32(dy), np.float32(dd))
        up = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(self.device)
        self.dx, self.dy, self.dd = up(self.dx_host), up(self.dy_host), up(self.dd_host)
        if da is None:
            da = np.float64(xres) * np.float64(yres)
        # pixel_area in A's units (d8_global.py:1078-1081)
        scale = 1e6 if 'km' in A_units.lower() else 1.0
        if np.ndim(da) == 0:
            self.pixel_area = float(np.float64(da) / scale) if scale != 1.0 else float(da)
            self.pixel_area_row = None
        else:
            rows = np.asarray(da)[:, 0] if np.ndim(da) == 2 else np.asarray(da)
            self.pixel_area = 0.0
            self.pixel_area_row = up(np.float64(rows) / scale if scale != 1.0 else np.float64(rows))
        z = lambda dt: torch.zeros((self.ny, self.nx), dtype=dt, device=self.device)
        self.code16 = z(torch.int16)
        self._scratch16 = None
        self.code = z(torch.uint8)
        self.ds = z(torch.float32)
        self.dw = z(torch.float32)
        self.A = z(torch.float32)
        self.count = z(torch.int64)
        self._pending = None
        self.n_reps = 0
        self.n_area_cells = 0

    def _st(self):
        return L.stream_ptr()

    # --- stages (d8_global.py:181-353, :355-372, :374-576, :139-146) -----------
    def start_codes(self):
        L.check(self.lib.tfb_d8_codes(
            self.DEM.data_ptr(), self.ny, self.nx, 0, 0, self.ny,
            float(self.dx_host[0]), float(self.dy_host[0]), float(self.dd_host[0]),
            self.nodata, 1 if self.link_flats else 0, self.code16.data_ptr(), self._st()),
            'tfb_d8_codes')

    def break_ties(self):
        L.check(self.lib.tfb_d8_break_ties(self.code16.data_ptr(), self.code16.numel(), self._st()),
                'tfb_d8_break_ties')

    def link_flats_loop(self, max_reps=0):
        if self._scratch16 is None:
            self._scratch16 = torch.empty_like(self.code16)
        n = C.c_int32(0)
        L.check(self.lib.tfb_d8_link_flats(self.code16.data_ptr(), self._scratch16.data_ptr(),
                                           self.ny, self.nx, int(max_reps), C.byref(n), self._st()),
                'tfb_d8_link_flats')
        self.n_reps = n.value

    def finalize_codes(self):
        L.check(self.lib.tfb_d8_finalize_codes(self.code16.data_ptr(), self.code.data_ptr(),
                                               self.code.numel(), self._st()), 'tfb_d8_finalize_codes')

    def update_flow_grid(self):
        self.start_codes()
        self.break_ties()
        if self.link_flats:
            self.link_flats_loop()
        self.finalize_codes()

    def update_ds_dw(self):
        L.check(self.lib.tfb_d8_ds_dw(self.code.data_ptr(), self.dx.data_ptr(), self.dy.data_ptr(),
                                      self.dd.data_ptr(), self.ny, self.nx, self.ds.data_ptr(),
                                      self.dw.data_ptr(), self._st()), 'tfb_d8_ds_dw')

    def update_area(self, with_counts=True):
        if self._pending is None:
            self._pending = torch.empty((self.ny, self.nx), dtype=torch.int32, device=self.device)
        n = C.c_int64(0)
        L.check(self.lib.tfb_d8_area(
            self.code.data_ptr(), self.ny, self.nx, self.pixel_area,
            None if self.pixel_area_row is None else self.pixel_area_row.data_ptr(),
            self.A.data_ptr(), self.count.data_ptr() if with_counts else None,
            self._pending.data_ptr(), C.byref(n), self._st()), 'tfb_d8_area')
        self.n_area_cells = n.value

    def update(self, with_counts=True):
        """d8_component.update() (d8_base.py:358-416) minus file output."""
        self.update_flow_grid()
        self.update_ds_dw()
        self.update_area(with_counts=with_counts)

    def parent_ids(self):
        pid = torch.empty((self.ny, self.nx), dtype=torch.int32, device=self.device)
        L.check(self.lib.tfb_d8_parent_ids(self.code.data_ptr(), self.ny, self.nx, pid.data_ptr(),
                                           self._st()), 'tfb_d8_parent_ids')
        return pid

    def slope(self):
        S = torch.empty((self.ny, self.nx), dtype=torch.float32, device=self.device)
        L.check(self.lib.tfb_d8_slope(self.DEM.data_ptr(), self.code.data_ptr(), self.dx.data_ptr(),
                                      self.dy.data_ptr(), self.dd.data_ptr(), self.ny, self.nx,
                                      S.data_ptr(), self._st()), 'tfb_d8_slope')
        return S

    def new_slope(self):
        """utils/new_slopes.get_new_slope_grid (:174-316): drop to the first LOWER cell
        down the flow path over the path length, fp64.  Returns (slope, n_reps)."""
        n = self.ny * self.nx
        scratch = torch.empty(2 * n + 1, dtype=torch.int32, device=self.device)
        S = torch.empty((self.ny, self.nx), dtype=torch.float64, device=self.device)
        reps = C.c_int32(0)
        L.check(self.lib.tfb_d8_new_slope(self.DEM.data_ptr(), self.code.data_ptr(),
                                          self.dx.data_ptr(), self.dy.data_ptr(), self.dd.data_ptr(),
                                          self.ny, self.nx, C.c_float(self.nodata), scratch.data_ptr(),
                                          S.data_ptr(), C.byref(reps), self._st()), 'tfb_d8_new_slope')
        return S, int(reps.value)

    def aspect(self):
        """d8_global.update_aspect_grid (:1350-1376): flow angle [rad], float32."""
        a = torch.empty((self.ny, self.nx), dtype=torch.float32, device=self.device)
        L.check(self.lib.tfb_d8_aspect(self.code.data_ptr(), self.ny * self.nx, a.data_ptr(),
                                       self._st()), 'tfb_d8_aspect')
        return a
