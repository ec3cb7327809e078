"""GPU parity of the D8 toolkit (C ABI) -- bit-exact against the reference
fixtures (RiverTools grids shipped with the reference + live-reference dumps)."""
import ctypes as C

import numpy as np
import pytest

from oracle import np_oracle as o

pytestmark = pytest.mark.gpu


def _d8(DEM, link_flats=True, xres=30.0, yres=30.0, pixel_area=None):
    from topoflow36_b200.d8 import D8Toolkit
    return D8Toolkit(DEM, xres=xres, yres=yres, link_flats=link_flats)


def test_resolve_table(d8_synth):
    from topoflow36_b200 import lib
    buf = (C.c_uint8 * 256)()
    lib.check(lib.load().tfb_d8_resolve_table(buf))
    assert np.array_equal(np.frombuffer(buf, dtype=np.uint8), d8_synth['resolve'])


def test_treynor_codes_and_area_bit_exact(treynor):
    t = _d8(treynor['DEM'], xres=float(treynor['xres']), yres=float(treynor['yres']))
    t.update()
    assert np.array_equal(t.code.cpu().numpy(), treynor['flow'])
    assert np.array_equal(t.A.cpu().numpy(), treynor['area'])
    assert np.array_equal(t.A.cpu().numpy(), treynor['d8_area'])
    assert int(t.count.max()) == 554


@pytest.mark.parametrize('lf', [1, 0])
def test_synth_every_stage_bit_exact(d8_synth, lf):
    g = d8_synth
    t = _d8(g['DEM'], link_flats=bool(lf))
    t.start_codes()
    assert np.array_equal(t.code16.cpu().numpy(), g['lf%d_start' % lf])
    t.break_ties()
    assert np.array_equal(t.code16.cpu().numpy(), g['lf%d_ties' % lf])
    t.update()
    assert np.array_equal(t.code.cpu().numpy(), g['lf%d_code' % lf])
    assert np.array_equal(t.ds.cpu().numpy(), g['lf%d_ds' % lf])
    assert np.array_equal(t.dw.cpu().numpy(), g['lf%d_dw' % lf])
    assert np.array_equal(t.A.cpu().numpy(), g['lf%d_area' % lf])
    assert np.array_equal(t.parent_ids().cpu().numpy(), g['lf%d_pid' % lf])
    assert np.array_equal(t.slope().cpu().numpy(), g['lf%d_slope' % lf], equal_nan=True)


def test_larger_dem_vs_oracle_with_counts():
    """1024 x 768 fractal DEM with quantised elevations (many flats): codes,
    fp32 areas and exact integer cell counts against the O(N) oracle; the count
    grid obeys conservation: every flow cell is counted once at its sink."""
    from topoflow36_b200 import synth
    ny, nx = 1024, 768
    DEM = synth.fractal_dem(ny, nx, seed=77, sigma=3.0, sy=0.003, sx=0.001)
    DEM = np.float32(np.round(DEM * 2.0) / 2.0)
    dx, dy, dd = o.pixel_dims(ny, 30.0, 30.0)
    code, _ = o.d8_flow_grid(DEM, dx, dy, dd)
    A, N = o.d8_area_kahn(code, np.float64(900.0) / 1e6, with_counts=True)
    t = _d8(DEM)
    t.update()
    assert np.array_equal(t.code.cpu().numpy(), code)
    assert np.array_equal(t.A.cpu().numpy(), A)
    assert np.array_equal(t.count.cpu().numpy(), N)
    # idempotence: a second update gives the same grids
    t.update()
    assert np.array_equal(t.A.cpu().numpy(), A)


@pytest.mark.parametrize('n', [2048, 8192])
def test_d8_large_bit_exact_vs_c_oracle_and_exact_properties(n):
    """Grids too large for the numpy oracle (2048^2; 8192^2 = 67 M cells, ~7 s of C oracle
    on the host): the C restatement (itself pinned to the reference fixtures in
    test_c_oracle.py) must agree bit for bit in codes, areas and counts, and the
    size-independent identities used at 16384^2 (scripts/bench_d8.py) must hold."""
    import torch
    from oracle import c_oracle
    from scripts import bench_d8
    from topoflow36_b200.d8 import D8Toolkit
    DEM = bench_d8.device_dem(n, seed=77)
    DEM = (torch.round(DEM * 8.0) / 8.0).contiguous()        # quantise: ties and flats appear
    tk = D8Toolkit(DEM)
    tk.start_codes()
    start = tk.code16.clone()
    tk.update()
    want = bench_d8.torch_start_codes(DEM, float(tk.dx_host[0]), float(tk.dy_host[0]),
                                      float(tk.dd_host[0]))
    assert torch.equal(start, want)
    dx, dy, dd = o.pixel_dims(n, 30.0, 30.0)
    code_o, _ = c_oracle.d8_flow_grid(DEM.cpu().numpy(), dx, dy, dd)
    assert np.array_equal(tk.code.cpu().numpy(), code_o)
    A_o, N_o = c_oracle.d8_area(code_o, np.float64(900.0) / 1e6, with_counts=True)
    assert np.array_equal(tk.A.cpu().numpy(), A_o)
    assert np.array_equal(tk.count.cpu().numpy(), N_o)
    assert tk.n_reps > 1                                       # flats were actually linked
    pid = tk.parent_ids().to(torch.int64).reshape(-1)
    cnt = tk.count.reshape(-1)
    flow = (tk.code.reshape(-1) != 0) & (pid > 0)
    acc = torch.zeros_like(cnt)
    acc.index_add_(0, pid[flow], cnt[flow])
    has_area = tk.A.reshape(-1) > 0
    assert torch.equal(cnt[has_area], (acc + 1)[has_area])


@pytest.mark.parametrize('n_slabs', [1, 2, 3, 5])
def test_area_across_row_slabs_equals_one_grid(n_slabs):
    """The row-slab form of the contributing-area pass (tfb_d8_area_slab_*: walkers stop at a
    slab boundary, the {area, count} words of the boundary rows are exchanged, the walk
    resumes on the other side) driven for `n_slabs` virtual ranks on ONE device -- the halo
    exchange of slab.slab_d8_area is replaced by row copies -- against the single-grid pass:
    areas (fp32) and counts bit for bit, on a pit-filled DEM whose rivers cross every
    boundary."""
    import ctypes as C
    import torch
    from topoflow36_b200 import lib as L, prep, synth
    from topoflow36_b200.d8 import D8Toolkit
    lib = L.load()
    ny, nx = 301, 257
    DEM = synth.fractal_dem(ny, nx, seed=21, sigma=1.5, sy=0.004, sx=0.001)
    t = torch.as_tensor(DEM).cuda()
    prep.fill_pits_tensor(t)                       # long flow paths, large basins
    tk = D8Toolkit(t.cpu().numpy())
    tk.update(with_counts=True)
    code = tk.code
    pa = float(tk.pixel_area)
    bounds = [round(k * ny / n_slabs) for k in range(n_slabs + 1)]
    st = L.stream_ptr
    slabs = []
    for k in range(n_slabs):
        r0, r1 = bounds[k], bounds[k + 1]
        n = r1 - r0
        c1 = torch.zeros((n + 2, nx), dtype=torch.uint8, device='cuda')
        lo, hi = max(r0 - 1, 0), min(r1 + 1, ny)
        c1[lo - (r0 - 1):hi - (r0 - 1)] = code[lo:hi]
        cell = torch.zeros((n + 2, nx), dtype=torch.int64, device='cuda')
        word = torch.zeros((n + 2, nx), dtype=torch.int32, device='cuda')
        own = lambda x: x.data_ptr() + nx * x.element_size()
        L.check(lib.tfb_d8_area_slab_init(own(c1), n, nx, r0, ny, own(cell), own(word), st()))
        L.check(lib.tfb_d8_area_slab_walk(n, nx, pa, None, own(cell), own(word), 0, st()))
        slabs.append(dict(n=n, r0=r0, cell=cell, word=word, c1=c1, own=own))
    rounds = 0
    sent = None
    while True:
        edge = torch.cat([torch.stack((s['cell'][1], s['cell'][s['n']])) for s in slabs])
        if sent is not None and torch.equal(edge, sent):
            break
        sent = edge
        for k, s in enumerate(slabs):              # the halo exchange
            if k > 0:
                s['cell'][0].copy_(slabs[k - 1]['cell'][slabs[k - 1]['n']])
            if k < n_slabs - 1:
                s['cell'][s['n'] + 1].copy_(slabs[k + 1]['cell'][1])
        for s in slabs:
            L.check(lib.tfb_d8_area_slab_walk(s['n'], nx, pa, None, s['own'](s['cell']),
                                              s['own'](s['word']), 1, st()))
        rounds += 1
        assert rounds < 200
    A = torch.empty((ny, nx), dtype=torch.float32, device='cuda')
    cnt = torch.empty((ny, nx), dtype=torch.int64, device='cuda')
    for s in slabs:
        nd = C.c_int64(0)
        a_, c_ = A[s['r0']:s['r0'] + s['n']], cnt[s['r0']:s['r0'] + s['n']]
        L.check(lib.tfb_d8_area_unpack(s['own'](s['cell']), s['n'] * nx, a_.data_ptr(), c_.data_ptr(),
                                       C.byref(nd), st()))
    assert torch.equal(A, tk.A), int((A != tk.A).sum())
    assert torch.equal(cnt, tk.count)
    assert int(cnt.max()) > 2000                   # real rivers
    if n_slabs > 1:
        assert rounds >= 2                         # flow really crossed the boundaries
