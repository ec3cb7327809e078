"""DEM / geometry preparation on the GPU against fixtures written by
the REFERENCE (tests/golden/prep.npz, oracle/gen_golden.py gen_prep): fill_pits(),
get_new_slope_grid(), get_grid_from_TCA().  Elevations and float32 files: bit-exact;
power-law grids: <= 1 float32 ulp (numpy's pow vs CUDA's)."""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def prep():
    return load_golden('prep.npz')


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_fill_pits_vs_reference_fixture(prep, tag):
    """Depressions of all sizes, nodata holes (interior, border, interior rim), NaN, Inf,
    a closed-basin code; 'b' is taller than wide, so the reference's row-numbers-as-IDs
    closed cells fall into rows 1-2."""
    import torch
    from topoflow36_b200 import prep as P
    from oracle import np_oracle as o
    DEM = prep['fill_%s_in' % tag].copy()
    ny, nx = DEM.shape
    n = P.fill_pits(DEM, 'FLOAT', nx, ny, SILENT=True)           # numpy array, in place
    want = prep['fill_%s_out' % tag]
    assert np.array_equal(DEM, want)
    assert n == int((want > np.where(np.isfinite(prep['fill_%s_in' % tag]),
                                     prep['fill_%s_in' % tag], -9999.0)).sum())
    t = torch.as_tensor(prep['fill_%s_in' % tag].copy()).cuda()   # device tensor, in place
    n2, rounds = P.fill_pits_tensor(t)
    assert n2 == n and rounds >= 1 and np.array_equal(t.cpu().numpy(), want)
    # filling again: the nodata cells are gone, so are their seeds -- same as the oracle;
    # a third pass changes nothing
    again = want.copy()
    n_again = o.fill_pits(again)
    assert P.fill_pits_tensor(t)[0] == n_again and np.array_equal(t.cpu().numpy(), again)
    assert P.fill_pits_tensor(t)[0] == 0 and np.array_equal(t.cpu().numpy(), again)
    # the documented intent (nodata stays closed) against the restatement
    D1, D2 = prep['fill_%s_in' % tag].copy(), prep['fill_%s_in' % tag].copy()
    assert P.fill_pits(D1, close_nodata=True) == o.fill_pits(D2, close_nodata=True)
    assert np.array_equal(D1, D2) and not np.array_equal(D1, want)


@pytest.mark.parametrize('shape', [(1, 1), (2, 7), (3, 3), (33, 32), (64, 65), (5, 400)])
def test_fill_pits_ragged_shapes(shape):
    from topoflow36_b200 import prep as P
    from oracle import np_oracle as o
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    DEM = np.float32(np.round(rng.random(shape) * 40.0) / 4.0)
    if DEM.size > 20:
        DEM.reshape(-1)[rng.integers(0, DEM.size, 3)] = -9999.0
    D1, D2 = DEM.copy(), DEM.copy()
    n1, n2 = P.fill_pits(D1), o.fill_pits(D2)
    assert np.array_equal(D1, D2) and n1 == n2


def test_fill_pits_large_vs_c_oracle():
    """2048^2 rough DEM without regional tilt: ~40 % of the cells lie in depressions,
    lakes span many tiles (the hard case for a tiled relaxation)."""
    import torch
    from topoflow36_b200 import prep as P, synth
    from oracle import c_oracle as co
    DEM = synth.fractal_dem(2048, 2048, seed=3, sigma=3.0, sy=0.0005, sx=0.0002)
    DEM[700:720, 900:960] = -9999.0
    want = DEM.copy()
    n_want = co.fill_pits(want)
    t = torch.as_tensor(DEM).cuda()
    n, rounds = P.fill_pits_tensor(t)
    assert n == n_want and n > 1000000
    assert np.array_equal(t.cpu().numpy(), want)
    # size-independent property: no interior cell of the filled DEM is a strict pit
    z = t
    nb = torch.stack([torch.roll(z, (dr, dc), (0, 1)) for dr in (-1, 0, 1) for dc in (-1, 0, 1)
                      if dr or dc]).amin(0)
    assert int((z[1:-1, 1:-1] < nb[1:-1, 1:-1]).sum()) == 0


def test_new_slope_grid_vs_reference_fixture(prep):
    import torch
    from topoflow36_b200.d8 import D8Toolkit
    from oracle import np_oracle as o
    DEM = prep['slope_DEM']
    tk = D8Toolkit(DEM)
    tk.update()
    assert np.array_equal(tk.code.cpu().numpy(), prep['slope_code'])
    S, reps = tk.new_slope()
    S = S.cpu().numpy()
    assert np.array_equal(np.float32(S), prep['slope_new'])          # the reference's RTG file
    ds, _ = o.d8_ds_dw(prep['slope_code'], tk.dx_host, tk.dy_host, tk.dd_host)
    S_o, reps_o = o.new_slope_grid(DEM, prep['slope_code'], ds)
    assert reps == reps_o and np.array_equal(S, S_o)                 # fp64, bit for bit
    assert (S > 0).sum() > 0.9 * (prep['slope_code'] != 0).sum()
    assert np.array_equal(tk.aspect().cpu().numpy(), prep['aspect'])    # update_aspect_grid


def test_new_slope_grid_flat_paths_and_cell_00():
    """Terraced DEM: long runs without a positive drop, walkers that run past their
    noflow cell and are compared with DEM[0, 0]; DEM[0, 0] low, high, nodata."""
    from topoflow36_b200.d8 import D8Toolkit
    from topoflow36_b200 import synth
    from oracle import np_oracle as o
    base = synth.fractal_dem(90, 70, seed=9, sigma=1.0, sy=0.01, sx=0.004)
    for z00, q in ((-50.0, 2.0), (1e4, 1.0), (-9999.0, 0.5)):
        DEM = np.float32(np.round(base / q) * q)
        DEM[0, 0] = z00
        tk = D8Toolkit(DEM)
        tk.update()
        code = tk.code.cpu().numpy()
        ds, _ = o.d8_ds_dw(code, tk.dx_host, tk.dy_host, tk.dd_host)
        S_o, reps_o = o.new_slope_grid(DEM, code, ds)
        S, reps = tk.new_slope()
        assert reps == reps_o
        assert np.array_equal(S.cpu().numpy(), S_o), (z00, q)


def test_grid_from_TCA_vs_reference_fixture(prep, tmp_path):
    from topoflow36_b200 import prep as P, grid_io
    A = prep['tca_area']
    cases = {'w': dict(g1=25.0, g2=1.5), 'n': dict(g1=0.03, p=-0.2),
             'd': dict(A1=50.0, g1=3.0, A2=0.01, g2=0.2)}
    for tag, kw in cases.items():
        g, c, p = P.grid_from_TCA(A, **kw)
        got = g.float().cpu().numpy()
        want = prep['tca_' + tag]
        assert g.dtype.is_floating_point and got.shape == want.shape
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
        assert ulp.max() <= 1, (tag, ulp.max())
        assert np.array_equal(got[A <= 0], np.ones((A <= 0).sum(), np.float32))
    # file-level mirror: same arguments and files as utils/parameterize.get_grid_from_TCA
    info = grid_io.make_rti(A.shape[1], A.shape[0], 30.0, 30.0)
    topo = str(tmp_path) + os.sep
    grid_io.write_rti(topo + 'Synth.rti', info)
    grid_io.write_rtg(topo + 'Synth_area.rtg', A, info)
    out = P.get_grid_from_TCA(site_prefix='Synth', topo_dir=topo, area_file='Synth_area.rtg',
                              out_file='Synth_chan-w.rtg', g1=25.0, g2=1.5, REPORT=False)
    back = grid_io.read_rtg(out, grid_io.read_rti(topo + 'Synth.rti'))
    ulp = np.abs(back.view(np.int32).astype(np.int64) - prep['tca_w'].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1
