"""Row-slab multi-GPU routing == single-GPU routing, bit for bit (needs >= 2 GPUs;
skipped on a 1-GPU box).  Each rank builds its slab of a (2*ny, nx) synthetic
grid (distributed D8 with flat linking across the slab boundary), steps it with
NCCL halo exchange, and compares its rows with a single-GPU run of the whole
grid made on the same device."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kinematic, infil, q, devgen=False, p2p=False, d0=0.0, packed=True,
            engine_kw=None):
    import torch
    import torch.distributed as dist
    from topoflow36_b200 import cases
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        ny, nx, steps = 96, 160, 120
        if devgen:
            slab = cases.build_engine_device(ny, nx, kinematic=kinematic, infil=infil, rank=rank,
                                             world=world, device=dev, dt=2.0, lean_memory=False)
            one = cases.build_engine_device(ny * world, nx, kinematic=kinematic, infil=infil,
                                            device=dev, dt=2.0, lean_memory=False)
            lay = slab.layout
        else:
            slab, parts = cases.build_engine(ny, nx, kinematic=kinematic, infil=infil, rank=rank,
                                             world=world, device=dev, return_parts=True, dt=2.0,
                                             p2p=p2p, d0=d0, packed=packed, **(engine_kw or {}))
            assert (slab.peer is not None) == bool(p2p)
            one = cases.build_engine(ny * world, nx, kinematic=kinematic, infil=infil, device=dev,
                                     tile_rows=ny, dt=2.0, d0=d0, packed=packed, **(engine_kw or {}))
            lay = parts['layout']
        rows = slice(lay.row0, lay.row0 + lay.ny)
        fails = []
        ok = bool(torch.equal(slab.engine.own(slab.engine.code), one.code[rows]))
        if not ok:
            fails.append('code (%d cells differ)' % int((slab.engine.own(slab.engine.code) != one.code[rows]).sum()))
        if not torch.equal(slab.engine.own(slab.engine.S_bed), one.S_bed[rows]):
            fails.append('S_bed')
        if not devgen:
            # contributing area across the slab boundary == single grid, bit for bit
            from topoflow36_b200.slab import SlabLayout, slab_d8_area
            A, cnt, rounds = slab_d8_area(parts['d8']['code'], lay, 900.0 / 1e6)
            A1, cnt1, _ = slab_d8_area(one.own(one.code).contiguous(),
                                       SlabLayout(ny * world, nx, 0, 1, 0), 900.0 / 1e6)
            if not (torch.equal(A, A1[rows]) and torch.equal(cnt, cnt1[rows])):
                fails.append('D8 area across slabs (%d cells differ, %d rounds)'
                             % (int((cnt != cnt1[rows]).sum()), rounds))
            # path-length slopes (new_slopes.get_new_slope_grid) of the sharded grid == one grid
            from topoflow36_b200.slab import slab_new_slope
            from topoflow36_b200.d8 import D8Toolkit
            S_new, n_path = slab_new_slope(parts['DEM'], parts['d8']['code'], lay, lay.halo + 1,
                                           device=dev)
            tile = cases.noise_tile(ny, nx, seed=1234)
            tk = D8Toolkit(cases.slab_dem(ny * world, nx, 0, ny * world, 0, tile), device=dev)
            tk.update_flow_grid()
            S1, n1 = tk.new_slope()
            if not (torch.equal(S_new, S1[rows]) and n_path == n1 and float(S1.max()) > 0):
                fails.append('new slopes across slabs (%d cells differ, longest path %d vs %d)'
                             % (int((S_new != S1[rows]).sum()), n_path, n1))
        for i in range(steps):
            if i == 80:
                slab.engine.set_input('P_rain', 0.0)
                one.set_input('P_rain', 0.0)
            slab.step(i * 2.0 / 60.0)
            one.step(i * 2.0 / 60.0)
        e = slab.engine
        extra = []
        if engine_kw and engine_kw.get('flood'):
            extra += [('d_flood', e.own(e.d_flood), one.own(one.d_flood)),
                      ('vol_flood', e.own(e.vol_flood), one.own(one.vol_flood))]
            if not float(one.d_flood.max()) > 0:
                fails.append('no overbank flow')
        if engine_kw and engine_kw.get('attenuate'):
            extra += [('vol_stored', e.vol_stored, one.vol_stored)]
        if engine_kw and e.last_kernel != 'ext':
            fails.append('kernel %s, expected ext' % e.last_kernel)
        for nm, a, b in [('vol', e.vol, one.vol), ('d', e.own(e.d), one.own(one.d)),
                         ('u', e.own(e.u), one.own(one.u)), ('Q', e.Q, one.Q)] + extra:
            if not torch.equal(a, b[rows]):
                ok = False
                fails.append('%s (%d cells, max %g)' % (nm, int((a != b[rows]).sum()),
                                                        float((a - b[rows]).abs().max())))
        moved = float(one.Q[rows].abs().max()) > 0
        sg, s1 = slab.read_scalars(), one.read_scalars()
        rel = lambda x, y: abs(x - y) <= 1e-12 * max(abs(y), 1e-300)
        for nm in ('vol_R', 'vol_edge', 'vol_Q') + (('vol_IN',) if infil else ()):
            if not rel(getattr(sg, nm), getattr(s1, nm)):
                fails.append('%s %r vs %r' % (nm, getattr(sg, nm), getattr(s1, nm)))
        for nm in ('Q_outlet', 'Q_peak', 'T_peak'):
            if getattr(sg, nm) != getattr(s1, nm):
                fails.append('%s %r vs %r' % (nm, getattr(sg, nm), getattr(s1, nm)))
        q.put((rank, not fails, moved, fails))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('kinematic', [True, False])
def test_two_slabs_flood_and_attenuate_on_the_extended_kernels(kinematic):
    """FLOOD_OPTION + ATTENUATE across a slab boundary (extended packed kernels, NCCL halo):
    the diffusive pass B reads the neighbour's flood depths too (free surface = channel +
    flood depth); 2 slabs must equal one grid bit for bit."""
    test_two_slabs_equal_one_grid(kinematic, False, False, False,
                                  engine_kw=dict(flood=True, d_bankfull=0.001, attenuate=True,
                                                 n_days=0.002))


def test_two_slabs_initial_depth_generic_diffusive():
    """d0 = 5 cm with the GENERIC diffusive kernel, which re-derives the parent's depth from
    the halo rows of the state: they must hold the neighbours' initial volumes before the
    first step (SlabChannels starts with stale state halos)."""
    test_two_slabs_equal_one_grid(False, True, False, False, d0=0.05, packed=False)


@pytest.mark.parametrize('kinematic,infil,devgen,p2p', [
    (True, False, False, False), (True, True, False, False), (False, False, False, False),
    (True, True, True, False), (True, False, False, True), (True, True, False, True),
    (False, False, False, True), (False, True, False, True)])
def test_two_slabs_equal_one_grid(kinematic, infil, devgen, p2p, d0=0.0, packed=True, engine_kw=None):
    """p2p = True: the halo rows travel through NVLink peer stores inside the routing
    kernel (tfb_channels_step_packed_p2p; diffusive wave: ..._dw_a_p2p stores the boundary
    depths, ..._dw_b_p2p the boundary discharges) instead of NCCL send/recv."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kinematic, infil, q, devgen, p2p,
                                               d0, packed, engine_kw))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert any(m for _, _, m, _ in res)


def _comp_worker(rank, world, port, case_dir, kinematic, q):
    """The BMI component under torch.distributed: each rank initialises from the
    SAME CFG/RTG files, reads only its rows, and the gathered grids must equal a
    single-GPU component run (made on rank 0's device)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world,
                            device_id=torch.device('cuda', rank))
    try:
        from topoflow36_b200 import channels_kinematic_wave as ckw, channels_diffusive_wave as cdw
        mod = ckw if kinematic else cdw
        cfg = os.path.join(case_dir, 'Case1' + mod.channels_component._att_map['cfg_extension'])
        c = mod.channels_component()
        c.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
        assert c.layout is not None and c.layout.world_size == world
        ny, nx = c.ny, c.nx
        rng = np.random.default_rng(4)
        ETg, INg = 2e-6 * rng.random((ny, nx)), 3e-6 * rng.random((ny, nx))
        ET = ETg.copy()
        z = lambda v: np.array(v, dtype='float64')
        L = {n: ln for ln, n in c._var_name_map.items()}
        for k, v in dict(P_rain=z(80 / 3.6e6), MR=z(0), GW=z(1e-7), SM=z(2e-7), rho_H2O=z(1000.0),
                         ET=ET, IN=INg.copy()).items():
            c.set_value(L[k], v)
        for i in range(80):
            c.update()
        grids = {n: c.gather_grid(n) for n in ('vol', 'd', 'u', 'Q')}
        scal = {n: float(getattr(c, n)) for n in ('vol_R', 'vol_edge', 'vol_Q', 'Q_outlet', 'Q_peak',
                                                   'T_peak')}
        c.finalize()
        dist.barrier()                            # every rank's rows of the output frames are on disk
        scal['vol_chan_sum'] = float(c.vol_chan_sum)
        scal['Q_max'] = float(c.Q_max)
        ok, fails = True, []
        if rank == 0:
            dist.destroy_process_group()          # the comparison run is a plain 1-GPU component
            one = mod.channels_component()
            one.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
            assert one.layout is None
            ET1 = ETg.copy()
            for k, v in dict(P_rain=z(80 / 3.6e6), MR=z(0), GW=z(1e-7), SM=z(2e-7),
                             rho_H2O=z(1000.0), ET=ET1, IN=INg.copy()).items():
                one.set_value(L[k], v)
            for i in range(80):
                one.update()
            for n, g in grids.items():
                if not np.array_equal(g, getattr(one, n)):
                    fails.append('%s differs in %d cells' % (n, int((g != getattr(one, n)).sum())))
            one.finalize()
            # output files: every rank wrote its own rows of each frame into ONE stack; the
            # single-GPU run found those names taken and wrote <name>_1 (BMI "_unique")
            import glob
            for stem in ('Case1_2D-Q', 'Case1_2D-d', 'Case1_0D-Q'):
                ext = '.txt' if '0D' in stem else '.rts'
                fa, fb = (os.path.join(case_dir, stem + sfx + ext) for sfx in ('', '_1'))
                if not (os.path.exists(fa) and os.path.exists(fb)):
                    fails.append('missing output %s: %s' % (stem, sorted(glob.glob(case_dir + '/Case1_*D*'))))
                elif open(fa, 'rb').read() != open(fb, 'rb').read():
                    fails.append('output file %s differs between 2 slabs and 1 GPU' % stem)
                elif ext == '.rts' and os.path.getsize(fa) != 3 * ny * nx * 4:
                    fails.append('%s: %d bytes, expected 3 frames' % (stem, os.path.getsize(fa)))
            for n, v in scal.items():
                w = float(getattr(one, n))
                if abs(v - w) > 1e-12 * max(abs(w), 1e-300):
                    fails.append('%s %r vs %r' % (n, v, w))
            lay_rows = slice(0, c.layout.ny)
            if not np.array_equal(ET[lay_rows], ET1[lay_rows]):
                fails.append('ET mask of the owned rows')
            if not float(one.Q.max()) > 0:
                fails.append('nothing moved')
        q.put((rank, not fails, fails))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize('kinematic', [True, False])
def test_component_two_slabs_equal_one_gpu(tmp_path, kinematic):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    from topoflow36_b200 import cases, synth
    from topoflow36_b200.d8 import D8Toolkit
    ny, nx = 130, 96
    DEM = synth.fractal_dem(ny, nx, seed=9, sigma=1.0)
    DEM = np.float32(np.round(DEM * 4) / 4)                 # flats across the slab boundary
    width, angle, nval = synth.channel_params(ny, nx, seed=9)
    tk = D8Toolkit(DEM)
    tk.update_flow_grid()
    slope = tk.slope().cpu().numpy()
    cfg = synth.write_case(str(tmp_path), DEM, slope, width, angle, nval=nval, dt=2.0,
                           kinematic=kinematic, outlet=(ny - 2, nx // 2), check_stability='Yes')
    txt = open(cfg).read()
    for k in ('SAVE_Q_GRIDS        | No', 'SAVE_D_GRIDS        | No', 'SAVE_Q_PIXELS       | No'):
        assert k in txt
        txt = txt.replace(k, k[:-2] + 'Yes')
    open(cfg, 'w').write(txt)
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_comp_worker, args=(r, world, port, str(tmp_path), kinematic, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
