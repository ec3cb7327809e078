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


def _worker(rank, world, port, kinematic, infil, q, devgen=False):
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
                                             world=world, device=dev, return_parts=True, dt=2.0)
            one = cases.build_engine(ny * world, nx, kinematic=kinematic, infil=infil, device=dev,
                                     tile_rows=ny, dt=2.0)
            lay = parts['layout']
        rows = slice(lay.row0, lay.row0 + lay.ny)
        fails = []
        ok = bool(torch.equal(slab.engine.own(slab.engine.code), one.code[rows]))
        if not ok:
            fails.append('code (%d cells differ)' % int((slab.engine.own(slab.engine.code) != one.code[rows]).sum()))
        if not torch.equal(slab.engine.own(slab.engine.S_bed), one.S_bed[rows]):
            fails.append('S_bed')
        for i in range(steps):
            if i == 80:
                slab.engine.set_input('P_rain', 0.0)
                one.set_input('P_rain', 0.0)
            slab.step(i * 2.0 / 60.0)
            one.step(i * 2.0 / 60.0)
        e = slab.engine
        for nm, a, b in (('vol', e.vol, one.vol), ('d', e.own(e.d), one.own(one.d)),
                         ('u', e.own(e.u), one.own(one.u)), ('Q', e.Q, one.Q)):
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


@pytest.mark.parametrize('kinematic,infil,devgen', [(True, False, False), (True, True, False),
                                                    (False, False, False), (True, True, True)])
def test_two_slabs_equal_one_grid(kinematic, infil, devgen):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kinematic, infil, q, devgen))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert any(m for _, _, m, _ in res)
