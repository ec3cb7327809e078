"""Host logic of the multi-GPU path on CPU: row partition, halo exchange over
gloo with world_size 2 and 3 (the device kernels are not involved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from topoflow36_b200.slab import SlabLayout, exchange_halos, partition_rows


def test_partition_rows():
    assert partition_rows(10, 3) == [(0, 4), (4, 3), (7, 3)]
    assert partition_rows(8, 8) == [(i, 1) for i in range(8)]
    for ny, w in ((65536, 8), (4097, 4), (29, 2)):
        parts = partition_rows(ny, w)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == ny
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_take_pads_outside_rows_with_zeros():
    g = np.arange(12 * 5, dtype=np.float64).reshape(12, 5) + 1
    lay = SlabLayout(12, 5, rank=0, world_size=3, halo=2)
    t = lay.take(g)
    assert t.shape == (4 + 4, 5) and (t[:2] == 0).all() and np.array_equal(t[2:], g[0:6])
    lay = SlabLayout(12, 5, rank=2, world_size=3, halo=2)
    t = lay.take(g)
    assert (t[-2:] == 0).all() and np.array_equal(t[:-2], g[6:12])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, halo, ny_global, nx, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        g = torch.arange(ny_global * nx, dtype=torch.float64).reshape(ny_global, nx) + 1.0
        g2 = -g
        lay = SlabLayout(ny_global, nx, rank, world, halo)
        # local tensors know only their OWNED rows; halos start as garbage
        a = torch.full((lay.ny + 2 * halo, nx), -777.0, dtype=torch.float64)
        b = torch.full((lay.ny + 2 * halo, nx), -777.0, dtype=torch.float64)
        a[halo:halo + lay.ny] = g[lay.row0:lay.row0 + lay.ny]
        b[halo:halo + lay.ny] = g2[lay.row0:lay.row0 + lay.ny]
        exchange_halos([a, b], lay, rows=[halo, 1])
        want_a = lay.take(g)
        ok = True
        # interior halos must equal the neighbour's rows; outer (global-border)
        # halos are left untouched
        for t, want, k in ((a, want_a, halo), (b, lay.take(g2), 1)):
            lo = slice(halo - k, halo) if lay.up is not None else None
            hi = slice(halo + lay.ny, halo + lay.ny + k) if lay.down is not None else None
            for sl in (lo, hi):
                if sl is not None:
                    ok &= bool(torch.equal(t[sl], want[sl]))
            if lay.up is None:
                ok &= bool((t[:halo] == -777.0).all())
            ok &= bool(torch.equal(t[halo:halo + lay.ny], want[halo:halo + lay.ny]))
        s = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(s)
        ok &= (s.item() == world * (world + 1) / 2)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world,halo', [(2, 1), (2, 2), (3, 2)])
def test_halo_exchange_gloo(world, halo):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, halo, 23, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res


def test_peer_halo_descriptor_addresses():
    """PeerHalos.descriptor() is pure address arithmetic: the neighbours' halo rows of the
    discharge buffer a launch writes (or of the depth grid) and the flag words.  Checked
    here without a device for halo widths 1 (kinematic) and 2 (diffusive)."""
    from topoflow36_b200 import slab as S
    for halo in (1, 2):
        lay = S.SlabLayout(3 * 10, 16, 1, 3, halo)              # the middle rank of three
        ph = object.__new__(S.PeerHalos)
        ph.layout = lay
        ph._mine = [0x1000, 0x2000, 0x3000, 0x4000]             # Q0, Q1, flags, d
        ph.nbr = {'up': dict(q=[0x10000, 0x20000], flags=0x30000, ny=10, d=0x40000),
                  'down': dict(q=[0x50000, 0x60000], flags=0x70000, ny=10, d=0x80000)}
        row = lay.nx * 8
        d = ph.descriptor(1, 7)
        assert d.up_Q == 0x20000 + (halo + 10) * row            # its first lower halo row
        assert d.down_Q == 0x60000 + (halo - 1) * row           # its last upper halo row
        assert d.up_flag == 0x30000 + 4 and d.down_flag == 0x70000
        assert d.from_up == 0x3000 and d.from_down == 0x3000 + 4 and d.step == 7
        d = ph.descriptor(0, 14, depth=True)                    # pass A of the diffusive step
        assert d.up_Q == 0x40000 + (halo + 10) * row and d.down_Q == 0x80000 + (halo - 1) * row
        assert d.step == 14
        # the first rank has no neighbour above
        ph.nbr.pop('up')
        d = ph.descriptor(0, 0)
        assert not d.up_Q and not d.up_flag and not d.from_up and d.down_Q


def _board_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    import numpy as np
    import torch.distributed as dist
    from topoflow36_b200 import slab as S
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lay = S.SlabLayout(world * 8, 16, rank, world, 1)
        ok = S.ScalarBoard.usable(lay)
        board = S.ScalarBoard(lay, 232)
        dist.barrier()
        ok &= not os.path.exists(board._path)               # unlinked once everyone has mapped it
        rng = np.random.default_rng(rank)
        for k in range(1, 301):
            mine = np.arange(29, dtype=np.float64) + 1000.0 * rank + k
            if rank == k % world:
                # a slow rank: the others must wait for it and must not overwrite its bank
                import time
                time.sleep(0.0005 * rng.random())
            got = board.exchange(mine)
            want = np.stack([np.arange(29, dtype=np.float64) + 1000.0 * r + k for r in range(world)])
            ok &= bool(np.array_equal(got, want))
        board.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_scalar_board_shared_memory(world):
    """slab.ScalarBoard: the ranks' scalar blocks exchanged through one shared-memory segment
    (two banks, sequence words) -- every read returns every rank's block of THAT read, also
    when ranks arrive at different times."""
    if not os.path.isdir('/dev/shm'):
        pytest.skip('no /dev/shm')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_board_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == list(range(world))
    assert all(ok for _, ok in res), res
