"""Quick device timing of the routing kernel variants (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from topoflow36_b200.routing import ChannelsEngine
from topoflow36_b200.d8 import D8Toolkit
from topoflow36_b200 import synth

def build(n, kinematic=True, diag=False, infil=False):
    dev = torch.device('cuda:0')
    g = torch.Generator(device=dev); g.manual_seed(1)
    DEM = torch.as_tensor(synth.fractal_dem(n, n, seed=1234)).to(dev) if n <= 4096 else None
    if DEM is None:
        r = torch.arange(n, device=dev, dtype=torch.float32)[:, None]; c = torch.arange(n, device=dev, dtype=torch.float32)[None, :]
        DEM = (1e4 - 0.6*r - 0.15*c + 2*torch.rand((n, n), device=dev, generator=g)).float()
    t = D8Toolkit(DEM); t0 = time.time(); t.update(); torch.cuda.synchronize(); print('d8 total s', time.time()-t0, 'reps', t.n_reps, 'area cells', t.n_area_cells)
    S = t.slope().double(); S[S <= 0] = S[S > 0].min()
    width = (1 + 4*torch.rand((n, n), device=dev, dtype=torch.float64, generator=g)).float().double()
    angle = torch.full((n, n), float(np.pi/4), device=dev, dtype=torch.float64)
    nval = (0.03 + 0.02*torch.rand((n, n), device=dev, dtype=torch.float64, generator=g)).float().double()
    e = ChannelsEngine(n, n, dt=1.0, code=t.code, dx=t.dx, dy=t.dy, dd=t.dd, S_bed=S, width=width, angle=angle, rough=nval,
                       kinematic=kinematic, diag=diag, outlet=(n-2, n//2))
    e.set_input('P_rain', 50/3.6e6)
    if infil:
        e.enable_fused_sources(infil=True)
        for k, v in dict(Ks=7.2e-6, Ki=1e-7, qs=0.485, qi=0.375808, G=0.724).items(): e.set_input(k, v)
    return e

def timeit(e, steps=50, warm=10):
    for _ in range(warm): e.step()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(steps): e.step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)/steps

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    for name, kw, B in (('kw lean', dict(), 81), ('kw+GA', dict(infil=True), 97), ('kw diag', dict(diag=True), 137), ('dw lean', dict(kinematic=False), 81)):
        e = build(n, **kw)
        ms = timeit(e)
        cells = n*n
        print(e.last_path, '%-8s n=%d  %.3f ms/step  %.2f Gcell/s  alg %.0f GB/s (%d B/cell)' % (name, n, ms, cells/ms/1e6, cells*B/ms/1e6, B))
        s = e.read_scalars(); print('   Q_outlet', s.Q_outlet, 'vol_R', s.vol_R, 'bad', s.n_neg_d, s.n_nan_u)
        del e; torch.cuda.empty_cache()
