"""Seeded synthetic routing workloads built straight on the device
for benchmarks and large-grid tests: fractal DEM -> GPU D8 toolkit -> channel
parameters -> ChannelsEngine.  The same generator feeds every rank of a
row-slab run: rank k owns rows [k*ny, (k+1)*ny) of a (world*ny, nx) grid whose
noise tile repeats every `ny` rows on top of a global tilt, so the slabs join
into ONE drainage network (flow crosses the slab boundaries).

Parameters follow SURVEY.md 8d: dx = dy = 30 m, width = 1 + 4 U(0,1) m, bank
angle 45 deg, Manning n = 0.03 + 0.02 U(0,1) -- float32 values, as they would
sit in RTG files -- uniform rain 50 mm/h, dt = 1 s, Green-Ampt scalars of the
Treynor CFG (Ks=7.2e-6, Ki=1e-7, qs=0.485, qi=0.375808, G=0.724).  The benchmark
starts from a pre-wetted soil (cumulative infiltration I0 = 0.1 m) so that the
infiltration capacity is already below the rain rate and runoff, routing and
infiltration are all active from the first step (with the CFG's I0 = 1e-6 m the
first ~6000 steps infiltrate everything and no water moves).
"""
import numpy as np

from . import synth

GA_TREYNOR = dict(Ks=7.2e-6, Ki=1e-7, qs=0.485, qi=0.375808, G=0.724)
RAIN_50MMPH = 50.0 / 3.6e6
I0_WET = 0.1
# BASELINE configs[3] source terms: T_air = 2 C +- 3 C noise (grid), degree-day
# c0 = 2.7 mm/day/C, T0 = -0.2 C, h0_swe = 0.1 m, rho_snow = 300; Priestley-Taylor alpha = 0.95,
# K_soil = 0.45, soil_x = 0.05, T_soil_x = 20 (cfg_templates/Test1); Green-Ampt as above
FULL_SOURCES = dict(met=True, snow=True, evap=True, infil=True, T_rain_snow=1.0, rho_snow=300.0,
                    alpha=0.95, K_soil=0.45, soil_x=0.05, T_soil_x=20.0, h0_swe=0.1, I0=I0_WET)
FULL_SCALARS = dict(c0=2.7, T0=-0.2, T_surf=5.0, Qn_SW=300.0, Qn_LW=-40.0)


def apply_sources(eng, mode, row0=0, halo=0, seed=1234):
    """mode: 'none' | 'ga' (fused Green-Ampt) | 'full' (met + snow + evap + infil fused,
    T_air a grid keyed by the GLOBAL cell so slabs agree)."""
    import torch
    eng.set_input('P_rain', RAIN_50MMPH)
    if mode == 'none':
        return
    if mode == 'grid':
        # what a coupled run delivers: infiltration and evaporation rates as (ny, nx) grids
        # (keyed by the GLOBAL cell); they ride the extended packed kernels
        ny, nx = eng.ny, eng.nx
        dev = eng.device
        ri = torch.arange(row0 - halo, row0 + ny + halo, device=dev, dtype=torch.int64)[:, None]
        ci = torch.arange(nx, device=dev, dtype=torch.int64)[None, :]
        for name, amp, sd in (('IN', 4e-6, 31), ('ET', 1e-7, 37)):
            g = amp * _hash01(ri, ci, seed + sd)
            eng.set_input(name, g if halo else g.contiguous())
        return
    if mode == 'ga':
        eng.enable_fused_sources(infil=True, I0=I0_WET)
    else:
        eng.enable_fused_sources(**FULL_SOURCES)
        ny, nx = eng.ny, eng.nx
        dev = eng.device
        ri = torch.arange(row0 - halo, row0 + ny + halo, device=dev, dtype=torch.int64)[:, None]
        ci = torch.arange(nx, device=dev, dtype=torch.int64)[None, :]
        T_air = 2.0 + 3.0 * (2.0 * _hash01(ri, ci, seed + 5) - 1.0)
        eng.set_input('T_air', T_air if halo else T_air.contiguous())
        for k, v in FULL_SCALARS.items():
            eng.set_input(k, v)
    for k, v in GA_TREYNOR.items():
        eng.set_input(k, v)



def noise_tile(ny, nx, seed=1234, sigma=2.0, beta=2.2):
    """Periodic fractal noise (zero mean) of one slab, float64 on the host."""
    z = synth.fractal_dem(ny, nx, seed=seed, sy=0.0, sx=0.0, sigma=sigma, beta=beta, z0=0.0)
    return np.float64(z)


def slab_dem(ny, nx, row0, ny_global, halo, tile, sy=0.02, sx=0.005, dx=30.0, dy=30.0,
             sigma=2.0):
    """float32 DEM rows [row0 - halo, row0 + ny + halo) of the global grid; rows
    outside the grid repeat the nearest valid row's formula (they are never
    owned, only read by the border stencil, whose result is forced to code 0)."""
    r = np.arange(row0 - halo, row0 + ny + halo, dtype=np.float64)[:, None]
    c = np.arange(nx, dtype=np.float64)[None, :]
    z0 = sy * dy * ny_global + sx * dx * nx + 10.0 * sigma + 100.0
    t = tile[np.mod(np.arange(row0 - halo, row0 + ny + halo), tile.shape[0])]
    return np.float32(z0 - sy * dy * r - sx * dx * c + t)


def slab_params(ny, nx, row0, halo, seed=1234):
    """width, angle [deg], nval for the rows of a slab incl. halo (float32).
    Halo rows only matter for the diffusive wave (parent depth re-derivation), so
    they must equal what the owning rank generates: the stream is keyed by the
    GLOBAL row."""
    rows = np.arange(row0 - halo, row0 + ny + halo)
    width = np.empty((rows.size, nx), dtype=np.float32)
    nval = np.empty((rows.size, nx), dtype=np.float32)
    # one Philox stream per 1024-row band keeps generation O(slab)
    band = 1024
    for b in np.unique(rows // band):
        rng = np.random.Generator(np.random.Philox(key=int(seed) + 7919 * (int(b) + 1)))
        w = rng.random((band, nx), dtype=np.float32)
        n = rng.random((band, nx), dtype=np.float32)
        sel = (rows // band) == b
        idx = rows[sel] - b * band
        width[sel] = np.float32(1.0) + np.float32(4.0) * w[idx]
        nval[sel] = np.float32(0.03) + np.float32(0.02) * n[idx]
    angle = np.full((rows.size, nx), 45.0, dtype=np.float32)
    return width, angle, nval


# ---------------------------------------------------------------------------
# device-side generator for slabs too large to synthesise on the host (65536^2)
# ---------------------------------------------------------------------------
def _hash01(r, c, seed):
    """Uniform [0, 1) float64 from global (row, col): a 64-bit mix evaluated with
    torch int64 ops (wrapping multiply), identical on every rank."""
    import torch
    x = (r * 73856093) ^ (c * 19349663) ^ int(seed)
    x = x * -7046029254386353131            # 0x9E3779B97F4A7C15 as int64
    x = x ^ (x >> 29)
    x = x * -4658895280553007687            # 0xBF58476D1CE4E5B9
    x = x ^ (x >> 32)
    return ((x >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))


def device_slab_dem(ny, nx, row0, ny_global, halo, device, seed=1234, sy=0.02, sx=0.005,
                    dx=30.0, dy=30.0, sigma=2.0, n_modes=24, jitter=0.25, chunk=512):
    """float32 DEM rows [row0 - halo, row0 + ny + halo) as a closed-form field of
    the GLOBAL cell coordinates: tilt + n_modes plane waves with power-law
    amplitudes (wavelengths 6 .. 6000 cells) + hashed per-cell jitter."""
    import torch
    rng = np.random.default_rng(seed)
    lam = np.exp(rng.uniform(np.log(6.0), np.log(6000.0), n_modes))
    th = rng.uniform(0, 2 * np.pi, n_modes)
    ph = rng.uniform(0, 2 * np.pi, n_modes)
    amp = lam ** 0.6
    amp *= sigma * np.sqrt(2.0) / np.sqrt((amp ** 2).sum())
    kx, ky = 2 * np.pi / lam * np.cos(th), 2 * np.pi / lam * np.sin(th)
    z0 = sy * dy * ny_global + sx * dx * nx + 10.0 * sigma + 100.0
    out = torch.empty((ny + 2 * halo, nx), dtype=torch.float32, device=device)
    c = torch.arange(nx, device=device, dtype=torch.float64)[None, :]
    ci = torch.arange(nx, device=device, dtype=torch.int64)[None, :]
    for a in range(0, ny + 2 * halo, chunk):
        b = min(ny + 2 * halo, a + chunk)
        ri = torch.arange(row0 - halo + a, row0 - halo + b, device=device, dtype=torch.int64)[:, None]
        r = ri.to(torch.float64)
        z = z0 - (sy * dy) * r - (sx * dx) * c
        for m in range(n_modes):
            z = z + amp[m] * torch.sin(kx[m] * c + ky[m] * r + ph[m])
        z = z + jitter * (_hash01(ri, ci, seed) - 0.5)
        out[a:b] = z.to(torch.float32)
    return out


def device_slab_params(ny, nx, row0, halo, device, seed=1234, chunk=2048):
    """width and Manning n (float32-exact float64 tensors) for the rows of a slab."""
    import torch
    width = torch.empty((ny + 2 * halo, nx), dtype=torch.float64, device=device)
    nval = torch.empty_like(width)
    ci = torch.arange(nx, device=device, dtype=torch.int64)[None, :]
    for a in range(0, ny + 2 * halo, chunk):
        b = min(ny + 2 * halo, a + chunk)
        ri = torch.arange(row0 - halo + a, row0 - halo + b, device=device, dtype=torch.int64)[:, None]
        width[a:b] = (1.0 + 4.0 * _hash01(ri, ci, seed + 11)).float().double()
        nval[a:b] = (0.03 + 0.02 * _hash01(ri, ci, seed + 23)).float().double()
    return width, nval


def build_engine_device(ny, nx, kinematic=True, infil=False, dt=1.0, rank=0, world=1, seed=1234,
                        device=None, group=None, link_flats=True, lean_memory=True, sources=None,
                        p2p=False):
    """Like build_engine, but every array is generated on the GPU (any size that
    fits HBM) and the generic fp64 static inputs are released once the packed
    encoding exists."""
    import torch
    from .routing import ChannelsEngine
    from .slab import SlabLayout, SlabChannels, slab_d8
    ny_global = ny * world
    halo = 0 if world == 1 else (1 if kinematic else 2)
    lay = SlabLayout(ny_global, nx, rank, world, halo)
    hd = halo + 1 if world > 1 else 0
    DEM = device_slab_dem(ny, nx, lay.row0, ny_global, hd, device, seed=seed)
    d8 = slab_d8(DEM, lay, hd, link_flats=link_flats, kinematic=kinematic, device=device,
                 group=group)
    del DEM
    width, nval = device_slab_params(ny, nx, lay.row0, halo, device, seed=seed)
    ang = float(np.float64(np.float32(45.0)) * (np.pi / np.float64(180)))
    peer = None
    if p2p and world > 1:
        from .slab import make_peer_halos
        peer = make_peer_halos(lay, device, group, with_depth=not kinematic)
    eng = ChannelsEngine(ny, nx, dt=dt, code=d8['code'], dx=d8['dx'], dy=d8['dy'], dd=d8['dd'],
                         S_bed=d8['S_bed'], width=width, angle=ang, tan_angle=float(np.tan(ang)),
                         cos_angle=float(np.cos(ang)), rough=nval, kinematic=kinematic,
                         outlet=(ny_global - 2, nx // 2), row0=lay.row0, ny_global=ny_global,
                         halo=halo, device=device, finalize=True, n_pixels_global=ny * nx,
                         q_buffers=None if peer is None else peer.q_buffers,
                         d_buffer=None if peer is None else peer.d_buffer)
    del width, nval, d8
    apply_sources(eng, sources or ('ga' if infil else 'none'), row0=lay.row0, halo=halo, seed=seed)
    if lean_memory and eng.packed is not None:
        eng.release_generic_inputs()
    torch.cuda.current_stream().synchronize()
    torch.cuda.empty_cache()
    return SlabChannels(lay, eng, group, peer=peer) if world > 1 else eng


def build_engine(ny, nx, kinematic=True, infil=False, diag=False, dt=1.0, rank=0, world=1,
                 seed=1234, device=None, group=None, link_flats=True, return_parts=False,
                 tile_rows=None, sources=None, p2p=False, d0=0.0, packed=True, **engine_kw):
    """Engine (and SlabChannels wrapper when world > 1) for rank `rank` of a
    (world*ny, nx) synthetic grid."""
    import torch
    from .routing import ChannelsEngine
    from .slab import SlabLayout, SlabChannels, slab_d8
    ny_global = ny * world
    halo = 0 if world == 1 else (1 if kinematic else 2)
    lay = SlabLayout(ny_global, nx, rank, world, halo)
    tile = noise_tile(tile_rows or ny, nx, seed=seed)
    hd = halo + 1 if world > 1 else 0             # DEM halo: one more row than the codes need
    DEM = slab_dem(ny, nx, lay.row0, ny_global, hd, tile)
    d8 = slab_d8(DEM, lay, hd, link_flats=link_flats, kinematic=kinematic, device=device,
                 group=group)
    width, angle, nval = slab_params(ny, nx, lay.row0, halo, seed=seed)
    DEG = np.pi / np.float64(180)
    ang = np.float64(angle) * DEG
    peer = None
    if p2p and world > 1:
        from .slab import make_peer_halos
        peer = make_peer_halos(lay, device, group, with_depth=not kinematic)
    eng = ChannelsEngine(ny, nx, dt=dt, code=d8['code'], dx=d8['dx'], dy=d8['dy'], dd=d8['dd'],
                         S_bed=d8['S_bed'], width=np.float64(width), angle=ang,
                         rough=np.float64(nval), kinematic=kinematic, diag=diag,
                         outlet=(ny_global - 2, nx // 2), row0=lay.row0, ny_global=ny_global,
                         halo=halo, device=device, finalize=True,
                         n_pixels_global=ny * nx, d0=d0, packed=packed,
                         q_buffers=None if peer is None else peer.q_buffers,
                         d_buffer=None if peer is None else peer.d_buffer, **engine_kw)
    apply_sources(eng, sources or ('ga' if infil else 'none'), row0=lay.row0, halo=halo, seed=seed)
    torch.cuda.current_stream().synchronize()
    stepper = SlabChannels(lay, eng, group, peer=peer) if world > 1 else eng
    if return_parts:
        return stepper, dict(DEM=DEM, d8=d8, width=width, angle=angle, nval=nval, layout=lay)
    return stepper
