"""numpy restatement of the TopoFlow 3.6 hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *oracle of record* for the floating-point (fp64) routing
path: it re-states, expression by expression and in the same evaluation
order, what the reference's numpy code computes, so that on the same inputs
it reproduces the reference bit-for-bit (checked live against the unmodified
reference in ``tests/test_oracle_vs_reference.py`` whenever /root/reference is
present, and against the committed fixtures in ``tests/golden/`` otherwise).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package
(``topoflow36_b200``) never does; it has no CPU path.

Parity status: PINNED -- D8 codes/area against the reference's own RiverTools
fixtures (Treynor_flow.rtg, Treynor_area.rtg: bit-exact) and against live
reference runs; routing against live reference runs (Treynor June-20-1967
storm, 1592 steps: bit-exact in every fp64 grid).

All ``file:line`` citations are relative to /root/reference/topoflow/.
"""
import numpy as np

# ---------------------------------------------------------------------------
# D8 constants  (components/d8_base.py:561-607, :624-640)
# neighbour index k = 0..7  <->  NE, E, SE, S, SW, W, NW, N
# ---------------------------------------------------------------------------
CODE_LIST = np.int16([1, 2, 4, 8, 16, 32, 64, 128])
CODE_OPPS = np.int16([16, 32, 64, 128, 1, 2, 4, 8])
# (drow, dcol) of the cell a code-k cell drains to / of neighbour k
NBR_DROW = np.array([-1, 0, 1, 1, 1, 0, -1, -1])
NBR_DCOL = np.array([1, 1, 1, 0, -1, -1, -1, 0])


_RESOLVE_HEX = (
    "0001020204010202080802020808020210010202040402020808080208080802"
    "2020020220200202080808020808080220202020202002020808080808080808"
    "4040020204010202080802020808020210400202100102020808080208080802"
    "2020202020200202202020020808080220202020202020022020202008080808"
    "8080808080800202088002020880020280808080808002020808080208080802"
    "2080808020800202208002020808080220202080202020022020200808080802"
    "8080808080808080808080808080020280808080808080800880808008080802"
    "2080808020808080208080802080028020202080202020802020202020200808")


def resolve_table():
    """256-entry tie-break table: index = sum of the tied direction codes,
    value = the single code RiverTools keeps.  It is DATA, not a rule (e.g.
    {E,S,SW}->S but {NE,E,S,SW}->E), so it is carried as a literal dumped from
    the reference's get_resolve_array (components/d8_base.py:711-772) by
    oracle/gen_golden.py and pinned by tests: sum 10161, 255 non-zero entries,
    sha256 b0cd03c82342c315...."""
    return np.frombuffer(bytes.fromhex(_RESOLVE_HEX), dtype=np.uint8).copy()


def pixel_dims(ny, xres, yres):
    """dx, dy, dd per row as float32 (fixed-length pixels).
    components/d8_base.py:505-518, utils/pixels.py:71-171."""
    dx = np.float64(xres)
    dy = np.float64(yres)
    dd = np.sqrt(dx ** 2 + dy ** 2)
    one = np.zeros(ny, dtype=np.float64)
    return (np.float32(one + dx), np.float32(one + dy), np.float32(one + dd))


def d8_start_codes(DEM, dx, dy, dd, nodata=-9999.0, link_flats=True):
    """components/d8_global.py:181-353 (non-periodic). Returns int16 grid."""
    ny, nx = DEM.shape
    dx0, dy0, dd0 = dx[0], dy[0], dd[0]
    r = np.roll
    with np.errstate(invalid='ignore', over='ignore'):
        s1 = (DEM - r(r(DEM, 1, axis=0), -1, axis=1)) / dd0
        s2 = (DEM - r(DEM, -1, axis=1)) / dx0
        s3 = (DEM - r(r(DEM, -1, axis=0), -1, axis=1)) / dd0
        s4 = (DEM - r(DEM, -1, axis=0)) / dy0
        s5 = (DEM - r(r(DEM, -1, axis=0), 1, axis=1)) / dd0
        s6 = (DEM - r(DEM, 1, axis=1)) / dx0
        s7 = (DEM - r(r(DEM, 1, axis=0), 1, axis=1)) / dd0
        s8 = (DEM - r(DEM, 1, axis=0)) / dy0
        mx = np.maximum(s1, s2)
        for s in (s3, s4, s5, s6, s7, s8):
            np.maximum(mx, s, mx)
        g = CODE_LIST
        code = ((s1 == mx) * g[0] + (s2 == mx) * g[1] + (s3 == mx) * g[2] +
                (s4 == mx) * g[3] + (s5 == mx) * g[4] + (s6 == mx) * g[5] +
                (s7 == mx) * g[6] + (s8 == mx) * g[7])
        if not link_flats:
            code[mx <= 0] = 0
        else:
            flats = (mx == 0)
            code[flats] = -1 * code[flats]
            code[mx < 0] = -300
        bad = np.logical_or(DEM <= nodata, np.isfinite(DEM) != 1)
    code[bad] = 0
    code[:, 0] = 0
    code[:, nx - 1] = 0
    code[0, :] = 0
    code[ny - 1, :] = 0
    return code.astype(np.int16)


def d8_break_ties(code, resolve):
    """components/d8_global.py:355-372."""
    w = code > 0
    code[w] = resolve[code[w]]
    return code


def d8_link_flats(code, resolve):
    """components/d8_global.py:374-576.  Jacobi sweeps; returns (code, n_reps)."""
    ny, nx = code.shape
    g, h = CODE_LIST, CODE_OPPS
    not_edge = np.ones((ny, nx), dtype=bool)
    not_edge[0, :] = not_edge[-1, :] = False
    not_edge[:, 0] = not_edge[:, -1] = False
    n_reps = 0
    while True:
        n_reps += 1
        rows, cols = np.where((code < 0) & (code != -300) & not_edge)
        n_flats = rows.size
        if n_flats == 0:
            break
        cur = (-1 * code[rows, cols]).astype(np.int64)
        ready = np.zeros(n_flats, dtype=np.int64)
        for k in range(8):
            nb = code[(rows + NBR_DROW[k]) % ny, (cols + NBR_DCOL[k]) % nx]
            valid = (nb > 0) & (nb != h[k])
            allowed = (cur & int(g[k])) != 0
            ready += (valid & allowed) * int(g[k])
        fix = ready > 0
        n_fix = int(fix.sum())
        if n_fix:
            code[rows[fix], cols[fix]] = resolve[ready[fix]]
        if n_fix == n_flats or n_fix == 0:
            break
    return code, n_reps


def d8_finalize(code):
    """components/d8_global.py:139-146: drop leftovers, map to uint8."""
    code = code.copy()
    code[(code < 0) | (code > 128)] = 0
    vmap = np.zeros(256, dtype=np.uint8)
    vmap[CODE_LIST] = np.uint8(CODE_LIST)
    return vmap[code]


def d8_flow_grid(DEM, dx, dy, dd, nodata=-9999.0, link_flats=True):
    res = resolve_table()
    code = d8_start_codes(DEM, dx, dy, dd, nodata, link_flats)
    code = d8_break_ties(code, res)
    n_reps = 0
    if link_flats:
        code, n_reps = d8_link_flats(code, res)
    return d8_finalize(code), n_reps


def d8_parent_ids(code):
    """components/d8_global.py:578-688 (non-periodic): (rows, cols) of parent;
    (0, 0) for noflow / edge cells."""
    ny, nx = code.shape
    inc = np.zeros(129, dtype=np.int32)
    inc[CODE_LIST] = np.array([-nx + 1, 1, nx + 1, nx, nx - 1, -1, -nx - 1, -nx],
                              dtype=np.int32)
    ids = np.arange(nx * ny, dtype=np.int32).reshape(ny, nx)
    pid = ids + inc[code]
    pid[:, 0] = 0
    pid[:, -1] = 0
    pid[0, :] = 0
    pid[-1, :] = 0
    pid[code <= 0] = 0
    return divmod(pid, nx)


def d8_ds_dw(code, dx, dy, dd):
    """components/d8_global.py:873-1027 -> (ds, dw) float32 grids."""
    ny, nx = code.shape
    rows = np.arange(ny)[:, None] + np.zeros((1, nx), dtype=np.int64)
    diag = (code == 1) | (code == 4) | (code == 16) | (code == 64)
    horiz = (code == 2) | (code == 32)
    vert = (code == 8) | (code == 128)
    none = (code == 0)
    ds = np.zeros((ny, nx), dtype=np.float32)
    dw = np.zeros((ny, nx), dtype=np.float32)
    ds[diag] = dd[rows[diag]]
    dw[diag] = dd[rows[diag]]
    ds[horiz] = dx[rows[horiz]]
    dw[horiz] = dy[rows[horiz]]
    ds[vert] = dy[rows[vert]]
    dw[vert] = dx[rows[vert]]
    ds[none] = dx[rows[none]]
    dw[none] = dx[rows[none]]
    return ds, dw


def d8_area_levels(code, pixel_area):
    """components/d8_global.py:1029-1332, whole-grid level sweeps exactly as the
    reference does them (O(levels x N)); use only at small sizes."""
    ny, nx = code.shape
    h = CODE_OPPS
    A = np.zeros((ny, nx), dtype=np.float32)
    pa = np.asarray(pixel_area)
    n_reps = 0
    while True:
        n_reps += 1
        rows, cols = np.where((A == 0) & (code != 0))
        if rows.size == 0:
            break
        a = []
        d = []
        for k in range(8):
            rr = (rows + NBR_DROW[k]) % ny
            cc = (cols + NBR_DCOL[k]) % nx
            a.append(A[rr, cc])
            d.append(code[rr, cc])
        ready = np.ones(rows.size, dtype=bool)
        for k in range(8):
            ready &= ~((d[k] == h[k]) & (a[k] == 0))
        WR = np.where(ready)[0]
        if WR.size == 0:
            break
        r2, c2 = rows[WR], cols[WR]
        A[r2, c2] = pa if pa.size == 1 else pa[r2, c2]
        for k in range(8):
            w = np.where(d[k][WR] == h[k])[0]
            if w.size:
                A[r2[w], c2[w]] += a[k][WR[w]]
        if WR.size == rows.size:
            break
    return A, n_reps


def d8_area_kahn(code, pixel_area, with_counts=False):
    """O(N) restatement of update_area_grid: same per-cell result because each
    cell's value is pixel_area (stored fp32) plus its children's FINAL values
    added in neighbour order k=0..7 in fp32 -- only the schedule differs
    (frontier of ready cells instead of whole-grid sweeps).  Optionally also
    returns exact integer contributing-cell counts (int64)."""
    ny, nx = code.shape
    h = CODE_OPPS
    pa = np.asarray(pixel_area)
    A = np.zeros((ny, nx), dtype=np.float32)
    N = np.zeros((ny, nx), dtype=np.int64)
    flow = code != 0
    # number of children (neighbours draining into the cell)
    nchild = np.zeros((ny, nx), dtype=np.int32)
    for k in range(8):
        nb = np.roll(np.roll(code, -NBR_DROW[k], axis=0), -NBR_DCOL[k], axis=1)
        nchild += (nb == h[k])
    prow, pcol = d8_parent_ids(code)
    rows, cols = np.where(flow & (nchild == 0))
    pending = nchild.copy()
    while rows.size:
        a = np.full(rows.size, 0, dtype=np.float32)
        a[:] = pa if pa.size == 1 else pa[rows, cols]
        n = np.ones(rows.size, dtype=np.int64)
        for k in range(8):
            rr = (rows + NBR_DROW[k]) % ny
            cc = (cols + NBR_DCOL[k]) % nx
            w = code[rr, cc] == h[k]
            a[w] += A[rr[w], cc[w]]
            n[w] += N[rr[w], cc[w]]
        A[rows, cols] = a
        N[rows, cols] = n
        pr, pc = prow[rows, cols], pcol[rows, cols]
        ok = flow[pr, pc] & ~((pr == 0) & (pc == 0))
        pr, pc = pr[ok], pc[ok]
        np.subtract.at(pending, (pr, pc), 1)
        lin = np.unique(pr.astype(np.int64) * nx + pc)
        r2, c2 = np.divmod(lin, nx)
        done = pending[r2, c2] == 0
        rows, cols = r2[done], c2[done]
    if with_counts:
        return A, N
    return A


def d8_aspect(code):
    """components/d8_global.py:1350-1376: flow angle, radians CCW from east, float32."""
    a = np.linspace(0.0, 2 * np.pi, 9)
    aspect = np.zeros(code.shape, dtype='float32')
    for cd, k in ((1, 1), (2, 0), (4, 7), (8, 6), (16, 5), (32, 4), (64, 3), (128, 2)):
        aspect[code == cd] = a[k]
    return aspect


def d8_slope(DEM, code, ds):
    """components/d8_global.py:1334-1348."""
    pr, pc = d8_parent_ids(code)
    with np.errstate(invalid='ignore', divide='ignore'):
        S = (DEM - DEM[pr, pc]) / ds
    S[code <= 0] = 0.0
    return S


# ---------------------------------------------------------------------------
# DEM / geometry preparation  (utils/fill_pits.py, new_slopes.py, parameterize.py)
# ---------------------------------------------------------------------------
def fill_pits_seeds(DEM, nodata=np.float32(-9999), closed_basin_code=-32000):
    """utils/fill_pits.py get_start_pixels :178-356 -> boolean seed grid.  Non-finite
    values of DEM are set to nodata IN PLACE (:232-235).  The 8 'neighbours' of an
    interior cell are rolls of the interior sub-array (they wrap inside it)."""
    ny, nx = DEM.shape
    DEM[~np.isfinite(DEM)] = nodata
    seed = np.zeros((ny, nx), dtype=bool)
    seed[0, :] = DEM[0, :] > nodata
    seed[-1, :] = DEM[-1, :] > nodata
    seed[1:-1, 0] = DEM[1:-1, 0] > nodata
    seed[1:-1, -1] = DEM[1:-1, -1] > nodata
    zc = DEM[1:ny - 1, 1:nx - 1]
    zs = np.full(zc.shape, np.inf, dtype=DEM.dtype)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr or dc:
                zs = np.minimum(zs, np.roll(np.roll(zc, dr, axis=0), dc, axis=1))
    seed[1:-1, 1:-1] = (zc == closed_basin_code) | ((zc > nodata) & (zs <= nodata))
    return seed


def fill_pits(DEM, nodata=np.float32(-9999), closed_basin_code=-32000, close_nodata=False):
    """utils/fill_pits.py fill_pits :358-640 with the heapq branch (USE_MY_HEAP = False):
    priority flood from the seeds; an OPEN neighbour of the popped cell is raised to the
    popped elevation if lower, then pushed.  In place; returns the number raised.
    Neighbour IDs are flat offsets checked against [0, n) only (:523-541), as there.

    What the reference DOES with nodata (close_nodata=False, the default here): :404-409
    takes ``where(DEM <= nodata)`` of the 2-D grid and keeps only its ROW numbers, then
    closes the cells whose FLAT index equals one of those row numbers.  So the nodata
    cells themselves stay open, lie lowest and are flooded up to their rim like any
    depression, while a handful of cells with flat index < ny are walled off instead
    (unless they are seeds, :437).  close_nodata=True is the documented intent (nodata
    cells closed)."""
    import heapq
    ny, nx = DEM.shape
    seed = fill_pits_seeds(DEM, nodata, closed_basin_code)
    flat = DEM.reshape(-1)
    n = flat.size
    OPEN, CLOSED, ON_HEAP = 0, 1, 2
    C = np.zeros(n, dtype=np.uint8)
    if close_nodata:
        C[flat <= nodata] = CLOSED
    else:
        C[np.where(DEM <= nodata)[0]] = CLOSED          # row numbers used as flat IDs
    ids = np.flatnonzero(seed.reshape(-1))
    heap = list(zip(flat[ids].tolist(), ids.tolist()))
    heapq.heapify(heap)
    C[ids] = ON_HEAP
    incs = (-nx - 1, -nx, -nx + 1, -1, 1, nx - 1, nx, nx + 1)
    zl = flat.tolist()
    Cl = C.tolist()
    n_raised = 0
    while heap:
        zmin, k = heapq.heappop(heap)
        Cl[k] = CLOSED
        for d in incs:
            j = k + d
            if 0 <= j < n and Cl[j] == OPEN:
                if zl[j] < zmin:
                    zl[j] = zmin
                    n_raised += 1
                Cl[j] = ON_HEAP
                heapq.heappush(heap, (zl[j], j))
    flat[:] = np.array(zl, dtype=DEM.dtype)
    return n_raised


def new_slope_grid(DEM, code, ds, nodata=-9999.0):
    """utils/new_slopes.py get_new_slope_grid :174-316 -> (slope float64, n_reps).
    All cells walk down their flow paths together; a cell takes the first positive drop
    from ITS OWN elevation over the accumulated path length.  Walkers that pass a noflow
    cell continue on cell (0, 0) (parent ID 0) until the longest path is done."""
    ny, nx = code.shape
    pr, pc = d8_parent_ids(code)
    pid_grid = (pr * nx + pc).astype(np.int32)
    z2 = DEM
    pIDs = (pr, pc)
    distance = np.float64(ds)
    slope = np.zeros((ny, nx), dtype='float64')
    n_reps = 0
    DONE = False
    while not DONE:
        z1 = DEM[pIDs]
        with np.errstate(invalid='ignore'):
            dz = z2 - z1
            w1 = np.logical_or(z1 <= nodata, code == 0)
            dz[w1] = 0
            wg = np.logical_and(slope == 0, dz > 0)
        nb = np.invert(wg).sum()
        if wg.sum() > 0:
            slope[wg] = dz[wg] / distance[wg]
        distance += ds[pIDs]
        pID_vals = pid_grid[pIDs]
        pIDs = divmod(pID_vals, nx)
        n_reps += 1
        DONE = (nb == 0) or (pID_vals.max() == 0) or (n_reps > 4 * nx)
    return slope, n_reps


def grid_from_TCA(A, A1=None, g1=None, A2=None, g2=None, p=None):
    """utils/parameterize.py get_grid_from_TCA :93-215 without the file IO ->
    (grid, c, p); numpy's own dtype rules decide float32 vs float64."""
    if g1 is not None and g2 is not None and A1 is not None and A2 is not None:
        p = np.log(g1 / g2) / np.log(A1 / A2)
        c = g1 / (A1 ** p)
    elif p is not None and g1 is not None:
        A1 = A.max()
        c = g1 / (A1 ** p)
    elif g1 is not None and g2 is not None:
        A1 = A.max()
        A2 = A[A > 0].min()
        p = np.log(g1 / g2) / np.log(A1 / A2)
        c = g1 / (A1 ** p)
    else:
        raise ValueError('not enough input parameters to compute c & p')
    w1 = (A <= 0)
    if p < 0:
        Ac = A.copy()
        Ac[w1] = 1.0
        grid = c * Ac ** p
    else:
        grid = c * A ** p
    grid[w1] = 1.0
    return grid, c, p


# ---------------------------------------------------------------------------
# Channel routing  (components/channels_base.py)
# ---------------------------------------------------------------------------
G_ACCEL = np.float64(9.81)
AVAL = np.float64(0.476)
KAPPA = np.float64(0.408)
LAW_CONST = np.sqrt(G_ACCEL) / KAPPA
ONE_THIRD = np.float64(1.0) / 3.0
TWO_THIRDS = np.float64(2.0) / 3.0
DEG_TO_RAD = np.pi / np.float64(180)


def remove_bad_slopes(slope):
    """components/channels_base.py:4164-4260 (in place)."""
    bad = np.isnan(slope) | (slope == 0.0) | (slope < 0.0) | np.isinf(slope)
    if bad.any():
        slope[bad] = slope[~bad].min()
    return slope


class Channels(object):
    """State + step of channels_component for the non-flood path.

    Mirrors initialize_computed_vars (:1137-1405) and update (:803-962).
    Inputs follow the reference's conventions: grids are (ny, nx) float64;
    anything that may be 'Scalar' in the CFG is a 0-D float64 array.
    """

    def __init__(self, code, ds32, dw32, slope, width, angle_deg, nval=None,
                 z0val=None, sinu=1.0, d0=0.0, da=900.0, dt=6.0,
                 kinematic=True, manning=True, outlet=(0, 0),
                 rho_H2O=1000.0, check_stability=True, flood=False, d_bankfull=10.0,
                 attenuate=False, n_days=5.0):
        f64 = np.float64
        self.code = code
        self.ny, self.nx = code.shape
        self.KINEMATIC_WAVE = kinematic
        self.MANNING = manning
        self.LAW_OF_WALL = not manning
        self.dt = np.array(dt, dtype=f64)
        self.da = np.array(da, dtype=f64)
        self.rho_H2O = np.array(rho_H2O, dtype=f64)
        self.outlet_ID = outlet
        self.CHECK_STABILITY = check_stability
        self.noflow = np.where(code <= 0)
        self.parent_IDs = d8_parent_ids(code)
        self.n_pixels = self.ny * self.nx
        # ---- :1196  angle to radians (in place on a float64 copy)
        self.angle = np.array(angle_deg, dtype=f64)
        self.angle *= DEG_TO_RAD
        # ---- :1206-1209 zero widths -> dw
        self.width = np.array(width, dtype=f64)
        if self.width.ndim == 2:
            w1 = (self.width == 0)
            if w1.sum() > 0:
                self.width[w1] = dw32[w1]
        self.sinu = np.array(sinu, dtype=f64)
        self.nval = None if nval is None else np.array(nval, dtype=f64)
        self.z0val = None if z0val is None else np.array(z0val, dtype=f64)
        # ---- :1242-1249
        self.ds = (self.sinu * ds32)           # float64 under numpy 2
        self.slope = (np.array(slope, dtype=f64) / self.sinu)
        self.S_bed = self.slope
        self.S_free = self.S_bed.copy()
        z = lambda: np.zeros((self.ny, self.nx), dtype=f64)
        self.u, self.f, self.d = z(), z(), z()
        self.d += np.array(d0, dtype=f64)
        self.R = z()
        self.tau, self.u_star, self.froude = z(), z(), z()
        self.vol_R = np.array(0, dtype=f64)
        self.vol_Q = np.array(0, dtype=f64)
        self.vol_edge = np.array(0, dtype=f64)
        if kinematic:
            remove_bad_slopes(self.slope)
        # ---- :1331-1336
        L2 = self.d * np.tan(self.angle)
        self.A_wet = self.d * (self.width + L2)
        self.P_wet = self.width + (f64(2) * self.d / np.cos(self.angle))
        self.vol = self.A_wet * self.ds
        self.Q = z()
        self.Rh = z()
        self.ATTENUATE, self.n_days = bool(attenuate), n_days
        self.vol_stored = z()                                  # :1343
        # ---- flood option :1346-1388 (rectangle over trapezoid, Sep. 2019)
        self.FLOOD_OPTION = bool(flood)
        self.d_flood = z()
        self.vol_flood = z()
        if self.FLOOD_OPTION:
            self.vol_chan = self.vol
            self.vol = self.vol_chan.copy()
            self.Qc, self.Qf = z(), z()
            self.flood_manning_n = 0.10
            self.width_ratio = 5.0
            self.d_bankfull = np.array(d_bankfull, dtype=f64)
            L3 = self.d_bankfull * np.tan(self.angle)
            Ac_bankfull = self.d_bankfull * (self.width + L3)
            self.vol_bankfull = Ac_bankfull * self.ds
        self.vol_chan_sum0 = self.vol.sum()
        for n in ('Q_outlet', 'u_outlet', 'd_outlet', 'f_outlet', 'Q_peak',
                  'T_peak', 'u_peak', 'Tu_peak', 'd_peak', 'Td_peak'):
            setattr(self, n, np.array(0, dtype=f64))
        self.time = np.array(0, dtype=f64)
        self.time_min = np.array(0, dtype=f64)
        self.time_index = 0
        zero = np.array(0, dtype=f64)
        self.P_rain, self.SM, self.GW, self.MR = zero, zero, zero, zero
        self.ET, self.IN = zero, zero

    # ---------------------------------------------------------------- update
    def step(self):
        f64 = np.float64
        dt = self.dt
        # update_R :1548-1655
        P, SM, GW, ET, IN, MR = (self.P_rain, self.SM, self.GW, self.ET,
                                 self.IN, self.MR)
        w = (self.vol < (ET * self.da * self.dt))
        if np.ndim(ET) == self.vol.ndim:
            ET[w] = 0.0
        else:
            ET = 0.0
        self.R = (P + SM + GW + MR) - (ET + IN)
        # update_R_integral :1657-1672
        volume = f64(self.R * self.da * self.dt)
        if np.size(volume) == 1:
            self.vol_R += (volume * self.n_pixels)
        else:
            self.vol_R += np.sum(volume)
        # update_channel_discharge :1697 (+ update_flood_discharge :1701-1750,
        # update_discharge :1791-1862 when FLOOD_OPTION)
        if self.FLOOD_OPTION:
            self.Qc[:] = self.u * self.A_wet
            Sf = self.S_bed if self.KINEMATIC_WAVE else self.S_free
            uf = (self.d_flood ** TWO_THIRDS) * np.sqrt(np.abs(Sf)) / self.flood_manning_n
            Af = self.width_ratio * self.width * self.d_flood
            self.Qf[:] = uf * Af
            self.Q[:] = self.Qc + self.Qf
        else:
            self.Q[:] = self.u * self.A_wet
        # update_flow_volume :2077-2145
        if self.ATTENUATE:
            # all runoff passes through a linear (Nash) reservoir per cell :2077-2084
            self.vol_stored += (self.R * self.da) * dt
            t_drain = 3600.0 * 24.0 * self.n_days
            Q_sides = (self.vol_stored / t_drain)
            vol_sides = np.minimum(Q_sides * dt, self.vol_stored)
            self.vol += vol_sides
            self.vol_stored -= vol_sides
            np.maximum(self.vol_stored, 0.0, self.vol_stored)
        else:
            self.vol += (self.R * self.da) * dt
        pr, pc = self.parent_IDs
        for k in range(8):
            wk = np.where(self.code == CODE_LIST[k])
            if wk[0].size:
                self.vol[pr[wk], pc[wk]] += (dt * self.Q[wk])
        self.vol -= (self.Q * dt)
        np.maximum(self.vol, 0.0, self.vol)
        if self.FLOOD_OPTION:
            # update_channel_volume :2149, update_flood_volume :2158-2185
            np.minimum(self.vol, self.vol_bankfull, self.vol_chan)
            np.maximum(self.vol - self.vol_bankfull, 0.0, self.vol_flood)
        # update_channel_depth :2250-2391
        self._update_depth()
        if self.FLOOD_OPTION:
            # update_flood_depth :2393-2448
            denom = (self.width_ratio * self.width * self.ds)
            np.maximum(self.vol_flood / denom, 0.0, self.d_flood)
        # update_trapezoid_Rh :2643-2712
        d, wb = self.d, self.width
        L2 = d * np.tan(self.angle)
        A_wet = d * (wb + L2)
        P_wet = wb + (f64(2) * d / np.cos(self.angle))
        P_wet[self.noflow] = f64(1)
        Rh = (A_wet / P_wet)
        Rh[self.noflow] = f64(0)
        self.Rh[:] = Rh
        self.A_wet[:] = A_wet
        self.P_wet[:] = P_wet
        # update_free_surface_slope :2583-2618
        if not self.KINEMATIC_WAVE:
            if self.FLOOD_OPTION:
                total_d = (self.d + self.d_flood)
                delta_d = total_d - total_d[self.parent_IDs]
            else:
                delta_d = (self.d - self.d[self.parent_IDs])
            self.S_free[:] = self.S_bed + (delta_d / self.ds)
        # shear stress / speed :2620-2641
        slope = self.S_bed if self.KINEMATIC_WAVE else np.abs(self.S_free)
        self.tau[:] = self.rho_H2O * G_ACCEL * self.d * slope
        self.u_star[:] = np.sqrt(self.tau / self.rho_H2O)
        # update_friction_factor :2714-2833
        wg = self.d_is_pos
        wbad = self.d_is_zero
        if self.MANNING:
            if self.nval.size > 1:
                n2 = self.nval[wg] ** f64(2)
            else:
                n2 = self.nval ** f64(2)
            self.f[wg] = G_ACCEL * (n2 / (self.d[wg] ** ONE_THIRD))
            self.f[wbad] = f64(0)
        if self.LAW_OF_WALL:
            smooth = (AVAL / self.z0val) * self.d
            np.maximum(smooth, f64(1.1), out=smooth)
            self.f[wg] = (KAPPA / np.log(smooth[wg])) ** f64(2)
            self.f[wbad] = f64(0)
        # update_velocity (kinematic_wave.py:84-127, :3960-4063)
        S = self.S_bed if self.KINEMATIC_WAVE else self.S_free
        S2 = np.abs(S)
        if self.MANNING:
            self.u[:] = (self.Rh ** TWO_THIRDS) * np.sqrt(S2) / self.nval
        if self.LAW_OF_WALL:
            smooth = (AVAL / self.z0val) * self.d
            np.maximum(smooth, f64(1.1), out=smooth)
            self.u[:] = LAW_CONST * np.sqrt(self.Rh * S2) * np.log(smooth)
        # update_velocity_on_edges :2851
        self.u[self.noflow] = f64(0)
        # update_froude_number :2867
        self.froude[wg] = self.u[wg] / np.sqrt(G_ACCEL * self.d[wg])
        self.froude[wbad] = f64(0)
        # outlet, peaks, integrals :2884-2936
        o = self.outlet_ID
        self.Q_outlet.fill(self.Q[o])
        self.u_outlet.fill(self.u[o])
        self.d_outlet.fill(self.d[o])
        self.f_outlet.fill(self.f[o])
        if self.Q_outlet > self.Q_peak:
            self.Q_peak.fill(self.Q_outlet)
            self.T_peak.fill(self.time_min)
        if self.u_outlet > self.u_peak:
            self.u_peak.fill(self.u_outlet)
            self.Tu_peak.fill(self.time_min)
        if self.d_outlet > self.d_peak:
            self.d_peak.fill(self.d_outlet)
            self.Td_peak.fill(self.time_min)
        self.vol_Q += (self.Q_outlet * self.dt)
        # update_edge_values :2938-2984
        self.vol_edge += self.vol[self.noflow].sum()
        self.vol[self.noflow] = 0.0
        self.d[self.noflow] = 0.0
        if self.FLOOD_OPTION:
            self.vol_flood[self.noflow] = 0.0
            self.d_flood[self.noflow] = 0.0
        # checks :3193-3454
        self.n_bad_depth = int((self.d < 0.0).sum() + np.isinf(self.d).sum())
        self.n_bad_vel = int((self.u < 0.0).sum() + np.isnan(self.u).sum()
                             + np.isinf(self.u).sum())
        if self.CHECK_STABILITY:
            if self.n_bad_depth:
                raise RuntimeError('Negative or NaN depth found.')
            if self.n_bad_vel:
                raise RuntimeError('Found bad velocities.')
        # update_time (utils/BMI_base.py:1029-1090)
        self.time_index += 1
        self.time += self.dt
        self.time_min = self.time / f64(60)

    def _update_depth(self):
        vol = self.vol_chan if self.FLOOD_OPTION else self.vol            # :2251
        width, angle = self.width, self.angle
        d = self.d
        with np.errstate(invalid='ignore', divide='ignore'):
            if np.size(angle) == 1:
                if angle == 0.0:
                    d = vol / (width * self.ds)
                else:
                    denom = 2.0 * np.tan(angle)
                    arg = 2.0 * denom * vol / self.ds
                    arg += width ** (2.0)
                    d = (np.sqrt(arg) - width) / denom
            else:
                w1 = (angle == 0)
                w2 = np.invert(w1)
                A_top = width[w1] * self.ds[w1]
                d[w1] = vol[w1] / A_top
                denom = 2.0 * np.tan(angle[w2])
                arg = 2.0 * denom * vol[w2] / self.ds[w2]
                arg += width[w2] ** (2.0)
                d[w2] = (np.sqrt(arg) - width[w2]) / denom
        np.maximum(d, 0.0, self.d)
        self.d_is_pos = (self.d > 0)
        self.d_is_zero = np.invert(self.d_is_pos)


# ---------------------------------------------------------------------------
# Source terms (each a separate BMI component in the reference)
# ---------------------------------------------------------------------------
def met_split_precip(P, T_air, T_rain_snow):
    """components/met_base.py:1300-1388."""
    return P * (T_air > T_rain_snow), P * (T_air <= T_rain_snow)


def green_ampt_step(I, P_rain, SM, Ks, Ki, qs, qi, G, dt,
                    h_table=None, elev=None):
    """infil_green_ampt.py:511-578 + infil_base.py:562-571, :654-709,
    :954-995, :782-811.  Returns (IN, I_new)."""
    P_total = (P_rain + SM)
    dK = (Ks - Ki)
    dq = (qs - qi)
    fc = (G * dK * dq / I) + Ks
    IN = np.minimum(fc, P_total)
    if np.size(P_total) == 1 and np.size(Ks) == 1:
        if np.size(IN) == 1:
            if P_total < Ks:
                IN = P_total
        else:
            if P_total < Ks:
                IN = np.zeros_like(IN) + P_total
    else:
        IN = np.array(IN, dtype=np.float64)
        w = (P_total < Ks)
        w = np.broadcast_to(w, IN.shape)
        if np.size(P_total) > 1:
            IN[w] = np.broadcast_to(P_total, IN.shape)[w]
        else:
            IN[w] = P_total
    if np.size(IN) > 1 and h_table is not None and elev is not None:
        IN[np.broadcast_to(h_table >= elev, IN.shape)] = 0.0
    I_new = I + (IN * dt)
    return IN, I_new


def degree_day_step(h_swe, P_snow, T_air, c0, T0, dt, rho_H2O=1000.0,
                    rho_snow=300.0):
    """snow_degree_day.py:293-360; snow_base.py:490-512, :559-609, :631-703.
    Returns (SM, h_swe_new, h_snow_new)."""
    f64 = np.float64
    M = (c0 / f64(8.64E7)) * (T_air - T0)
    SM = np.array(M, dtype=f64) + np.zeros_like(h_swe)
    SM_max = h_swe / dt
    np.minimum(SM, SM_max, out=SM)
    np.maximum(SM, f64(0), out=SM)
    h_new = np.array(h_swe, dtype=f64, copy=True)
    h_new += (P_snow * dt)
    h_new -= (SM * dt)
    np.maximum(h_new, f64(0), h_new)
    h_snow = h_new * (rho_H2O / rho_snow)
    return SM, h_new, h_snow


def priestley_taylor_step(T_air, T_surf, Qn_SW, Qn_LW, alpha, K_soil, soil_x,
                          T_soil_x):
    """evap_base.py:326-353; evap_priestley_taylor.py:308-413. Returns ET."""
    f64 = np.float64
    delta_T = (T_surf - T_soil_x)
    Qc = K_soil * delta_T / soil_x
    Q_net = Qn_SW + Qn_LW
    Qet = alpha * (f64(0.406) + (f64(0.011) * T_air)) * (Q_net + Qc)
    ET = (Qet / f64(2.5E+9))
    return np.maximum(ET, f64(0))


# ---------------------------------------------------------------------------
# Meteorology energy chain (components/met_base.py update() :790-809 with
# PRECIP_ONLY = No) and the clear-sky radiation model it calls
# (components/solar_funcs.py).  Point-wise; every function cites its source.
# ---------------------------------------------------------------------------
I_SC = np.float64(1361.5)                                   # solar_funcs.py:133
OMEGA_H = (np.float64(360) / np.float64(24)) * (np.pi / np.float64(180))   # :236-246


def solar_day_terms(julian_day):
    """Day angle, declination, eccentricity correction (solar_funcs.py:147-234)."""
    f = np.float64
    G = (2 * np.pi) * julian_day / f(365)
    E0 = f(1.000110) + (f(0.034221) * np.cos(G)) + (f(0.001280) * np.sin(G)) + \
        (f(0.000719) * np.cos(f(2) * G)) + (f(0.000077) * np.sin(f(2) * G))
    delta = f(0.006918) - (f(0.399912) * np.cos(G)) + (f(0.070257) * np.sin(G)) - \
        (f(0.006758) * np.cos(f(2) * G)) + (f(0.000907) * np.sin(f(2) * G)) - \
        (f(0.002697) * np.cos(f(3) * G)) + (f(0.001480) * np.sin(f(3) * G))
    return G, delta, E0


def _zenith(lat_deg, delta, th):                             # solar_funcs.py:248-269
    lat_rad = lat_deg * (np.pi / np.float64(180))
    term1 = np.sin(lat_rad) * np.sin(delta)
    term2 = np.cos(lat_rad) * np.cos(delta) * np.cos(OMEGA_H * th)
    return np.arccos(term1 + term2)


def _optical_air_mass(lat_deg, delta, th):                   # :477-551
    a, b, c = 0.50572, 6.07995, 1.6364
    Z_deg = _zenith(lat_deg, delta, th) * (180 / np.pi)
    gamma = np.maximum(90.0 - Z_deg, 0.0)
    term1 = np.sin(gamma * (np.pi / 180))
    term2 = a / (gamma + b) ** c
    return np.float64(1) / (term1 + term2)


def _et_flux(lat_deg, delta, E0, th):                        # :354-411
    lat_rad = lat_deg * (np.pi / np.float64(180))
    term1 = np.cos(delta) * np.cos(lat_rad) * np.cos(OMEGA_H * th)
    term2 = np.sin(delta) * np.sin(lat_rad)
    return np.maximum(I_SC * E0 * (term1 + term2), 0.0)


def _lon_offset(lat_deg, alpha, beta):                       # :701-721
    lat_rad = lat_deg * (np.pi / np.float64(180))
    term1 = np.sin(beta) * np.sin(alpha)
    term2 = np.cos(beta) * np.cos(lat_rad)
    term3 = np.sin(beta) * np.sin(lat_rad) * np.cos(alpha)
    return np.arctan(term1 / (term2 - term3))


def _eq_lat(lat_deg, alpha, beta):                           # :723-751 (radians)
    lat_rad = lat_deg * (np.pi / np.float64(180))
    term1 = np.sin(beta) * np.cos(alpha) * np.cos(lat_rad)
    term2 = np.cos(beta) * np.sin(lat_rad)
    return np.arcsin(term1 + term2)


def _sun_offset(lat_deg, delta, sign):                       # :286-340
    lat_rad = lat_deg * (np.pi / np.float64(180))
    arg = -np.float64(1) * np.tan(lat_rad) * np.tan(delta)
    arg = np.minimum(np.maximum(-1, arg), 1)
    return sign * np.arccos(arg) / OMEGA_H


def clear_sky_radiation(lat_deg, julian_day, W_p, th, alpha, beta, albedo, dust):
    """solar_funcs.Clear_Sky_Radiation :870-942 and everything below it."""
    f = np.float64
    G, delta, E0 = solar_day_terms(julian_day)
    M_opt = _optical_air_mass(lat_deg, delta, th)
    a_sa = -f(0.1240) - (f(0.0207) * W_p)                     # :567-594
    b_sa = -f(0.0682) - (f(0.0248) * W_p)
    tau = np.minimum(np.maximum(np.exp(a_sa + (b_sa * M_opt)) - dust, 0), 1)
    lat_eq = _eq_lat(lat_deg, alpha, beta)                    # :822-868
    dlon = _lon_offset(lat_deg, alpha, beta)
    term1 = np.cos(delta) * np.cos(lat_eq)
    term2 = np.cos((OMEGA_H * th) + dlon)
    term3 = np.sin(lat_eq) * np.sin(delta)
    K_ET_slope = np.maximum(I_SC * E0 * ((term1 * term2) + term3), 0)
    a_s = -f(0.0363) - (f(0.0084) * W_p)                      # :619-636
    b_s = -f(0.0572) - (f(0.0173) * W_p)
    gam_s = (1 - np.exp(a_s + (b_s * M_opt))) + dust
    K_ET = _et_flux(lat_deg, delta, E0, th)
    K_dif = f(0.5) * gam_s * K_ET                             # :638-651
    K_global = (tau * K_ET) + K_dif                           # :653-669
    K_bs = f(0.5) * gam_s * albedo * K_global                 # :671-699
    K_cs = (tau * K_ET_slope) + K_dif + K_bs
    eq_lat_deg = lat_eq * (f(180) / np.pi)
    t_noon = -f(1) * dlon / OMEGA_H                           # :753-761
    T_sr = np.maximum(_sun_offset(eq_lat_deg, delta, -f(1)) + t_noon,
                      _sun_offset(lat_deg, delta, -f(1)))     # :763-786
    T_ss = np.minimum(_sun_offset(eq_lat_deg, delta, f(1)) + t_noon,
                      _sun_offset(lat_deg, delta, f(1)))      # :788-811
    dark = np.logical_or((th <= T_sr), (th >= T_ss))
    K_cs = np.where(dark, f(0), K_cs)
    return K_cs


def albedo_update(st, P_snow, T_air, h_snow, h_ice, dt, rho_H2O=1000.0, rho_snow=300.0,
                  method='aging'):
    """met_base.update_albedo :1995-2045.  `st` = dict(albedo, n, ring) with ring of shape
    (int(259200/dt), ny, nx) (met_base.py:1121-1133); updated in place, returns albedo."""
    albedo = st['albedo']
    if method == 'aging':
        r = np.where((T_air > 0), 0.12, 0.05)
        K, alpha0 = 0.44, 0.4
        st['ring'] = np.roll(st['ring'], -1, axis=0)
        st['ring'][st['ring'].shape[0] - 1] = P_snow * dt * (rho_H2O / rho_snow)
        total = np.sum(st['ring'], axis=0)
        st['n'] = np.where((total >= 0.03), 0, st['n'])
        st['n'] = np.where((total < 0.03), st['n'] + dt / 86400, st['n'])
        snow_albedo = alpha0 + K * np.exp(-st['n'] * r)
        albedo = np.where((h_snow > 0), snow_albedo, albedo)
    else:
        albedo = np.where((h_snow > 0), np.float64(0.75), albedo)
    albedo = np.where(((h_snow == 0) & (h_ice > 0)), np.float64(0.3), albedo)
    albedo = np.where(((h_snow == 0) & (h_ice == 0)), np.float64(0.15), albedo)
    st['albedo'] = albedo
    return albedo


def met_energy_chain(T_air, RH, uz, z, z0_air, h_snow, h_ice, p0, albedo, em_surf, lat_deg,
                     alpha, beta, julian_day, TSN_offset, dust_atten=0.08, cloud_factor=0.0,
                     canopy_factor=0.0, Qa=0.0, Qc=0.0, satterlund=False):
    """met_base.update() :790-809 in its own order (albedo is an input: the 'aging'
    method's 3-day snowfall ring buffer is not part of this path).  Returns a dict."""
    f = np.float64
    g, kappa, rho_air, Cp_air, Lv = f(9.81), f(0.408), f(1.2614), f(1005.7), f(2500000)
    sigma, C_to_K = f(5.67E-8), f(273.15)

    def e_sat_mbar(T):                                        # :1657-1741
        if not satterlund:
            term1 = (f(17.3) * T) / (T + f(237.3))
            e = f(0.611) * np.exp(term1)
        else:
            term1 = f(2353) / (T + f(273.15))
            e = f(10) ** (f(11.4) - term1)
            e = (e / f(1000))
        return e * f(10)
    o = {}
    o['e_sat_air'] = e_sat_mbar(T_air)
    o['e_air'] = (RH * o['e_sat_air'])                        # :1743-1788
    log_term = np.log(o['e_air'] / 6.1121)                    # :1790-1835
    o['T_dew'] = 257.14 * log_term / (18.678 - log_term)
    o['T_surf'] = np.where(((h_snow > 0) | (h_ice > 0)), np.minimum(o['T_dew'], f(0)),
                           o['T_dew'])                        # :1837-1853
    T_surf = o['T_surf']
    o['e_sat_surf'] = e_sat_mbar(T_surf)
    top = g * z * (T_air - T_surf)                            # :1448-1475
    bot = (uz) ** 2.0 * (T_air + f(273.15))
    o['Ri'] = (top / bot)
    arg = kappa / np.log((z - h_snow) / z0_air)               # :1477-1625
    Dn = uz * (arg) ** 2.0
    Dn = Dn + np.zeros(np.shape(o['Ri']))
    Ri = o['Ri']
    with np.errstate(all='ignore'):
        Dh = np.where(Ri > 0, Dn / (f(1) + (f(10) * Ri)), Dn * (f(1) - (f(10) * Ri)))
    o['Dn'], o['Dh'] = Dn, Dh
    o['Qh'] = (rho_air * Cp_air) * Dh * (T_air - T_surf)      # :1627-1655
    o['W_p'] = f(1.12) * np.exp(f(0.0614 * o['T_dew']))       # :1855-1873
    o['e_surf'] = (RH * o['e_sat_surf'])
    factor = (rho_air * Lv * Dh)                              # :1875-1936
    o['Qe'] = factor * (o['e_air'] - o['e_surf']) * (f(0.622) / p0)
    K_cs = clear_sky_radiation(lat_deg, julian_day, o['W_p'], TSN_offset, alpha, beta, albedo,
                               dust_atten)
    o['Qn_SW'] = K_cs * (1 - albedo)                          # :2048-2088
    T_air_K = T_air + C_to_K                                  # :2090-2161
    if not satterlund:
        e_air_kPa = o['e_air'] / f(10)
        term1 = (1.0 - canopy_factor) * 1.72 * (e_air_kPa / T_air_K) ** (f(1) / 7)
        term2 = (1.0 + (0.22 * cloud_factor ** 2.0))
        o['em_air'] = (term1 * term2) + canopy_factor
    else:
        eterm = np.exp(-1 * (o['e_air']) ** (T_air_K / 2016))
        o['em_air'] = 1.08 * (1.0 - eterm)
    T_surf_K = T_surf + C_to_K                                # :2163-2224
    LW_in = o['em_air'] * sigma * (T_air_K) ** 4.0
    LW_out = em_surf * sigma * (T_surf_K) ** 4.0
    LW_out = LW_out + ((1.0 - em_surf) * LW_in)
    o['Qn_LW'] = (LW_in - LW_out)
    o['Q_sum'] = o['Qn_SW'] + o['Qn_LW'] + o['Qh'] + o['Qe'] + Qa + Qc   # :2305-2306
    return o
