"""Hydrograph parity over a LONG storm on a grid far larger than Treynor: the packed CUDA path
against the C restatement of the reference path (oracle/tf_oracle.c, pinned to the
reference's own fixtures by tests/test_c_oracle.py), same inputs, from the initial state:
a heavy storm (rain on for the first 60 % of the run, then recession), checkpoints every
`every` steps with the max-norm and the per-cell relative error of vol / d / u / Q and the
outlet series.  Test infrastructure (it executes oracle/): never on the product path.

    python scripts/long_parity.py [n] [steps] [kw_ga|kw|dw] [every] [rain_mm_h] [dt_s] [probe]  ->  one JSON line

`probe`: also step a SECOND copy of the C oracle whose Manning n differs by ONE ULP in every
cell and report how far the two CPU runs drift apart -- the explicit diffusive-wave scheme
amplifies rounding differences where the time step is near its stability limit ("Time step
may be too large for stability", channels_base.py:3275), and then no two implementations
(or compilers, or libm versions) agree to 1e-9.  `generic` as the 8th argument steps the
exact-order generic kernel instead of the packed path.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from oracle import c_oracle, np_oracle as o
from topoflow36_b200 import cases
from topoflow36_b200.d8 import D8Toolkit



def cell_rel(a, b, floor_frac=1e-12):
    """max over cells of |a - b| / (|b| + floor), floor = floor_frac * max|b|."""
    floor = floor_frac * float(np.abs(b).max()) + 1e-300
    return float((np.abs(a - b) / (np.abs(b) + floor)).max())


if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    mode = sys.argv[3] if len(sys.argv) > 3 else 'kw_ga'
    every = int(sys.argv[4]) if len(sys.argv) > 4 else 500
    rain_mmh = float(sys.argv[5]) if len(sys.argv) > 5 else 200.0   # developed flow within a few hundred steps
    DT = float(sys.argv[6]) if len(sys.argv) > 6 else 2.0
    probe = len(sys.argv) > 7 and sys.argv[7] == 'probe'
    generic = len(sys.argv) > 8 and sys.argv[8] == 'generic'
    RAIN = rain_mmh / 3.6e6
    kin = mode.startswith('kw')
    infil = (mode == 'kw_ga')
    dev = torch.device('cuda', 0)
    eng, parts = cases.build_engine(n, n, kinematic=kin, sources='ga' if infil else 'none', dt=DT,
                                    device=dev, return_parts=True, packed=not generic)
    tk = D8Toolkit(parts['DEM'], device=dev)
    tk.update_flow_grid()
    S32 = tk.slope().cpu().numpy()
    del tk
    code = eng.own(eng.code).cpu().numpy()
    dx, dy, dd = o.pixel_dims(n, 30.0, 30.0)
    def oracle(nval=None):
        c = c_oracle.Channels(code, dx, dy, dd, S32, parts['width'], parts['angle'],
                              parts['nval'] if nval is None else nval,
                              dt=DT, kinematic=kin, outlet=(n - 2, n // 2), diag=False)
        c.nthreads = os.cpu_count() or 1
        if infil:
            c.enable_fused(infil=True, I0=cases.I0_WET)
            for k, v in cases.GA_TREYNOR.items():
                c.set_input(k, v)
        return c
    ch = oracle()
    eng.set_input('P_rain', RAIN)
    ch.set_input('P_rain', RAIN)
    ch2 = None
    if probe:
        ch2 = oracle(np.nextafter(np.float64(parts['nval']), 1.0))
        ch2.set_input('P_rain', RAIN)
    t_off = int(0.6 * steps)
    checks, q_gpu, q_cpu = [], [], []
    t0 = time.perf_counter()
    worst = 0.0
    for i in range(steps):
        if i == t_off:
            eng.set_input('P_rain', 0.0)
            ch.set_input('P_rain', 0.0)
            if ch2 is not None:
                ch2.set_input('P_rain', 0.0)
        tm = i * DT / 60.0
        eng.step(tm)
        ch.time_min = tm
        ch.step()
        if ch2 is not None:
            ch2.time_min = tm
            ch2.step()
        if (i + 1) % 25 == 0:
            q_gpu.append(eng.read_scalars().Q_outlet)
            q_cpu.append(float(ch.Q_outlet))
        if (i + 1) % every == 0 or i + 1 == steps:
            row = {'step': i + 1}
            for nm in ('vol', 'd', 'u', 'Q'):
                got = (eng.Q if nm == 'Q' else eng.vol if nm == 'vol' else eng.own(getattr(eng, nm))).cpu().numpy()
                want = getattr(ch, nm)
                sc = float(np.abs(want).max())
                row[nm] = {'max_norm': float(np.abs(got - want).max() / sc) if sc > 0 else 0.0,
                           'per_cell': cell_rel(got, want), 'max': sc}
                worst = max(worst, row[nm]['max_norm'], row[nm]['per_cell'])
                if ch2 is not None:
                    row[nm]['cpu_vs_cpu_one_ulp_n'] = cell_rel(getattr(ch2, nm), want)
            s = eng.read_scalars()
            for nm in ('vol_R', 'vol_edge', 'vol_Q', 'Q_peak') + (('vol_IN',) if infil else ()):
                want = float(getattr(ch, nm))
                row[nm] = abs(getattr(s, nm) - want) / max(abs(want), 1e-300)
                worst = max(worst, row[nm])
            checks.append(row)
    q_gpu, q_cpu = np.array(q_gpu), np.array(q_cpu)
    out = {'what': 'long-storm parity, packed CUDA path vs oracle/tf_oracle.c', 'grid': [n, n], 'mode': mode,
           'steps': steps, 'dt_s': DT, 'rain_mm_per_h': rain_mmh, 'rain_off_at_step': t_off,
           'path': eng.last_path, 'kernel': eng.last_kernel,
           'outlet_hydrograph': {'samples': int(q_cpu.size), 'Q_max': float(q_cpu.max()),
                                 'max_abs_diff_over_Q_max': float(np.abs(q_gpu - q_cpu).max() / max(q_cpu.max(), 1e-300))},
           'checkpoints': checks, 'worst_rel_err': worst, 'tolerance_north_star': 1e-9,
           'ok': bool(worst <= 1e-9), 'wall_s': round(time.perf_counter() - t0, 1)}
    print(json.dumps(out))
