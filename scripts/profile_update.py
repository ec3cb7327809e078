"""Where does the host time of one BMI update() go?  A small grid (the kernel takes a few
microseconds) stepped through the public API exactly as bench.py's e2e leg does it
(set_value x7 + update() + Q_outlet read), under cProfile.
    python scripts/profile_update.py [n] [steps]"""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from topoflow36_b200 import cases, synth, channels_kinematic_wave as ckw

if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    dev = torch.device('cuda', 0)
    eng, parts = cases.build_engine(n, n, kinematic=True, sources='ga', device=dev, return_parts=True)
    S32 = parts['d8']['S_bed'].float().cpu().numpy()
    work = tempfile.mkdtemp(prefix='tfb_prof_')
    cfg = synth.write_case(work, parts['DEM'], np.float32(S32), parts['width'], parts['angle'],
                           nval=parts['nval'], dt=1.0, kinematic=True, manning=True,
                           n_steps=10 ** 9, outlet=(n - 2, n // 2))
    del eng
    comp = ckw.channels_component()
    comp.device = dev
    comp.enable_fused_sources(infil=True, I0=cases.I0_WET)
    comp.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
    for k, v in cases.GA_TREYNOR.items():
        comp.set_source_input(k, v)
    LONG = {s: ln for ln, s in comp._var_name_map.items()}
    host_in = {s: np.array(0.0) for s in ('MR', 'GW', 'ET', 'IN', 'SM')}
    host_in['P_rain'] = np.array(cases.RAIN_50MMPH)
    host_in['rho_H2O'] = np.array(1000.0)

    def step():
        for s, v in host_in.items():
            comp.set_value(LONG[s], v)
        comp.update()
        return float(comp.Q_outlet)

    for _ in range(50):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    print('us per update (no profiler): %.2f' % ((time.perf_counter() - t0) / steps * 1e6))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(25)
