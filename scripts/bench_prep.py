"""DEM preparation on a synthetic n x n DEM (the recipe of BASELINE configs[2]): depression
filling, then D8, then the path-length slope grid and a power-law width grid -- timed with
CUDA events and checked

  * bit for bit against the C restatement of the reference's priority flood
    (oracle/tf_oracle.c) when --oracle is given, and always
  * through exact size-independent properties: the filled DEM is >= the raw DEM, equal
    outside depressions, a fixed point of the filling, and has no strict interior pit.

    python scripts/bench_prep.py [n] [--rough] [--oracle]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from scripts.bench_d8 import device_dem
from topoflow36_b200 import prep as P
from topoflow36_b200.d8 import D8Toolkit


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 4096
    rough = '--rough' in sys.argv
    DEM = device_dem(n, sigma=3.0, sy=0.0005, sx=0.0002) if rough else device_dem(n)
    raw = DEM.clone()
    P.fill_pits_tensor(DEM.clone())                    # warm-up (module load, allocator)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    e0.record()
    n_raised, rounds = P.fill_pits_tensor(DEM)
    e1.record()
    torch.cuda.synchronize()
    ms = {'fill_pits': e0.elapsed_time(e1)}
    checks = {'n_raised': n_raised, 'rounds': rounds}
    checks['filled_ge_raw'] = bool((DEM >= raw).all())
    nb = None
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            if dr or dc:
                r = torch.roll(DEM, (dr, dc), (0, 1))
                nb = r if nb is None else torch.minimum(nb, r)
    checks['no_strict_interior_pit'] = int((DEM[1:-1, 1:-1] < nb[1:-1, 1:-1]).sum()) == 0
    del nb, r
    again = DEM.clone()
    checks['idempotent'] = P.fill_pits_tensor(again)[0] == 0 and bool(torch.equal(again, DEM))
    del again
    if '--oracle' in sys.argv:
        from oracle import c_oracle as co
        want = raw.cpu().numpy()
        t0 = time.time()
        n_o = co.fill_pits(want)
        checks['oracle_host_s'] = round(time.time() - t0, 2)
        checks['bit_exact_vs_c_oracle'] = bool(np.array_equal(DEM.cpu().numpy(), want)) and n_o == n_raised
    del raw
    tk = D8Toolkit(DEM)
    tk.update()
    tk.new_slope()                                     # warm-up
    P.grid_from_TCA(tk.A, g1=25.0, g2=1.5, out_dtype=torch.float32)
    ea, eb = ev(), ev()
    ea.record()
    tk.update_area(with_counts=True)                   # basin-wide networks on the filled DEM
    eb.record()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    S, reps = tk.new_slope()
    e1.record()
    W, c, p = P.grid_from_TCA(tk.A, g1=25.0, g2=1.5, out_dtype=torch.float32)
    e2.record()
    torch.cuda.synchronize()
    ms['d8_area_filled_dem'] = ea.elapsed_time(eb)
    checks['max_cells_draining_through_one_cell'] = int(tk.count.max())
    ms['new_slope'] = e0.elapsed_time(e1)
    ms['grid_from_TCA'] = e1.elapsed_time(e2)
    checks['longest_flow_path'] = reps
    checks['cells_with_positive_new_slope'] = int((S > 0).sum())
    checks['flow_cells'] = int((tk.code != 0).sum())
    checks['cells_without_area'] = int(((tk.A == 0) & (tk.code != 0)).sum())   # 0 after filling + flat linking? (cycles)
    out = {'metric': 'DEM preparation', 'n': n, 'cells': n * n, 'rough': rough, 'ms': ms,
           'fill_Gcells_per_s': n * n / ms['fill_pits'] / 1e6, 'checks': checks}
    print(json.dumps(out))
    bad = [k for k, v in checks.items() if v is False]
    if bad:
        raise SystemExit('prep checks failed: %s' % bad)


if __name__ == '__main__':
    main()
