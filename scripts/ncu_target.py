"""Tiny fixed workload for ncu captures: N routing steps of a bench workload at n x n.
    python scripts/ncu_target.py [n] [kw|kw_ga|dw|dw_full|kw_grid|dw_grid] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from topoflow36_b200 import cases
if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    mode = sys.argv[2] if len(sys.argv) > 2 else 'kw_ga'
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    src = {'kw_ga': 'ga', 'kw': 'none', 'dw': 'none', 'dw_full': 'full', 'kw_grid': 'grid',
           'dw_grid': 'grid'}[mode]
    e = cases.build_engine(n, n, kinematic=mode.startswith('kw'), sources=src,
                           device=torch.device('cuda', 0))
    for i in range(steps):
        e.step(i / 60.0)
    torch.cuda.synchronize()
    print('done', e.last_path, e.read_scalars().vol_R)
