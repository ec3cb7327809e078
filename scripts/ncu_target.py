"""Tiny fixed workload for ncu captures: N routing steps at n x n."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scripts.probe_kernel import build
if __name__ == '__main__':
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    mode = sys.argv[2] if len(sys.argv) > 2 else 'kw'
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    kw = dict(kw=dict(), ga=dict(infil=True), diag=dict(diag=True), dw=dict(kinematic=False))[mode]
    e = build(n, **kw)
    for _ in range(steps):
        e.step()
    torch.cuda.synchronize()
    print('done', e.read_scalars().vol_R)
