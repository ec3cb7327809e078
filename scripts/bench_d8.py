"""D8 preprocessing on a synthetic n x n DEM (BASELINE.json configs[2]: 16384^2):
flow codes (+ tie breaking, flat linking), ds/dw, total contributing area with
exact cell counts, slope -- timed per stage with CUDA events, and checked

  * bit for bit against the C restatement of the reference (oracle/tf_oracle.c)
    when --oracle is given (about a minute of host time at 16384^2), and always
  * through size-independent exact properties: every flow cell's count equals
    1 + the counts of the cells that drain into it (recomputed with an integer
    scatter over the parent IDs), the start codes equal an independent torch
    restatement of the fp32 steepest-descent stencil, and area == count *
    pixel_area to fp32 rounding.

    python scripts/bench_d8.py [n] [--oracle] [--json]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from topoflow36_b200.d8 import D8Toolkit


def device_dem(n, seed=1234, sigma=2.0, beta=2.2, sy=0.02, sx=0.005, dx=30.0):
    """Same recipe as synth.fractal_dem, evaluated on the GPU (fp64 FFT)."""
    dev = torch.device('cuda')
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    white = torch.randn((n, n), device=dev, dtype=torch.float32, generator=g)
    F = torch.fft.rfft2(white)
    ky = torch.fft.fftfreq(n, device=dev)[:, None]
    kx = torch.fft.rfftfreq(n, device=dev)[None, :]
    k = torch.sqrt(ky * ky + kx * kx)
    k[0, 0] = 1.0
    amp = k ** (-beta / 2.0)
    amp[0, 0] = 0.0
    F = torch.fft.irfft2(F * amp, s=(n, n))
    del white, k, amp
    F *= sigma / F.std()
    r = torch.arange(n, device=dev, dtype=torch.float64)[:, None]
    c = torch.arange(n, device=dev, dtype=torch.float64)[None, :]
    z0 = sy * dx * n + sx * dx * n + 10.0 * sigma + 100.0
    return (z0 - sy * dx * r - sx * dx * c + F.double()).float().contiguous()


def torch_start_codes(DEM, dx, dy, dd):
    """Independent restatement of d8_global.start_new_d8_codes (:181-353) with
    torch ops (fp32, LINK_FLATS on): ties as code sums, flats negated, pits -300."""
    z = DEM
    # (1, 1) device tensors: a python-scalar divisor would let torch multiply by the
    # reciprocal instead of dividing
    dx, dy, dd = (torch.full((1, 1), float(v), dtype=z.dtype, device=z.device) for v in (dx, dy, dd))
    nb = lambda dr, dc: torch.roll(z, shifts=(-dr, -dc), dims=(0, 1))
    s = [(z - nb(-1, 1)) / dd, (z - nb(0, 1)) / dx, (z - nb(1, 1)) / dd, (z - nb(1, 0)) / dy,
         (z - nb(1, -1)) / dd, (z - nb(0, -1)) / dx, (z - nb(-1, -1)) / dd, (z - nb(-1, 0)) / dy]
    mx = s[0]
    for t in s[1:]:
        mx = torch.maximum(mx, t)
    code = torch.zeros_like(z, dtype=torch.int16)
    for kbit, t in enumerate(s):
        code += (t == mx).to(torch.int16) * (1 << kbit)
    code = torch.where(mx == 0, -code, code)
    code = torch.where(mx < 0, torch.full_like(code, -300), code)
    code[0, :] = 0; code[-1, :] = 0; code[:, 0] = 0; code[:, -1] = 0
    return code


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16384
    use_oracle = '--oracle' in sys.argv
    DEM = device_dem(n)
    tk = D8Toolkit(DEM)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stages, marks = [], [ev()]
    marks[0].record()

    def stage(name, fn):
        fn()
        e = ev(); e.record(); marks.append(e); stages.append(name)

    tk.update(); tk.slope()                 # warm-up (allocations, module load)
    torch.cuda.synchronize()
    marks = [ev()]; marks[0].record(); stages = []
    stage('start_codes', tk.start_codes)
    start = tk.code16.clone()
    stage('(clone for check)', lambda: None)
    stage('break_ties', tk.break_ties)
    stage('link_flats', tk.link_flats_loop)
    stage('finalize', tk.finalize_codes)
    stage('ds_dw', tk.update_ds_dw)
    stage('area+counts', lambda: tk.update_area(with_counts=True))
    S = [None]
    stage('slope', lambda: S.__setitem__(0, tk.slope()))
    torch.cuda.synchronize()
    ms = {nm: marks[i].elapsed_time(marks[i + 1]) for i, nm in enumerate(stages)}
    ms.pop('(clone for check)')
    total = sum(ms.values())
    cells = n * n
    # ---- exact properties at full size -----------------------------------------
    checks = {}
    want = torch_start_codes(DEM, float(tk.dx_host[0]), float(tk.dy_host[0]), float(tk.dd_host[0]))
    checks['start_codes_equal_torch_stencil'] = bool(torch.equal(start, want))
    del want, start
    pid = tk.parent_ids().to(torch.int64).reshape(-1)
    cnt = tk.count.reshape(-1)
    flow = (tk.code.reshape(-1) != 0) & (pid > 0)
    acc = torch.zeros_like(cnt)
    acc.index_add_(0, pid[flow], cnt[flow])
    has_area = tk.A.reshape(-1) > 0
    checks['count_is_1_plus_children'] = bool(torch.equal(cnt[has_area], (acc + 1)[has_area]))
    checks['cells_with_area'] = int(has_area.sum())
    checks['max_count'] = int(cnt.max())
    relA = ((tk.A.reshape(-1).double() - cnt.double() * tk.pixel_area).abs()
            / (cnt.double() * tk.pixel_area).clamp_min(1e-30))[has_area].max()
    checks['area_vs_count_max_rel'] = float(relA)
    checks['link_flats_sweeps'] = int(tk.n_reps)
    del pid, acc, flow
    if use_oracle:
        from oracle import c_oracle, np_oracle as o
        t0 = time.time()
        dxh, dyh, ddh = o.pixel_dims(n, 30.0, 30.0)
        code_o, _ = c_oracle.d8_flow_grid(DEM.cpu().numpy(), dxh, dyh, ddh)
        checks['codes_bit_exact_vs_c_oracle'] = bool(np.array_equal(tk.code.cpu().numpy(), code_o))
        A_o, N_o = c_oracle.d8_area(code_o, np.float64(900.0) / 1e6, with_counts=True)
        checks['area_bit_exact_vs_c_oracle'] = bool(np.array_equal(tk.A.cpu().numpy(), A_o))
        checks['counts_exact_vs_c_oracle'] = bool(np.array_equal(tk.count.cpu().numpy(), N_o))
        checks['oracle_host_s'] = round(time.time() - t0, 1)
    out = {'metric': 'D8 preprocessing', 'n': n, 'cells': cells, 'ms': ms, 'total_ms': total,
           'Gcells_per_s': cells / total / 1e6, 'checks': checks}
    print(json.dumps(out))
    bad = [k for k, v in checks.items() if v is False]
    if bad:
        raise SystemExit('D8 checks failed: %s' % bad)


if __name__ == '__main__':
    main()
