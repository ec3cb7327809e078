#!/usr/bin/env python
"""Routing throughput benchmark (BASELINE.json metric: routing Gcell-updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one `update()` of the channel-routing component over the whole grid
(one cell-update per cell).  Workload (BASELINE.json configs[1]): synthetic
fractal DEM, 4096 x 4096 cells per GPU, kinematic wave + Manning, fused
Green-Ampt infiltration, uniform 50 mm/h rain, fp64, dt = 1 s.  With N > 1 the
grid is (N*4096) x 4096, one 4096-row slab per GPU (weak scaling), boundary
discharge rows exchanged over NCCL every step.

Prints ONE JSON line (rank 0).  `value` is timed with the state resident in
HBM; `e2e` drives the BMI component (`channels_kinematic_wave.channels_component`
read from CFG/RTG files it wrote itself) with host inputs through `set_value()`
and a device->host read of the scalar block (Q_outlet, ...) every update.
`--impl reference` times the CPU restatement of the reference's numpy path
(oracle/tf_oracle.c, all host threads) -- the reference itself is pure Python
and absent on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NY_PER_GPU = 4096
NX = 4096
BYTES_PER_CELL = {'kw': 81, 'kw_ga': 97, 'dw': 106, 'dw_full': 146}
METRIC = 'routing Gcell-updates/sec'
WORKLOAD = ('synthetic fractal DEM %dx%d per GPU, kinematic wave + Manning, fused Green-Ampt '
            'infiltration (pre-wetted soil, I0=0.1 m), uniform 50 mm/h rain, fp64, dt=1s'
            % (NY_PER_GPU, NX))


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi samples of SM clock and throttle reasons DURING a timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index=0, period_ms=50):
        self.index, self.period_ms = index, period_ms
        self.rows, self.proc, self.th = [], None, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                 '-i', str(self.index), '-lms', str(self.period_ms)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(',')]
            if len(p) >= 7:
                self.rows.append(p)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(self.period_ms / 1000.0)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        for p in self.rows:
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, p[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)),
                'power_w_max': float(max(pw)), 'samples': len(sm), 'reasons': sorted(reasons)}


# ------------------------------------------------------------- CPU baseline
def cpu_port_rate(budget_s, n_steps=None, warmup=1, threads=None):
    """Time the C restatement of the reference path (oracle/tf_oracle.c) on the
    host cores over a bounded sample of the workload: the first `ny_s` rows of the
    same synthetic generator (own D8 on the CPU), KW + fused Green-Ampt.  Returns
    (Gcell/s, description, cores, ms_per_step, steps)."""
    from oracle import c_oracle, np_oracle as o
    from topoflow36_b200 import cases
    threads = threads or (os.cpu_count() or 1)

    def build(ny_s):
        tile = cases.noise_tile(ny_s, NX, seed=1234)
        DEM = cases.slab_dem(ny_s, NX, 0, ny_s, 0, tile)
        dx, dy, dd = o.pixel_dims(ny_s, 30.0, 30.0)
        code, _ = c_oracle.d8_flow_grid(DEM, dx, dy, dd)
        ds32, _ = o.d8_ds_dw(code, dx, dy, dd)
        with np.errstate(all='ignore'):
            slope = o.d8_slope(DEM, code, ds32)
        width, angle, nval = cases.slab_params(ny_s, NX, 0, 0, seed=1234)
        ch = c_oracle.Channels(code, dx, dy, dd, slope, width, angle, nval, dt=1.0,
                               outlet=(ny_s - 2, NX // 2), diag=False)
        ch.nthreads = threads
        ch.set_input('P_rain', cases.RAIN_50MMPH)
        ch.enable_fused(infil=True, I0=cases.I0_WET)
        for k, v in cases.GA_TREYNOR.items():
            ch.set_input(k, v)
        return ch

    # calibrate on a thin strip, then size the sample to the budget
    ch = build(128)
    ch.step()
    t0 = time.perf_counter()
    for _ in range(3):
        ch.step()
    rate = 3 * 128 * NX / (time.perf_counter() - t0)        # cells/s
    if n_steps is None:
        n_steps = 10
    cells_per_step = rate * budget_s / float(n_steps + warmup)
    ny_s = int(min(NY_PER_GPU, max(64, cells_per_step // NX)))
    ch = build(ny_s)
    for _ in range(warmup):
        ch.step()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        ch.step()
    dt = time.perf_counter() - t0
    val = n_steps * ny_s * NX / dt / 1e9
    desc = ('first %d of %d rows x %d cols of the same synthetic case (own D8), %d steps, '
            'C restatement of the numpy path, %d pthreads' % (ny_s, NY_PER_GPU, NX, n_steps, threads))
    return val, desc, threads, dt / n_steps * 1e3, n_steps


def cpu_numpy_rate(n=1024, n_steps=6):
    """The reference's OWN CPU path is whole-grid numpy (one core).  It cannot travel to
    the GPU box, so its numpy restatement (oracle/np_oracle.py, the same expressions in
    the same order, bit-exact against it) is timed instead on an n x n sub-case: KW +
    Green-Ampt, like the benchmark.  Returns (Gcell/s, description)."""
    from oracle import c_oracle, np_oracle as o
    from topoflow36_b200 import cases
    tile = cases.noise_tile(n, n, seed=1234)
    DEM = cases.slab_dem(n, n, 0, n, 0, tile)
    dx, dy, dd = o.pixel_dims(n, 30.0, 30.0)
    code, _ = c_oracle.d8_flow_grid(DEM, dx, dy, dd)
    ds32, dw32 = o.d8_ds_dw(code, dx, dy, dd)
    with np.errstate(all='ignore'):
        slope = o.d8_slope(DEM, code, ds32)
    width, angle, nval = cases.slab_params(n, n, 0, 0, seed=1234)
    ch = o.Channels(code, ds32, dw32, slope, width, angle, nval=nval, dt=1.0,
                    outlet=(n - 2, n // 2), check_stability=True)
    ch.P_rain = np.array(cases.RAIN_50MMPH)
    I = np.zeros((n, n)) + cases.I0_WET

    def step():
        nonlocal I
        ch.IN, I = o.green_ampt_step(I, ch.P_rain, 0.0, dt=1.0, **cases.GA_TREYNOR)
        ch.step()
    with np.errstate(all='ignore'):
        step()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            step()
        dt = time.perf_counter() - t0
    return n_steps * n * n / dt / 1e9, ('%d x %d sub-case, %d steps, numpy restatement of the '
                                        'reference path (np_oracle), 1 core' % (n, n, n_steps))


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # keep the whole run within ~2 minutes whatever K the driver asks for
    val, desc, cores, ms, steps = cpu_port_rate(100.0, n_steps=args.steps, warmup=args.warmup)
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'Gcell-updates/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': desc},
        'cpu_baseline': {'value': val, 'unit': 'Gcell-updates/s', 'cores': cores, 'kind': 'port',
                         'sample': desc},
        'e2e': {'value': val, 'unit': 'Gcell-updates/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': ('the reference is pure Python/numpy (5-8 Mcell-updates/s on one core, SURVEY.md '
                 'section 6) and cannot travel to the GPU box; this arm times oracle/tf_oracle.c, '
                 'the same arithmetic in C on all host threads -- a far stronger CPU baseline'),
    }
    print(json.dumps(out))


# -------------------------------------------------------------------- main
def write_component_case(work, eng_parts, dt=1.0):
    """Lay the synthetic case out in the reference's file formats so the BMI
    component can be initialised exactly as EMELI would (CFG + RTI + RTG)."""
    from topoflow36_b200 import synth
    p = eng_parts
    return synth.write_case(work, p['DEM'], p['slope32'], p['width'], p['angle'], nval=p['nval'],
                            dt=dt, kinematic=True, manning=True, n_steps=10 ** 9,
                            check_stability='No')


def main():
    global NX, WORKLOAD
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--rows-per-gpu', type=int, default=NY_PER_GPU)
    ap.add_argument('--nx', type=int, default=NX)
    ap.add_argument('--workload', default='kw_ga', choices=['kw_ga', 'kw', 'dw', 'dw_full'],
                    help='kw_ga = BASELINE configs[1] (default); kw = configs[4]; dw_full = '
                         'configs[3] (diffusive wave + met/snow/evap/infil fused); dw = diffusive only')
    ap.add_argument('--device-gen', action='store_true',
                    help='synthesise the case on the GPU (needed for 65536-wide slabs)')
    ap.add_argument('--halo', default='p2p', choices=['p2p', 'nccl', 'p2p-dw'],
                    help='N > 1: boundary rows by NVLink peer stores fused into the kernel '
                         '(kinematic wave, default) or by NCCL send/recv after it')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from topoflow36_b200 import cases, lib as L
    L.load()
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (there is no CPU fallback); '
                         'use --impl reference for the CPU arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    ny = args.rows_per_gpu
    NX = args.nx
    infil = (args.workload == 'kw_ga')
    kin = args.workload.startswith('kw')
    sources = {'kw_ga': 'ga', 'kw': 'none', 'dw': 'none', 'dw_full': 'full'}[args.workload]
    WORKLOAD = ('synthetic %s DEM %dx%d per GPU, %s wave + Manning, %suniform 50 mm/h rain, '
                'fp64, dt=1s' % ('closed-form (device-generated)' if args.device_gen else 'fractal',
                                 ny, NX, 'kinematic' if kin else 'diffusive',
                                 {'ga': 'fused Green-Ampt infiltration (pre-wetted soil, I0=0.1 m), ',
                                  'none': '', 'full': 'fused rain/snow split + degree-day melt + '
                                  'Priestley-Taylor ET + Green-Ampt (T_air grid), '}[sources]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the diffusive passes also run with the peer halo (--halo p2p-dw), but measured slower than
    # NCCL on 2 GPUs (0.367 vs 0.358 ms/step): NCCL stays their default
    use_p2p = (world > 1 and (args.halo == 'p2p-dw' or (args.halo == 'p2p' and kin)))
    if args.device_gen:
        stepper = cases.build_engine_device(ny, NX, kinematic=kin, sources=sources, rank=rank,
                                            world=world, device=dev, p2p=use_p2p)
        parts = None
        args.no_e2e = args.no_e2e or world == 1
    else:
        stepper, parts = cases.build_engine(ny, NX, kinematic=kin, sources=sources, rank=rank,
                                            world=world, device=dev, return_parts=True,
                                            p2p=use_p2p)
        args.no_e2e = args.no_e2e or (world == 1 and args.workload != 'kw_ga' and args.workload != 'kw')
    eng = stepper.engine if world > 1 else stepper
    cells = ny * NX * world
    launches_per_step = 1 if kin else 3       # DW: depth pass, velocity pass, noflow zeroing

    # ---- `value`: state resident in HBM, no host round trip per step ----------
    tmin = 0.0
    for _ in range(args.warmup):
        stepper.step(tmin); tmin += 1.0 / 60.0
    sampler = ClockSampler(index=local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        stepper.step(tmin); tmin += 1.0 / 60.0
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    by_rank = [float(ms.item()) / args.steps]
    if world > 1:
        every = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(every, ms)
        by_rank = [float(v) / args.steps for v in every.tolist()]
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if sampler else None
    ms_per_step = float(ms.item()) / args.steps
    value = cells / (ms_per_step * 1e-3) / 1e9
    sc = stepper.read_scalars()
    path_used = eng.last_path
    stored = eng.vol.sum().reshape(1)
    qmax = eng.Q.max().reshape(1)
    if world > 1:
        dist.all_reduce(stored)
        dist.all_reduce(qmax, op=dist.ReduceOp.MAX)
    sanity = {'Q_outlet': sc.Q_outlet, 'Q_max': float(qmax), 'vol_R': sc.vol_R, 'vol_IN': sc.vol_IN,
              'mass_balance_rel_err': abs(sc.vol_R - (float(stored) + sc.vol_edge)) / max(sc.vol_R, 1e-300),
              'bad_values': int(sc.n_neg_d + sc.n_inf_d + sc.n_neg_u + sc.n_nan_u + sc.n_inf_u)}

    # ---- `e2e`: through the public API with host inputs and a host read -------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            work = tempfile.mkdtemp(prefix='tfb_bench_')
            from topoflow36_b200 import channels_kinematic_wave as ckw
            from topoflow36_b200.d8 import D8Toolkit
            S32 = parts['d8'].get('slope32')
            if S32 is None:
                tk = D8Toolkit(parts['DEM'], device=dev)
                tk.update_flow_grid()
                S32 = tk.slope().cpu().numpy()
                del tk
            parts['slope32'] = S32
            del stepper, eng
            torch.cuda.empty_cache()
            cfg = write_component_case(work, parts)
            comp = ckw.channels_component()
            comp.device = dev
            if infil:
                comp.enable_fused_sources(infil=True, I0=cases.I0_WET)
            comp.initialize(cfg_file=cfg, mode='nondriver', SILENT=True)
            if infil:
                for k, v in cases.GA_TREYNOR.items():
                    comp.set_source_input(k, v)
            LONG = {n: ln for ln, n in comp._var_name_map.items()}
            host_in = {n: np.array(0.0) for n in ('MR', 'GW', 'ET', 'IN', 'SM')}
            host_in['P_rain'] = np.array(cases.RAIN_50MMPH)
            host_in['rho_H2O'] = np.array(1000.0)

            def e2e_step():
                for n, v in host_in.items():                 # what EMELI does each step
                    comp.set_value(LONG[n], v)
                comp.update()                                # 1 launch + D2H of the scalar block
                return float(comp.Q_outlet)
            h2d = 8 * len(host_in)
            import ctypes
            d2h = ctypes.sizeof(L.ChannelsScalars)
            import shutil
        else:
            def e2e_step():
                eng.set_input('P_rain', float(np.array(cases.RAIN_50MMPH)))
                stepper.step(0.0)
                return stepper.read_scalars().Q_outlet       # all-reduce + D2H
            h2d = 8
            import ctypes
            d2h = ctypes.sizeof(L.ChannelsScalars)
        n_e2e = max(10, args.steps // 4)
        for _ in range(args.warmup):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            q = e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {'value': cells * n_e2e / float(dt.item()) / 1e9, 'unit': 'Gcell-updates/s',
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'steps': n_e2e,
               'ms_per_step': float(dt.item()) / n_e2e * 1e3,
               'api': ('channels_kinematic_wave.channels_component: set_value() x7 + update() + '
                       'Q_outlet read per step' if world == 1 else
                       'SlabChannels.step() + collective read_scalars() per step'),
               'init_h2d_bytes': int(ny * NX * (4 + 8 * 6)) if world == 1 else None}
        if world == 1:
            comp.finalize()
            shutil.rmtree(work, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the one kernel of the step --------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    bpc = BYTES_PER_CELL[args.workload]
    cells_per_launch = ny * NX
    achieved = cells_per_launch * bpc / (ms_per_step * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, 'profiles', 'kernel_traffic.json')))
        traffic = (prof.get(args.workload + '_4096', {}).get('dram_bytes_per_launch')
                   if (ny, NX) == (4096, 4096) else None)
    except (OSError, ValueError):
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (measured)' if peaks
                else 'B200_PROFILING.md fallback 6.65 TB/s',
                'kernel': ('channels_packed_kernel<SRC=%s, %s>%s (TMA ring)'
                           % (sources, 'KW' if kin else 'DW pass A', '' if kin else
                              ' + channels_packed_dwb_kernel (both launches of the step)')),
                'path': path_used,
                'algorithmic_bytes_per_cell': bpc, 'cells_per_launch': cells_per_launch,
                'launch_ms': ms_per_step,
                'note': '%d launch(es) per step; time = timed region / steps (CUDA events)' % launches_per_step}

    cpu = None
    if world == 1 and not args.no_cpu_baseline and infil and NX == 4096:
        v, desc, cores, _, _ = cpu_port_rate(15.0)
        cpu = {'value': v, 'unit': 'Gcell-updates/s', 'cores': cores, 'kind': 'port', 'sample': desc}
        nv, ndesc = cpu_numpy_rate()
        cpu['numpy_path'] = {'value': nv, 'unit': 'Gcell-updates/s', 'cores': 1, 'sample': ndesc}

    out = {
        'metric': METRIC, 'value': value, 'unit': 'Gcell-updates/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
        'ms_per_step_by_rank': by_rank,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'grid': [ny * world, NX], 'rows_per_gpu': ny,
                   'partition': ('row slabs, 1 halo row of Q per neighbour per step, %s'
                                 % ('NVLink peer stores fused into the routing kernel(s), flag words '
                                    'polled by the boundary tiles (no NCCL call per step)' if use_p2p
                                    else 'NCCL send/recv after the kernel'))
                   if world > 1 else 'single GPU',
                   'l2': 'working set %.2f GB per step per GPU >> 126 MB L2, no flush needed'
                         % (ny * NX * bpc / 1e9)},
        'e2e': e2e, 'gpu_launches': launches_per_step * args.steps,
        'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks, 'sanity': sanity,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
