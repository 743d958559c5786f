#!/usr/bin/env python
"""bench.py -- full pdf_update + resample + opt_setting cycles per second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], fits one B200): synthetic Lorentzian cloud, 1e8 particles x
3 parameters, 1e5 settings, n_draws = 30, resample forced every cycle (resample_threshold > 1).
A "step" is one full cycle.  Prints ONE JSON line (rank 0).

  value      cycles/s with everything resident in HBM, no host synchronisation inside the timed
             region (run_cycle_async), CUDA events, max over ranks
  e2e        the same cycle through the reference-shaped API (pdf_update(record) -> opt_setting()),
             closed loop: the record goes host->device every step, the stats block and the chosen
             index come back every step
  roofline   dominant kernel (the fused systematic resample) against the measured HBM peak
  cpu_baseline / --impl reference: the numpy restatement of the reference (oracle/), timed on the
             host on a bounded sample and scaled linearly in N to the full workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

TRUE_PARS = (3.14, -1200.0, 50400.0)
SIGMA = 500.0
CONS = (0.1,)


def lorentz(x, p):
    return p[2] + p[1] / (((x - p[0]) / CONS[0]) ** 2 + 1)


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                     r[5:9]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU baseline: the numpy restatement of the reference, bounded sample, scaled to the workload
# -------------------------------------------------------------------------------------------------
def cpu_cycle_rate(n_full, n_settings, n_draws, n_sample=1_000_000, budget_s=20.0, max_cycles=8):
    from oracle import obe_oracle as orc
    rng = np.random.default_rng(1001)
    prior = np.array([rng.uniform(2, 4, n_sample), rng.uniform(-2000, -400, n_sample),
                      rng.normal(50000, 1000, n_sample)])
    settings = (np.linspace(1.5, 4.5, n_settings),)
    eng = orc.OracleOBE(orc.model_lorentzian_hwhm, settings, prior, CONS, n_draws=n_draws, scale=False,
                        default_noise_std=SIGMA, resample_threshold=2.0, rng=np.random.default_rng(1003))
    meas = np.random.default_rng(1002)
    t_n, t_grid, cycles = 0.0, 0.0, 0
    t_start = time.perf_counter()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        while cycles < max_cycles and (time.perf_counter() - t_start) < budget_s:
            t0 = time.perf_counter()
            # design half: the O(N) weighted draw ...
            draws, _ = orc.randdraw(eng.particles, eng.particle_weights, eng.rng.random(n_draws))
            t1 = time.perf_counter()
            # ... and the O(K S) grid evaluation + argmax
            var_p, _ = orc.yvar_from_draws(eng.model, eng.allsettings, draws, CONS, 1)
            util = orc.utility_variance(var_p, orc.noise_var_default(SIGMA, 1))
            best = orc.opt_index(util)
            t2 = time.perf_counter()
            x = (eng.allsettings[0, best],)
            y = float(lorentz(x[0], TRUE_PARS) + SIGMA * meas.standard_normal())
            t3 = time.perf_counter()
            eng.pdf_update((x, y, SIGMA))          # update + forced multinomial resample
            t4 = time.perf_counter()
            t_n += (t1 - t0) + (t4 - t3)
            t_grid += (t2 - t1)
            cycles += 1
    per_cycle_sample = (t_n + t_grid) / cycles
    per_cycle_full = (t_n / cycles) * (n_full / n_sample) + t_grid / cycles
    import threadpoolctl
    blas = sum(i.get('num_threads', 0) for i in threadpoolctl.threadpool_info() if i.get('user_api') == 'blas')
    return {
        'value': 1.0 / per_cycle_full, 'unit': 'cycles/s', 'cores': 1, 'kind': 'port',
        'sample': (f'oracle (numpy restatement of the reference, multinomial resample as the reference does) on '
                   f'{n_sample} particles x {n_settings} settings, {cycles} full cycles, '
                   f'{per_cycle_sample * 1e3:.1f} ms/cycle measured; O(N) part scaled x{n_full / n_sample:g} to '
                   f'{n_full} particles. numpy elementwise/cumsum/searchsorted are single-threaded (1 core; '
                   f'BLAS threads available to cov/matmul: {blas}); host has {os.cpu_count()} logical cores'),
        'ms_per_cycle_sample': per_cycle_sample * 1e3,
    }


# -------------------------------------------------------------------------------------------------
def bench_c5(args, rank, world):
    """BASELINE configs[4]: 4096 independent lock-in engines x 1e4 particles, batched; instances are split over
    the ranks with no collective at all (weak in nothing: the total is fixed -> strong scaling)."""
    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from optbayesexpt_b200.batched import BatchedOptBayesExpt
    from oracle import obe_oracle as orc
    B_total, n = 4096, 10000
    B = B_total // world
    g = torch.Generator(device='cuda')
    g.manual_seed(1001 + rank)
    prior = torch.empty((B, 4, n), dtype=torch.float64, device='cuda')
    for j, sc_ in enumerate((1e-3, 10.0, 1e-5, 10.0)):
        prior[:, j] = torch.empty((B, n), dtype=torch.float64, device='cuda').exponential_(1.0, generator=g) * sc_
    settings = (2 * np.pi * np.logspace(2, 6, 200),)
    eng = BatchedOptBayesExpt('lockin_coil', settings, prior, (), noise_parameter_index=(3, 3),
                              constraint_lt=(0, 1, 2, 3), cost_of_changing_setting=5.0, scale=False, seed=1003)
    del prior
    truth = (1.2e-3, 8.0, 0.9e-5)
    meas = np.random.default_rng(1002 + rank)

    def step(sync):
        out = eng.opt_setting(sync=sync)
        if sync:
            z = orc.model_lockin_coil((out[1][0],), truth, ())
            y = z.T + 5.0 * meas.standard_normal((B, 2))
        else:
            y = step.y
        eng.pdf_update(y, force_resample=args.force_resample)
        step.y = y
    step.y = None
    # device-resident leg: the on-device MeasurementSimulator writes every instance's record (no H2D at all)
    eng.set_simulator(np.tile(np.array(truth), (B, 1)), 5.0, seed=1002 + rank)
    for _ in range(max(args.warmup, 3)):
        step(True)
    for _ in range(3):
        eng.closed_loop_cycle(force_resample=args.force_resample)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(args.steps):
        eng.closed_loop_cycle(force_resample=args.force_resample)   # select -> simulate -> update/resample, no host
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(True)                        # e2e: chosen settings D2H, measurements H2D every cycle
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        tt = torch.tensor([ms, e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(tt[0]), float(tt[1])
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    d = 4
    b_cycle = 8.0 * n * B_total * ((d + 2) + (args.force_resample and (2 * d + 2) + (d + 1) or 0))
    line = {'metric': 'batched pdf_update+resample+opt_setting cycles/sec (4096 engines)', 'value': 1e3 / ms,
            'unit': 'batched cycles/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'batched demos/lockin: 4096 OBE instances x 1e4 particles, d=4, 2 channels, S=200 '
                                   '(BASELINE configs[4])', 'force_resample': bool(args.force_resample)},
            'instance_cycles_per_s': B_total * 1e3 / ms,
            'e2e': {'value': 1.0 / e2e_s, 'unit': 'batched cycles/s', 'h2d_bytes_per_step': B * 12 * 8,
                    'd2h_bytes_per_step': B * 8},
            'roofline': {'bound': 'hbm', 'achieved': b_cycle / world / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': b_cycle / world / (ms * 1e-3) / 1e9 / peak, 'traffic': None, 'peak_source': peak_src,
                         'kernel': 'whole batched cycle (update + resample of flagged instances + select)'},
            'reference_note': 'one reference instance runs 488 cycles/s on one CPU core (BASELINE.md): 4096 instances '
                              '~ 8.4 s per batched cycle'}
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
def bench_sweeper(args):
    """the sweeper's inference half.  One sweep = 64 measured points digested by
    OptBayesExptSweeper.pdf_update on 1e7 particles (d = 4, sigma unknown); fused multi-point kernel against the
    reference's loop of one update per point.  Resample test on (threshold 0.5), systematic resampling."""
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import optbayesexpt_b200 as obe
    from oracle import obe_oracle as orc
    n, m = int(args.particles) if args.particles < 1e8 else 10_000_000, 64
    g = torch.Generator(device='cuda')
    g.manual_seed(1001)
    f64 = dict(dtype=torch.float64, device='cuda')
    xvals = np.linspace(1.5, 4.5, 1000)
    truth, noise = (3.2, 1500.0, 300.0), 300.0

    def make():
        prior = torch.empty((4, n), **f64)
        prior[0] = 2 + 2 * torch.rand(n, generator=g, **f64)
        prior[1] = 400 + 1600 * torch.rand(n, generator=g, **f64)
        prior[2] = 500 + 1000 * torch.randn(n, generator=g, **f64)
        prior[3] = torch.empty(n, **f64).exponential_(1.0, generator=g) * 500
        return obe.OptBayesExptSweeper('lorentzian_hwhm', (xvals,), prior, (0.1,), noise_parameter_index=3,
                                       scale=False, seed=7)
    out = {}
    for mode in ('fused', 'point_by_point'):
        g.manual_seed(1001)                       # same cloud, same sweeps, same noise for both modes
        meas = np.random.default_rng(1002)
        eng = make()
        eng.fused_sweep = (mode == 'fused')
        times, n_res = [], 0
        for it in range(args.warmup + args.steps):
            start = int(meas.integers(0, len(xvals) - m))
            xs = xvals[start:start + m]
            ys = orc.model_lorentzian_hwhm((xs,), truth, (0.1,)) + noise * meas.standard_normal(m)
            e0 = eng._epoch
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.pdf_update(((xs,), ys))
            torch.cuda.synchronize()
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
                n_res += eng._epoch - e0
        out[mode] = (float(np.mean(times)), n_res / max(1, args.steps))
        del eng
        torch.cuda.empty_cache()
    t_f, r_f = out['fused']
    t_p, r_p = out['point_by_point']
    line = {'metric': 'sweeper pdf_update: measured points digested per second', 'value': m / t_f, 'unit': 'points/s',
            'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t_f * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': f'demos/sweeper inference half: {n:.0e} particles, d=4, sweeps of {m} points, '
                                   'resample test after every point (threshold 0.5)'},
            'point_by_point': {'value': m / t_p, 'unit': 'points/s', 'ms_per_sweep': t_p * 1e3,
                               'resamples_per_sweep': r_p},
            'resamples_per_sweep': r_f, 'speedup_vs_point_by_point': t_p / t_f,
            'e2e': {'value': m / t_f, 'unit': 'points/s', 'h2d_bytes_per_step': m * 96, 'd2h_bytes_per_step': 16},
            'note': 'wall clock around pdf_update incl. host synchronisation (the resample decisions are host-side)'}
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--particles', type=float, default=1e8)
    ap.add_argument('--settings', type=int, default=100000)
    ap.add_argument('--draws', type=int, default=30)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--force-resample', action='store_true')
    ap.add_argument('--workload', default='c4', choices=['c4', 'c5', 'sweeper'],
                    help='c4: 1e8-particle Lorentzian cloud (default, the metric); c5: 4096 batched lock-in engines; '
                         'sweeper: the sweeper demo\'s multi-point inference (1 GPU)')
    args = ap.parse_args()
    n_total = int(args.particles)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    config = {'workload': f'synthetic Lorentzian scale-out: {n_total:.0e} particles x {args.settings} settings, '
                          f'n_draws={args.draws}, d=3, resample forced every cycle (BASELINE configs[3])',
              'particles': n_total, 'settings': args.settings, 'n_draws': args.draws, 'n_params': 3,
              'l2': 'inputs (3.2 GB per cycle) exceed the 126 MB L2, no flush needed'}

    if args.workload == 'c5' and args.impl == 'ours':
        return bench_c5(args, rank, world)
    if args.workload == 'sweeper' and args.impl == 'ours':
        return bench_sweeper(args) if rank == 0 else None
    if args.impl == 'reference':
        if rank != 0:
            return
        base = cpu_cycle_rate(n_total, args.settings, args.draws, max_cycles=max(args.steps, 1) if args.steps < 8 else 8)
        line = {'impl': 'reference', 'metric': 'pdf_update+resample+opt_setting cycles/sec', 'value': base['value'],
                'unit': 'cycles/s', 'n_gpus': 0, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': 1e3 / base['value'], 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': config,
                'cpu_baseline': base,
                'e2e': {'value': base['value'], 'unit': 'cycles/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    import optbayesexpt_b200 as obe

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:
        from optbayesexpt_b200.sharded import ShardedOptBayesExpt
    n_local = n_total // world + (1 if rank < n_total % world else 0)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1001 + rank)
    prior = torch.empty((3, n_local), dtype=torch.float64, device='cuda')
    prior[0] = 2 + 2 * torch.rand(n_local, generator=gen, dtype=torch.float64, device='cuda')
    prior[1] = -2000 + 1600 * torch.rand(n_local, generator=gen, dtype=torch.float64, device='cuda')
    prior[2] = 50000 + 1000 * torch.randn(n_local, generator=gen, dtype=torch.float64, device='cuda')
    settings = (np.linspace(1.5, 4.5, args.settings),)
    kw = dict(n_draws=args.draws, scale=False, default_noise_std=SIGMA, seed=1003, resample_threshold=2.0)
    if world > 1:
        eng = ShardedOptBayesExpt('lorentzian_hwhm', settings, prior, CONS, **kw)
    else:
        eng = obe.OptBayesExpt('lorentzian_hwhm', settings, prior, CONS, **kw)
    del prior
    meas = np.random.default_rng(1002)
    xs = settings[0]

    def record_for(x):
        return ((float(x),), float(lorentz(x, TRUE_PARS) + SIGMA * meas.standard_normal()), SIGMA)

    # ---------------- device-resident throughput: no host sync inside the timed region -----------
    fixed = [record_for(xs[(7919 * t + 50000) % len(xs)]) for t in range(args.warmup + args.steps)]
    for t in range(args.warmup):
        eng.run_cycle_async(fixed[t])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e_start.record()
    for t in range(args.steps):
        rec = fixed[args.warmup + t]
        ev[t][0].record()
        eng.run_cycle_async(rec, resample=False, select=False)
        ev[t][1].record()
        eng.resample()
        ev[t][2].record()
        eng._utility_dev_run()
        ev[t][3].record()
    e_stop.record()
    barrier()
    ms_total = e_start.elapsed_time(e_stop)
    t_upd = float(np.mean([ev[t][0].elapsed_time(ev[t][1]) for t in range(args.steps)]))
    t_res = float(np.mean([ev[t][1].elapsed_time(ev[t][2]) for t in range(args.steps)]))
    t_sel = float(np.mean([ev[t][2].elapsed_time(ev[t][3]) for t in range(args.steps)]))
    if world > 1:
        tt = torch.tensor([ms_total, t_upd, t_res, t_sel], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, t_upd, t_res, t_sel = [float(v) for v in tt.cpu()]
    ms_step = ms_total / args.steps

    # ---------------- the cycle without a resample (cycle_noresample) ------------------
    # update + select only; the weight row is explicit here, so the update moves 8N(d+2) bytes.  Single GPU
    # only (the sharded plan kernel would need a no-op resample to keep the ranks in step).
    ms_nores = None
    if world == 1:
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for t in range(3):
            eng.run_cycle_async(fixed[t], resample=False, select=True)
        n0.record()
        for t in range(args.steps):
            eng.run_cycle_async(fixed[args.warmup + t], resample=False, select=True)
        n1.record()
        torch.cuda.synchronize()
        ms_nores = n0.elapsed_time(n1) / args.steps
        eng.resample()                       # back to a healthy cloud for the closed loop below

    # ---------------- end to end through the reference-shaped API, closed loop -------------------
    x = eng.opt_setting()
    for _ in range(max(3, args.warmup // 2)):
        eng.pdf_update(record_for(x[0]))
        x = eng.opt_setting()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for _ in range(e2e_steps):
        eng.pdf_update(record_for(x[0]))     # H2D: the record; D2H: the stats block (N_eff decision)
        x = eng.opt_setting()                # H2D: 30 uniforms; D2H: the chosen index
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    d = 3
    # The resample is forced every cycle, so the update always finds the offspring's implicit uniform weights
    # (never stored, never read): it moves d particle rows in and one weight row out, 8N(d+1), not the 8N(d+2)
    # of an update that follows an update.
    b_upd = 8.0 * n_total * (d + 1)
    b_res = 8.0 * n_total * (2 * d + 2)
    b_sel = 8.0 * args.settings * 2
    b_cycle = b_upd + b_res + b_sel
    gbs_res = b_res / world / (t_res * 1e-3) / 1e9
    line = {
        'metric': 'pdf_update+resample+opt_setting cycles/sec', 'value': 1e3 / ms_step, 'unit': 'cycles/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config,
        'e2e': {'value': e2e_steps / e2e_s, 'unit': 'cycles/s',
                'h2d_bytes_per_step': 8 * (1 + 1 + 1 + 8 + 1) + 8 * args.draws,
                'd2h_bytes_per_step': 8 * 64 + 16,
                'note': 'record, pivot and uniforms travel as kernel arguments; stats block + argmax come back'},
        # update, plan, ancestors, move, draw, utility; sharded: + shard plan (peer exchange fused in) + draw collect
        'gpu_launches': (6 if world == 1 else (8 if getattr(eng, '_peer', None) is not None else 7)) * args.steps,
        'roofline': {'bound': 'hbm', 'kernel': 'resample step = k_sys_plan + k_sys_ancestors + k_sys_move (one event bracket)',
                     'achieved': gbs_res, 'peak': peak, 'unit': 'GB/s', 'frac': gbs_res / peak,
                     # dram__bytes_read+write of the three kernels, ncu --set full (profiles/r1_final_ncu_summary.md):
                     # 63.8 bytes per particle at d = 3
                     'traffic': 63.8 * n_total / world if d == 3 else None,
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': b_res / world},
        'kernels_ms': {'update': t_upd, 'resample': t_res, 'draw+utility+argmax': t_sel},
        'kernels_gbs': {'update': b_upd / world / (t_upd * 1e-3) / 1e9, 'resample': gbs_res},
        'exchange': None if world == 1 else ('peer (CUDA IPC over NVLink)' if getattr(eng, '_peer', None) is not None
                                             else 'nccl'),
        'cycle_hbm_frac': b_cycle / world / (ms_step * 1e-3) / 1e9 / peak,
        'cycle_noresample': None if ms_nores is None else {
            'ms_per_step': ms_nores, 'cycles_per_s': 1e3 / ms_nores,
            'hbm_frac': (8.0 * n_total * (d + 2) + b_sel) / (ms_nores * 1e-3) / 1e9 / peak},
        'clocks': clocks,
    }
    if not args.no_cpu_baseline and world == 1:       # the CPU arm is timed at N = 1 only
        line['cpu_baseline'] = cpu_cycle_rate(n_total, args.settings, args.draws)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
