"""Where does the closed loop (pdf_update -> opt_setting, forced resample) lose time against the device-resident cycle?
Per cycle: host time inside pdf_update, host time inside opt_setting (includes the wait for the GPU), GPU busy time
(events at both ends of the cycle's work on the stream) and the idle gap between consecutive cycles on the device.
python tools/e2e_gap.py [n_particles] [n_settings] [cycles]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import torch
    import optbayesexpt_b200 as obe
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    n_set = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
    cycles = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1001)
    prior = torch.empty((3, n), dtype=torch.float64, device='cuda')
    prior[0] = 2 + 2 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
    prior[1] = -2000 + 1600 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
    prior[2] = 50000 + 1000 * torch.randn(n, generator=gen, dtype=torch.float64, device='cuda')
    settings = (np.linspace(1.5, 4.5, n_set),)
    eng = obe.OptBayesExpt('lorentzian_hwhm', settings, prior, (0.1,), n_draws=30, scale=False, default_noise_std=500.0,
                           seed=1003, resample_threshold=2.0)
    del prior
    eng.eager_select = eng.async_update = True
    meas = np.random.default_rng(1002)

    def rec(x):
        return ((float(x),), float(50400.0 - 1200.0 / (((x - 3.14) / 0.1) ** 2 + 1) + 500.0 * meas.standard_normal()), 500.0)
    import warnings
    warnings.simplefilter('ignore')
    x = eng.opt_setting()
    for _ in range(5):
        eng.pdf_update(rec(x[0]))
        x = eng.opt_setting()
    torch.cuda.synchronize()
    from optbayesexpt_b200 import _lib
    lib = _lib.load()

    def closed_loop(x):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(cycles)]
        t_upd, t_opt, t_rec = [], [], []
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(cycles):
            a = time.perf_counter()
            r = rec(x[0])
            b = time.perf_counter()
            ev[t][0].record()
            eng.pdf_update(r)
            ev[t][1].record()
            c = time.perf_counter()
            x = eng.opt_setting()
            d = time.perf_counter()
            t_rec.append(b - a); t_upd.append(c - b); t_opt.append(d - c)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / cycles
        busy = np.array([e0.elapsed_time(e1) for e0, e1 in ev]) * 1e3
        gap = np.array([ev[t][1].elapsed_time(ev[t + 1][0]) for t in range(cycles - 1)]) * 1e3
        us = lambda v: float(np.median(v)) * 1e6
        return x, (f'wall {wall * 1e6:.1f} us/cycle | host: record {us(t_rec):.1f}, pdf_update {us(t_upd):.1f}, opt_setting '
                   f'(incl. wait) {us(t_opt):.1f} | device: busy {np.median(busy):.1f}, idle between cycles {np.median(gap):.1f}')

    def plain_loop(x):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in range(cycles):
            eng.pdf_update(rec(x[0]))
            x = eng.opt_setting()
        torch.cuda.synchronize()
        return x, (time.perf_counter() - t0) / cycles * 1e6
    print(f'n={n:.3g} settings={n_set} cycles={cycles}')
    x, line = closed_loop(x)
    print('  instrumented:', line)
    for rep in range(3):                 # A/B inside one process: the boxes differ in host speed from call to call
        for side in (1, 0):
            lib.obe_set_option(b'copy_out_side', side)
            x, w = plain_loop(x)
            print(f'  copy_out_side={side}: {w:.1f} us/cycle')
    lib.obe_set_option(b'copy_out_side', 1)
    # the same cycles enqueued back to back, no synchronisation
    recs = [rec(3.0 + 0.01 * t) for t in range(cycles)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in recs:
        eng.run_cycle_async(r)
    e1.record()
    torch.cuda.synchronize()
    print(f'  device-resident: {e0.elapsed_time(e1) / cycles * 1e3:.1f} us/cycle')


if __name__ == '__main__':
    main()
