"""Where does a small-cloud cycle spend its host time?  cProfile of the closed loop pdf_update -> opt_setting on the
c1 shape (1e4 particles x 200 settings).  python tools/profile_small.py [n_cycles] [c1|c2|c3] [sync|fast] [threshold]
(fast: eager_select + async_update, i.e. one C call per cycle with the resample test on the device)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import torch
    import optbayesexpt_b200 as obe
    from baseline import reference_arm as ra
    wl = ra.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else 'c1']
    n = wl['n_particles']
    prior = wl['prior'](np.random.default_rng(1001), n)
    kw = dict(n_draws=30, scale=False, seed=1003)
    if wl['kind'] == 'noise':
        eng = obe.OptBayesExptNoiseParameter(wl['device_model'], wl['settings'](), prior, wl['cons'],
                                             noise_parameter_index=wl['noise_parameter_index'], **kw)
    else:
        eng = obe.OptBayesExpt(wl['device_model'], wl['settings'](), prior, wl['cons'],
                               default_noise_std=wl['default_noise_std'], **kw)
    if len(sys.argv) > 3 and sys.argv[3] == 'fast':
        eng.eager_select = eng.async_update = True
    if len(sys.argv) > 4:
        eng.tuning_parameters['resample_threshold'] = float(sys.argv[4])
    meas = np.random.default_rng(1002)
    cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    import warnings
    warnings.simplefilter('ignore')
    x = eng.opt_setting()
    for _ in range(50):
        eng.pdf_update(ra.simulate(wl, x, meas))
        x = eng.opt_setting()
    recs = [ra.simulate(wl, x, meas) for _ in range(cycles)]

    def loop():
        xx = x
        for r in recs:
            eng.pdf_update((xx,) + tuple(r[1:]))
            xx = eng.opt_setting()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'{cycles / dt:.0f} cycles/s, {dt / cycles * 1e6:.1f} us per cycle (no profiler)')
    if len(sys.argv) > 5:                     # A/B of a library option inside one process (boxes differ in host speed)
        from optbayesexpt_b200 import _lib
        lib = _lib.load()
        for rep in range(3):
            for val in (1, 0):
                if sys.argv[5] in ('split_cycle', 'early_select'):
                    setattr(eng, sys.argv[5], bool(val))
                elif sys.argv[5] == 'one_stream':
                    eng.two_stream_min_particles = 10 ** 12 if val else 0
                else:
                    lib.obe_set_option(sys.argv[5].encode(), val)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                loop()
                torch.cuda.synchronize()
                print(f'  {sys.argv[5]}={val}: {(time.perf_counter() - t0) / cycles * 1e6:.1f} us per cycle')
        return
    pr = cProfile.Profile()
    pr.enable()
    loop()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(28)


if __name__ == '__main__':
    main()
