"""Sweep the run-time tuning options of the systematic resample on one GPU.

    python tools/sweep_resample.py [n_particles ...]

For every cloud size: one engine, one update, then per option set the resample step (plan + one-kernel resample) is
timed with CUDA events over REPS launches (the options are library globals read at launch time).  Prints one JSON
line per (size, option set)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

REPS = 20


def main():
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import optbayesexpt_b200 as obe
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    sizes = [int(float(a)) for a in sys.argv[1:]] or [12_500_000, 100_000_000]
    for n in sizes:
        gen = torch.Generator(device='cuda')
        gen.manual_seed(1001)
        prior = torch.empty((3, n), dtype=torch.float64, device='cuda')
        prior[0] = 2 + 2 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
        prior[1] = -2000 + 1600 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
        prior[2] = 50000 + 1000 * torch.randn(n, generator=gen, dtype=torch.float64, device='cuda')
        eng = obe.OptBayesExpt('lorentzian_hwhm', (np.linspace(1.5, 4.5, 100000),), prior, (0.1,), n_draws=30,
                               scale=False, default_noise_std=500.0, seed=1003, resample_threshold=2.0)
        del prior
        rec = ((3.1,), 49800.0, 500.0)
        for units in (16, 32, 64, 128, 256):
            for cluster_min in (8192, 1024):
                lib.obe_set_option(b'resample_units_per_sm', units)
                lib.obe_set_option(b'plan_cluster_min_tiles', cluster_min)
                for _ in range(3):
                    eng.run_cycle_async(rec, resample=True, select=False)
                ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(REPS)]
                for t in range(REPS):
                    ev[t][0].record()
                    eng.run_cycle_async(rec, resample=False, select=False)
                    ev[t][1].record()
                    eng.resample()
                    ev[t][2].record()
                torch.cuda.synchronize()
                upd = float(np.median([e[0].elapsed_time(e[1]) for e in ev]))
                res = float(np.median([e[1].elapsed_time(e[2]) for e in ev]))
                print(json.dumps({'n': n, 'units_per_sm': units, 'plan_cluster_min_tiles': cluster_min,
                                  'update_ms': round(upd, 4), 'resample_ms': round(res, 4)}), flush=True)
        del eng
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
