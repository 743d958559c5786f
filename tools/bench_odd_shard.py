"""Cost of a shard whose first global slot is odd/even in the one-kernel systematic resample (one GPU).
    python tools/bench_odd_shard.py [n]
Times obe_resample_systematic_sharded for slot_begin in (0, 1, 2, 3)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main():
    import torch
    import __graft_entry__
    __graft_entry__.build()
    import optbayesexpt_b200 as obe
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 12_500_000
    gen = torch.Generator(device='cuda')
    gen.manual_seed(3)
    prior = torch.randn((3, n), generator=gen, dtype=torch.float64, device='cuda')
    pdf = obe.ParticlePDF(prior, scale=False, resampling='systematic', seed=5)
    pdf._ensure_moments()
    total = float(pdf._fetch_stats()[_lib.ST_TOTAL])
    factor = np.eye(3) * 0.01
    mean = np.zeros(3)
    alt = pdf._buf.empty_like()
    for begin in (0, 1, 2, 3):
        for shift in (1,):
            n_total = n + begin
            cdf_total = total / (1.0 - begin / n_total)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(22)]
            for t in range(21):
                ev[t].record()
                _lib.check(lib.obe_resample_systematic_sharded(
                    pdf._cs(), C.byref(alt.struct()), 0.4142, n_total, begin, n_total, cdf_total - total, cdf_total, 1,
                    _lib.darr(factor.reshape(-1)), _lib.darr(mean), 99, t, 0.98, 0, None, None, pdf._stream()))
            ev[21].record()
            torch.cuda.synchronize()
            ms = float(np.median([ev[t].elapsed_time(ev[t + 1]) for t in range(1, 21)]))
            print(json.dumps({'n': n, 'slot_begin': begin, 'resample_ms': round(ms, 4)}), flush=True)


if __name__ == '__main__':
    main()
