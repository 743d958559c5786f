"""GPU test of the sharded engine: two ranks (gloo, both on cuda:0 so it runs on a one-GPU box; the
collectives are staged through the host) against the single-cloud engine on the same inputs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import warnings
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import optbayesexpt_b200 as obe
        from optbayesexpt_b200.sharded import ShardedOptBayesExpt
        from oracle.scenarios import build_inputs, by_name
        sc = by_name('c1_find_peak')
        n = 50000
        inp = build_inputs(sc, n)
        cut = [0, 29000, n]
        lo, hi = cut[rank], cut[rank + 1]
        kw = dict(scale=False, default_noise_std=500.0, seed=77)
        eng = ShardedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'][:, lo:hi], inp['cons'], **kw)
        ref = obe.OptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'], inp['cons'], **kw)
        assert eng.n_total == n
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            for eng_ in (eng, ref):
                eng_.tuning_parameters['auto_resample'] = False
            # design half: same uniforms -> same draws -> same utility -> same argmax
            s1, s2 = eng.opt_setting(), ref.opt_setting()
            assert eng.last_setting_index == ref.last_setting_index and s1 == s2
            # inference half
            rec = (s1, 49600.0, 500.0)
            eng.pdf_update(rec)
            ref.pdf_update(rec)
            np.testing.assert_allclose(eng.mean(), ref.mean(), rtol=1e-12)
            np.testing.assert_allclose(eng.covariance(), ref.covariance(), rtol=1e-9)
            np.testing.assert_allclose(eng.std(), ref.std(), rtol=1e-10)
            np.testing.assert_allclose(eng.n_eff(), ref.n_eff(), rtol=1e-12)
            np.testing.assert_allclose(eng.particle_weights, ref.particle_weights[lo:hi], rtol=1e-12)
            s1, s2 = eng.opt_setting(), ref.opt_setting()
            assert eng.last_setting_index == ref.last_setting_index
            np.testing.assert_allclose(eng.utility(), ref.utility(), rtol=1e-9)   # fresh draws, same stream
            # second update on lazily normalised weights, then a resample on both
            rec = (s1, 49900.0, 500.0)
            eng.pdf_update(rec)
            ref.pdf_update(rec)
            np.testing.assert_allclose(eng.particle_weights, ref.particle_weights[lo:hi], rtol=1e-12)
            eng.rng = np.random.default_rng(5)
            ref.rng = np.random.default_rng(5)
            eng._philox_seed = ref._philox_seed = 4242
            eng._epoch = ref._epoch = 0
            eng.resample()
            ref.resample()
        # shard lengths float; together the shards are the single engine's cloud, in order
        counts = eng._counts
        assert counts.sum() == n and eng.n_particles == counts[rank]
        start = int(counts[:rank].sum())
        got = eng.particles
        want = ref.particles[:, start:start + eng.n_particles]
        spread = want.std(axis=1, keepdims=True)
        err = np.abs(got - want) / (np.abs(want) * 1e-12 + spread * 1e-9)
        assert err.max() <= 1.0, f'resampled shard differs from the single cloud: {err.max():.3g}'
        np.testing.assert_array_equal(eng.particle_weights, np.full(eng.n_particles, 1.0 / n))
        np.testing.assert_allclose(eng.mean(), ref.mean(), rtol=1e-10)
        # and it keeps going
        s1 = eng.opt_setting()
        eng.pdf_update((s1, 50000.0, 500.0))
        assert abs(eng._comm.allreduce_sum(torch.tensor([eng.particle_weights.sum()], dtype=torch.float64)).item()
                   - 1.0) < 1e-12
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def test_two_shards_match_single_cloud(obe_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'
