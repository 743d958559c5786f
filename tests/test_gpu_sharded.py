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


def _init_group(rank, world):
    """OBE_TEST_BACKEND=nccl: one GPU per rank, NCCL collectives and real NVLink peer buffers (boxes with >= 2 GPUs);
    default: both ranks on cuda:0 with gloo staging the collectives through the host."""
    import torch
    import torch.distributed as dist
    if os.environ.get('OBE_TEST_BACKEND') == 'nccl':
        torch.cuda.set_device(rank)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    else:
        torch.cuda.set_device(0)
        dist.init_process_group('gloo', rank=rank, world_size=world)


def _worker(rank, world, port, out):
    import warnings
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    _init_group(rank, world)
    try:
        import optbayesexpt_b200 as obe
        from optbayesexpt_b200 import sharded as _sh
        from optbayesexpt_b200.sharded import ShardedOptBayesExpt
        from oracle.scenarios import build_inputs, by_name
        _sh.REPLICATE_GRID_MAX = 0      # this test slices the setting grid over the ranks (the other replicates it)
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
            def same_stream(seed):
                eng.rng = np.random.default_rng(seed)
                ref.rng = np.random.default_rng(seed)
            # design half: same uniforms -> same draws -> same utility -> same argmax
            same_stream(1)
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
            # the device-side shard plan against its host restatement
            from optbayesexpt_b200 import sharded as sh, _lib
            gs_dev = eng._fetch_plan()
            # (peer-exchange mode keeps no gathered copy: gather the local blocks again; combine_stats does not
            # read the normaliser word the plan has overwritten since)
            gathered = eng._keep if eng._keep is not None else eng._comm.allgather(eng._buf.stats)
            gs_host = sh.combine_stats(gathered.cpu().numpy(), eng.n_dims)
            for key in ('totals', 'offsets', 'm1', 'm2', 'pivot'):
                np.testing.assert_array_equal(gs_dev[key], gs_host[key], err_msg=key)
            assert gs_dev['total'] == gs_host['total'] and gs_dev['sumsq'] == gs_host['sumsq']
            bounds = sh.shard_slot_bounds(gs_host['offsets'], gs_host['total'], eng._u0, n, _lib.load().obe_comb_count)
            np.testing.assert_array_equal(eng._plan_host[_lib.PLAN_COUNTS:_lib.PLAN_COUNTS + world], np.diff(bounds))
            mean_h, cov_h, _, _ = sh.moments_from(gs_host, eng.n_dims)
            f_host = np.linalg.cholesky((1 - 0.98 ** 2) * cov_h).T
            np.testing.assert_allclose(eng._plan_host[16:16 + 9].reshape(3, 3), f_host, rtol=1e-12, atol=1e-300)
            same_stream(2)
            s1, s2 = eng.opt_setting(), ref.opt_setting()
            assert eng.last_setting_index == ref.last_setting_index
            same_stream(3)
            np.testing.assert_array_equal(eng.utility(), ref.utility())   # same draws -> bit-identical utility
            # second update on lazily normalised weights, then a resample on both
            rec = (s1, 49900.0, 500.0)
            eng.pdf_update(rec)
            ref.pdf_update(rec)
            np.testing.assert_allclose(eng.particle_weights, ref.particle_weights[lo:hi], rtol=1e-12)
            eng.rng = np.random.default_rng(5)
            ref.rng = np.random.default_rng(5)
            eng._philox_seed = ref._philox_seed = 4242
            eng._epoch = ref._epoch = 0
            ref.rng = np.random.default_rng(5)
            ref.rng.random()                    # the comb offset the sharded plan drew at the update
            ref.rng = type('R', (), {'random': staticmethod(lambda *a: eng._u0)})()
            eng.resample()
            ref.resample()
        # shard lengths float; together the shards are the single engine's cloud, in order
        counts = eng.shard_counts
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
        assert abs(eng._comm.allreduce_sum(torch.tensor([eng.particle_weights.sum()], dtype=torch.float64,
                                                        device=eng._buf.device)).item()
                   - 1.0) < 1e-12
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def _worker_noise(rank, world, port, out):
    """noise-parameter engine (config c2): sigma is a particle coordinate, positivity constraint after a resample"""
    import warnings
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    _init_group(rank, world)
    try:
        import optbayesexpt_b200 as obe
        from optbayesexpt_b200.sharded import ShardedOptBayesExptNoiseParameter
        from oracle.scenarios import build_inputs, by_name
        sc = by_name('c2_line_noise')
        n = 40000
        inp = build_inputs(sc, n)
        prior = inp['prior'].copy()
        cut = [0, 17000, n]
        lo, hi = cut[rank], cut[rank + 1]
        # a small a_param makes the Liu-West nudge large enough to push some noise parameters below zero
        kw = dict(scale=False, seed=5, noise_parameter_index=2, a_param=0.5)
        eng = ShardedOptBayesExptNoiseParameter('line', inp['setting_values'], prior[:, lo:hi], (), **kw)
        ref = obe.OptBayesExptNoiseParameter('line', inp['setting_values'], prior, (), **kw)
        for e in (eng, ref):
            e.tuning_parameters['auto_resample'] = False
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            np.testing.assert_allclose(eng.yvar_noise_model(), ref.yvar_noise_model(), rtol=1e-12)
            for seed, y in ((1, 0.1), (2, 0.25)):
                eng.rng = np.random.default_rng(seed)
                ref.rng = np.random.default_rng(seed)
                s1, s2 = eng.opt_setting(), ref.opt_setting()
                assert eng.last_setting_index == ref.last_setting_index
                eng.rng = np.random.default_rng(seed)
                ref.rng = np.random.default_rng(seed)
                np.testing.assert_allclose(eng.utility(), ref.utility(), rtol=1e-11)
                eng.pdf_update((s1, y))
                ref.pdf_update((s1, y))
                np.testing.assert_allclose(eng.particle_weights, ref.particle_weights[lo:hi], rtol=1e-12, atol=1e-300)
                np.testing.assert_allclose(eng.n_eff(), ref.n_eff(), rtol=1e-12)
                np.testing.assert_allclose(eng.yvar_noise_model(), ref.yvar_noise_model(), rtol=1e-12)
            # resample on both with the same comb offset and normals, then the positivity constraint
            eng._philox_seed = ref._philox_seed = 99
            eng._epoch = ref._epoch = 0
            u0 = eng._u0                                     # the comb offset of the current plan
            ref.rng = type('R', (), {'random': staticmethod(lambda *a: u0)})()
            eng.resample()
            eng.enforce_parameter_constraints()
            ref.resample()
            ref.enforce_parameter_constraints()
        counts = eng.shard_counts
        start = int(counts[:rank].sum())
        want_w = ref.particle_weights[start:start + eng.n_particles]
        want_p = ref.particles[:, start:start + eng.n_particles]
        spread = want_p.std(axis=1, keepdims=True)
        perr = np.abs(eng.particles - want_p) / (np.abs(want_p) * 1e-12 + spread * 1e-9)
        assert perr.max() <= 1.0, f'resampled shard differs from the single cloud: {perr.max():.3g} counts {counts}'
        np.testing.assert_allclose(eng.particle_weights, want_w, rtol=1e-12, atol=0)
        assert (eng.particle_weights == 0).sum() == (want_w == 0).sum()
        zeros = eng._comm.allreduce_sum(torch.tensor([float((eng.particle_weights == 0).sum())], dtype=torch.float64,
                                                     device=eng._buf.device))
        assert zeros.item() > 0, 'the constraint never bit: the test does not exercise it'
        np.testing.assert_allclose(eng.yvar_noise_model(), ref.yvar_noise_model(), rtol=1e-11)
        np.testing.assert_allclose(eng.mean(), ref.mean(), rtol=1e-10)
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def _worker_early(rank, world, port, out):
    """early select over two shards: run_cycle_async takes the K draws from the resample plan (every rank the slots it
    owns, exchanged by peer writes or an all-reduce) and overlaps the utility pass with the resample"""
    import warnings
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    _init_group(rank, world)
    try:
        import optbayesexpt_b200 as obe
        from optbayesexpt_b200.sharded import ShardedOptBayesExpt
        from oracle.scenarios import build_inputs, by_name
        sc = by_name('c1_find_peak')
        n = 60000
        inp = build_inputs(sc, n)
        cut = [0, 26000, n]
        lo, hi = cut[rank], cut[rank + 1]
        kw = dict(scale=False, default_noise_std=500.0, seed=77)
        eng = ShardedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'][:, lo:hi], inp['cons'], **kw)
        ref = obe.OptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'], inp['cons'], **kw)
        assert eng._early_select_ok() and ref._early_select_ok()
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            for cycle, y in enumerate((49600.0, 49900.0, 50100.0)):
                rec = ((3.0 + 0.1 * cycle,), y, 500.0)
                eng.rng = np.random.default_rng(40 + cycle)
                ref.rng = np.random.default_rng(40 + cycle)
                twin = np.random.default_rng(40 + cycle)
                eng._philox_seed = ref._philox_seed = 4242 + cycle
                eng._epoch = ref._epoch = cycle
                eng.run_cycle_async(rec)
                ref.run_cycle_async(rec)
                twin.random()
                u = twin.random(eng.N_DRAWS)
                slots = np.minimum((u * n).astype(np.int64), n - 1)
                counts = eng.shard_counts
                assert counts.sum() == n
                start = int(counts[:rank].sum())
                draws = eng._draws_dev.cpu().numpy()
                mine = (slots >= start) & (slots < start + counts[rank])
                # the draws this rank owns are ITS offspring, bit for bit; all ranks hold all draws
                np.testing.assert_array_equal(draws[:, mine], eng.particles[:, slots[mine] - start])
                want = ref._draws_dev.cpu().numpy()
                spread = ref.particles.std(axis=1, keepdims=True)
                err = np.abs(draws - want) / (np.abs(want) * 1e-12 + spread * 1e-9)
                assert err.max() <= 1.0, f'sharded draws differ from the single cloud: {err.max():.3g}'
                assert int(eng.best_index_dev.cpu()[0]) == int(ref.best_index_dev.cpu()[0])
                got, wantp = eng.particles, ref.particles[:, start:start + eng.n_particles]
                perr = np.abs(got - wantp) / (np.abs(wantp) * 1e-12 + spread * 1e-9)
                assert perr.max() <= 1.0
        eng._fetch_plan()
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def _worker_rebalance(rank, world, port, out):
    """rebalance(): one all-to-all-v per row moves the cut points back to n_total*g/G; the global cloud (order, values,
    weights) is unchanged and the engine carries on as the single-cloud engine does.  Also the failure mode: a shard
    that outgrows its buffer in ONE resample is flagged and the next host look raises."""
    import warnings
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    _init_group(rank, world)
    try:
        import optbayesexpt_b200 as obe
        from optbayesexpt_b200.sharded import ShardedOptBayesExpt
        from oracle.scenarios import build_inputs, by_name
        sc = by_name('c1_find_peak')
        n = 50000
        inp = build_inputs(sc, n)
        cut = [0, 31000, n]
        lo, hi = cut[rank], cut[rank + 1]
        kw = dict(scale=False, default_noise_std=500.0, seed=77, auto_resample=False)
        eng = ShardedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'][:, lo:hi], inp['cons'],
                                  slack=0.5, **kw)
        ref = obe.OptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'], inp['cons'], **kw)

        def global_cloud(e):
            cap = n                               # (the same shape on every rank: the collective needs it)
            pad = torch.zeros((e.n_dims + 1, cap), dtype=torch.float64, device=e._buf.device)
            m = e.n_particles
            pad[:e.n_dims, :m] = torch.from_numpy(e.particles).to(pad.device)
            pad[e.n_dims, :m] = torch.from_numpy(e.particle_weights).to(pad.device)
            allp = e._comm.allgather(pad).cpu().numpy()
            counts = e.shard_counts
            return np.concatenate([allp[g][:, :counts[g]] for g in range(world)], axis=1), counts
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            assert eng.rebalance() is False                       # nothing near its capacity: no-op without force
            # (1) explicit, non-uniform weights travel with their particles
            rec = ((3.0,), 49700.0, 500.0)
            eng.pdf_update(rec)
            ref.pdf_update(rec)
            before, c0 = global_cloud(eng)
            assert list(c0) == [31000, 19000]
            assert eng.rebalance(force=True) is True
            after, c1 = global_cloud(eng)
            assert list(c1) == [25000, 25000] and eng.n_particles == 25000
            np.testing.assert_array_equal(after, before)
            np.testing.assert_allclose(after[-1], ref.particle_weights, rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(eng.mean(), ref.mean(), rtol=1e-12)
            np.testing.assert_allclose(eng.n_eff(), ref.n_eff(), rtol=1e-12)
            eng.rng = np.random.default_rng(3)
            ref.rng = np.random.default_rng(3)
            s1, s2 = eng.opt_setting(), ref.opt_setting()
            assert eng.last_setting_index == ref.last_setting_index
            # (2) after a resample: implicit uniform weights, lengths drifted again
            rec = (s1, 49900.0, 500.0)
            eng.pdf_update(rec)
            ref.pdf_update(rec)
            eng.resample()
            before, c2 = global_cloud(eng)
            assert c2.sum() == n and c2[0] != c2[1]
            assert eng.rebalance(force=True) is True
            after, c3 = global_cloud(eng)
            assert list(c3) == [25000, 25000]
            np.testing.assert_array_equal(after, before)
            np.testing.assert_array_equal(after[-1], np.full(n, 1.0 / n))
            # and the engine carries on: update + resample + selection
            eng.pdf_update((s1, 50100.0, 500.0))
            eng.resample()
            eng.opt_setting()
            assert eng.shard_counts.sum() == n
        # (3) failure mode: no slack at all and all the weight on rank 0's particles -> rank 0 must hold every offspring
        tight = ShardedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'][:, lo:hi], inp['cons'],
                                    slack=0.0, **kw)
        w = np.zeros(hi - lo)
        if rank == 0:
            w[:] = 1.0 / (hi - lo)
        tight.particle_weights = w
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            tight._plan_valid = False
            tight._moments_valid = False
            try:
                tight._fetch_plan()
                raised = False
            except RuntimeError as exc:
                raised = 'raise `slack`' in str(exc)
        assert raised, 'a shard that cannot hold its offspring must be reported'
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def _spawn(worker, peer, backend='gloo'):
    """Two ranks on the one GPU (gloo carries the host-side collectives).  peer='1': the stats and the draws travel
    by peer writes into CUDA-IPC-mapped buffers + flags instead of collectives (the NVLink path of a real node).
    backend='nccl': one GPU per rank."""
    import torch.multiprocessing as mp
    os.environ['OBE_PEER_EXCHANGE'] = peer          # inherited by the spawned ranks
    os.environ['OBE_TEST_BACKEND'] = backend
    try:
        ctx = mp.get_context('spawn')
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=worker, args=(r, 2, port, out)) for r in range(2)]
        for p in procs:
            p.start()
        results = []
        try:
            for _ in procs:
                results.append(out.get(timeout=240))
                if results[-1][1] != 'ok':
                    break                        # the peer of a failed rank would wait in a collective for ever
        except Exception:                        # queue.Empty: a rank hangs
            results.append((-1, 'timeout: a rank did not report within 240 s'))
        if all(msg == 'ok' for _, msg in results):
            for p in procs:
                p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.terminate()
        os.environ.pop('OBE_PEER_EXCHANGE', None)
        os.environ.pop('OBE_TEST_BACKEND', None)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'


@pytest.mark.parametrize('peer', ['0', '1'], ids=['collectives', 'peer_exchange'])
def test_two_shards_noise_parameter_engine(obe_lib, peer):
    _spawn(_worker_noise, peer)


@pytest.mark.parametrize('peer', ['0', '1'], ids=['collectives', 'peer_exchange'])
def test_two_shards_match_single_cloud(obe_lib, peer):
    _spawn(_worker, peer)


@pytest.mark.parametrize('peer', ['0', '1'], ids=['collectives', 'peer_exchange'])
def test_two_shards_early_select(obe_lib, peer):
    _spawn(_worker_early, peer)


@pytest.mark.parametrize('peer', ['0', '1'], ids=['collectives', 'peer_exchange'])
def test_two_shards_rebalance(obe_lib, peer):
    _spawn(_worker_rebalance, peer)


def _two_gpus():
    import torch
    return torch.cuda.is_available() and torch.cuda.device_count() >= 2


@pytest.mark.parametrize('peer', ['0', '1'], ids=['nccl_collectives', 'nvlink_peer_exchange'])
@pytest.mark.parametrize('worker', ['base', 'noise', 'early', 'rebalance'])
def test_two_gpus_nccl(obe_lib, peer, worker):
    """The same two-shard comparisons on TWO GPUs: NCCL collectives / peer buffers mapped over NVLink (CUDA IPC between
    devices).  Skipped on a one-GPU box (the driver's GPU test box); bench.py --gpus N runs the same invariance check
    on every multi-GPU run (`invariance` in its JSON line)."""
    if not _two_gpus():
        pytest.skip('needs >= 2 GPUs')
    _spawn({'base': _worker, 'noise': _worker_noise, 'early': _worker_early, 'rebalance': _worker_rebalance}[worker], peer, backend='nccl')
