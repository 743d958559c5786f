"""CPU tests (gloo, world_size 2) of the host-side logic of the sharded engine: the three small
collectives and the pure functions that turn gathered stats into global moments, shard slot bounds,
draw ownership and the global argmax.  No GPU involved."""
import os
import socket

import numpy as np
import pytest

from oracle import obe_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _fake_stats(particles, t, pivot, d):
    """What the update kernel leaves in a shard's stats block, computed with numpy."""
    from optbayesexpt_b200 import _lib
    st = np.zeros(_lib.STATS_LEN)
    st[_lib.ST_TOTAL] = t.sum()
    st[_lib.ST_SUMT] = t.sum()
    st[_lib.ST_SUMSQ] = (t * t).sum()
    dx = particles - pivot[:, None]
    st[_lib.ST_M1:_lib.ST_M1 + d] = (dx * t).sum(axis=1)
    q = _lib.ST_M2
    for j in range(d):
        for k in range(j, d):
            st[q] = (t * dx[j] * dx[k]).sum()
            q += 1
    st[_lib.ST_PIVOT:_lib.ST_PIVOT + d] = pivot
    return st


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from optbayesexpt_b200 import _lib
        from optbayesexpt_b200 import sharded as sh
        lib = _lib.load()
        comm = sh.Comm()
        assert comm.rank == rank and comm.world == world
        # ---- collectives
        g = comm.allgather(torch.tensor([float(rank), 10.0 + rank], dtype=torch.float64))
        assert g.shape == (world, 2) and g[1, 0].item() == 1.0 and g[0, 1].item() == 10.0
        r = comm.allreduce_sum(torch.ones(3, dtype=torch.float64) * (rank + 1))
        assert r[0].item() == sum(range(1, world + 1))
        # ---- a cloud split unevenly over the ranks; dyadic weights make every sum exact
        n, d = 6000, 3
        rng = np.random.default_rng(5)
        particles = rng.standard_normal((d, n)) * np.array([[1.0], [50.0], [0.01]]) + np.array([[3.0], [-900.0], [5e4]])
        t = rng.integers(0, 64, n).astype(np.float64) / 4096.0
        cut = [0, 2500, n]
        lo, hi = cut[rank], cut[rank + 1]
        pivot = particles.mean(axis=1)
        mine = torch.from_numpy(_fake_stats(particles[:, lo:hi], t[lo:hi], pivot, d))
        gathered = comm.allgather(mine).numpy()
        gs = sh.combine_stats(gathered, d)
        mean, cov, var, n_eff = sh.moments_from(gs, d)
        w = t / t.sum()
        np.testing.assert_allclose(mean, orc.weighted_mean(particles, w), rtol=1e-12)
        np.testing.assert_allclose(cov, orc.weighted_covariance_longdouble(particles, w), rtol=1e-10)
        np.testing.assert_allclose(np.sqrt(var), orc.std_centered(particles, w), rtol=1e-10)
        np.testing.assert_allclose(n_eff, orc.n_effective(w), rtol=1e-12)
        assert gs['offsets'][0] == 0.0 and gs['offsets'][1] == t[:2500].sum() and gs['total'] == t.sum()
        # ---- shard slot bounds: consistent with the single-cloud systematic resample of the oracle
        for u0 in (0.0, 0.37, 0.999999):
            bounds = sh.shard_slot_bounds(gs['offsets'], gs['total'], u0, n, lib.obe_comb_count)
            assert bounds[0] == 0 and bounds[-1] == n and bounds[1] >= 0
            cdf = np.cumsum(t) * (1.0 / t.sum())
            cdf[-1] = 1.0
            anc = orc.search_cdf(cdf, orc.systematic_uniforms(u0, n))
            # slots below the bound descend from shard 0's particles, the rest from shard 1's
            assert np.all(anc[:bounds[1]] < 2500) and np.all(anc[bounds[1]:] >= 2500), u0
        # ---- draws: every uniform has exactly one owner and a local uniform that picks the same particle
        u = np.random.default_rng(9).random(64)
        owner, local = sh.assign_draws(u, gs['offsets'], gs['totals'], gs['total'])
        want = orc.choice_indices(w, u)
        assert np.array_equal(owner, (want >= 2500).astype(int))
        for g_ in range(world):
            sel = owner == g_
            wl = t[cut[g_]:cut[g_ + 1]]
            got = orc.choice_indices(wl / wl.sum(), local[sel]) + cut[g_]
            assert np.array_equal(got, want[sel])
        # ---- argmax reduce: first maximum over the concatenated grid, NaN is the maximum
        assert sh.reduce_best([(7, 1.0), (3, 1.0)])[0] == 3
        assert sh.reduce_best([(7, 2.0), (3, 1.0)])[0] == 7
        assert sh.reduce_best([(7, 2.0), (9, float('nan'))])[0] == 9
        assert sh.reduce_best([(-1, 0.0), (4, -5.0)])[0] == 4
        assert sh.setting_slice(10, 0, 3) == (0, 3) and sh.setting_slice(10, 2, 3) == (6, 10)
        out.put((rank, 'ok'))
    except Exception as exc:  # pragma: no cover
        import traceback
        out.put((rank, traceback.format_exc()))
        raise exc
    finally:
        dist.destroy_process_group()


def test_sharded_host_logic_gloo_world2(obe_lib):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'


def test_comb_count_matches_numpy_definition(obe_lib):
    """obe_comb_count(c, u0, n) == #{i : (i + u0) * (1/n) < c} with numpy's IEEE arithmetic."""
    rng = np.random.default_rng(1)
    for n in (1, 5, 2048, 99991):
        u = orc.systematic_uniforms(0.3, n)
        for c in list(rng.random(50)) + [0.0, 1.0, u[n // 2], np.nextafter(u[n // 2], 1.0)]:
            assert obe_lib.obe_comb_count(float(c), 0.3, n) == int(np.sum(u < c)), (n, c)
