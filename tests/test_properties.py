"""Property tests (hypothesis, CPU): the size-independent invariants the GPU tests check at scale, stated on
the oracle and on the host-side sharding logic -- comb ancestors, shard slot bounds, draw ownership, argmax
combination, sweep pairs."""
import numpy as np
from hypothesis import given, settings, strategies as st
from numpy.testing import assert_array_equal

from oracle import obe_oracle as orc
from optbayesexpt_b200 import sharded as sh

FAST = settings(max_examples=60, deadline=None)


def _weights(rng, n, dead):
    w = rng.random(n) ** 4
    w[rng.random(n) < dead] = 0.0
    if w.sum() == 0.0:
        w[rng.integers(0, n)] = 1.0
    return w / w.sum()


def _host_comb_count(c, u0, n):
    """#{i in [0,n) : (i + u0) * (1/n) < c} by the definition (the library's obe_comb_count is tested against
    the same definition in test_sharded_logic.py)."""
    return int(np.count_nonzero(orc.systematic_uniforms(u0, n) < c))


@FAST
@given(n=st.integers(1, 5000), seed=st.integers(0, 2 ** 31), dead=st.floats(0.0, 0.9), u0=st.floats(0.0, 0.999999))
def test_comb_ancestors_are_monotone_and_counts_are_floor_or_ceil(n, seed, dead, u0):
    rng = np.random.default_rng(seed)
    w = _weights(rng, n, dead)
    idx = orc.search_cdf(orc.normalized_cdf(w), orc.systematic_uniforms(u0, n))
    assert idx.shape == (n,) and idx.min() >= 0 and idx.max() <= n - 1
    assert np.all(np.diff(idx) >= 0)                       # ancestors of consecutive comb teeth never go back
    counts = np.bincount(idx, minlength=n)
    assert counts.sum() == n
    assert np.all(np.abs(counts - n * w) < 1.0 + 1e-9)     # systematic: floor(n w) or ceil(n w) offspring
    assert np.all(counts[w == 0.0] == 0) or idx[-1] == n - 1   # dead particles get nothing (clamp aside)


@FAST
@given(world=st.integers(1, 8), seed=st.integers(0, 2 ** 31), u0=st.floats(0.0, 0.999999), n=st.integers(8, 20000))
def test_shard_slot_bounds_cover_the_comb_exactly_once(world, seed, u0, n):
    rng = np.random.default_rng(seed)
    totals = rng.random(world) * (rng.random(world) > 0.2)
    if totals.sum() == 0.0:
        totals[0] = 1.0
    offsets = np.concatenate(([0.0], np.cumsum(totals)[:-1]))
    total = float(totals.sum())
    b = sh.shard_slot_bounds(offsets, total, u0, n, _host_comb_count)
    assert b[0] == 0 and b[-1] == n and len(b) == world + 1
    assert all(b[g] <= b[g + 1] for g in range(world))     # monotone: every slot has exactly one owner
    # the owner of tooth i is the shard whose CDF interval holds it
    teeth = orc.systematic_uniforms(u0, n)
    ends = (offsets + totals) / total
    for g in range(world):
        mine = teeth[b[g]:b[g + 1]]
        if len(mine) and g + 1 < world:
            assert mine.max() < offsets[g + 1] / total + 1e-15
        if len(mine) and g > 0:
            assert mine.min() >= offsets[g] / total - 1e-15
    assert ends[-1] > 0


@FAST
@given(world=st.integers(1, 8), seed=st.integers(0, 2 ** 31), k=st.integers(1, 64))
def test_every_draw_has_exactly_one_owner_with_weight(world, seed, k):
    rng = np.random.default_rng(seed)
    totals = rng.random(world) * (rng.random(world) > 0.3)
    if totals.sum() == 0.0:
        totals[-1] = 0.5
    offsets = np.concatenate(([0.0], np.cumsum(totals)[:-1]))
    u = rng.random(k)
    owner, local = sh.assign_draws(u, offsets, totals, float(totals.sum()))
    assert owner.shape == (k,) and np.all((owner >= 0) & (owner < world))
    assert np.all(totals[owner] > 0)                       # never an empty shard
    assert np.all((local >= 0.0) & (local < 1.0))
    # the shard-local uniform maps back to the global one
    back = (offsets[owner] + local * totals[owner]) / totals.sum()
    assert np.all(np.abs(back - u) < 1e-12)


@FAST
@given(seed=st.integers(0, 2 ** 31), world=st.integers(1, 8), s=st.integers(1, 300), nans=st.booleans())
def test_reduce_best_is_numpy_argmax_over_the_whole_grid(seed, world, s, nans):
    rng = np.random.default_rng(seed)
    u = np.round(rng.random(s), 2)                          # ties on purpose
    if nans and s > 2:
        u[rng.integers(0, s)] = np.nan
    pairs = []
    for r in range(world):
        lo, hi = sh.setting_slice(s, r, world)
        if hi > lo:
            j = int(np.argmax(u[lo:hi]))
            pairs.append((lo + j, float(u[lo + j])))
        else:
            pairs.append((-1, 0.0))
    best, val = sh.reduce_best(pairs)
    assert best == int(np.argmax(u))
    assert (val != val) if np.isnan(u[best]) else val == u[best]


@FAST
@given(n=st.integers(2, 400), sub=st.integers(1, 9))
def test_sweep_pairs(n, sub):
    pairs = orc.sweep_start_stop_indices(n, sub)
    grid = sorted(set(range(0, n, sub)) | {n - 1})
    assert len(pairs) == len(grid) * (len(grid) - 1) // 2
    assert np.all(pairs[:, 1] > pairs[:, 0])
    assert set(pairs.ravel().tolist()) <= set(grid)
    assert_array_equal(pairs[0], [grid[0], grid[1]])
    # the utility of a pair is additive along the sweep
    u = np.random.default_rng(n).random(n)
    su = orc.sweep_utility(u, pairs, 5.0)
    k = len(pairs) // 2
    a, b = pairs[k]
    assert abs(su[k] * (b - a + 5.0) - u[a + 1:b + 1].sum()) < 1e-9
