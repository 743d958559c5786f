"""Host-side pieces of the closed loop that need no GPU: the plain-float pivot / N_eff arithmetic of
OptBayesExpt._adopt_stats must be bit-identical to the numpy expressions of ParticlePDF._mean_from / _n_eff_from
(the pivot feeds the next update kernel, so a one-ulp difference would change the moments' last bits), and the ctypes
mirror of obe_cycle_t must expose the fields the two-phase call and the completion word use."""
import types
import warnings

import numpy as np
import pytest

from optbayesexpt_b200 import _lib
from optbayesexpt_b200.obe_base import OptBayesExpt
from optbayesexpt_b200.particlepdf import ParticlePDF


def _fake_engine(d, n_total):
    e = types.SimpleNamespace(n_dims=d, _pivot=np.full(d, -1.0), n_particles=n_total)
    e._n_total_for_test = lambda: n_total
    e._mean_from = lambda st: ParticlePDF._mean_from(e, st)
    return e


@pytest.mark.parametrize('d', [1, 2, 3, 4, 8])
def test_adopt_stats_matches_the_numpy_expressions(d):
    g = np.random.default_rng(d)
    for trial in range(200):
        st = np.zeros(_lib.STATS_LEN)
        st[_lib.ST_SUMT] = g.uniform(1e-3, 1e3)
        st[_lib.ST_INVS] = 1.0 / g.uniform(1e-3, 1e3)
        st[_lib.ST_SUMSQ] = g.uniform(1e-9, 1e3)
        st[_lib.ST_PIVOT:_lib.ST_PIVOT + d] = g.normal(0, 1e3, d)
        st[_lib.ST_M1:_lib.ST_M1 + d] = g.normal(0, 1e2, d)
        e = _fake_engine(d, 10 ** 6)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            OptBayesExpt._adopt_stats(e, st)
        np.testing.assert_array_equal(e._pivot, ParticlePDF._mean_from(e, st))


def test_adopt_stats_keeps_the_pivot_when_the_block_is_degenerate():
    e = _fake_engine(3, 1000)
    st = np.zeros(_lib.STATS_LEN)                      # sum of weights 0: no mean
    OptBayesExpt._adopt_stats(e, st)
    np.testing.assert_array_equal(e._pivot, np.full(3, -1.0))
    st[_lib.ST_SUMT] = 1.0
    st[_lib.ST_M1] = np.nan                            # a NaN moment: keep the old pivot
    st[_lib.ST_INVS], st[_lib.ST_SUMSQ] = 1.0, 1.0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        OptBayesExpt._adopt_stats(e, st)
    np.testing.assert_array_equal(e._pivot, np.full(3, -1.0))


def test_adopt_stats_warns_on_impoverishment():
    e = _fake_engine(2, 10 ** 4)
    st = np.zeros(_lib.STATS_LEN)
    st[_lib.ST_SUMT], st[_lib.ST_INVS] = 1.0, 1.0
    st[_lib.ST_SUMSQ] = 1.0 / 500.0                    # N_eff = 500 < 0.1 * 1e4
    with pytest.warns(RuntimeWarning, match='Particle filter rejected'):
        OptBayesExpt._adopt_stats(e, st)
    st[_lib.ST_SUMSQ] = 1.0 / 5000.0                   # N_eff = 5000: quiet
    with warnings.catch_warnings():
        warnings.simplefilter('error')
        OptBayesExpt._adopt_stats(e, st)


def test_cycle_struct_has_the_closed_loop_fields():
    names = [n for n, _ in _lib.Cycle._fields_]
    for field in ('resample_threshold', 'stats_host', 'stats_src_dev', 'best_host', 'phase', 'seq', 'seq_host'):
        assert field in names
    assert _lib.ST_FIRED == _lib.STATS_LEN - 1
    assert 'obe_stream_sync' in _lib.SIGNATURES
