"""Sweeper workload (demos/sweeper/obe_sweeper.py) through the C ABI on the GPU.

The golden file holds the closed loop of the UNMODIFIED reference class (oracle/pin_against_reference.py:
run_sweeper_lockstep): 10 sweeps, 210 point updates, 19 multinomial resamples.
"""
import os
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import obe_oracle as orc
from oracle.scenarios import SWEEPER, build_inputs

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(scope='module')
def obe():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import optbayesexpt_b200 as pkg
    return pkg


def _sweeper(obe, **kw):
    sc = SWEEPER
    inp = build_inputs(sc)
    args = dict(n_draws=sc['n_draws'], scale=sc['scale'], a_param=sc['a_param'],
                resample_threshold=sc['resample_threshold'], resampling='multinomial')
    args.update(kw)
    eng = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'],
                                  noise_parameter_index=sc['noise_parameter_index'], **args)
    eng.rng = np.random.default_rng(sc['seed_rng'])
    return eng, inp


def test_sweeper_attributes_match_reference(obe):
    eng, inp = _sweeper(obe)
    g = np.load(os.path.join(GOLDEN, SWEEPER['name'] + '.npz'))
    assert eng.start_stop_subsample == 3 and eng.cost_of_new_sweep == 5.0
    assert_array_equal(eng.start_stop_indices, g['start_stop_indices'])            # obe_sweeper.py:213-232
    assert_array_equal(eng.start_stop_choice_indices, np.arange(len(g['start_stop_indices'])))
    assert_array_equal(eng.start_stop_values, inp['setting_values'][0][g['start_stop_indices']])
    assert_array_equal(eng.sweep_cost_estimate(), g['start_stop_indices'][:, 1] - g['start_stop_indices'][:, 0] + 5.0)
    assert eng.cost_estimate() == 1.0


def test_sweep_utility_kernel_matches_oracle(obe):
    """cumsum + pair kernel against numpy on the device's own point utility; argmax identical."""
    eng, _ = _sweeper(obe)
    su = eng.sweep_utility()
    pu = eng._utility_dev.cpu().numpy()
    want = orc.sweep_utility(pu, eng.start_stop_indices, eng.cost_of_new_sweep)
    # a pair utility is a difference of two cumsum values: tolerance relative to the cumsum's size
    cost = eng.sweep_cost_estimate()
    assert_allclose(su, want, rtol=1e-12, atol=1e-13 * np.cumsum(pu)[-1] / cost.min())
    assert_allclose(eng._cumsum_dev.cpu().numpy(), np.cumsum(pu), rtol=1e-13)
    best = eng._pair_best_dev.cpu().numpy()
    assert int(best[0]) == int(np.argmax(su))
    # a long setting axis (many chunks of the single-CTA scan) and a hand-made pair list
    import torch, ctypes as C
    from optbayesexpt_b200 import _lib
    n, rng = 300_001, np.random.default_rng(5)
    u = rng.random(n)
    pairs = np.sort(rng.integers(0, n, size=(50_000, 2)), axis=1).astype(np.int32)
    pairs[:, 1] = np.minimum(pairs[:, 1] + 1, n - 1)
    ud, pd = torch.from_numpy(u).cuda(), torch.from_numpy(pairs).cuda()
    cum = torch.empty(n, dtype=torch.float64, device='cuda')
    out = torch.empty(len(pairs), dtype=torch.float64, device='cuda')
    best = torch.zeros(2, dtype=torch.int64, device='cuda')
    lib = _lib.load()
    scratch = torch.zeros(int(lib.obe_select_scratch_bytes(n)), dtype=torch.uint8, device='cuda')
    _lib.check(lib.obe_sweep_utility(C.c_void_p(ud.data_ptr()), n, C.c_void_p(pd.data_ptr()), len(pairs), 2.5,
                                     C.c_void_p(cum.data_ptr()), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(best.data_ptr()), C.c_void_p(scratch.data_ptr()), None))
    torch.cuda.synchronize()
    ref_cum = np.cumsum(u)
    assert_allclose(cum.cpu().numpy(), ref_cum, rtol=1e-13)
    got = out.cpu().numpy()
    mine = (cum.cpu().numpy()[pairs[:, 1]] - cum.cpu().numpy()[pairs[:, 0]]) / ((pairs[:, 1] - pairs[:, 0]) + 2.5)
    assert_array_equal(got, mine)                               # bit-exact on the device's own cumsum
    assert int(best[0].item()) == int(np.argmax(got))           # first maximum


@pytest.mark.parametrize('fused', [False, True], ids=['point_by_point', 'fused_multi_point'])
def test_sweeper_golden_trajectory(obe, fused):
    """Seeded like the reference run: every chosen (start, stop) pair identical, sweep utility, moments and
    final weights within the condition-aware tolerance -- with one launch per point as the reference loops,
    and with the multi-point kernel (one pass per segment between resamples)."""
    sc = SWEEPER
    g = np.load(os.path.join(GOLDEN, sc['name'] + '.npz'))
    eng, inp = _sweeper(obe)
    eng.fused_sweep = fused
    n_res = 0
    xvals = inp['setting_values'][0]
    ofs = np.concatenate(([0], np.cumsum(g['y_lengths'])))
    tol = sc['traj_rtol']
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for t in range(sc['n_sweeps']):
            pair = eng.opt_setting()
            assert eng.last_setting_index == g['pair_index'][t], f'pair index differs at sweep {t}'
            assert_array_equal(pair, g['pairs'][t])
            su = eng._pair_utility_dev.cpu().numpy()
            assert_allclose(su, g['sweep_utility'][t], rtol=tol if t else 1e-12,
                            atol=1e-13 * np.sum(g['point_utility'][t]) / 8.0, err_msg=f'sweep utility t={t}')
            assert_allclose(eng._utility_dev.cpu().numpy(), g['point_utility'][t], rtol=tol if t else 1e-12)
            y = g['y_concat'][ofs[t]:ofs[t + 1]]
            e0 = eng._epoch
            eng.pdf_update(((xvals[pair[0]:pair[1]],), y))
            assert eng._epoch - e0 == g['n_resamples'][t], f'number of resamples in sweep {t}'
            assert_allclose(eng.mean(), g['mean'][t], rtol=tol, err_msg=f'mean t={t}')
            assert_allclose(eng.std(), g['std'][t], rtol=1e-6, err_msg=f'std t={t}')
    w = eng.particle_weights
    assert_allclose(w, g['final_weights'], rtol=max(tol, 1e-8), atol=1e-15 * g['final_weights'].max())


def test_sweeper_good_and_random_setting(obe):
    from optbayesexpt_b200 import obe_sweeper
    eng, _ = _sweeper(obe, selection_method='good', pickiness=20)
    obe_sweeper.rng = np.random.default_rng(77)
    pair = eng.get_setting()                                    # good_setting through the selection_method kwarg
    su = eng._pair_utility_dev.cpu().numpy()
    u = np.random.default_rng(77).random()
    want = orc.good_index(su, 20, u)
    assert eng.last_setting_index == want
    assert_array_equal(pair, eng.start_stop_indices[want])
    pair = eng.random_setting()
    assert pair[1] > pair[0]


def test_sweeper_default_systematic_converges(obe):
    """Default (systematic, device RNG) sweeper: the posterior lands on the truth."""
    sc = SWEEPER
    inp = build_inputs(sc)
    eng = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'],
                                  noise_parameter_index=3, scale=False, seed=3)
    meas = np.random.default_rng(4)
    xvals = inp['setting_values'][0]
    model = orc.MODELS[sc['model']][0]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for _ in range(25):
            a, b = eng.opt_setting()
            xs = xvals[a:b]
            y = model((xs,), sc['true_pars'], inp['cons']) + sc['noise'] * meas.standard_normal(len(xs))
            eng.pdf_update(((xs,), y))
    truth = np.array(list(sc['true_pars']) + [sc['noise']])
    err = np.abs(eng.mean() - truth) / eng.std()
    assert np.all(err < 5), (eng.mean(), eng.std())


def test_multi_point_update_matches_point_by_point(obe):
    """obe_update_multi against M single updates on the same cloud (no resample in between): weights to
    1e-12 after normalisation, per-point N_eff from the kernel's sums against the engine's own."""
    import ctypes as C
    import torch
    from optbayesexpt_b200 import _lib
    sc = SWEEPER
    inp = build_inputs(sc, 300_000)
    xs = inp['setting_values'][0][20:75]
    model = orc.MODELS[sc['model']][0]
    ys = model((xs,), sc['true_pars'], inp['cons']) + sc['noise'] * np.random.default_rng(3).standard_normal(len(xs))
    a = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'], noise_parameter_index=3,
                                scale=False, seed=1, auto_resample=False)
    b = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'], noise_parameter_index=3,
                                scale=False, seed=1, auto_resample=False)
    a.fused_sweep, b.fused_sweep = True, False
    a._sweep_chunk = 128               # (without the resample test a launch is capped at 8 points: 7 launches)
    neff = []
    for x, y in zip(xs, ys):
        b.pdf_update(((np.array([x]),), np.array([y])))
        neff.append(b.n_eff())
    a.pdf_update(((xs,), ys))
    wa, wb = a.particle_weights, b.particle_weights
    assert_allclose(wa, wb, rtol=1e-11, atol=1e-15 * wb.max())
    assert_allclose(a.mean(), b.mean(), rtol=1e-11)
    assert_allclose(a.std(), b.std(), rtol=1e-8)
    # the whole sweep in ONE launch (the kernel entry itself, below the chunk policy): per-point sums -> N_eff
    d = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'], noise_parameter_index=3,
                                scale=False, seed=1, auto_resample=False)
    first, _ = d._multi_update(xs, ys)
    assert first == -1
    sums = d._multi_sums.cpu().numpy()[:len(xs)]
    assert_allclose(sums[:, 0] ** 2 / sums[:, 1], neff, rtol=1e-10)
    # with the resample test on, the kernel reports the first point whose N_eff falls below the threshold
    c = obe.OptBayesExptSweeper(sc['model'], inp['setting_values'], inp['prior'], inp['cons'], noise_parameter_index=3,
                                scale=False, seed=1)
    first, ratio = c._multi_update(xs, ys)                      # (one launch regardless of the chunk policy)
    want = int(np.argmax(np.array(neff) / 300_000 < 0.5))
    assert first == want and abs(ratio - neff[want] / 300_000) < 1e-10
