"""Early select: the K draws of opt_setting taken from the resample PLAN (k_sys_resample_warp<D, true>) before the
cloud is streamed, the utility pass overlapped with the resample on a second stream.

The contract: draw q is the offspring in output slot floor(u_q * N); its value is bit-identical to what the streaming
kernel stores in that slot; the utility / argmax computed from those draws equal the oracle's on the same draws; and a
closed loop with ``eager_select`` takes the same decisions as the classic order (resample, then draw from the
offspring cloud through its CDF)."""
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def obe():
    import torch
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import optbayesexpt_b200 as pkg
    return pkg


def _engine(obe, n, seed=3, n_set=400, **kw):
    g = np.random.default_rng(seed)
    prior = np.array([g.uniform(2, 4, n), g.uniform(-2000, -400, n), g.normal(50000, 1000, n)])
    settings = (np.linspace(1.5, 4.5, n_set),)
    eng = obe.OptBayesExpt('lorentzian_hwhm', settings, prior, (0.1,), scale=False, default_noise_std=500.0, seed=11,
                           **kw)
    return eng, prior, settings


@pytest.mark.parametrize('n', [1000, 2048, 10000, 250_007, 3_000_001])
@pytest.mark.parametrize('sharp', [False, True], ids=['flat', 'sharp'])
def test_picked_draws_are_the_offspring(obe, n, sharp):
    """run_cycle_async with early select: the draws equal the resampled cloud at slots floor(u*N), bit for bit, whatever
    the weights look like (flat: ~1 slot per particle; sharp: a few tiles own most of the slots -> multi-unit tiles)."""
    from oracle import obe_oracle as orc
    eng, prior, settings = _engine(obe, n)
    assert eng._early_select_ok()
    sigma = 20.0 if sharp else 500.0
    rec = ((3.1,), 49500.0, sigma)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for cycle in range(3):
            eng.rng = np.random.default_rng(100 + cycle)
            twin = np.random.default_rng(100 + cycle)
            eng.run_cycle_async(rec)
            twin.random()                                   # the comb offset
            u = twin.random(eng.N_DRAWS)
            slots = np.minimum((u * n).astype(np.int64), n - 1)
            draws = eng._draws_dev.cpu().numpy()
            cloud = eng.particles
            np.testing.assert_array_equal(draws, cloud[:, slots])
            # the selection made from them: oracle utility on the same draws
            var_p, _ = orc.yvar_from_draws(orc.model_lorentzian_hwhm, orc.make_allsettings(settings), draws, (0.1,), 1)
            util = orc.utility_variance(var_p, orc.noise_var_default(500.0, 1))
            got = eng._utility_dev.cpu().numpy()
            np.testing.assert_allclose(got, util, rtol=1e-12)
            assert int(eng.best_index_dev.cpu()[0]) == orc.opt_index(got)
            assert abs(eng.particle_weights.sum() - 1.0) < 1e-12


def test_early_select_equals_classic_order(obe):
    """Same seeds, early select on / off: identical offspring clouds; the draws (slot rule vs the offspring CDF) and
    therefore the utilities agree."""
    n = 200_000
    a, _, _ = _engine(obe, n)
    b, _, _ = _engine(obe, n)
    b.early_select = False
    rec = ((2.9,), 49800.0, 300.0)
    for cycle in range(3):
        a.rng = np.random.default_rng(7 + cycle)
        b.rng = np.random.default_rng(7 + cycle)
        a.run_cycle_async(rec)
        b.run_cycle_async(rec)
        np.testing.assert_array_equal(a.particles, b.particles)
        np.testing.assert_array_equal(a._draws_dev.cpu().numpy(), b._draws_dev.cpu().numpy())
        np.testing.assert_array_equal(a._utility_dev.cpu().numpy(), b._utility_dev.cpu().numpy())
        assert int(a.best_index_dev.cpu()[0]) == int(b.best_index_dev.cpu()[0])


def test_eager_select_closed_loop(obe):
    """pdf_update + opt_setting with eager_select: the resample inside pdf_update starts the selection; opt_setting
    only fetches the argmax.  Same trajectory as the classic engine."""
    n = 50_000
    a, _, settings = _engine(obe, n, resample_threshold=0.9)
    b, _, _ = _engine(obe, n, resample_threshold=0.9)
    a.eager_select = True
    a.rng = np.random.default_rng(5)
    b.rng = np.random.default_rng(5)
    meas = np.random.default_rng(9)
    resamples = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        xa, xb = a.opt_setting(), b.opt_setting()
        for t in range(25):
            assert xa == xb and a.last_setting_index == b.last_setting_index, f'cycle {t}'
            y = 50400.0 - 1200.0 / (((xa[0] - 3.14) / 0.1) ** 2 + 1) + 500.0 * meas.standard_normal()
            a.pdf_update((xa, y, 500.0))
            b.pdf_update((xb, y, 500.0))
            assert a.just_resampled == b.just_resampled
            if a.just_resampled:
                resamples += 1
                assert a._select_ready
            xa, xb = a.opt_setting(), b.opt_setting()
            assert not a._select_ready
    assert resamples >= 3, 'the loop never resampled: the test does not exercise the eager path'
    np.testing.assert_array_equal(a.particles, b.particles)
    np.testing.assert_allclose(a.mean(), b.mean(), rtol=1e-13)


def test_selection_is_invalidated_by_a_new_cloud(obe):
    n = 20_000
    a, prior, _ = _engine(obe, n, resample_threshold=2.0)
    a.eager_select = True
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        a.pdf_update(((3.0,), 49900.0, 500.0))
        assert a.just_resampled and a._select_ready
        a.set_pdf(prior)                         # a new cloud: the prefetched selection must not be used
        assert not a._select_ready
        a.opt_setting()
        a.pdf_update(((3.0,), 49900.0, 500.0))
        assert a._select_ready
        a.set_n_draws(12)
        assert not a._select_ready
        a.opt_setting()
        assert a._draws_dev.shape == (3, 12)


def test_noise_parameter_engine_keeps_the_classic_order(obe):
    """sigma as a parameter: the positivity constraint re-weights the offspring, so the draws cannot come from the plan."""
    g = np.random.default_rng(2)
    n = 20_000
    prior = np.array([g.normal(0, 2, n), g.normal(0, 2, n), g.exponential(1.0, n)])
    eng = obe.OptBayesExptNoiseParameter('line', (np.linspace(-1, 1, 101),), prior, (), noise_parameter_index=2,
                                         scale=False, seed=4)
    assert not eng._early_select_ok()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        eng.run_cycle_async(((0.3,), 0.2))
    assert abs(eng.particle_weights.sum() - 1.0) < 1e-12


@pytest.mark.parametrize('kind', ['base', 'noise', 'base_classic_select'])
def test_cycle_entry_equals_the_stepwise_path(obe, kind):
    """run_cycle_async through the one-call C entry (obe_cycle) against the same cycle enqueued step by step from
    Python: identical clouds, weights, draws, utilities and argmax, for every resample/select combination."""
    g = np.random.default_rng(8)
    n = 30_000

    def make():
        if kind == 'noise':
            prior = np.array([g0.normal(0, 2, n), g0.normal(0, 2, n), g0.exponential(1.0, n)])
            return obe.OptBayesExptNoiseParameter('line', (np.linspace(-1, 1, 101),), prior, (),
                                                  noise_parameter_index=2, scale=False, seed=4, a_param=0.6)
        eng, _, _ = _engine(obe, n)
        if kind == 'base_classic_select':
            eng.early_select = False
        return eng
    g0 = np.random.default_rng(21)
    a = make()
    g0 = np.random.default_rng(21)
    b = make()
    b.use_cycle_entry = False
    rec = ((0.3,), 0.2) if kind == 'noise' else ((3.05,), 49700.0, 400.0)
    combos = [(True, True), (False, True), (True, False), (False, False), (True, True)]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for t, (res, sel) in enumerate(combos):
            a.rng = np.random.default_rng(60 + t)
            b.rng = np.random.default_rng(60 + t)
            a.run_cycle_async(rec, resample=res, select=sel)
            b.run_cycle_async(rec, resample=res, select=sel)
            np.testing.assert_array_equal(a.particles, b.particles)
            np.testing.assert_array_equal(a.particle_weights, b.particle_weights)
            if sel:
                np.testing.assert_array_equal(a._draws_dev.cpu().numpy(), b._draws_dev.cpu().numpy())
                np.testing.assert_array_equal(a._utility_dev.cpu().numpy(), b._utility_dev.cpu().numpy())
                assert int(a.best_index_dev.cpu()[0]) == int(b.best_index_dev.cpu()[0])
            np.testing.assert_allclose(a.mean(), b.mean(), rtol=1e-14)
            assert a.just_resampled == b.just_resampled or not res


@pytest.mark.parametrize('slot_begin', [1, 2, 3, 5, 4098])
@pytest.mark.parametrize('d', [1, 3, 4])
def test_shard_whose_first_slot_is_not_a_multiple_of_4(obe, slot_begin, d):
    """The one-kernel resample aligns its emission groups to the shard's OUTPUT; on a shard whose first global slot is
    not a multiple of 4 the groups straddle the 4-slot blocks of the packed normal stream (jitter_group4_w, one more
    Philox call).  Same offspring, bit for bit, as the two-kernel path (k_sys_ancestors + k_sys_move, which groups by
    GLOBAL slots), ancestors included -- the normals of a slot must not depend on the grouping."""
    import ctypes as C
    import torch
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    n = 300_001
    g = np.random.default_rng(17)
    prior = g.normal(0, 1, (d, n))
    w = g.exponential(1.0, n)
    w[1000:5000] *= 50.0
    w /= w.sum()
    pdf = obe.ParticlePDF(prior, scale=False, resampling='systematic', seed=5)
    pdf.particle_weights = w
    pdf._ensure_moments()
    cov, mean = pdf.covariance(), pdf.mean()
    factor = np.ascontiguousarray(np.linalg.cholesky((1 - 0.98 ** 2) * np.atleast_2d(cov)).T)
    total = float(pdf._fetch_stats()[_lib.ST_TOTAL])
    # the cloud as the LAST shard of a bigger one: slot_begin virtual offspring belong to the shards before it
    n_total = n + slot_begin
    cdf_total = total / (1.0 - slot_begin / n_total)
    cdf_offset = cdf_total - total
    out = {}
    for fused in (1, 0):
        alt = pdf._buf.empty_like()
        idx = torch.full((n,), -1, dtype=torch.int64, device='cuda')
        zout = torch.zeros((n, d), dtype=torch.float64, device='cuda')
        _lib.check(lib.obe_set_option(b'resample_fused', fused))
        try:
            _lib.check(lib.obe_resample_systematic_sharded(
                pdf._cs(), C.byref(alt.struct()), 0.4142, n_total, slot_begin, n_total, cdf_offset, cdf_total, 1,
                _lib.darr(factor.reshape(-1)), _lib.darr(mean), 99, 3, 0.98, 0, C.c_void_p(idx.data_ptr()),
                C.c_void_p(zout.data_ptr()), pdf._stream()))
        finally:
            _lib.check(lib.obe_set_option(b'resample_fused', 1))
        out[fused] = (alt.particles[:, :n].clone(), idx.clone(), zout.clone())
    assert torch.equal(out[1][1], out[0][1])
    assert torch.equal(out[1][2], out[0][2])
    assert torch.equal(out[1][0], out[0][0])
    idx = out[1][1]
    assert int(idx.min()) >= 0 and int(idx.max()) < n and bool((idx[1:] >= idx[:-1]).all())
    # the normals are the packed stream indexed by GLOBAL slot (restated in the oracle)
    from oracle import obe_oracle as orc
    want = orc.device_normals_packed(n, d, 99, 3, slot_begin=slot_begin)
    np.testing.assert_allclose(out[1][2].cpu().numpy(), want, rtol=0, atol=1e-4)
    moved = (out[1][0] - torch.from_numpy(prior).cuda()[:, idx]).abs().max().item()
    assert 0.0 < moved < 10.0 * float(np.abs(factor).max()) + 1e-12
