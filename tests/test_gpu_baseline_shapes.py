"""GPU parity at the BASELINE shapes themselves, the distribution of the device normals,
and the host-side additions of round 2 (utility methods by name, device-side multinomial resample, constraints in
the asynchronous cycle, sweeper re-initialisation, explicit device checks)."""
import ctypes as C
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import obe_oracle as orc
from oracle.scenarios import build_inputs, by_name

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def obe():
    import optbayesexpt_b200 as pkg
    return pkg


@pytest.fixture(scope='module')
def torch():
    import torch as t
    assert t.cuda.is_available(), 'GPU tests need a CUDA device'
    return t


def wclose(actual, desired, rtol=1e-12):
    assert_allclose(actual, desired, rtol=rtol, atol=1e-15 * np.max(np.abs(desired)))


def cov_close(actual, desired, tol):
    sd = np.sqrt(np.diag(desired))
    err = np.abs(actual - desired) / np.outer(sd, sd)
    assert err.max() < tol, f'covariance off by {err.max():.3g} (correlation units)'


# =================================================================================================
# c2 at 1e5 particles (line + unknown sigma) and c3 at 1e6 (Rabi, 101 x 101 'ij' grid): update weights, moments,
# utility to 1e-12 and the chosen index against the oracle, on the real shapes of BASELINE configs[1] / [2]
# =================================================================================================
@pytest.mark.parametrize('name,n', [('c2_line_noise', 100_000), ('c3_pipulse', 1_000_000)])
def test_baseline_shape_matches_oracle(obe, name, n):
    sc = by_name(name)
    inp = build_inputs(sc, n)
    kw = dict(n_draws=sc['n_draws'], scale=False, resampling='multinomial')
    if sc['kind'] == 'noise':
        eng = obe.OptBayesExptNoiseParameter(sc['model'], inp['setting_values'], inp['prior'], inp['cons'],
                                             noise_parameter_index=sc['noise_parameter_index'], **kw)
        rec = ((0.4,), 0.1)
    else:
        eng = obe.OptBayesExpt(sc['model'], inp['setting_values'], inp['prior'], inp['cons'],
                               default_noise_std=sc['default_noise_std'], **kw)
        rec = ((0.31, 1.2), 99700.0, 315.0)
    eng.tuning_parameters['auto_resample'] = False
    model, _, _, _, nch = orc.MODELS[sc['model']]
    assert eng.allsettings.shape == orc.make_allsettings(inp['setting_values']).shape
    assert_array_equal(eng.allsettings, orc.make_allsettings(inp['setting_values']))     # 'ij' flattening
    w = np.ones(n) / n
    for step in range(2):
        y = (model(rec[0], inp['prior'], inp['cons']),)
        if sc['kind'] == 'noise':
            lik = orc.likelihood_noise_parameter(y, rec[1], inp['prior'], sc['noise_parameter_index'])
        else:
            lik = orc.likelihood_known_sigma(y, rec[1], rec[2])
        w = orc.normalized_product(w, lik)
        eng.pdf_update(rec)
        wclose(eng.particle_weights, w)
        assert_allclose(eng.n_eff(), orc.n_effective(w), rtol=1e-12)
        assert_allclose(eng.mean(), orc.weighted_mean(inp['prior'], w), rtol=1e-12)
        cov_close(eng.covariance(), orc.weighted_covariance_longdouble(inp['prior'], w), 1e-12)
        # design half with the same uniforms
        eng.rng = np.random.default_rng(100 + step)
        g = np.random.default_rng(100 + step)
        draws, _ = orc.randdraw(inp['prior'], w, g.random(sc['n_draws']))
        var_p, _ = orc.yvar_from_draws(model, orc.make_allsettings(inp['setting_values']), draws, inp['cons'], nch)
        if sc['kind'] == 'noise':
            var_n = orc.noise_var_noise_parameter(inp['prior'], w, sc['noise_parameter_index'])
        else:
            var_n = orc.noise_var_default(sc['default_noise_std'], nch)
        want_u = orc.utility_variance(var_p, var_n)
        got = eng.opt_setting()
        assert_allclose(eng._utility_dev.cpu().numpy(), want_u, rtol=1e-12)
        assert eng.last_setting_index == orc.opt_index(want_u)
        assert got == tuple(orc.make_allsettings(inp['setting_values'])[:, orc.opt_index(want_u)])
        rec = (got,) + tuple(rec[1:])                                # measure where the engine says


# =================================================================================================
# c4 at its full size: 1e8 particles built on the device; host slices checked against the oracle
# =================================================================================================
def test_c4_full_size_against_oracle_slices(obe, torch):
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    n = 100_000_000
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip('needs ~12 GB of device memory')
    gen = torch.Generator(device='cuda')
    gen.manual_seed(1001)
    prior = torch.empty((3, n), dtype=torch.float64, device='cuda')
    prior[0] = 2 + 2 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
    prior[1] = -2000 + 1600 * torch.rand(n, generator=gen, dtype=torch.float64, device='cuda')
    prior[2] = 50000 + 1000 * torch.randn(n, generator=gen, dtype=torch.float64, device='cuda')
    eng = obe.OptBayesExpt('lorentzian_hwhm', (np.linspace(1.5, 4.5, 100000),), prior, (0.1,), scale=False,
                           default_noise_std=500.0, seed=7)
    del prior
    eng.tuning_parameters['auto_resample'] = False
    rec = ((3.1,), 49600.0, 500.0)
    eng.pdf_update(rec)
    # (a) the un-normalised device weights of a 1e6 prefix and of a 1e6 random slice against the oracle likelihood
    #     (prior weight exactly 1/n; the normaliser is checked separately against numpy's pairwise sum of the row)
    t_dev = eng.weights_dev
    sel = torch.randint(0, n, (1_000_000,), device='cuda', generator=gen)
    for idx in (torch.arange(1_000_000, device='cuda'), sel):
        p = eng.particles_dev[:, idx].cpu().numpy()
        lik = orc.likelihood_known_sigma((orc.model_lorentzian_hwhm(rec[0], p, (0.1,)),), rec[1], rec[2])
        want_t = np.nan_to_num((1.0 / n) * lik)
        wclose(t_dev[idx].cpu().numpy(), want_t)
    t_host = t_dev.cpu().numpy()
    assert_allclose(float(eng._buf.stats[_lib.ST_TOTAL].item()), np.sum(t_host), rtol=1e-12)
    assert_allclose(eng.n_eff(), np.sum(t_host) ** 2 / np.sum(t_host * t_host), rtol=1e-12)
    w_host = t_host / np.sum(t_host)
    del t_host
    # (b) the ancestors contract at full size with the cluster plan active (no option override): monotone, every
    #     particle's offspring count within 1 of n*w, and on sampled tiles bit-equal to searchsorted on the GPU's own
    #     materialised CDF
    cdf = torch.empty(n, dtype=torch.float64, device='cuda')
    _lib.check(lib.obe_cdf(eng._cs(), C.c_void_p(cdf.data_ptr()), eng._stream()))
    alt = eng._buf.empty_like()
    idx = torch.empty(n, dtype=torch.int64, device='cuda')
    u0 = 0.4375
    _lib.check(lib.obe_resample_systematic(eng._cs(), C.byref(alt.struct()), u0, None, None, 99, 1, 0.98, 0,
                                           C.c_void_p(idx.data_ptr()), None, eng._stream()))
    assert bool((idx[1:] >= idx[:-1]).all()) and int(idx[0]) >= 0 and int(idx[-1]) < n
    counts = torch.bincount(idx, minlength=n).to(torch.float64)
    assert float((counts - n * torch.from_numpy(w_host).cuda()).abs().max()) < 1.0 + 1e-6
    rng = np.random.default_rng(3)
    tiles = np.concatenate([[0, 1, n // 2048 - 1, (n - 1) // 2048], rng.integers(2, n // 2048 - 2, 60)])
    inv_n = 1.0 / np.float64(n)
    for k in tiles:
        a, b = int(k) * 2048, min(n, (int(k) + 1) * 2048)
        c_host = cdf[max(a - 1, 0):b].cpu().numpy()
        c_lo = c_host[0] if a > 0 else 0.0
        seg = c_host[1:] if a > 0 else c_host
        i_lo = max(0, int(np.floor(c_lo * n)) - 3)
        i_hi = min(n, int(np.ceil(seg[-1] * n)) + 3)
        teeth_i = np.arange(i_lo, i_hi)
        u = (teeth_i.astype(np.float64) + np.float64(u0)) * inv_n
        keep = (u >= c_lo) & (u < seg[-1]) if b < n else (u >= c_lo)
        want = a + np.minimum(np.searchsorted(seg, u[keep], side='right'), b - a - 1)
        got = idx[torch.from_numpy(teeth_i[keep]).cuda()].cpu().numpy()
        assert_array_equal(got, want, err_msg=f'tile {k}')


# =================================================================================================
# the device normals (Philox4x32-10 + float32 Box-Muller on 23-bit uniforms) as a distribution
# =================================================================================================
def test_device_normals_distribution(obe, torch):
    """1.2e7 normals straight out of the resample kernel (z_out): Kolmogorov-Smirnov against N(0,1), moments up to
    the fourth, tail counts at 3/4/5 sigma, no value beyond the 23-bit Box-Muller bound sqrt(-2 ln 2^-24) = 5.77,
    independence of the d coordinates of a slot."""
    from scipy import stats
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    n, d = 4_000_000, 3
    prior = torch.randn((d, n), dtype=torch.float64, device='cuda')
    pdf = obe.ParticlePDF(prior, scale=False, resampling='systematic', seed=3)
    pdf._ensure_moments()
    alt = pdf._buf.empty_like()
    z = torch.empty((n, d), dtype=torch.float64, device='cuda')
    _lib.check(lib.obe_resample_systematic(pdf._cs(), C.byref(alt.struct()), 0.3, None, None, 20261017, 5, 0.98, 0,
                                           None, C.c_void_p(z.data_ptr()), pdf._stream()))
    zh = z.cpu().numpy()
    flat = zh.reshape(-1)
    m = flat.size
    ks = stats.kstest(flat, 'norm')
    assert ks.statistic < 1.95 / np.sqrt(m), f'KS statistic {ks.statistic:.3g} (99.9 % bound {1.95 / np.sqrt(m):.3g})'
    assert abs(flat.mean()) < 5 / np.sqrt(m)
    assert abs(flat.var() - 1) < 5 * np.sqrt(2 / m)
    assert abs(stats.skew(flat)) < 5 * np.sqrt(6 / m)
    assert abs(stats.kurtosis(flat)) < 5 * np.sqrt(24 / m)          # excess kurtosis
    for thr in (3.0, 4.0, 5.0):
        expect = m * 2 * stats.norm.sf(thr)
        got = int((np.abs(flat) > thr).sum())
        assert abs(got - expect) < 5 * np.sqrt(expect) + 3, f'|z| > {thr}: {got} vs {expect:.1f}'
    assert np.abs(flat).max() <= np.sqrt(-2 * np.log(2.0 ** -24)) + 1e-3
    corr = np.corrcoef(zh.T)
    assert np.abs(corr - np.eye(d)).max() < 5 / np.sqrt(n)
    # neighbouring slots are independent too
    assert abs(np.corrcoef(zh[:-1, 0], zh[1:, 0])[0, 1]) < 5 / np.sqrt(n)


# =================================================================================================
# host-side additions of round 2
# =================================================================================================
def test_utility_methods_dispatch_by_name(obe):
    """utility_variance / utility_max_min / utility_pseudo / utility_full_kld are four methods
    (obe_base.py:579-720), whatever utility_method the engine was built with."""
    sc = by_name('c1_find_peak')
    inp = build_inputs(sc, 5000)
    engines = {m: obe.OptBayesExpt('lorentzian_hwhm', inp['setting_values'], inp['prior'], inp['cons'], scale=False,
                                   default_noise_std=500.0, utility_method=m, seed=1)
               for m in ('variance_approx', 'max_min', 'pseudo_utility')}
    base = engines['variance_approx']
    for name, call in (('variance_approx', 'utility_variance'), ('max_min', 'utility_max_min'),
                       ('pseudo_utility', 'utility_pseudo')):
        base.rng = np.random.default_rng(9)
        got = getattr(base, call)()
        engines[name].rng = np.random.default_rng(9)                # (may be the same object as base)
        assert_array_equal(got, engines[name].utility())
    base.rng = np.random.default_rng(9)
    u_var = base.utility()                                           # the constructor's choice is untouched
    base.rng = np.random.default_rng(9)
    assert_array_equal(u_var, base.utility_variance())
    base.rng = np.random.default_rng(9)
    assert not np.array_equal(u_var, base.utility_max_min())


def test_multinomial_device_resample(obe, torch):
    """The reference's algorithm with device-made randomness: ancestors == searchsorted(cdf_gpu, u_gpu), offspring
    = ancestors + Liu-West jitter with the device Cholesky factor (mean preserved, covariance inflated by 2 - a^2)."""
    from optbayesexpt_b200 import _lib
    n, d = 300_001, 3
    g = np.random.default_rng(4)
    prior = np.array([g.uniform(2, 4, n), g.uniform(-2000, -400, n), g.normal(5e4, 1e3, n)])
    w = g.random(n) ** 3
    w /= w.sum()
    pdf = obe.ParticlePDF(prior, scale=False, resampling='multinomial_device', seed=8)
    pdf.particle_weights = w
    mean0, cov0 = pdf.mean(), pdf.covariance()
    cdf = torch.empty(n, dtype=torch.float64, device='cuda')
    _lib.check(_lib.load().obe_cdf(pdf._cs(), C.c_void_p(cdf.data_ptr()), pdf._stream()))
    cdf_h = cdf.cpu().numpy()
    pdf.resample()
    u, _, idx = pdf._multinomial_bufs
    assert_array_equal(idx.cpu().numpy(), orc.search_cdf(cdf_h, u.cpu().numpy()))
    assert_array_equal(pdf.particle_weights, np.full(n, 1.0 / n))
    sd = np.sqrt(np.diag(cov0))
    assert np.all(np.abs(pdf.mean() - mean0) < 6 * sd / np.sqrt(n))
    assert_allclose(np.diag(pdf.covariance()) / np.diag(cov0), 1 + (1 - 0.98 ** 2), rtol=2e-2)


def test_async_cycle_enforces_noise_constraint(obe):
    """run_cycle_async on a noise-parameter engine applies the positivity constraint after its resample, like
    pdf_update does (obe_base.py:396-397, obe_noiseparam.py:57-79) -- without synchronising."""
    sc = by_name('c2_line_noise')
    inp = build_inputs(sc, 60000)
    eng = obe.OptBayesExptNoiseParameter('line', inp['setting_values'], inp['prior'], (), noise_parameter_index=2,
                                         scale=False, a_param=0.5, seed=2)       # a big nudge: some sigma go negative
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        eng.run_cycle_async(((0.4,), 0.1))
    p, w = eng.particles, eng.particle_weights
    bad = p[2] <= 0
    assert bad.sum() > 0, 'the constraint never bit: the test does not exercise it'
    assert np.all(w[bad] == 0.0) and np.all(w[~bad] > 0.0)
    assert abs(w.sum() - 1.0) < 1e-12
    assert 0 <= int(eng.best_index_dev.cpu()[0]) < len(eng.setting_indices)


def test_sweeper_set_pdf_with_a_larger_cloud(obe):
    """the fused sweep's cached weight row and noise scale must follow set_pdf to a different size."""
    g = np.random.default_rng(1)

    def prior(n):
        return np.array([g.uniform(2, 4, n), g.uniform(400, 2000, n), g.normal(500, 1000, n), g.exponential(500, n)])
    xs = np.linspace(1.5, 4.5, 100)
    eng = obe.OptBayesExptSweeper('lorentzian_hwhm', (xs,), prior(3000), (0.1,), noise_parameter_index=3, scale=False,
                                  seed=1)
    ys = orc.model_lorentzian_hwhm((xs[10:20],), (3.2, 1500.0, 300.0), (0.1,)) + 300 * g.standard_normal(10)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        eng.pdf_update(((xs[10:20],), ys))
        eng.set_pdf(prior(50000))
        assert eng._multi_w is None and eng._sigma_ref is None
        eng.pdf_update(((xs[10:20],), ys))
    assert eng.n_particles == 50000
    assert abs(eng.particle_weights.sum() - 1.0) < 1e-12
    assert eng._multi_w is None or eng._multi_w.numel() == eng._buf.ld


def test_explicit_device_must_be_current(obe, torch):
    """device= that is not the current CUDA device raises instead of launching on the wrong GPU."""
    cur = torch.cuda.current_device()
    pdf = obe.ParticlePDF((np.arange(4.0),), device=f'cuda:{cur}')
    assert pdf._buf.device.index == cur
    with pytest.raises(ValueError):
        obe.ParticlePDF((np.arange(4.0),), device=f'cuda:{cur + 1}')
    with pytest.raises(ValueError):
        obe.ParticlePDF((np.arange(4.0),), device='cpu')
