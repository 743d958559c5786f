"""GPU tests of the batched engine (BASELINE config c5): instance b of a BatchedOptBayesExpt must
reproduce a single engine (itself parity-tested against the oracle / the reference goldens) that is
fed the same uniforms -- every chosen setting and resample decision identical, moments and weights to
1e-9 -- for the lock-in model (2 channels, noise parameter, positivity constraints, sticky cost) and
for the known-sigma Lorentzian."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import obe_oracle as orc
from oracle.scenarios import build_inputs, by_name

pytestmark = pytest.mark.gpu


def _run_pair(obe, sc, B, n, cycles, make_single, batched_kwargs, meas):
    from optbayesexpt_b200.batched import BatchedOptBayesExpt
    inp = build_inputs(sc, n)
    model = orc.MODELS[sc['model']][0]
    nch = orc.MODELS[sc['model']][4]
    rng = np.random.default_rng(7)
    priors = np.stack([sc['prior'](np.random.default_rng(100 + b), n) for b in range(B)])
    truths = [tuple(np.asarray(sc['true_pars']) * (1 + 0.05 * rng.standard_normal(len(sc['true_pars'])))) for _ in range(B)]
    beng = BatchedOptBayesExpt(sc['model'], inp['setting_values'], priors, inp['cons'], n_draws=30, scale=False,
                               seed=4321, **batched_kwargs)
    singles = []
    for b in range(B):
        e = make_single(priors[b], inp)
        e.rng = orc.ReplayRng(4321, b, 30)
        e._philox_seed = beng._seed_normal + b
        e._epoch = 0
        singles.append(e)
    n_res = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for t in range(cycles):
            idx, settings = beng.opt_setting()
            for b, e in enumerate(singles):
                e.opt_setting()
                assert e.last_setting_index == idx[b], f'cycle {t} instance {b}: setting index differs'
            ys, sig = meas(model, settings, truths, inp, rng)
            beng.pdf_update(ys, sigma=sig)
            flags = beng.just_resampled
            means, stds, neff = beng.mean(), beng.std(), beng.n_eff()
            for b, e in enumerate(singles):
                rec_y = tuple(ys[b]) if nch > 1 else float(ys[b, 0])
                e.pdf_update((tuple(settings[:, b]), rec_y, None if sig is None else float(sig)))
                assert bool(e.just_resampled) == bool(flags[b]), f'cycle {t} instance {b}: resample decision differs'
                np.testing.assert_allclose(means[b], e.mean(), rtol=1e-9, err_msg=f'mean {t} {b}')
                np.testing.assert_allclose(stds[b], e.std(), rtol=1e-7, err_msg=f'std {t} {b}')
                np.testing.assert_allclose(neff[b], e.n_eff(), rtol=1e-9)
                e.rng.next_cycle()
            n_res += int(flags.sum())
    assert n_res > 0, 'the scenario never resampled: it does not exercise the batched resample'
    for b, e in enumerate(singles):
        w = e.particle_weights
        np.testing.assert_allclose(beng.particle_weights(b), w, rtol=1e-9, atol=1e-15 * w.max())
        p = e.particles
        spread = p.std(axis=1, keepdims=True)
        err = np.abs(beng.particles(b) - p) / (np.abs(p) * 1e-10 + spread * 1e-9)
        assert err.max() <= 1.0, f'instance {b}: particles differ ({err.max():.3g}x tol)'
    return beng


def test_batched_lockin_matches_single_engines():
    import optbayesexpt_b200 as obe
    sc = by_name('c5_lockin')

    class Lockin(obe.OptBayesExptNoiseParameter):       # demos/lockin/lockin_of_coil.py:107-153
        def enforce_parameter_constraints(self):
            self._apply_constraint_masks(mask_lt=(1 << self.n_dims) - 1)

        def cost_estimate(self):
            cost = np.ones_like(self.allsettings[0]) * 5.0
            cost[self.last_setting_index] = 1.0
            return cost

    def make_single(prior, inp):
        return Lockin('lockin_coil', inp['setting_values'], prior, inp['cons'], noise_parameter_index=(3, 3),
                      scale=False, n_draws=30)

    def meas(model, settings, truths, inp, rng):
        ys = np.stack([np.asarray(model((settings[0, b],), truths[b], inp['cons'])) for b in range(len(truths))])
        return ys + 5.0 * rng.standard_normal(ys.shape), None

    _run_pair(obe, sc, B=5, n=10000, cycles=25, make_single=make_single,
              batched_kwargs=dict(noise_parameter_index=(3, 3), constraint_lt=(0, 1, 2, 3), cost_of_changing_setting=5.0),
              meas=meas)


def test_batched_lorentzian_matches_single_engines():
    import optbayesexpt_b200 as obe
    sc = by_name('c1_find_peak')

    def make_single(prior, inp):
        return obe.OptBayesExpt('lorentzian_hwhm', inp['setting_values'], prior, inp['cons'], scale=False, n_draws=30,
                                default_noise_std=500.0)

    def meas(model, settings, truths, inp, rng):
        ys = np.array([[model((settings[0, b],), truths[b], inp['cons'])] for b in range(len(truths))])
        return ys + 500.0 * rng.standard_normal(ys.shape), 500.0

    # n = 5000 is not a whole number of tiles: exercises the padding of the batched layout
    _run_pair(obe, sc, B=4, n=5000, cycles=30, make_single=make_single,
              batched_kwargs=dict(default_noise_std=500.0), meas=meas)


def test_batched_full_config_c5_runs():
    """The BASELINE shape (scaled to 512 instances to keep the test short): closed loop on the device,
    every instance's posterior mean of R ends near its truth."""
    from optbayesexpt_b200.batched import BatchedOptBayesExpt
    sc = by_name('c5_lockin')
    B, n = 512, 10000
    inp = build_inputs(sc, n)
    rng = np.random.default_rng(3)
    import torch
    g = torch.Generator(device='cuda')
    g.manual_seed(5)
    prior = torch.empty((B, 4, n), dtype=torch.float64, device='cuda')
    for j, scale_ in enumerate((1e-3, 10.0, 1e-5, 10.0)):
        prior[:, j] = torch.empty((B, n), dtype=torch.float64, device='cuda').exponential_(1.0, generator=g) * scale_
    beng = BatchedOptBayesExpt('lockin_coil', inp['setting_values'], prior, (), noise_parameter_index=(3, 3),
                               constraint_lt=(0, 1, 2, 3), cost_of_changing_setting=5.0, scale=False, seed=11)
    truth = np.array(sc['true_pars'])
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for t in range(60):
            idx, settings = beng.opt_setting()
            z = orc.model_lockin_coil((settings[0],), truth, ())
            y = z.T + 5.0 * rng.standard_normal((B, 2))
            beng.pdf_update(y)
    mean, std = beng.mean(), beng.std()
    assert np.all(np.isfinite(mean))
    err = np.abs(mean[:, 1] - truth[1]) / std[:, 1]
    assert np.median(err) < 2.0 and np.mean(err < 5.0) > 0.95


# ---- on-device MeasurementSimulator (obe_utils.py:8-53), SURVEY 8(f) row 4 ------------------------------
def test_batched_simulator_matches_oracle():
    """The simulated records of one cycle against the numpy restatement: model value to 1e-12, the noise to the
    float32 accuracy of the device normals; known-sigma models also get sigma written."""
    from optbayesexpt_b200.batched import BatchedOptBayesExpt
    sc = by_name('c1_find_peak')
    B, n = 96, 2048
    inp = build_inputs(sc, n)
    rng = np.random.default_rng(8)
    prior = np.stack([sc['prior'](rng, n) for _ in range(B)])
    truths = np.stack([rng.uniform(2.2, 3.8, B), rng.uniform(-1800, -600, B), rng.normal(50000, 500, B)], axis=1)
    beng = BatchedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], prior, inp['cons'], scale=False,
                               default_noise_std=500.0, seed=4)
    beng.set_simulator(truths, 500.0, seed=99)
    for cycle in range(3):
        idx, settings = beng.opt_setting()
        rec = beng.simulate_measurement().cpu().numpy()
        want = orc.simulate_batch_measurement(orc.model_lorentzian_hwhm, settings.T, truths, inp['cons'], 500.0,
                                              99, beng.cycle, 1)
        assert_array_equal(rec[:, 0], settings[0])
        assert_allclose(rec[:, 4], want[:, 0], rtol=0, atol=500.0 * 1e-5)
        assert_array_equal(rec[:, 8], np.full(B, 500.0))
        noise_free = np.array([orc.model_lorentzian_hwhm((settings[0, b],), truths[b], inp['cons']) for b in range(B)])
        z = (rec[:, 4] - noise_free) / 500.0
        assert abs(z.mean()) < 0.5 and 0.6 < z.std() < 1.4            # it is noise, and of the right size
        beng._update_from_record(1, 1, False)                           # consume the simulated record
    # per-instance noise levels, two channels (lock-in): sigma is a particle coordinate, not written
    sc5 = by_name('c5_lockin')
    inp5 = build_inputs(sc5, 2048)
    prior5 = np.stack([sc5['prior'](rng, 2048) for _ in range(8)])
    b5 = BatchedOptBayesExpt('lockin_coil', inp5['setting_values'], prior5, (), noise_parameter_index=(3, 3),
                             constraint_lt=(0, 1, 2, 3), scale=False, seed=2)
    levels = np.linspace(1.0, 8.0, 8)
    truths5 = np.tile(np.array(sc5['true_pars']), (8, 1))
    b5.set_simulator(truths5, levels, seed=7)
    idx, settings = b5.opt_setting()
    rec = b5.simulate_measurement().cpu().numpy()
    want = orc.simulate_batch_measurement(orc.model_lockin_coil, settings.T, truths5, (), levels, 7, b5.cycle, 2)
    assert_allclose(rec[:, 4:6], want, rtol=1e-12, atol=8.0 * 1e-5)
    assert_array_equal(rec[:, 8:10], 0.0)


def test_batched_closed_loop_on_device_converges():
    """opt_setting -> simulated measurement -> pdf_update without a host round trip: every instance's
    posterior lands on its own truth."""
    from optbayesexpt_b200.batched import BatchedOptBayesExpt
    sc = by_name('c1_find_peak')
    B, n = 64, 10000
    inp = build_inputs(sc, n)
    rng = np.random.default_rng(18)
    prior = np.stack([sc['prior'](rng, n) for _ in range(B)])
    truths = np.stack([rng.uniform(2.3, 3.7, B), rng.uniform(-1800, -800, B), rng.normal(50000, 500, B)], axis=1)
    beng = BatchedOptBayesExpt('lorentzian_hwhm', inp['setting_values'], prior, inp['cons'], scale=False,
                               default_noise_std=500.0, seed=6)
    beng.set_simulator(truths, 500.0, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        for _ in range(150):
            beng.closed_loop_cycle()
    mean, std = beng.mean(), beng.std()
    err = np.abs(mean - truths) / std
    assert np.mean(err[:, 0] < 5.0) > 0.95, err[:, 0]
    assert np.median(std[:, 0]) < 0.02
