"""The resample test on the device (obe_cycle, resample == 2; OptBayesExpt.device_resample_test).

pdf_update + the selection of the next opt_setting() are ONE C call: the update kernel's finishing block evaluates the
test of particlepdf.py:236-258 and writes stats[FIRED]; plan / pick / streaming resample run gated on it, the plain K
draws on its complement.  The host learns the outcome at the one synchronisation of the cycle.  Contract: every
decision, chosen setting and particle is identical to the synchronous path (host decision) given the same random
numbers, and every way of looking at the engine while the cycle is pending settles it first."""
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


class ScriptedRng:
    """random() -> the u0 of cycle t, random(k) -> the K uniforms of cycle t: the device path draws u0 every cycle, the
    host path only when it resamples, so the numbers are keyed by the cycle instead of by the order of consumption."""

    def __init__(self, seed, cycles, k):
        g = np.random.default_rng(seed)
        self.u0 = g.random(cycles + 1)
        self.uk = g.random((cycles + 1, k))
        self.t = 0

    def random(self, size=None):
        if size is None:
            return self.u0[self.t]
        return self.uk[self.t, :size].copy()


def _engine(obe, n, model='lorentzian_hwhm', n_set=400, thr=0.5, **kw):
    g = np.random.default_rng(3)
    prior = np.array([g.uniform(2, 4, n), g.uniform(-2000, -400, n), g.normal(50000, 1000, n)])
    settings = (np.linspace(1.5, 4.5, n_set),)
    return obe.OptBayesExpt(model, settings, prior, (0.1,), scale=False, default_noise_std=500.0, seed=11,
                            resample_threshold=thr, **kw)


def _measure(x, meas):
    return 50400.0 - 1200.0 / (((x[0] - 3.14) / 0.1) ** 2 + 1) + 500.0 * meas.standard_normal()


@pytest.mark.parametrize('n', [1000, 10_000, 50_000, 1_000_003])
def test_device_test_equals_the_host_decision(obe, n):
    cycles = 40
    a, b = _engine(obe, n), _engine(obe, n)
    a.eager_select = a.async_update = True                 # device-side test
    b.eager_select = True                                  # host decision (synchronous pdf_update)
    assert a._device_test_ok()
    a.rng, b.rng = ScriptedRng(5, cycles, 30), ScriptedRng(5, cycles, 30)
    meas = np.random.default_rng(9)
    fired = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        xa, xb = a.opt_setting(), b.opt_setting()
        for t in range(cycles):
            a.rng.t = b.rng.t = t + 1
            assert xa == xb, f'cycle {t}'
            y = _measure(xa, meas)
            a.pdf_update((xa, y, 500.0))
            assert a._pending_cycle, 'the device-test path did not run'
            b.pdf_update((xb, y, 500.0))
            if t % 3 == 0:                                   # sometimes look at the outcome before the selection
                assert a.just_resampled == b.just_resampled
                assert not a._pending_cycle
            xa, xb = a.opt_setting(), b.opt_setting()
            assert a.just_resampled == b.just_resampled, f'cycle {t}'
            assert a.last_setting_index == b.last_setting_index, f'cycle {t}'
            fired += int(a.just_resampled)
            if t % 10 == 7:
                np.testing.assert_array_equal(a.particle_weights, b.particle_weights)
    assert 2 <= fired < cycles, f'{fired} resamples in {cycles} cycles: the test needs both outcomes'
    np.testing.assert_array_equal(a.particles, b.particles)
    np.testing.assert_array_equal(a.particle_weights, b.particle_weights)
    np.testing.assert_array_equal(a.mean(), b.mean())
    np.testing.assert_array_equal(a.covariance(), b.covariance())


def test_pending_cycle_is_settled_by_any_look_at_the_engine(obe):
    n = 20_000
    looks = [lambda e: e.mean(), lambda e: e.std(), lambda e: e.covariance(), lambda e: e.n_eff(),
             lambda e: e.particles, lambda e: e.particle_weights, lambda e: e.particles_dev, lambda e: e.weights_dev,
             lambda e: e.just_resampled, lambda e: e.randdraw(5), lambda e: e.good_setting(),
             lambda e: e.utility(), lambda e: e.resample(), lambda e: e.pdf_update(((3.0,), 49800.0, 500.0))]
    for i, look in enumerate(looks):
        for thr in (0.05, 1.0):                            # the test does not fire / fires
            a, b = _engine(obe, n, thr=thr), _engine(obe, n, thr=thr)
            a.eager_select = a.async_update = True
            b.eager_select = True
            a.rng, b.rng = ScriptedRng(5, 4, 30), ScriptedRng(5, 4, 30)
            with warnings.catch_warnings():
                warnings.simplefilter('ignore', RuntimeWarning)
                a.opt_setting(), b.opt_setting()
                a.rng.t = b.rng.t = 1
                a.pdf_update(((3.1,), 49700.0, 500.0))
                b.pdf_update(((3.1,), 49700.0, 500.0))
                assert a._pending_cycle
                ra, rb = look(a), look(b)             # (same cycle key: both paths see the same u0 / K uniforms)
                assert not a._pending_cycle or i == len(looks) - 1     # (the last look starts another cycle)
                assert a.just_resampled == b.just_resampled
                if isinstance(ra, np.ndarray):
                    np.testing.assert_array_equal(ra, rb)
                np.testing.assert_array_equal(a.particles, b.particles)
                np.testing.assert_array_equal(a.particle_weights, b.particle_weights)


def test_device_test_not_used_where_it_does_not_apply(obe):
    n = 5000
    g = np.random.default_rng(2)
    # sigma as a parameter: constraint masks follow the resample -> host decision
    prior = np.array([g.uniform(-1, 1, n), g.uniform(0, 2, n), g.uniform(0.1, 2, n)])
    e = obe.OptBayesExptNoiseParameter('line', (np.linspace(-1, 1, 50),), prior, (), noise_parameter_index=2, scale=False)
    e.eager_select = e.async_update = True
    assert not e._device_test_ok()
    # multinomial resampling, a forced decision, the switch
    e2 = _engine(obe, n, resampling='multinomial')
    e2.eager_select = e2.async_update = True
    assert not e2._device_test_ok()
    e3 = _engine(obe, n)
    e3.eager_select = e3.async_update = True
    e3.device_resample_test = False
    assert not e3._device_test_ok()
    x = e3.opt_setting()
    e3.pdf_update((x, 49900.0, 500.0))
    assert not e3._pending_cycle


def test_closed_loop_converges_with_the_device_test(obe):
    n = 100_000
    e = _engine(obe, n)
    e.eager_select = e.async_update = True
    meas = np.random.default_rng(4)
    fired = 0
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        x = e.opt_setting()
        for _ in range(300):
            e.pdf_update((x, _measure(x, meas), 500.0))
            x = e.opt_setting()
            fired += int(e.just_resampled)
    assert 5 <= fired <= 150
    m, s = e.mean(), e.std()
    assert abs(m[0] - 3.14) < 5 * s[0] + 1e-3 and s[0] < 0.02
    assert abs(m[1] + 1200.0) < 5 * s[1] and abs(m[2] - 50400.0) < 5 * s[2]
    assert abs(e.particle_weights.sum() - 1.0) < 1e-12


@pytest.mark.parametrize('variant', ['copy_engine', 'single_call', 'one_stream', 'stream_sync'])
@pytest.mark.parametrize('thr', [0.5, 2.0], ids=['natural', 'forced'])
def test_result_delivery_variants_agree(obe, variant, thr):
    """The kernels store stats + argmax into the pinned host block themselves and the cycle call is split in two phases
    (defaults); asynchronous D2H copies, one whole call and the early order on one stream must give the same trajectory."""
    from optbayesexpt_b200 import _lib
    lib = _lib.load()
    n, cycles = 30_000, 25
    a, b = _engine(obe, n, thr=thr), _engine(obe, n, thr=thr)
    for e in (a, b):
        e.eager_select = e.async_update = True
        e.rng = ScriptedRng(5, cycles, 30)
    if variant == 'copy_engine':
        b.poll_results = False                  # (no completion word without the kernels' own stores)
    elif variant == 'single_call':
        b.split_cycle = False
    elif variant == 'one_stream':
        b.two_stream_min_particles = 10 ** 12
    elif variant == 'stream_sync':
        b.poll_results = False
    meas = np.random.default_rng(9)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            xa, xb = a.opt_setting(), b.opt_setting()
            for t in range(cycles):
                a.rng.t = b.rng.t = t + 1
                assert xa == xb, f'cycle {t}'
                y = _measure(xa, meas)
                a.pdf_update((xa, y, 500.0))
                if variant == 'copy_engine':
                    lib.obe_set_option(b'zero_copy_out', 0)
                b.pdf_update((xb, y, 500.0))
                lib.obe_set_option(b'zero_copy_out', 1)
                xa, xb = a.opt_setting(), b.opt_setting()
                assert a.just_resampled == b.just_resampled
                assert a.last_setting_index == b.last_setting_index
    finally:
        lib.obe_set_option(b'zero_copy_out', 1)
    np.testing.assert_array_equal(a.particles, b.particles)
    np.testing.assert_array_equal(a.particle_weights, b.particle_weights)
    np.testing.assert_array_equal(a.mean(), b.mean())
