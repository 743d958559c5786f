"""Seeded closed-loop scenarios shared by the pinning script and the parity tests.

TEST INFRASTRUCTURE.  Inputs follow SURVEY.md 8(d): priors seeded 1001, measurement
noise 1002, engine rng 1003, "truth" fixed per scenario.  Shapes are the BASELINE
configs c1/c2/c3/c5 scaled so the CPU side finishes in seconds.
"""
import numpy as np


def _c1_prior(rng, n):
    # demos/find_peak/sequentialLorentzian.py:88-98
    return np.array([rng.uniform(2, 4, n), rng.uniform(-2000, -400, n), rng.normal(50000, 1000, n)])


def _c2_prior(rng, n):
    # demos/line_plus_noise/line_plus_noise.py:65-69
    return np.array([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.exponential(0.1, n)])


def _c3_prior(rng, n):
    # demos/pipulse/pipulse.py:69-75
    return np.array([rng.uniform(1, 5, n), rng.uniform(-7, 7, n)])


def _c5_prior(rng, n):
    # demos/lockin/lockin_of_coil.py:168-182
    return np.array([rng.exponential(1e-3, n), rng.exponential(10, n),
                     rng.exponential(1e-5, n), rng.exponential(10, n)])


SCENARIOS = [
    dict(name='c1_find_peak', kind='base', model='lorentzian_hwhm', n_particles=10000,
         prior=_c1_prior, settings=lambda: (np.linspace(1.5, 4.5, 200),), cons=(0.1,),
         true_pars=(3.14, -1200.0, 50400.0), noise=500.0, sigma_record='noise',
         n_cycles=60, selection='opt', n_draws=30, scale=False, a_param=0.98,
         resample_threshold=0.5, default_noise_std=500.0,
         seed_prior=1001, seed_meas=1002, seed_rng=1003),
    dict(name='c1_good_scale', kind='base', model='lorentzian_hwhm', n_particles=4000,
         prior=_c1_prior, settings=lambda: (np.linspace(1.5, 4.5, 200),), cons=(0.1,),
         true_pars=(2.71, -900.0, 49500.0), noise=500.0, sigma_record='noise',
         n_cycles=40, selection='good', pickiness=15, n_draws=30, scale=True, a_param=0.98,
         resample_threshold=0.5, default_noise_std=500.0, choke=0.8,
         seed_prior=2001, seed_meas=2002, seed_rng=2003),
    dict(name='c2_line_noise', kind='noise', model='line', n_particles=10000,
         prior=_c2_prior, settings=lambda: (np.linspace(0, 1, 101),), cons=(),
         true_pars=(0.7, -0.3), noise=0.2, sigma_record='noise', noise_parameter_index=2,
         n_cycles=40, selection='opt', n_draws=30, scale=False, a_param=0.98,
         resample_threshold=0.5, seed_prior=1001, seed_meas=1002, seed_rng=1003),
    dict(name='c3_pipulse', kind='base', model='rabi', n_particles=10000,
         prior=_c3_prior, settings=lambda: (np.linspace(0, 1, 101), np.linspace(-10, 10, 101)),
         cons=(100000.0, 0.01, 0.5), true_pars=(3.3, 1.7), noise='sqrt', sigma_record='sqrt',
         n_cycles=12, selection='opt', n_draws=30, scale=False, a_param=0.98,
         resample_threshold=0.5, default_noise_std=300.0,
         seed_prior=1001, seed_meas=1002, seed_rng=1003),
    dict(name='c5_lockin', kind='lockin', model='lockin_coil', n_particles=10000,
         prior=_c5_prior, settings=lambda: (2 * np.pi * np.logspace(2, 6, 200),), cons=(),
         true_pars=(1.2e-3, 8.0, 0.9e-5), noise=5.0, sigma_record='noise',
         noise_parameter_index=(3, 3), cost_of_changing_setting=5.0, traj_rtol=1e-7,
         n_cycles=40, selection='opt', n_draws=30, scale=False, a_param=0.98,
         resample_threshold=0.5, seed_prior=1001, seed_meas=1002, seed_rng=1003),
]


def _sweeper_prior(rng, n):
    # demos/sweeper/sweeper.py:60-69
    return np.array([rng.uniform(2, 4, n), rng.uniform(400, 2000, n), rng.normal(500, 1000, n),
                     rng.exponential(500, n)])


# demos/sweeper/sweeper.py: Lorentzian peak, unknown noise, settings are (start, stop) sweeps
SWEEPER = dict(name='sweeper_lorentzian', kind='sweeper', model='lorentzian_hwhm', n_particles=10000,
               prior=_sweeper_prior, settings=lambda: (np.linspace(1.5, 4.5, 100),), cons=(0.1,),
               true_pars=(3.2, 1500.0, 300.0), noise=300.0, noise_parameter_index=3,
               n_sweeps=10, n_draws=30, scale=False, a_param=0.98, resample_threshold=0.5,
               start_stop_subsample=3, cost_of_new_sweep=5.0, traj_rtol=1e-8,
               seed_prior=1001, seed_meas=1002, seed_rng=1003)


def by_name(name):
    if name == SWEEPER['name']:
        return SWEEPER
    for sc in SCENARIOS:
        if sc['name'] == name:
            return sc
    raise KeyError(name)


def build_inputs(sc, n_particles=None):
    n = n_particles or sc['n_particles']
    prior = sc['prior'](np.random.default_rng(sc['seed_prior']), n)
    return dict(prior=prior, setting_values=sc['settings'](), cons=sc['cons'])


def simulate_measurement(sc, model, setting, inp, meas_rng):
    """Noise-free model at the true parameters + Gaussian noise.  Returns (y[C], sigma[C])."""
    y_true = np.atleast_1d(np.asarray(model(setting, sc['true_pars'], inp['cons']), dtype=np.float64))
    if sc['noise'] == 'sqrt':  # counting noise, demos/pipulse/pipulse.py:165-170
        sig_true = np.sqrt(np.abs(y_true))
    else:
        sig_true = np.full_like(y_true, sc['noise'])
    y = y_true + sig_true * meas_rng.standard_normal(y_true.shape)
    if sc['sigma_record'] == 'sqrt':
        sig = np.sqrt(np.abs(y))
    else:
        sig = sig_true
    return y, sig
