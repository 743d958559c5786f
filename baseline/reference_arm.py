"""The CPU arm of bench.py: the UNMODIFIED reference (usnistgov/optbayesexpt v1.2.0) timed on the host.

`baseline/_ref/` holds a verbatim install of the reference's pure-Python package (baseline/install_reference.sh;
git-ignored, travels to the GPU box).  Nothing of this repo's engine is on this path: the classes are the
reference's own `OptBayesExpt` / `OptBayesExptNoiseParameter`, driven through their public API
(`opt_setting()` -> simulated measurement -> `pdf_update(record)`, obe_base.py:340-399, 733-756) exactly as the
reference's demos drive them, with the model functions of those demos as plain numpy callables.

What is measured, never assumed:
  * c1 / c2 / c3: the full workload at its real size (1e4 / 1e5 / 1e6 particles), closed loop;
  * c4 (1e8 particles x 1e5 settings): full cycles at 1e6 AND 1e7 particles (and at 1e8 when OBE_REF_FULL=1 and
    the host has the ~25 GB it needs); the rate at 1e8 is then the power-law extrapolation through the two
    measured sizes, labelled `extrapolated: true` with the fitted exponent next to the measured points.
numpy's elementwise kernels, cumsum and searchsorted are single-threaded; only np.cov / np.dot may use BLAS
threads, so the reference runs on ~1 core whatever the host offers.  `cores` reports that honestly.
"""
import os
import sys
import tempfile
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


# -------------------------------------------------------------------------------------------------
# the demo models of the reference, as the numpy callables its API takes: model(sets, pars, cons)
# -------------------------------------------------------------------------------------------------
def model_lorentzian(sets, pars, cons):
    """demos/find_peak/sequentialLorentzian.py:66-75"""
    x, = sets
    x0, a, b = pars[0], pars[1], pars[2]
    d, = cons
    return b + a / (((x - x0) / d) ** 2 + 1)


def model_line(sets, pars, cons):
    """demos/line_plus_noise/line_plus_noise.py:46-54 (sigma is a parameter the model ignores)"""
    x, = sets
    return pars[0] * x + pars[1]


def model_rabi(sets, pars, cons):
    """demos/pipulse/pipulse.py:16-49"""
    pulsetime, delta_f = sets
    b1, f_center = pars[0], pars[1]
    baseline, contrast, t1 = cons
    zz = ((delta_f - f_center) / b1) ** 2
    f_rabi = np.hypot(delta_f - f_center, b1)
    return baseline * (1 - np.exp(-pulsetime / t1) * contrast / 2 * (1 - np.cos(np.pi * 2 * f_rabi * pulsetime)) / (zz + 1))


def _prior_c1(rng, n):
    return np.array([rng.uniform(2, 4, n), rng.uniform(-2000, -400, n), rng.normal(50000, 1000, n)])


def _prior_c2(rng, n):
    return np.array([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.exponential(0.1, n)])


def _prior_c3(rng, n):
    return np.array([rng.uniform(1, 5, n), rng.uniform(-7, 7, n)])


# priors seeded 1001, measurement noise 1002, engine rng 1003
WORKLOADS = {
    'c1': dict(config_index=0, device_model='lorentzian_hwhm', model=model_lorentzian, kind='base', n_particles=10_000,
               prior=_prior_c1, settings=lambda: (np.linspace(1.5, 4.5, 200),), cons=(0.1,),
               truth=(3.14, -1200.0, 50400.0), noise=500.0, default_noise_std=500.0,
               label='demos/find_peak Lorentzian (x0, amplitude, background), 1e4 particles, 200 settings'),
    'c2': dict(config_index=1, device_model='line', model=model_line, kind='noise', n_particles=100_000,
               prior=_prior_c2, settings=lambda: (np.linspace(0, 1, 101),), cons=(), truth=(0.7, -0.3), noise=0.2,
               noise_parameter_index=2,
               label='demos/line_plus_noise with OptBayesExptNoiseParameter (unknown sigma), 1e5 particles, 101 settings'),
    'c3': dict(config_index=2, device_model='rabi', model=model_rabi, kind='base', n_particles=1_000_000,
               prior=_prior_c3, settings=lambda: (np.linspace(0, 1, 101), np.linspace(-10, 10, 101)),
               cons=(100000.0, 0.01, 0.5), truth=(3.3, 1.7), noise='sqrt', default_noise_std=300.0,
               label='demos/pipulse Rabi model, 2-D setting grid 101 x 101 (pulse length x detuning), 1e6 particles'),
    'c4': dict(config_index=3, device_model='lorentzian_hwhm', model=model_lorentzian, kind='base',
               n_particles=100_000_000, prior=_prior_c1, settings=lambda: (np.linspace(1.5, 4.5, 100000),), cons=(0.1,),
               truth=(3.14, -1200.0, 50400.0), noise=500.0, default_noise_std=500.0,
               label='synthetic Lorentzian scale-out: 1e8 particles x 1e5 settings, n_draws=30, d=3'),
}


def simulate(wl, setting, meas_rng):
    """One measurement record at `setting`: the model at the true parameters + Gaussian noise."""
    y_true = float(wl['model'](setting, wl['truth'], wl['cons']))
    if wl['noise'] == 'sqrt':                       # counting noise, demos/pipulse/pipulse.py:165-170
        y = y_true + np.sqrt(abs(y_true)) * meas_rng.standard_normal()
        return (tuple(setting), y, float(np.sqrt(abs(y))))
    y = y_true + wl['noise'] * meas_rng.standard_normal()
    if wl['kind'] == 'noise':
        return (tuple(setting), y)                  # obe_noiseparam.py:110: the record carries no sigma
    return (tuple(setting), y, wl['noise'])


# -------------------------------------------------------------------------------------------------
def import_reference():
    """The reference package, unmodified: baseline/_ref on the GPU box, /root/reference in the build container.
    Returns (module, where) or (None, reason)."""
    os.environ.setdefault('NUMBA_CACHE_DIR', os.path.join(tempfile.gettempdir(), 'numba_cache_obe_ref'))
    for path in (os.path.join(HERE, '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(path, 'optbayesexpt')):
            if path not in sys.path:
                sys.path.insert(0, path)
            warnings.simplefilter('ignore', SyntaxWarning)
            try:
                import optbayesexpt
                return optbayesexpt, path
            except Exception as exc:                # pragma: no cover
                return None, f'import failed: {exc!r}'
    return None, 'baseline/_ref missing: run baseline/install_reference.sh in the build container'


def make_engine(ref, wl, n_particles, forced, n_draws=30):
    prior = wl['prior'](np.random.default_rng(1001), n_particles)
    kw = dict(n_draws=n_draws, scale=False)
    if forced:
        kw['resample_threshold'] = 2.0              # N_eff/N < 2 always: a resample every cycle
    if wl['kind'] == 'noise':
        eng = ref.OptBayesExptNoiseParameter(wl['model'], wl['settings'](), prior, wl['cons'],
                                             noise_parameter_index=wl['noise_parameter_index'], **kw)
    else:
        eng = ref.OptBayesExpt(wl['model'], wl['settings'](), prior, wl['cons'],
                               default_noise_std=wl['default_noise_std'], **kw)
    eng.rng = np.random.default_rng(1003)
    return eng


def time_cycles(ref, wl, n_particles, forced, warmup, steps, budget_s, n_draws=30):
    """Closed-loop cycles of the reference; returns dict(n, cycles, s_per_cycle (mean), best, resamples_per_cycle)."""
    eng = make_engine(ref, wl, n_particles, forced, n_draws)
    meas = np.random.default_rng(1002)
    times, n_res = [], 0
    t_begin = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            x = eng.opt_setting()
            rec = simulate(wl, x, meas)
            eng.pdf_update(rec)
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
                n_res += 1 if eng.just_resampled else 0
            if it >= warmup and (time.perf_counter() - t_begin) > budget_s and len(times) >= 1:
                break
    del eng
    return dict(n=int(n_particles), cycles=len(times), s_per_cycle=float(np.mean(times)), best_s=float(np.min(times)),
                resamples_per_cycle=n_res / max(1, len(times)))


def blas_threads():
    try:
        import threadpoolctl
        return sum(i.get('num_threads', 0) for i in threadpoolctl.threadpool_info() if i.get('user_api') == 'blas')
    except Exception:                               # pragma: no cover
        return None


def reference_rate(workload, forced=True, warmup=2, steps=8, budget_s=25.0, n_draws=30, settings=None):
    """cpu_baseline object for `workload` (cycles/s of the unmodified reference on this host), or a dict with
    `unavailable` when the reference cannot be imported."""
    ref, where = import_reference()
    if ref is None:
        return {'unavailable': where}
    wl = dict(WORKLOADS[workload])
    if settings is not None and workload == 'c4':
        wl['settings'] = lambda: (np.linspace(1.5, 4.5, int(settings)),)
    n_full = wl['n_particles']
    note_threads = (f'numpy elementwise/cumsum/searchsorted are single-threaded (1 core used; BLAS threads available to '
                    f'np.cov/np.dot: {blas_threads()}; host has {os.cpu_count()} logical cores)')
    if workload != 'c4':
        m = time_cycles(ref, wl, n_full, forced, warmup, steps, budget_s, n_draws)
        return {'value': 1.0 / m['s_per_cycle'], 'unit': 'cycles/s', 'cores': 1, 'kind': 'reference',
                'sample': (f'unmodified reference ({where}), {wl["label"]}: the FULL workload, closed loop, '
                           f'{m["cycles"]} cycles after {warmup} warm-up, resample '
                           f'{"forced every cycle" if forced else "at its natural rate"} '
                           f'({m["resamples_per_cycle"]:.2f}/cycle); {note_threads}'),
                'measured': [m], 'extrapolated': False, 'ms_per_cycle': m['s_per_cycle'] * 1e3,
                'timed_s': m['s_per_cycle'] * m['cycles']}
    # c4: measure at 1e6 and 1e7, extrapolate to 1e8 (or measure 1e8 when asked and possible)
    pts = [time_cycles(ref, wl, 1_000_000, forced, warmup, steps, budget_s * 0.5, n_draws)]
    pts.append(time_cycles(ref, wl, 10_000_000, forced, 0, 2, budget_s * 0.5, n_draws))
    expo = float(np.log(pts[1]['s_per_cycle'] / pts[0]['s_per_cycle']) / np.log(pts[1]['n'] / pts[0]['n']))
    extrapolated = True
    s_full = pts[1]['s_per_cycle'] * (n_full / pts[1]['n']) ** expo
    if os.environ.get('OBE_REF_FULL') == '1':
        try:
            pts.append(time_cycles(ref, wl, n_full, forced, 0, 1, 1.0, n_draws))
            s_full, extrapolated = pts[-1]['s_per_cycle'], False
        except MemoryError:                         # pragma: no cover
            pass
    return {'value': 1.0 / s_full, 'unit': 'cycles/s', 'cores': 1, 'kind': 'reference',
            'sample': (f'unmodified reference ({where}), closed loop, multinomial resample (its only resampler) forced '
                       f'every cycle, {wl["settings"]()[0].size} settings: measured at '
                       + ', '.join(f'{p["n"]:.0e} particles ({p["cycles"]} cycles, {p["s_per_cycle"] * 1e3:.0f} ms/cycle)'
                                   for p in pts)
                       + (f'; rate at {n_full:.0e} particles EXTRAPOLATED with the fitted exponent {expo:.3f} '
                          f'(t ~ N^p through the two measured sizes)' if extrapolated else '; 1e8 measured')
                       + f'; {note_threads}'),
            'measured': pts, 'measured_n': [p['n'] for p in pts], 'extrapolated': extrapolated, 'exponent': expo,
            'ms_per_cycle': s_full * 1e3, 'timed_s': float(sum(p['s_per_cycle'] * p['cycles'] for p in pts))}
