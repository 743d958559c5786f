"""CPU oracle for the optbayesexpt hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This module is a plain-numpy restatement of the reference's particle-filter
inference + setting-selection loop (usnistgov/optbayesexpt v1.2.0).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline.  The product package ``optbayesexpt_b200`` never imports it.

Parity status: PINNED.  ``oracle/pin_against_reference.py`` runs the unmodified
reference (imported from a scratch copy of /root/reference, seeded through
``obe.rng``) next to this restatement and asserts equality; the golden vectors it
writes live in ``tests/golden/``.  ``tests/test_oracle.py`` replays every
known-answer assertion of the reference's own tests for this path
(tests/test_particlepdf.py, tests/test_optbayesexpt.py,
tests/test_zinference.py::test_infer) against this module.

All randomness is SUPPLIED (uniforms ``u``, standard normals ``z``), in the
order the reference consumes its ``Generator``:
  resample      : N uniforms, then N*d standard normals (particle-major)
  randdraw(K)   : K uniforms
  good_setting  : 1 uniform
The arithmetic lives in numpy (un-vendored dependency of the reference,
setup.py:29 pins no version; numpy 2.3.5 here).  The equivalences relied on are
checked in pin_against_reference.py:
  Generator.choice(a, size=M, p=w)  ==  searchsorted(cumsum(w)/cumsum(w)[-1], random(M), 'right')
  Generator.multivariate_normal(0, S, N) == standard_normal(N*d).reshape(N,d) @ (u*sqrt(s)).T, (u,s,vh)=svd(S)
"""
import numpy as np

TILE = 2048  # canonical tile of the B200 layout (only used by the *_tiled helpers)


# ----------------------------------------------------------------------------
# settings grid  (obe_base.py:174-180)
# ----------------------------------------------------------------------------
def make_allsettings(setting_values):
    """meshgrid(indexing='ij') flattened, first knob slowest (obe_base.py:174-176)."""
    grids = np.meshgrid(*setting_values, indexing='ij')
    return np.array([g.flatten() for g in grids], dtype=np.float64)


# ----------------------------------------------------------------------------
# likelihood  (obe_base.py:263-271, 418-461; obe_noiseparam.py:81-120)
# ----------------------------------------------------------------------------
def gauss_noise_likelihood(y_model, y_meas, sigma):
    """exp(-((ym-y)/sigma)**2/2)/sigma, no 1/sqrt(2 pi)  (obe_base.py:264-271)."""
    return np.exp(-((y_model - y_meas) / sigma) ** 2 / 2) / sigma


def likelihood_known_sigma(y_model, y_meas, sigma, choke=None):
    """Product over channels; zip() truncates to the shortest of
    (channels, y_meas, sigma)  (obe_base.py:451-461)."""
    lky = 1.0
    for y_m, y, s in zip(y_model, np.atleast_1d(y_meas), np.atleast_1d(sigma)):
        lky = lky * gauss_noise_likelihood(y_m, y, s)
    if choke is not None:
        return np.power(lky, choke)
    return lky


def likelihood_noise_parameter(y_model, y_meas, particles, noise_index, choke=None):
    """sigma_c is the particle coordinate noise_index[c]  (obe_noiseparam.py:110-120)."""
    lky = 1.0
    sigma = particles[np.atleast_1d(noise_index)]
    for y_m, y, s in zip(y_model, np.atleast_1d(y_meas), sigma):
        lky = lky * gauss_noise_likelihood(y_m, y, s)
    if choke is not None:
        return np.power(lky, choke)
    return lky


# ----------------------------------------------------------------------------
# Bayes update + resample test  (particlepdf.py:136-139, 216-258)
# ----------------------------------------------------------------------------
def normalized_product(weights, likelihood):
    """nan_to_num(nan_to_num(w*l)/sum)  (particlepdf.py:136-139)."""
    tmp = np.nan_to_num(weights * likelihood)
    return np.nan_to_num(tmp / np.sum(tmp))


def n_effective(weights):
    """1/sum(nan_to_num(w*w))  (particlepdf.py:243-244)."""
    return 1.0 / np.sum(np.nan_to_num(weights * weights))


def resample_decision(weights, resample_threshold):
    """Returns (do_resample, impoverished_warning)  (particlepdf.py:243-258)."""
    n = len(weights)
    n_eff = n_effective(weights)
    if n_eff < 0.1 * n:
        return True, True
    if n_eff / n < resample_threshold:
        return True, False
    return False, False


# ----------------------------------------------------------------------------
# moments  (particlepdf.py:173-214)
# ----------------------------------------------------------------------------
def weighted_mean(particles, weights):
    """np.average(axis=1, weights) == sum(x*w)/sum(w)  (particlepdf.py:182-183)."""
    return np.average(particles, axis=1, weights=weights)


def weighted_covariance(particles, weights):
    """np.cov(X, aweights=w)  (particlepdf.py:194): two-pass centred,
    (Xc*w) @ Xc.T / (V1 - V2/V1) with V1=sum(w), V2=sum(w^2); the normalisation is
    pinned by tests/test_particlepdf.py:80-90.  numpy itself is called so the SVD
    factor downstream sees the reference's bits; weighted_covariance_formula is the
    spelled-out form the CUDA kernel implements."""
    particles = np.atleast_2d(particles)
    d = particles.shape[0]
    return np.cov(particles, aweights=weights).reshape(d, d)


def weighted_covariance_formula(particles, weights):
    particles = np.atleast_2d(particles)
    v1 = weights.sum()
    v2 = (weights * weights).sum()
    avg = (particles * weights).sum(axis=1) / v1
    xc = particles - avg[:, None]
    c = (xc * weights) @ xc.T / (v1 - v2 / v1)
    return c.reshape(particles.shape[0], particles.shape[0])


def weighted_covariance_longdouble(particles, weights):
    """Same quantity in extended precision, for condition-aware tolerances."""
    p = np.asarray(particles, dtype=np.longdouble)
    w = np.asarray(weights, dtype=np.longdouble)
    v1 = w.sum()
    v2 = (w * w).sum()
    avg = (p * w).sum(axis=1) / v1
    xc = p - avg[:, None]
    return np.asarray((xc * w) @ xc.T / (v1 - v2 / v1), dtype=np.float64)


def std_biased(particles, weights):
    """sqrt(sum(w x^2) - (sum w x)^2) via np.dot  (particlepdf.py:209-214).
    One-pass, cancels catastrophically; see std_biased_longdouble."""
    var = np.zeros(particles.shape[0])
    for i, p in enumerate(particles):
        mean = np.dot(p, weights)
        msq = np.dot(p * p, weights)
        var[i] = msq - mean ** 2
    return np.sqrt(var)


def std_biased_longdouble(particles, weights):
    """The same biased estimator evaluated in extended precision."""
    p = np.asarray(particles, dtype=np.longdouble)
    w = np.asarray(weights, dtype=np.longdouble)
    mean = (p * w).sum(axis=1)
    msq = (p * p * w).sum(axis=1)
    return np.asarray(np.sqrt(msq - mean * mean), dtype=np.float64)


def std_centered(particles, weights):
    """sqrt(sum w (x - mean)^2) for normalised weights: algebraically the estimator of
    particlepdf.py:209-214, without its cancellation (what the CUDA path's pivot-shifted
    accumulators compute).  Use as the truth when E[x^2]/var is huge."""
    p = np.asarray(particles, dtype=np.longdouble)
    w = np.asarray(weights, dtype=np.longdouble)
    w = w / w.sum()
    mean = (p * w).sum(axis=1)
    return np.asarray(np.sqrt((w * (p - mean[:, None]) ** 2).sum(axis=1)), dtype=np.float64)


# ----------------------------------------------------------------------------
# weighted draws  (particlepdf.py:312-345  ==  Generator.choice(p=w))
# ----------------------------------------------------------------------------
def normalized_cdf(weights):
    """cumsum(w)/cumsum(w)[-1]: what Generator.choice builds from p."""
    cdf = np.cumsum(weights)
    cdf /= cdf[-1]
    return cdf


def choice_indices(weights, u):
    """Ancestor indices of Generator.choice(arange(N), size=len(u), p=w) given its
    uniforms u  (particlepdf.py:330-331)."""
    return np.searchsorted(normalized_cdf(weights), u, side='right')


def search_cdf(cdf_normalized, u):
    """idx = #{k : cdf[k] <= u}, clamped to N-1 (u == 1.0 cannot come out of
    Generator.random but can out of the systematic comb's rounding)."""
    idx = np.searchsorted(cdf_normalized, u, side='right')
    return np.minimum(idx, len(cdf_normalized) - 1)


def systematic_uniforms(u0, n):
    """The B200 fast-path comb: u_i = (i + u0) * (1/n), two IEEE ops, monotone in i.
    (Opt-in replacement for the N i.i.d. uniforms of particlepdf.py:330.)"""
    inv_n = 1.0 / np.float64(n)
    return (np.arange(n, dtype=np.float64) + np.float64(u0)) * inv_n


def randdraw(particles, weights, u):
    """(d, K) weighted draws  (particlepdf.py:312-345)."""
    idx = choice_indices(weights, u)
    return particles[:, idx], idx


# ----------------------------------------------------------------------------
# resample  (particlepdf.py:260-310)
# ----------------------------------------------------------------------------
def mvn_factor_svd(cov):
    """F with x = z @ F distributed N(0, cov), as Generator.multivariate_normal's
    default method='svd' builds it (particlepdf.py:300).  numpy 2.3.5 computes
    (u, s, vh) = svd(cov); x = z @ (u*sqrt(s)).T  -- verified BIT-equal in
    pin_against_reference.check_equivalences (the vh-based form differs in the
    last bits)."""
    (u, s, _) = np.linalg.svd(cov)
    return (u * np.sqrt(s)).T


def mvn_factor_cholesky(cov):
    """x = z @ L.T  (method='cholesky'); the B200 fast path's factor."""
    return np.linalg.cholesky(cov).T


def liu_west(coords, z, factor, a_param, scale, old_center):
    """nudged = coords + (z @ F).T ; optionally a*nudged + (1-a)*mean
    (particlepdf.py:296-307).  z is (N, d) particle-major."""
    nudged = coords + (z @ factor).T
    if scale:
        return nudged * a_param + old_center.reshape(-1, 1) * (1 - a_param)
    return nudged


def resample(particles, weights, u, z, a_param=0.98, scale=True, factor='svd'):
    """Full resample given its randomness (particlepdf.py:286-310).
    Returns (new_particles, new_weights, ancestor_idx)."""
    n = particles.shape[1]
    d = particles.shape[0]
    idx = choice_indices(weights, u)
    coords = particles[:, idx]
    covar = weighted_covariance(particles, weights)
    center = weighted_mean(particles, weights)
    newcovar = (1 - a_param ** 2) * covar
    f = mvn_factor_svd(newcovar) if factor == 'svd' else mvn_factor_cholesky(newcovar)
    new = liu_west(coords, np.asarray(z).reshape(n, d), f, a_param, scale, center)
    return new, np.full(n, 1.0 / n), idx


# ----------------------------------------------------------------------------
# constraints  (obe_noiseparam.py:57-79; demos/lockin/lockin_of_coil.py:115-133)
# ----------------------------------------------------------------------------
def enforce_noise_positive(particles, weights, noise_index):
    """w=0 where a noise parameter <= 0, then renormalise (obe_noiseparam.py:67-79)."""
    w = weights.copy()
    changed = False
    for param in particles[np.atleast_1d(noise_index)]:
        bad, = np.nonzero(param <= 0)
        if len(bad) > 0:
            changed = True
            w[bad] = 0
    if changed:
        w = w / np.sum(w)
    return w


def enforce_all_nonnegative(particles, weights):
    """w=0 where ANY parameter < 0, then renormalise (lockin_of_coil.py:120-133)."""
    w = weights.copy()
    changed = False
    for param in particles:
        bad = np.argwhere(param < 0).flatten()
        if len(bad) > 0:
            changed = True
            w[bad] = 0
    if changed:
        w = w / np.sum(w)
    return w


# ----------------------------------------------------------------------------
# utility + selection  (obe_base.py:463-489, 542-577, 628-655, 733-789)
# ----------------------------------------------------------------------------
def yvar_from_draws(model, allsettings, draws, cons, n_channels):
    """K model curves over the whole grid, population variance over draws
    (obe_base.py:480-488).  draws is (d, K)."""
    k = draws.shape[1]
    y_space = np.zeros((k, n_channels, allsettings.shape[1]))
    for i, oneparamset in enumerate(draws.T):
        y = model(allsettings, oneparamset, cons)
        y_space[i] = y if n_channels > 1 else (y,)
    return np.var(y_space, axis=0), y_space


def noise_var_default(default_noise_std, n_channels):
    """(C,1) constant (obe_base.py:228-229, 564)."""
    return (np.ones((n_channels, 1)) * default_noise_std) ** 2


def noise_var_noise_parameter(particles, weights, noise_index):
    """Weighted mean of sigma^2 per channel (obe_noiseparam.py:132-136)."""
    s2 = particles[np.atleast_1d(noise_index)] ** 2
    return np.average(s2, weights=weights, axis=1).reshape(-1, 1)


def utility_variance(var_p, var_n, cost=1.0, log_form=False):
    """sum_c(var_p/var_n)/cost -- LINEAR at this commit (obe_base.py:650-655);
    log_form=True gives the commented-out log(1+var/sigma^2) (obe_base.py:653)."""
    if log_form:
        return np.sum(np.log(1 + var_p / var_n), axis=0) / cost
    return np.sum(var_p / var_n, axis=0) / cost


def utility_pseudo(y_space, var_n, cost=1.0):
    """Entropy-equivalent variance utility (obe_base.py:491-518, 657-686): the differential entropy of
    the K model outputs (scipy.stats.differential_entropy, which the reference imports, obe_base.py:7-10)
    cast as the variance of a Gaussian with that entropy."""
    from scipy.stats import differential_entropy
    ent = differential_entropy(y_space, axis=0)
    var_p = np.exp(2 * ent) / (2 * np.pi * np.e)
    return np.sum(var_p / var_n, axis=0) / cost


def utility_full_kld(y_space, noisevalues):
    """exp(H[y + noise] - H[noise]) - 1 (obe_base.py:706-720); noisevalues is (K, C)."""
    from scipy.stats import differential_entropy
    y_entropy = differential_entropy(y_space + noisevalues[:, :, None], axis=0)
    n_entropy = differential_entropy(noisevalues, axis=0)
    return np.exp(y_entropy - n_entropy) - 1.0


def opt_index(utility):
    """np.argmax: first maximum (obe_base.py:748)."""
    return int(np.argmax(utility))


def good_index(utility, pickiness, u):
    """Draw with p = nan_to_num(U**pickiness)/sum given its one uniform
    (obe_base.py:781-785)."""
    p = np.asarray(utility, dtype=np.float64) ** pickiness
    p = np.nan_to_num(p)
    p = p / np.sum(p)
    return int(choice_indices(p, np.atleast_1d(u))[0])


def lockin_cost(n_settings, last_setting_index, cost_of_changing_setting):
    """Sticky cost vector (lockin_of_coil.py:135-152)."""
    cost = np.ones(n_settings) * cost_of_changing_setting
    cost[last_setting_index] = 1.0
    return cost


# ----------------------------------------------------------------------------
# demo model functions, (sets, pars, cons) broadcasting contract (obe_base.py:50-66)
# ----------------------------------------------------------------------------
def model_lorentzian_hwhm(sets, pars, cons):
    """b + a/(((x-x0)/d)**2+1)   demos/find_peak/sequentialLorentzian.py:66-75"""
    x, = sets
    x0, a, b = pars[:3]
    d, = cons
    return b + a / (((x - x0) / d) ** 2 + 1)


def model_lorentzian_fwhm(sets, pars, cons):
    """a/((2*(x-x0)/d)**2+1)+b   demos/numba/numbaLorentzian.py:104"""
    x, = sets
    x0, a, b = pars[:3]
    d, = cons
    return a / ((2 * (x - x0) / d) ** 2 + 1) + b


def model_lorentzian_4p(sets, pars, cons):
    """b + a/(((x-x0)*2/d)**2+1), linewidth d is a parameter
    demos/find_peak/seqLor_pdfevolve.py:28"""
    x, = sets
    x0, a, b, d = pars[:4]
    return b + a / (((x - x0) * 2 / d) ** 2 + 1)


def model_lorentzian_dip(sets, pars, cons):
    """1 - A/(((f-f0)*2/lw)**2+1)   demos/server/server_script.py:33"""
    f, = sets
    f0, A, lw = pars[:3]
    return 1 - A / (((f - f0) * 2 / lw) ** 2 + 1)


def model_line(sets, pars, cons):
    """m*x + b   demos/line_plus_noise/line_plus_noise.py:54"""
    x, = sets
    m, b = pars[:2]
    return m * x + b


def model_rabi(sets, pars, cons):
    """Rabi counts   demos/pipulse/pipulse.py:18-49"""
    pulsetime, delta_f = sets
    B1, f_center = pars[:2]
    baseline, contrast, T1 = cons
    zz = ((delta_f - f_center) / B1) ** 2
    f_rabi = np.hypot(delta_f - f_center, B1)
    return baseline * (1 - np.exp(-pulsetime / T1) * contrast / 2 *
                       (1 - np.cos(np.pi * 2 * f_rabi * pulsetime)) / (zz + 1))


def model_lockin_coil(sets, pars, cons):
    """(Re Z, Im Z) of R-L in parallel with C   demos/lockin/lockin_of_coil.py:63-102"""
    w, = sets
    L, R, C = pars[:3]
    y1 = 1 / (R + 1j * w * L)
    y2 = 1j * w * C
    z = 1 / (y1 + y2)
    return np.array((np.real(z), np.imag(z)))


MODELS = {
    # name: (fn, n_settings, n_params(min), n_cons, n_channels)
    'lorentzian_hwhm': (model_lorentzian_hwhm, 1, 3, 1, 1),
    'lorentzian_fwhm': (model_lorentzian_fwhm, 1, 3, 1, 1),
    'lorentzian_4p': (model_lorentzian_4p, 1, 4, 0, 1),
    'lorentzian_dip': (model_lorentzian_dip, 1, 3, 0, 1),
    'line': (model_line, 1, 2, 0, 1),
    'rabi': (model_rabi, 2, 2, 3, 1),
    'lockin_coil': (model_lockin_coil, 1, 3, 0, 2),
}


# ----------------------------------------------------------------------------
# Counter-based RNG of the B200 device path (NOT in the reference: this is the
# restatement of optbayesexpt_b200/csrc Philox4x32-10 + Box-Muller so the device
# RNG mode can be checked; the reference-parity mode uses numpy Generators).
# ----------------------------------------------------------------------------
_PH_M0 = np.uint64(0xD2511F53)
_PH_M1 = np.uint64(0xCD9E8D57)
_PH_W0 = np.uint32(0x9E3779B9)
_PH_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al. 2011). Inputs broadcastable uint32."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3)]
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _PH_M0 * c0.astype(np.uint64)
            p1 = _PH_M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_PH_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_PH_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def _u24(x):
    """uniform in (0,1) from the top 23 bits, exact in float32: ((x >> 9) + 0.5) * 2^-23."""
    return ((x >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)


def device_normals(n, d, seed, epoch):
    """(n, d) standard normals as the device jitter kernels draw them (csrc/obe_b200.cu
    device_normals): for output slot i and call c (4 dims per call) ctr=(i_lo, i_hi, c, epoch),
    key=(seed_lo, seed_hi); the four Philox words give two Box-Muller pairs evaluated in
    float32.  The device uses CUDA's logf/sqrtf/sincospif, so values agree to float32 rounding
    (~1e-6), not bit for bit: parity tests feed the kernel's own normals (z_out) to liu_west and
    check this restatement only to that accuracy."""
    i = np.arange(n, dtype=np.uint64)
    ilo = (i & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ihi = (i >> np.uint64(32)).astype(np.uint32)
    z = np.empty((n, d))
    for c in range((d + 3) // 4):
        x = philox4x32_10(ilo, ihi, np.uint32(c), np.uint32(epoch), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for h in range(2):
            if 4 * c + 2 * h < d:
                rad = np.sqrt(np.float32(-2.0) * np.log(_u24(x[2 * h])))
                ang = np.float32(2.0) * _u24(x[2 * h + 1]).astype(np.float64) * np.pi
                z[:, 4 * c + 2 * h] = (rad * np.cos(ang).astype(np.float32)).astype(np.float64)
                if 4 * c + 2 * h + 1 < d:
                    z[:, 4 * c + 2 * h + 1] = (rad * np.sin(ang).astype(np.float32)).astype(np.float64)
    return z


def device_normals_packed(n, d, seed, epoch, slot_begin=0):
    """(n, d) standard normals of output slots slot_begin .. slot_begin+n-1 as the STREAMING resample kernels draw
    them (csrc/obe_b200.cu packed_normals4): the jitter normals of the whole cloud are one sequence, normal
    m = d*slot + j, four per Philox call: ctr = (q_lo, q_hi, 0x80000000 | d, epoch) with q = m >> 2, word m & 3;
    words (0,1) and (2,3) are two float32 Box-Muller pairs (cos, sin).  Same accuracy caveat as device_normals."""
    m = np.arange(n * d, dtype=np.uint64) + np.uint64(d) * np.uint64(slot_begin)
    q = m >> np.uint64(2)
    uq, inv = np.unique(q, return_inverse=True)
    x = philox4x32_10((uq & np.uint64(0xFFFFFFFF)).astype(np.uint32), (uq >> np.uint64(32)).astype(np.uint32),
                      np.uint32(0x80000000 | d), np.uint32(epoch), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    four = np.empty((len(uq), 4))
    for h in range(2):
        rad = np.sqrt(np.float32(-2.0) * np.log(_u24(x[2 * h])))
        ang = np.float32(2.0) * _u24(x[2 * h + 1]).astype(np.float64) * np.pi
        four[:, 2 * h] = (rad * np.cos(ang).astype(np.float32)).astype(np.float64)
        four[:, 2 * h + 1] = (rad * np.sin(ang).astype(np.float32)).astype(np.float64)
    return four[inv, (m & np.uint64(3)).astype(np.int64)].reshape(n, d)


# ----------------------------------------------------------------------------
# A whole engine, consuming a numpy Generator in the reference's order.
# Used for closed-loop parity and as bench.py's CPU baseline ("port").
# ----------------------------------------------------------------------------
class OracleOBE:
    """Restatement of OptBayesExpt / OptBayesExptNoiseParameter orchestration
    (obe_base.py:340-399, 733-789; obe_noiseparam.py) over the functions above."""

    def __init__(self, model, setting_values, parameter_samples, constants,
                 n_channels=1, n_draws=30, choke=None, pickiness=15,
                 default_noise_std=1.0, a_param=0.98, resample_threshold=0.5,
                 auto_resample=True, scale=True, noise_parameter_index=None,
                 nonneg_constraint=False, cost_of_changing_setting=None, rng=None,
                 resampling='multinomial'):
        self.model = model
        self.allsettings = make_allsettings(setting_values)
        self.particles = np.array(parameter_samples, dtype=np.float64)
        self.n_dims, self.n_particles = self.particles.shape
        self.particle_weights = np.ones(self.n_particles) / self.n_particles
        self.cons = constants
        self.n_channels = n_channels
        self.N_DRAWS = n_draws
        self.choke = choke
        self.pickiness = pickiness
        self.default_noise_std = default_noise_std
        self.tuning_parameters = dict(a_param=a_param, resample_threshold=resample_threshold,
                                      auto_resample=auto_resample, scale=scale)
        self.noise_parameter_index = (None if noise_parameter_index is None
                                      else np.atleast_1d(noise_parameter_index))
        self.nonneg_constraint = nonneg_constraint
        self.cost_of_changing_setting = cost_of_changing_setting
        self.rng = rng if rng is not None else np.random.default_rng()
        self.resampling = resampling
        self.just_resampled = False
        self.last_setting_index = 0
        self.last_ancestors = None
        self.last_utility = None

    # -- model in both orientations (obe_base.py:320, 338)
    def eval_over_all_parameters(self, onesetting):
        y = self.model(onesetting, self.particles, self.cons)
        return y if self.n_channels > 1 else (y,)

    def pdf_update(self, record, forced_resample=False):
        y_model = self.eval_over_all_parameters(record[0])
        if self.noise_parameter_index is not None:
            lik = likelihood_noise_parameter(y_model, record[1], self.particles,
                                             self.noise_parameter_index, self.choke)
        else:
            lik = likelihood_known_sigma(y_model, record[1], record[2], self.choke)
        self.particle_weights = normalized_product(self.particle_weights, lik)
        self.just_resampled = False
        if self.tuning_parameters['auto_resample'] or forced_resample:
            do, _ = resample_decision(self.particle_weights,
                                      self.tuning_parameters['resample_threshold'])
            if do or forced_resample:
                self.resample()
                self.just_resampled = True
        if self.just_resampled:
            self.enforce_parameter_constraints()
        return self.particles, self.particle_weights

    def resample(self):
        n, d = self.n_particles, self.n_dims
        if self.resampling == 'systematic':
            u = systematic_uniforms(self.rng.random(), n)
        else:
            u = self.rng.random(n)
        # covariance/mean of the OLD cloud, then N*d normals  (particlepdf.py:286-301)
        z = self.rng.standard_normal(n * d).reshape(n, d)
        self.particles, self.particle_weights, self.last_ancestors = resample(
            self.particles, self.particle_weights, u, z,
            self.tuning_parameters['a_param'], self.tuning_parameters['scale'])

    def enforce_parameter_constraints(self):
        if self.nonneg_constraint:
            self.particle_weights = enforce_all_nonnegative(self.particles, self.particle_weights)
        elif self.noise_parameter_index is not None:
            self.particle_weights = enforce_noise_positive(self.particles, self.particle_weights,
                                                           self.noise_parameter_index)

    def utility(self):
        draws, _ = randdraw(self.particles, self.particle_weights, self.rng.random(self.N_DRAWS))
        var_p, _ = yvar_from_draws(self.model, self.allsettings, draws, self.cons, self.n_channels)
        if self.noise_parameter_index is not None:
            var_n = noise_var_noise_parameter(self.particles, self.particle_weights,
                                              self.noise_parameter_index)
        else:
            var_n = noise_var_default(self.default_noise_std, self.n_channels)
        cost = 1.0
        if self.cost_of_changing_setting is not None:
            cost = lockin_cost(self.allsettings.shape[1], self.last_setting_index,
                               self.cost_of_changing_setting)
        self.last_utility = utility_variance(var_p, var_n, cost)
        return self.last_utility

    def opt_setting(self):
        best = opt_index(self.utility())
        self.last_setting_index = best
        return tuple(self.allsettings[:, best])

    def good_setting(self, pickiness=None):
        pk = self.pickiness if pickiness is None else pickiness
        util = self.utility()
        best = good_index(util, pk, self.rng.random())
        self.last_setting_index = best
        return tuple(self.allsettings[:, best])

    def mean(self):
        return weighted_mean(self.particles, self.particle_weights)

    def covariance(self):
        return weighted_covariance(self.particles, self.particle_weights)

    def std(self):
        return std_biased(self.particles, self.particle_weights)


# ----------------------------------------------------------------------------
# Sweeper workload (demos/sweeper/obe_sweeper.py): settings are (start, stop) index pairs into the
# swept setting array; a sweep is worth the point utility integrated along it, per unit cost.
# ----------------------------------------------------------------------------
def sweep_start_stop_indices(n_settings, subsample=3):
    """All [start, stop] pairs, stop > start, on every `subsample`-th setting index plus the last
    one (obe_sweeper.py:213-232)."""
    sub = list(range(0, n_settings, subsample))
    if sub[-1] != n_settings - 1:
        sub.append(n_settings - 1)
    pairs = [[a, b] for i, a in enumerate(sub[:-1]) for b in sub[i + 1:]]
    return np.array(pairs, dtype=np.int64)


def sweep_utility(point_utility, start_stop, cost_of_new_sweep):
    """(cumsum(U)[stop] - cumsum(U)[start]) / ((stop - start) + cost_of_new_sweep)
    (obe_sweeper.py:103-145)."""
    cum = np.cumsum(point_utility)
    ends = cum[start_stop]
    cost = start_stop[:, 1] - start_stop[:, 0] + cost_of_new_sweep
    return (ends[:, 1] - ends[:, 0]) / cost


class OracleSweeper(OracleOBE):
    """OptBayesExptSweeper (obe_sweeper.py:9-211) over OracleOBE: pdf_update runs one noise-parameter
    update per point of the sweep, the design half picks a (start, stop) pair."""

    def __init__(self, *a, start_stop_subsample=3, cost_of_new_sweep=5.0, **k):
        OracleOBE.__init__(self, *a, **k)
        self.start_stop_indices = sweep_start_stop_indices(self.allsettings.shape[1], start_stop_subsample)
        self.cost_of_new_sweep = cost_of_new_sweep
        self.last_sweep_utility = None

    def pdf_update(self, record, forced_resample=False):
        (setting_values,), result_values = record
        for setting, result in zip(setting_values, result_values):
            OracleOBE.pdf_update(self, ((setting,), result), forced_resample)
        return self.particles, self.particle_weights

    def sweep_utility(self):
        self.last_sweep_utility = sweep_utility(self.utility(), self.start_stop_indices, self.cost_of_new_sweep)
        return self.last_sweep_utility

    def opt_setting(self):
        index = opt_index(self.sweep_utility())
        self.last_setting_index = index
        return self.start_stop_indices[index]


def simulate_batch_measurement(model, settings_b, true_pars_b, cons, noise_level, seed, cycle, n_channels):
    """On-device MeasurementSimulator of the batched engines (csrc obe_bsimulate_body; the reference's
    obe_utils.py:8-53 with the library's counter-based normals): instance b measures at settings_b[b],
    y_c = model_c(setting, true_pars_b[b], cons) + noise_c * z_c, z = device_normals(counter b, key seed,
    epoch cycle).  settings_b (B, s), true_pars_b (B, p), noise_level scalar, (C,) or (B,).  -> (B, C)"""
    B = len(settings_b)
    z = device_normals(B, n_channels, seed, cycle)
    nl = np.asarray(noise_level, dtype=np.float64)
    y = np.empty((B, n_channels))
    for b in range(B):
        yb = np.atleast_1d(np.asarray(model(tuple(settings_b[b]), tuple(true_pars_b[b]), cons), dtype=np.float64))
        lvl = nl[b] if nl.shape == (B,) and B != n_channels else np.broadcast_to(nl, (n_channels,))
        y[b] = yb + lvl * z[b]
    return y


# ----------------------------------------------------------------------------
# Uniform streams of the batched engines (restatement of csrc obe_batch_uniform): the q-th
# uniform of instance b in cycle `cycle` is u53 of Philox4x32-10(ctr=(q, cycle, b, 0x0B5E0001), key=seed).
# ----------------------------------------------------------------------------
def batch_uniform(seed, b, cycle, q):
    q = np.atleast_1d(np.asarray(q, dtype=np.uint32))
    x = philox4x32_10(q, np.uint32(cycle), np.uint32(b), np.uint32(0x0B5E0001),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    v = (x[1].astype(np.uint64) << np.uint64(32)) | x[0].astype(np.uint64)
    return ((v >> np.uint64(11)).astype(np.float64) + 0.5) * (2.0 ** -53)


class ReplayRng:
    """Stands in for a numpy Generator inside a single engine so that it consumes exactly the
    uniforms instance b of a batched engine uses: random(K) -> uniforms 0..K-1 of the current cycle,
    random() -> uniform K (the comb offset of a resample).  Call next_cycle() after every cycle."""

    def __init__(self, seed, b, n_draws):
        self.seed, self.b, self.k, self.cycle = seed, b, n_draws, 0

    def random(self, n=None):
        if n is None:
            return float(batch_uniform(self.seed, self.b, self.cycle, self.k)[0])
        return batch_uniform(self.seed, self.b, self.cycle, np.arange(n))

    def next_cycle(self):
        self.cycle += 1
