"""Particle cloud sharded over the GPUs of one node: one process per GPU (torch.distributed).

Particles are exchangeable, so every O(N) step is local to a shard; the setting grid is sliced for
the utility pass.  Only small vectors cross NVLink, through NCCL:
  * after the update: one all-gather of the 64-double stats blocks -> global normaliser, N_eff,
    mean, covariance, and the exclusive scan of shard weight totals (the inter-GPU CDF offsets);
  * the K drawn parameter sets: each shard writes the draws it owns into a zeroed (d, K) buffer,
    one all-reduce(sum) makes them global;
  * the argmax: one all-gather of (value, index) pairs, lowest global index wins ties.
Resampling needs NO particle exchange: shard g's particles own the global comb slots
[H_g, H_g+1), so shard lengths float by a fraction of a percent per resample inside a slack
capacity.  The host-side decisions are pure functions (combine_stats, shard_slot_bounds,
assign_draws, reduce_best) so they can be tested without a GPU (gloo).
"""
import ctypes as C

import numpy as np

from . import _lib
from .obe_base import OptBayesExpt


# --------------------------------------------------------------------------------------------------
# pure host logic
# --------------------------------------------------------------------------------------------------
def combine_stats(gathered, d):
    """Global stats from the per-shard stats blocks, summed in rank order (identical on every rank).

    Returns a dict with totals (G,), offsets (G,), total, sumsq, sumt, m1 (d,), m2 packed, noise (4,),
    pivot (d,).  All shards must have used the same pivot."""
    g = np.asarray(gathered, dtype=np.float64)
    nm2 = d * (d + 1) // 2
    totals = g[:, _lib.ST_TOTAL].copy()
    offsets = np.zeros_like(totals)
    acc = 0.0
    for r in range(len(totals)):          # sequential: the canonical inter-GPU exclusive scan
        offsets[r] = acc
        acc = acc + totals[r]

    def rsum(cols):
        out = g[0, cols].copy()
        for r in range(1, g.shape[0]):
            out = out + g[r, cols]
        return out
    return dict(totals=totals, offsets=offsets, total=acc,
                sumsq=float(rsum([_lib.ST_SUMSQ])[0]), sumt=float(rsum([_lib.ST_SUMT])[0]),
                m1=rsum(list(range(_lib.ST_M1, _lib.ST_M1 + d))),
                m2=rsum(list(range(_lib.ST_M2, _lib.ST_M2 + nm2))),
                noise=rsum(list(range(_lib.ST_NOISE, _lib.ST_NOISE + 4))),
                pivot=g[0, _lib.ST_PIVOT:_lib.ST_PIVOT + d].copy())


def moments_from(gs, d):
    """(mean, covariance, biased variance, n_eff) from combined stats (same formulas as ParticlePDF)."""
    s = gs['sumt']
    mean = gs['pivot'] + gs['m1'] / s
    m2 = np.zeros((d, d))
    q = 0
    for j in range(d):
        for k in range(j, d):
            m2[j, k] = m2[k, j] = gs['m2'][q]
            q += 1
    cov = (m2 - np.outer(gs['m1'], gs['m1']) / s) / (s - gs['sumsq'] / s)
    var = np.diag(m2) / s - (gs['m1'] / s) ** 2
    n_eff = gs['total'] ** 2 / gs['sumsq']
    return mean, cov, var, n_eff


def shard_slot_bounds(offsets, total, u0, n_total, comb_count):
    """H_g: first global comb slot owned by shard g (H_0 = 0, H_G = n_total), monotone.
    comb_count(c, u0, n) must be the library's obe_comb_count so that neighbours agree to the bit."""
    inv_total = 1.0 / total
    bounds = [0]
    for g in range(1, len(offsets)):
        h = int(comb_count(float(offsets[g] * inv_total), float(u0), int(n_total)))
        bounds.append(min(max(h, bounds[-1]), int(n_total)))
    bounds.append(int(n_total))
    return bounds


def assign_draws(u, offsets, totals, total):
    """Owner shard and shard-local uniform of every global uniform u_k.
    Owner g: offset_g <= u*total < offset_g + total_g (last shard takes the remainder)."""
    u = np.asarray(u, dtype=np.float64)
    target = u * total
    ends = offsets + totals
    owner = np.searchsorted(ends, target, side='right')
    owner = np.minimum(owner, len(totals) - 1)
    # skip empty shards
    for i in range(len(owner)):
        while totals[owner[i]] <= 0 and owner[i] > 0:
            owner[i] -= 1
    local = (target - offsets[owner]) / np.where(totals[owner] > 0, totals[owner], 1.0)
    local = np.clip(local, 0.0, np.nextafter(1.0, 0.0))
    return owner, local


def reduce_best(pairs):
    """np.argmax over the concatenated grid from per-shard (global_index, value) pairs: the first
    maximum wins, NaN counts as the maximum (numpy semantics)."""
    best_i, best_v = -1, 0.0
    for idx, val in pairs:
        idx = int(idx)
        if idx < 0:
            continue
        if best_i < 0:
            take = True
        else:
            vn, bn = val != val, best_v != best_v
            if vn and bn:
                take = idx < best_i
            elif vn:
                take = True
            elif bn:
                take = False
            else:
                take = val > best_v or (val == best_v and idx < best_i)
        if take:
            best_i, best_v = idx, float(val)
    return best_i, best_v


def setting_slice(n_settings, rank, world):
    lo = n_settings * rank // world
    hi = n_settings * (rank + 1) // world
    return lo, hi


class Comm:
    """The three small collectives, over torch.distributed (nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

        self.backend = dist.get_backend(group)

    def _staged(self, t):
        # gloo moves host memory: stage device tensors through the CPU (tests on a single GPU)
        return self.backend == 'gloo' and t.is_cuda

    def allgather(self, vec):
        import torch
        src = vec.contiguous().view(-1)
        if self._staged(vec):
            src = src.cpu()
        parts = [torch.empty_like(src) for _ in range(self.world)]
        if self.backend == 'gloo':
            self.dist.all_gather(parts, src, group=self.group)
            out = torch.stack(parts)
        else:
            out = torch.empty((self.world, src.numel()), dtype=src.dtype, device=src.device)
            self.dist.all_gather_into_tensor(out.view(-1), src, group=self.group)
        return out.view((self.world,) + tuple(vec.shape)).to(vec.device)

    def allreduce_sum(self, t):
        if self._staged(t):
            c = t.cpu()
            self.dist.all_reduce(c, op=self.dist.ReduceOp.SUM, group=self.group)
            t.copy_(c)
        else:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t


# --------------------------------------------------------------------------------------------------
class ShardedOptBayesExpt(OptBayesExpt):
    """OptBayesExpt over a cloud sharded across the ranks of a process group.

    ``parameter_samples`` is THIS rank's shard (d, n_local); ``setting_values`` is the full grid on
    every rank.  All ranks must pass the same ``seed`` (the comb offset u0 and the K uniforms are
    drawn from identically seeded Generators so that every rank takes the same decisions)."""

    def __init__(self, measurement_model, setting_values, parameter_samples, constants, group=None,
                 slack=0.25, seed=0, **kwargs):
        import torch
        self._comm = Comm(group)
        kwargs['resampling'] = 'systematic'
        kwargs['seed'] = seed
        n_local = parameter_samples.shape[-1]
        self._capacity = int(n_local * (1.0 + slack)) + 2 * _lib.TILE     # shard lengths float
        self._gstats = None
        OptBayesExpt.__init__(self, measurement_model, setting_values, parameter_samples, constants, **kwargs)
        dev = self._buf.device
        counts = self._comm.allgather(torch.tensor([self.n_particles], dtype=torch.int64, device=dev))
        self._counts = counts.cpu().numpy().reshape(-1).astype(np.int64)
        self.n_total = int(self._counts.sum())
        self._check(self._lib.obe_set_uniform_total(self._cs(), self.n_total, self._stream()))
        # slice of the setting grid this rank evaluates
        n_set = len(self.setting_indices)
        self._s_lo, self._s_hi = setting_slice(n_set, self._comm.rank, self._comm.world)
        # a common pivot for the shifted moments: rank 0's estimate
        piv = torch.from_numpy(self._pivot.copy()).to(dev)
        self._pivot = self._comm.allgather(piv)[0].cpu().numpy()

    def _invalidate(self, particles=False, weights=True):
        OptBayesExpt._invalidate(self, particles, weights)
        if weights:
            self._gstats = None

    # ---- global stats
    def _sync_global_stats(self):
        """all-gather the stats blocks, combine, and install the global normaliser on the device."""
        gathered = self._comm.allgather(self._buf.stats).cpu().numpy()
        self._stats = gathered[self._comm.rank].copy()
        self._gstats = combine_stats(gathered, self.n_dims)
        self._gmom = moments_from(self._gstats, self.n_dims)
        return self._gstats

    def _install_global_normaliser(self):
        gs = self._sync_global_stats()
        self._buf.stats[_lib.ST_INVS:_lib.ST_INVS + 1].fill_((1.0 / gs['total']) if self._weights_lazy else 1.0)
        self._moments_valid = True
        return gs

    def _after_update(self):
        self._invalidate()
        self._weights_uniform = False
        self._weights_lazy = True
        self._install_global_normaliser()
        self._pivot = self._gmom[0].copy()
        if self.tuning_parameters['auto_resample']:
            self.resample_test()

    def _ensure_moments(self):
        if self._gstats is None:
            if not self._moments_valid:      # device stats are stale too: one refresh pass
                ni = self._noise_index
                self._check(self._lib.obe_refresh(self._cs(), 0, 0, _lib.iarr(ni), 0 if ni is None else len(ni),
                                                  _lib.darr(self._pivot, _lib.MAX_PARAMS), 0, self._stream()))
            self._install_global_normaliser()
        return self._stats

    def mean(self):
        self._ensure_moments()
        return self._gmom[0].copy()

    def covariance(self):
        self._ensure_moments()
        return self._gmom[1].copy()

    def std(self):
        self._ensure_moments()
        return np.sqrt(np.maximum(self._gmom[2], 0.0))

    def n_eff(self):
        self._ensure_moments()
        return float(self._gmom[3])

    def resample_test(self):
        import warnings
        n_eff = self.n_eff()
        if n_eff < 0.1 * self.n_total:
            warnings.warn(f"\nParticle filter rejected > 90 % of particles. N_eff = {n_eff:.2f}. "
                          "Particle impoverishment may lead to errors.", RuntimeWarning)
            self.resample()
            self.just_resampled = True
        elif n_eff / self.n_total < self.tuning_parameters['resample_threshold']:
            self.resample()
            self.just_resampled = True
        else:
            self.just_resampled = False

    # ---- resample: every shard keeps its own offspring
    def resample(self):
        self._ensure_moments()
        gs = self._gstats
        a_param = float(self.tuning_parameters['a_param'])
        scale = 1 if self.tuning_parameters['scale'] else 0
        self._epoch += 1
        u0 = float(self.rng.random())                       # identical on every rank
        bounds = shard_slot_bounds(gs['offsets'], gs['total'], u0, self.n_total, self._lib.obe_comb_count)
        r = self._comm.rank
        lo, hi = bounds[r], bounds[r + 1]
        if hi - lo < 1:
            raise RuntimeError('a shard lost all its particles in a resample; rebalance is not implemented')
        mean, cov = self._gmom[0], self._gmom[1]
        newcov = (1.0 - a_param ** 2) * cov
        try:
            factor = np.ascontiguousarray(np.linalg.cholesky(newcov).T)
        except np.linalg.LinAlgError:
            (uu, ss, _) = np.linalg.svd(newcov)
            factor = np.ascontiguousarray((uu * np.sqrt(ss)).T)
        if self._alt is None:
            self._alt = self._buf.empty_like()
        self._alt.resize(hi - lo)
        self._check(self._lib.obe_resample_systematic_sharded(
            self._cs(), self._cs(self._alt), u0, self.n_total, lo, hi, float(gs['offsets'][r]), float(gs['total']),
            1 if r == self._comm.world - 1 else 0, _lib.darr(factor.reshape(-1)), _lib.darr(mean),
            self._philox_seed, self._epoch, a_param, scale, None, None, self._stream()))
        self._buf, self._alt = self._alt, self._buf
        self.n_particles = self._buf.n
        self._counts = np.diff(np.asarray(bounds, dtype=np.int64))
        self._invalidate(particles=True)
        self._stats = None
        self._moments_valid = False
        self._weights_uniform = True
        self._weights_lazy = False

    def _cdf_totals(self):
        """(offsets, totals, total) of the current weights, for the assignment of the draws."""
        if self._gstats is None and self._weights_uniform:
            totals = self._counts.astype(np.float64) * (1.0 / self.n_total)
            offsets = np.zeros_like(totals)
            acc = 0.0
            for r in range(len(totals)):
                offsets[r] = acc
                acc = acc + totals[r]
            return offsets, totals, acc
        self._ensure_moments()
        return self._gstats['offsets'], self._gstats['totals'], self._gstats['total']

    # ---- K draws through the sharded CDF
    def _randdraw_dev(self, n_draws):
        import torch
        u = self.rng.random(n_draws)                        # identical on every rank
        offsets, totals, total = self._cdf_totals()
        owner, local = assign_draws(u, offsets, totals, total)
        order = np.argsort(owner, kind='stable')
        draws = torch.zeros((self.n_dims, n_draws), dtype=torch.float64, device=self._buf.device)
        mine = order[owner[order] == self._comm.rank]
        if len(mine):
            col0 = int(np.sum(owner < self._comm.rank))
            ptr = C.c_void_p(draws.data_ptr() + 8 * col0)
            self._check(self._lib.obe_draw_strided(self._cs(), _lib.darr(local[mine]), len(mine), ptr,
                                                   int(n_draws), None, self._stream()))
        self._comm.allreduce_sum(draws)
        return draws

    # ---- utility over this rank's slice of the grid
    def _utility_dev_run(self):
        draws = self._randdraw_dev(self.N_DRAWS)
        n_loc = self._s_hi - self._s_lo
        var_noise = _lib.darr(np.asarray(self.yvar_noise_model(), dtype=np.float64).reshape(-1), _lib.MAX_CHANNELS)
        cost = self.cost_estimate()
        cost_ptr = None
        if not (np.isscalar(cost) and float(cost) == 1.0):
            import torch
            cost_arr = np.array(np.broadcast_to(np.asarray(cost, dtype=np.float64), (len(self.setting_indices),)))
            self._cost_dev = torch.from_numpy(cost_arr[self._s_lo:self._s_hi].copy()).to(self._buf.device)
            cost_ptr = C.c_void_p(self._cost_dev.data_ptr())
        settings_ptr = C.c_void_p(self._settings_dev.data_ptr() + 8 * self._s_lo)
        util_ptr = C.c_void_p(self._utility_dev.data_ptr() + 8 * self._s_lo)
        self._check(self._lib.obe_utility(self._model, C.c_void_p(draws.data_ptr()), int(self.N_DRAWS), settings_ptr,
                                          self._lds, n_loc, self._cons_arr, var_noise, None, cost_ptr,
                                          self._utility_code, 1 if self.utility_log_form else 0, util_ptr,
                                          C.c_void_p(self._best_dev.data_ptr()),
                                          C.c_void_p(self._select_scratch.data_ptr()), self._stream()))

    def opt_setting(self):
        import torch
        self._utility_dev_run()
        pairs = self._comm.allgather(self._best_dev).cpu()
        vals = pairs[:, 1].contiguous().view(torch.float64).numpy()
        idxs = pairs[:, 0].numpy()
        lows = [setting_slice(len(self.setting_indices), r, self._comm.world)[0] for r in range(self._comm.world)]
        best, _ = reduce_best([(int(idxs[r]) + lows[r] if idxs[r] >= 0 else -1, float(vals[r]))
                               for r in range(self._comm.world)])
        self.last_setting_index = best
        return tuple(self.allsettings[:, best])

    def utility(self):
        import torch
        self._utility_dev_run()
        n_set = len(self.setting_indices)
        world = self._comm.world
        width = max(setting_slice(n_set, r, world)[1] - setting_slice(n_set, r, world)[0] for r in range(world))
        buf = torch.zeros(width, dtype=torch.float64, device=self._buf.device)
        buf[:self._s_hi - self._s_lo] = self._utility_dev[self._s_lo:self._s_hi]
        allu = self._comm.allgather(buf).cpu().numpy()
        out = np.empty(n_set)
        for r in range(world):
            lo, hi = setting_slice(n_set, r, world)
            out[lo:hi] = allu[r, :hi - lo]
        return out

    def good_setting(self, pickiness=None):
        import torch
        if pickiness is None:
            pickiness = self.pickiness
        full = torch.from_numpy(self.utility()).to(self._buf.device)
        self._utility_dev.copy_(full)
        u = float(self.rng.random())
        self._check(self._lib.obe_pick(C.c_void_p(self._utility_dev.data_ptr()), len(self.setting_indices),
                                       float(pickiness), u, C.c_void_p(self._pick_dev.data_ptr()),
                                       C.c_void_p(self._select_scratch.data_ptr()), self._stream()))
        goodindex = int(self._pick_dev.item())
        self.last_setting_index = goodindex
        return tuple(self.allsettings[:, goodindex])

    def run_cycle_async(self, measurement_record, resample=True, select=True):
        """Sharded cycle: the update kernels are asynchronous, but the global normaliser, the shard
        slot bounds and the argmax come from gathered stats on the host (one small collective +
        one synchronisation per phase)."""
        OptBayesExpt.run_cycle_async(self, measurement_record, resample=False, select=False)
        self._install_global_normaliser()
        if resample:
            self.resample()
            self.just_resampled = True
        if select:
            self._utility_dev_run()
