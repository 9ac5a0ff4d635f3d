"""Batched Monte-Carlo runner: thousands of independent particle-filter runs in one launch per step.

The reference loops `for run in range(n_runs): for config in filter_configs: perform_predict_update_cycles(...)`
strictly sequentially on one core (pyrecest/evaluation/iterate_configs_and_runs.py:33-49,
perform_predict_update_cycles.py:45-79).  Runs share the model (transition, noise kinds, likelihood kind) and
differ in measurements and draws, so here all runs of one filter configuration advance together:

  particles   (n_runs * n, D) run-major, one contiguous buffer
  weights     padded stride ceil(n / 2048) * 2048 per run so that scan tiles never straddle runs
  per step    pfb_predict_loglik_batched  (measurement of run r = mu_batch[r])
              pfb_scan_cdf_batched        (per-run sequential-exact CDF = numpy.cumsum of that run)
              pfb_resample_batched        (per-run guide table + search + gather)
              pfb_moments_batched         (per-run estimate, one CTA per run)

Step order is the reference's: update (each measurement of the time step), estimate, then predict if
`apply_sys_noise_times[t]` (perform_predict_update_cycles.py:45-79).  Runs shard over GPUs with no
communication: give every rank its slice of `measurements`.
"""
from __future__ import annotations

import copy
import ctypes as C_
import math
import time
import warnings

import numpy as np
import torch

from . import _lib as L
from . import filters as F
from . import random as prandom
from . import registry as R
from ._engine import engine

TWO_PI = 2.0 * math.pi


class BatchedParticleFilter:
    """n_runs independent filters of n_particles each on one manifold ("euclid" | "torus" | "sphere")."""

    def __init__(self, manifold, n_runs, n_particles, dim, dtype=torch.float64, device=None):
        cls = {"euclid": F.EuclideanParticleFilter, "torus": F.HypertoroidalParticleFilter,
               "sphere": F.HypersphericalParticleFilter}[manifold]
        self.manifold = manifold
        self.n_runs, self.n = int(n_runs), int(n_particles)
        # the flat filter over all runs' particles builds the predict parameter blocks and owns the buffers
        self.flat = cls(self.n_runs * self.n, dim, dtype=dtype, device=device)
        st = self.flat._filter_state
        self.device = st.device
        self.D = st._records().shape[1]
        self.eng = engine(self.device)
        self.run_pad = self.eng.run_pad(self.n)
        self._logw = torch.empty(self.n_runs * self.run_pad, dtype=torch.float64, device=self.device)
        self._cdf = torch.empty_like(self._logw)
        self._alt = torch.empty_like(st._records())
        self.scan_exact = True
        self.resampling = "multinomial"
        self.failed = torch.zeros(self.n_runs, dtype=torch.bool, device=self.device)

    # -- state ---------------------------------------------------------------------------------------------
    @property
    def d(self):
        """(n_runs, n, D) view of the particles"""
        self.flat._flush()
        return self.flat._filter_state._records().view(self.n_runs, self.n, self.D)

    def set_particles(self, d):
        self.flat._pending = None
        x = torch.as_tensor(d, device=self.device).to(self.flat._filter_state.d.dtype).reshape(self.n_runs * self.n, self.D)
        self.flat._filter_state.d = x.contiguous().reshape(self.flat._filter_state.d.shape)

    def set_prior(self, prior):
        """every run starts from n samples of the same prior (configure_for_filter.py:73,101-103)"""
        self.flat.filter_state = prior

    # -- predict / update ------------------------------------------------------------------------------------
    def predict_nonlinear(self, f, noise_distribution=None, shift_instead_of_add=None):
        self.flat.predict_nonlinear(f, noise_distribution, True, shift_instead_of_add)

    def predict_identity(self, noise_distribution):
        self.flat.predict_identity(noise_distribution)

    def update_identity(self, meas_noise, measurements):
        """measurements: (n_runs, d) -- run r is weighted with meas_noise.set_mode(measurements[r]).pdf
        (abstract_particle_filter.py:95-125 per run), then resampled n of n."""
        z = R._np(measurements).reshape(self.n_runs, -1)
        if z.shape[1] != self.D:
            raise AssertionError("measurement dimension does not match the filter state")
        flat = self.flat
        st = flat._filter_state
        x = st._records()
        dist0 = copy.deepcopy(meas_noise)
        dist0 = dist0.set_mode(np.zeros(self.D)) if R.kind_of(dist0) != "vmf" else dist0
        if R.kind_of(dist0) == "wn_multi":
            # the image table must hold for every measurement: prune against the widest possible offset box
            lp, keep = _wn_params_any_mean(dist0, self.D, self.device)
        else:
            lp, keep = R.lik_params(dist0, self.D, images_to_device=R.device_put(self.device))
        mu_dev = torch.as_tensor(z, dtype=torch.float64, device=self.device).contiguous()
        lp.mu_batch_dev = mu_dev.data_ptr()
        eng = self.eng
        if flat._pending is not None:
            p, noise = flat._pending
            flat._pending = None
            try:
                eng.predict_loglik_batched(p, lp, x, x, noise, self.n_runs, self.n, self._logw)
            except NotImplementedError:
                eng.predict(p, x, x, noise)
                eng.loglik_batched(lp, x, self.n_runs, self.n, self._logw)
        else:
            eng.loglik_batched(lp, x, self.n_runs, self.n, self._logw)
        eng.scan_batched(self._logw, self._cdf, self.n_runs, self.n, exact=self.scan_exact)
        totals = self._cdf.view(self.n_runs, self.run_pad)[:, -1]
        self.failed |= ~(totals > 0)  # all weights zero / NaN: the reference's AssertionError -> run_failed
        systematic = self.resampling == "systematic"
        src = prandom.source()
        if src.fused:
            eng.resample_batched(self._cdf, x, self._alt, self.n_runs, self.n, self.D, seed=src.seed,
                                 offset=src.next_offset(), systematic=systematic)
        else:
            u = src.uniforms((self.n_runs,) if systematic else (self.n_runs * self.n,), self.device)
            eng.resample_batched(self._cdf, x, self._alt, self.n_runs, self.n, self.D, u=u, systematic=systematic)
        del keep, mu_dev
        new, self._alt = self._alt, x
        st._d = new.reshape(st._d.shape)
        st._uniform = True
        st._est_cache = None

    # -- estimates ---------------------------------------------------------------------------------------------
    def get_point_estimates(self):
        """(n_runs, D): mean (R^d), mean direction (T^d, S^d) per run; NaN for failed runs."""
        x = self.d.reshape(self.n_runs * self.n, self.D)
        D = self.D
        if self.manifold == "torus":
            m = self.eng.moments_batched(L.MOM_TRIG, x, self.n_runs, self.n, D)[:, : 2 * D].reshape(self.n_runs, D, 2)
            est = torch.remainder(torch.atan2(m[:, :, 1], m[:, :, 0]), TWO_PI)
        else:
            est = self.eng.moments_batched(L.MOM_MEAN, x, self.n_runs, self.n, D)[:, :D].clone()
            if self.manifold == "sphere":
                est = est / torch.linalg.norm(est, dim=1, keepdim=True)
        return torch.where(self.failed[:, None], torch.full_like(est, float("nan")), est)


_ANY_MEAN_CACHE: dict = {}


def _wn_params_any_mean(dist, dim, device):
    """wrapped-normal likelihood parameters whose image table is valid for ANY mean in [0, 2 pi)^d: the offset
    box of registry.wn_image_table is taken for mu = 0 and mu = 2 pi and the kept images are united."""
    C = np.atleast_2d(R._np(dist.C))
    key = (C.tobytes(), dim, str(device))
    hit = _ANY_MEAN_CACHE.get(key)
    if hit is not None:
        p = L.LikParams()
        C_.memmove(C_.byref(p), C_.byref(hit[0]), C_.sizeof(L.LikParams))
        return p, hit[1]
    W, logdet = R.gaussian_whitening(C)
    P = W @ W.T
    rows = {}
    for corner in np.array(np.meshgrid(*[[0.0, TWO_PI]] * dim)).T.reshape(-1, dim):
        table, _ = R.wn_image_table(corner, P)
        for row in table:
            rows[tuple(np.round(row[dim + 1:] / TWO_PI).astype(int))] = row
    table = np.array([rows[k] for k in sorted(rows, key=lambda kk: rows[kk][dim])])
    p = L.LikParams()
    p.kind, p.dim = L.LIK_WN_MULTI, dim
    p.W[: dim * dim] = W.reshape(-1).tolist()
    p.logc = -0.5 * (dim * math.log(TWO_PI) + logdet)
    p.n_images = table.shape[0]
    keep = torch.as_tensor(np.ascontiguousarray(table), device=device).contiguous()
    p.images_dev = keep.data_ptr()
    if len(_ANY_MEAN_CACHE) > 64:
        _ANY_MEAN_CACHE.clear()
    tmpl = L.LikParams()
    C_.memmove(C_.byref(tmpl), C_.byref(p), C_.sizeof(L.LikParams))
    _ANY_MEAN_CACHE[key] = (tmpl, keep)
    return p, keep


def _manifold_name(name):
    if name in ("Euclidean", "euclidean"):
        return "euclid"
    if name in ("circle", "hypertorus"):
        return "torus"
    if name in ("hypersphere", "hypersphere_general"):
        return "sphere"
    raise NotImplementedError(f"Manifold {name} not supported yet")


def iterate_configs_and_runs(groundtruths, measurements, scenario_config, filter_configs, evaluation_config=None):
    """Same inputs / outputs as pyrecest.evaluation.iterate_configs_and_runs for {"name": "pf"} configurations;
    all runs of a configuration advance together on the GPU.

    measurements: (n_runs, n_timesteps, d) array (one measurement per time step)
    returns (last_filter_states, run_times, run_failed, groundtruths, measurements) where last_filter_states[c]
    is the BatchedParticleFilter of configuration c (its .d[r] are run r's particles), run_times[c, r] is the
    wall time of the batch divided by n_runs, and run_failed[c, r] marks runs whose weights vanished."""
    evaluation_config = evaluation_config or {}
    if evaluation_config.get("extract_all_point_estimates", False) or evaluation_config.get("convert_to_point_estimate_during_runtime", False):
        raise NotImplementedError("This is not implemented yet.")
    meas = R._np(measurements)
    n_runs, T = meas.shape[0], scenario_config["n_timesteps"]
    apply = scenario_config.get("apply_sys_noise_times", [True] * (T - 1) + [False])
    manifold = _manifold_name(scenario_config["manifold"])
    prior = scenario_config["initial_prior"]
    dim = prior.dim
    configs = []
    for fc in filter_configs:
        if fc["name"] != "pf":
            raise NotImplementedError("only particle filters run on the batched path")
        for par in np.atleast_1d(fc["parameter"]):
            if not par:
                raise ValueError("Using zero particles does not make sense")
            configs.append(int(par))
    run_times = np.empty((len(configs), n_runs))
    run_failed = np.zeros((len(configs), n_runs), dtype=bool)
    last_filter_states = np.empty(len(configs), dtype=object)
    last_estimates = []
    for c, n_particles in enumerate(configs):
        bf = BatchedParticleFilter(manifold, n_runs, n_particles, dim)
        bf.set_prior(prior)
        torch.cuda.synchronize(bf.device)
        t0 = time.time()
        for t in range(T):
            bf.update_identity(scenario_config["meas_noise"], meas[:, t].reshape(n_runs, -1))
            if apply[t]:
                bf.predict_identity(scenario_config["sys_noise"])
        est = bf.get_point_estimates()
        torch.cuda.synchronize(bf.device)
        run_times[c, :] = (time.time() - t0) / n_runs
        failed = bf.failed.cpu().numpy()
        run_failed[c, :] = failed
        if failed.any() and not evaluation_config.get("tolerate_failure", False):
            raise AssertionError("The sum of all weights should be greater than 0.")
        if failed.any() and evaluation_config.get("auto_warning_on_off", False):
            warnings.warn(f"Filter {c} config {n_particles}: runs {np.nonzero(failed)[0].tolist()} FAILED")
        last_filter_states[c] = bf
        last_estimates.append(est.cpu().numpy())
    iterate_configs_and_runs.last_estimates = np.stack(last_estimates)
    return last_filter_states, run_times, run_failed, groundtruths, measurements
