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
import datetime
import math
import os
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
        if isinstance(measurements, torch.Tensor):  # already on the device (evaluation.generate_measurements): stays there
            z = measurements.reshape(self.n_runs, -1)
        else:
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
        if R.kind_of(dist0) in ("wn_multi", "wn_1d", "vm"):
            # equivalent on the torus, and required: the pruned image table is built for centres in [0, 2 pi)^d
            mu_dev = torch.remainder(mu_dev, 2.0 * np.pi)
        lp.mu_batch_dev = mu_dev.data_ptr()
        eng = self.eng
        if flat._pending is not None:
            p, noise = flat._pending
            flat._pending = None
            try:
                eng.predict_loglik_batched(p, lp, x, x, noise, self.n_runs, self.n, self._logw, scaled=True)
            except NotImplementedError:
                eng.predict(p, x, x, noise)
                eng.loglik_batched(lp, x, self.n_runs, self.n, self._logw, scaled=True)
        else:
            eng.loglik_batched(lp, x, self.n_runs, self.n, self._logw, scaled=True)
        # (slice-scaled linear weights between the weighting kernel and the scan: one exp per particle and step)
        eng.scan_batched(self._logw, self._cdf, self.n_runs, self.n, exact=self.scan_exact, scaled=True)
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
    flat = np.ascontiguousarray(table).reshape(-1)
    # candidate grid over dev in [-3 pi, pi)^d (valid for every measurement): 2 G cells per dimension, G = 8 for
    # d <= 2 (256 cells), 4 for d = 3 (512 cells)
    G = R.WN_CELL_GRID if dim <= 2 else 4
    masks = R.wn_cell_masks(None, table, G=G, any_mean=True)
    cells = None if masks is None else R.wn_cell_table(None, table, masks, G=G, any_mean=True)
    if cells is not None:
        flat = np.concatenate([flat, cells.view(np.float64)])
        p.cell_grid, p.cell_span = G, 1
    keep = torch.as_tensor(flat, device=device).contiguous()
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
    meas = measurements if isinstance(measurements, torch.Tensor) else R._np(measurements)
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
            # perform_predict_update_cycles.py:57-65: every measurement of the time step is an update of its own
            # (measurements (n_runs, T, d): one per step; (n_runs, T, M, d): M per step)
            if meas.ndim == 4:
                for m in range(meas.shape[2]):
                    bf.update_identity(scenario_config["meas_noise"], meas[:, t, m].reshape(n_runs, -1))
            else:
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


# ---------------------------------------------------------------------------------------------------------------
# Either side of the runner: scenario generation and the error summary for all runs at once,
# on the device, so that a Monte-Carlo study never moves per-run data through the host.  These are (n_runs, d)
# tensor operations -- plumbing, not kernels; the draws come from the same sources as the filters' (pyrecest_b200.random).
# ---------------------------------------------------------------------------------------------------------------
def _sample_on_device(dist, n):
    s = R.as_distribution(dist).sample(n)
    if not isinstance(s, torch.Tensor):  # a reference (numpy-backend) distribution object: its draws come from the host
        dev = torch.device("cuda", torch.cuda.current_device())
        s = torch.as_tensor(np.asarray(s, dtype=np.float64), device=dev)
    return s.reshape(n, -1)


def generate_groundtruths(scenario_config, n_runs):
    """pyrecest/evaluation/generate_groundtruth.py:7-94 for every run at once (n_targets = 1, identity system model
    with additive `sys_noise`, the case the particle-filter scenarios use): x_0 ~ initial_prior,
    x_t = x_{t-1} + sys_noise.sample().  Like the reference, the walk is NOT wrapped on periodic manifolds.
    Returns a device tensor (n_runs, n_timesteps, d)."""
    if "gen_next_state_with_noise" in scenario_config or "gen_next_state_without_noise" in scenario_config:
        raise NotImplementedError("only the identity system model with additive sys_noise is generated on the device")
    if scenario_config.get("n_targets", 1) != 1:
        raise NotImplementedError("multi-target scenarios are outside the particle-filter path")
    if "sys_noise" not in scenario_config:
        raise ValueError("Cannot generate groundtruth.")
    T = int(scenario_config["n_timesteps"])
    x = _sample_on_device(scenario_config["initial_prior"], n_runs)
    out = [x]
    for _ in range(1, T):
        x = x + _sample_on_device(scenario_config["sys_noise"], n_runs)
        out.append(x)
    return torch.stack(out, dim=1)


def generate_measurements(groundtruths, scenario_config, reference_indexing=True):
    """pyrecest/evaluation/generate_measurements.py:147-184 for every run at once, one measurement per time step:
    hypertoroidal noise -> mod(x + noise, 2 pi); Gaussian -> x + noise; von Mises-Fisher -> a draw around x.
    reference_indexing: the reference reads groundtruth[t - 1] for the measurement of step t (so step 0 sees the LAST
    state); kept by default for identical behaviour, False pairs step t with groundtruth[t].
    Returns a device tensor (n_runs, n_timesteps, d)."""
    gt = groundtruths
    n_runs, T = gt.shape[0], gt.shape[1]
    noise = R.as_distribution(scenario_config["meas_noise"])
    kind = R.kind_of(noise)
    out = []
    for t in range(T):
        x = gt[:, (t - 1) % T if reference_indexing else t]
        if kind in ("wn_multi", "wn_1d", "vm"):
            z = x + _sample_on_device(noise, n_runs)
            zw = torch.empty_like(z)
            engine(z.device).wrap_2pi(z.contiguous(), zw)
            out.append(zw)
        elif kind == "gauss":
            out.append(x + _sample_on_device(noise, n_runs))
        elif kind == "vmf":
            p = L.PredictParams()
            p.mode, p.dim, p.noise_cols = L.PRED_SPHERE_VMF, 3, 3
            p.kappa = float(R._np(noise.kappa))
            p.exp_m2k = math.exp(-2.0 * p.kappa)
            src = prandom.source()
            modes = x.contiguous()
            z = torch.empty_like(modes)
            if src.fused:
                p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VMF, src.seed, src.next_offset()
                engine(modes.device).predict(p, modes, z, None)
            else:
                engine(modes.device).predict(p, modes, z, src.vmf_triples(n_runs, modes.device, modes.dtype))
            out.append(z)
        else:
            raise NotImplementedError(f"measurement noise {type(noise).__name__} is not generated on the device")
    return torch.stack(out, dim=1)


def get_distance_function(manifold_name):
    """pyrecest/evaluation/get_distance_function.py:14-56, vectorised over runs: (n_runs, d) x (n_runs, d) -> (n_runs,)"""
    if "Symm" in manifold_name:
        raise NotImplementedError("Not implemented yet")
    if "circle" in manifold_name or "hypertorus" in manifold_name:

        def distance_function(xest, xtrue):  # norm of AbstractHypertoroidalDistribution.angular_error (:146-167)
            diff = torch.abs(torch.remainder(xest, TWO_PI) - torch.remainder(xtrue, TWO_PI))
            return torch.linalg.norm(torch.minimum(diff, TWO_PI - diff), dim=-1)

    elif "hypersphere" in manifold_name:

        def distance_function(x1, x2):
            return torch.arccos(torch.sum(x1 * x2, dim=-1))

    elif ("euclidean" in manifold_name or "Euclidean" in manifold_name) and "MTT" not in manifold_name:

        def distance_function(x1, x2):
            return torch.linalg.norm(x1 - x2, dim=-1)

    else:
        raise NotImplementedError(f"distance on {manifold_name} is outside the particle-filter path")
    return distance_function


def determine_all_deviations(last_estimates, groundtruths, manifold_name):
    """pyrecest/evaluation/determine_all_deviations.py:9-52: distance between the final estimate of every
    (configuration, run) and the last groundtruth state.  last_estimates: (n_configs, n_runs, d)."""
    gt = groundtruths
    est = torch.as_tensor(last_estimates, device=gt.device, dtype=gt.dtype) if not isinstance(last_estimates, torch.Tensor) else last_estimates
    dist = get_distance_function(manifold_name)
    dev = torch.stack([dist(est[c].reshape(gt.shape[0], -1), gt[:, -1]) for c in range(est.shape[0])])
    if bool(torch.any(torch.isinf(dev))):
        warnings.warn("Some of the errors are infinite.")
    return dev


def summarize_filter_results(scenario_config, filter_configs, runtimes, groundtruths, run_failed, last_estimates):
    """pyrecest/evaluation/summarize_filter_results.py:11-58: error_mean / error_std / time_mean / failure_rate per
    configuration (population std, like numpy's)."""
    dev = determine_all_deviations(last_estimates, groundtruths, scenario_config["manifold"])
    err_mean = dev.mean(dim=1).cpu().numpy()
    err_std = dev.std(dim=1, unbiased=False).cpu().numpy()
    times = np.mean(R._np(runtimes), axis=1)
    failed = R._np(run_failed)
    rates = failed.sum(axis=1) / failed.shape[1]
    results = copy.deepcopy(list(filter_configs))
    for d, e, s, tm, fr in zip(results, err_mean, err_std, times, rates):
        d["error_mean"], d["error_std"], d["time_mean"], d["failure_rate"] = float(e), float(s), float(tm), float(fr)
    return results


# ---------------------------------------------------------------------------------------------------------------
# The reference's result files (checkpoint / resume of a study): evaluate_for_variables writes one .npy
# per study with a fixed set of keys, evaluate_for_file reads such a file back and evaluates filters on its
# ground truths and measurements.
# ---------------------------------------------------------------------------------------------------------------
RESULT_KEYS = ("groundtruths", "measurements", "run_failed", "last_filter_states", "runtimes", "scenario_config",
               "filter_configs", "evaluation_config")  # evaluate_for_variables.py:58-70


def _to_host(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a


def evaluate_for_variables(groundtruths, measurements, filter_configs, scenario_config=None, save_folder=".",
                           plot_each_step=False, convert_to_point_estimate_during_runtime=False,
                           extract_all_point_estimates=False, tolerate_failure=False, auto_warning_on_off=False):
    """pyrecest/evaluation/evaluate_for_variables.py:11-82 on the batched path: expand the filter configurations, run
    them (iterate_configs_and_runs), save `<name>_<date>.npy` with the reference's keys, return the reference's
    8-tuple.  Differences forced by the batched representation: last_filter_states[c] is the (n_runs, n_particles, D)
    array of configuration c's last particle sets (the reference pickles one distribution object per run), and
    runtimes[c, r] is the batch's wall time divided by the number of runs."""
    if scenario_config is None:
        scenario_config = {"name": "custom"}
    if plot_each_step:
        raise NotImplementedError("Plotting is not implemented yet.")
    evaluation_config = {
        "plot_each_step": plot_each_step,
        "convert_to_point_estimate_during_runtime": convert_to_point_estimate_during_runtime,
        "extract_all_point_estimates": extract_all_point_estimates,
        "tolerate_failure": tolerate_failure,
        "auto_warning_on_off": auto_warning_on_off,
    }
    filter_configs = [{"name": f["name"], "parameter": p} for f in filter_configs for p in (np.atleast_1d(f["parameter"]).tolist() or [None])]
    states, runtimes, run_failed, groundtruths, measurements = iterate_configs_and_runs(
        groundtruths, measurements, scenario_config, filter_configs, evaluation_config)
    last_filter_states = np.empty(len(states), dtype=object)
    for c, bf in enumerate(states):
        last_filter_states[c] = bf.d.cpu().numpy()
    stamp = datetime.datetime.now().strftime("%Y-%m-%d--%H-%M-%S")
    filename = os.path.join(save_folder, f"{scenario_config['name']}_{stamp}.npy")
    storable = {k: v for k, v in scenario_config.items() if not hasattr(v, "pdf") and not callable(v)}
    np.save(filename, {
        "groundtruths": _to_host(groundtruths), "measurements": _to_host(measurements), "run_failed": run_failed,
        "last_filter_states": last_filter_states, "runtimes": runtimes, "scenario_config": storable,
        "filter_configs": filter_configs, "evaluation_config": evaluation_config,
    }, allow_pickle=True)
    evaluate_for_variables.last_file = filename
    return (last_filter_states, runtimes, run_failed, groundtruths, measurements, scenario_config, filter_configs,
            evaluation_config)


def evaluate_for_file(input_file_name, filter_configs, scenario_config, save_folder=".", plot_each_step=False,
                      convert_to_point_estimate_during_runtime=False, extract_all_point_estimates=False,
                      tolerate_failure=False, auto_warning_on_off=False):
    """pyrecest/evaluation/evaluate_for_file.py:13-71: ground truths and measurements come from a result file (the
    reference's or ours: both are a pickled dict with RESULT_KEYS); scenario_config supplies the manifold, the prior
    and the noise models, and gets the reference's defaults (name, n_timesteps, n_meas_at_individual_time_step,
    apply_sys_noise_times)."""
    data = np.load(input_file_name, allow_pickle=True).item()
    if "name" not in scenario_config:
        scenario_config["name"] = os.path.splitext(os.path.basename(input_file_name))[0]
    gts, meas = np.asarray(data["groundtruths"]), data["measurements"]
    if "n_timesteps" in scenario_config:
        assert scenario_config["n_timesteps"] == gts.shape[1], \
            "n_timesteps in scenario_config does not match the number of measurements in the file"
    else:
        scenario_config["n_timesteps"] = gts.shape[1]
    meas = np.asarray(meas)
    if meas.dtype == object:  # the reference stores one array per (run, step): (d,) or (M, d)
        meas = np.array([[np.asarray(m, dtype=np.float64) for m in row] for row in meas])
    n_meas = np.full(meas.shape[1], meas.shape[2] if meas.ndim == 4 else 1, dtype=int)
    scenario_config.setdefault("n_meas_at_individual_time_step", n_meas)
    scenario_config.setdefault("apply_sys_noise_times",
                               np.concatenate([np.ones(scenario_config["n_timesteps"] - 1, dtype=bool), [False]]))
    return evaluate_for_variables(gts, meas, filter_configs, scenario_config, save_folder=save_folder,
                                  plot_each_step=plot_each_step,
                                  convert_to_point_estimate_during_runtime=convert_to_point_estimate_during_runtime,
                                  extract_all_point_estimates=extract_all_point_estimates,
                                  tolerate_failure=tolerate_failure, auto_warning_on_off=auto_warning_on_off)
