"""Particle filters with the reference's class names, constructor signatures and method semantics
(pyrecest/filters/*particle_filter.py, v0.9.1); predict / update / resample / estimate run in the
libpfb.so kernels on the GPU.

Differences a pyRecEst user must know (all from BASELINE.json's north_star):
  * transitions and likelihoods come from a registry (registry.IdentityTransition /
    LinearTransition; Gaussian, (hypertoroidal) wrapped-normal, von Mises, von Mises-Fisher
    distributions).  Any other Python callable raises NotImplementedError -- it is not run on the CPU.
  * extra, optional knobs: `resampling` ("multinomial" = reference order; "multinomial_sorted" = the same exact
    multinomial offspring counts emitted in CDF order, in-kernel draws only; or "systematic"), `scan_exact`
    (True = CDF bit-identical to numpy.cumsum), `validate` (read the device status word after each
    update to raise the reference's AssertionErrors; costs one tiny D2H sync).
  * HypersphericalParticleFilter exists (absent in the reference, configure_for_filter.py:90-100).
"""
from __future__ import annotations

import copy
import os
import math

import numpy as np
import torch

from . import _lib as L
from . import random as prandom
from . import registry as R
from ._engine import engine
from .distributions import (
    AbstractDiracDistribution,
    CircularDiracDistribution,
    HyperhemisphericalDiracDistribution,
    HypersphericalDiracDistribution,
    HypertoroidalDiracDistribution,
    LinearDiracDistribution,
    WrappedNormalDistribution,
    _as_device_tensor,
    _default_device,
)
from .registry import IdentityTransition, LinearNoiseTransition, LinearTransition  # noqa: F401  (re-export)

TWO_PI = 2.0 * math.pi


class AbstractParticleFilter:
    """abstract_particle_filter.py:14-129 (+ abstract_filter.py:7-37)."""

    _default_shift = True

    def __init__(self, initial_filter_state=None):
        self._filter_state = initial_filter_state
        self.resampling = "multinomial"
        self.scan_exact = True
        self.validate = True
        self._alt = None
        self._logw = None
        self._cdf = None
        self._est = None
        self.last_log_max = None
        # predict is deferred until the particles are looked at, so that a following update runs the fused
        # predict+weight kernel (one pass over the particles instead of two); reading `filter_state`
        # materialises it, so the reference's observable semantics are unchanged
        self.lazy_predict = True
        self._pending = None
        # random gather of the multinomial resampling: a 24-byte record straddles a 64-byte DRAM fetch 25 % of the
        # time (40 / 48 bytes: 50 / 62 %).  When the update runs a pending predict, the predict kernel also leaves
        # the particles as records padded to 32 / 64 bytes (pfb_predict_params.x_pad_out) and the gather reads those.
        self.padded_gather = True
        self.padded_gather_min = 1 << 20  # particles; below, the extra buffer is not worth it
        self._pad = None

    # ---- state ---------------------------------------------------------------------------------
    def _flush(self):
        if self._pending is not None:
            p, noise = self._pending
            self._pending = None
            st = self._filter_state
            x = st._records()
            self._eng().predict(p, x, x, noise)  # in place
            st._d = x.reshape(st._d.shape)
            st._est_cache = None

    @property
    def filter_state(self):
        self._flush()
        return self._filter_state

    @filter_state.setter
    def filter_state(self, new_state):
        self._pending = None  # a new state replaces whatever a deferred predict would have produced
        self._set_filter_state(new_state)

    def _set_filter_state(self, new_state):
        """abstract_particle_filter.py:75-93."""
        if self._filter_state is None:
            self._filter_state = copy.deepcopy(new_state)
        elif isinstance(new_state, type(self.filter_state)):
            assert self.filter_state.d.shape == new_state.d.shape
            self._filter_state = copy.deepcopy(new_state)
        else:
            n = self.filter_state.w.shape[0]
            samples = new_state.sample(n)
            samples = _as_device_tensor(samples, dtype=self._filter_state.d.dtype, device=self._filter_state.device)
            assert samples.shape == self.filter_state.d.shape
            self._filter_state._d = samples
            self._filter_state._w = torch.full((n,), 1.0 / n, dtype=torch.float64, device=samples.device)
            self._filter_state._uniform = True
            self._filter_state._est_cache = None

    @property
    def dim(self):
        return self.filter_state.dim

    def _eng(self):
        return engine(self._filter_state.device)

    def _scratch(self):
        st = self._filter_state
        x = st._records()
        N = x.shape[0]
        if self._alt is None or self._alt.shape != x.shape or self._alt.dtype != x.dtype or self._alt.device != x.device:
            self._alt = torch.empty_like(x)
            self._logw = torch.empty(N, dtype=torch.float64, device=x.device)
            self._cdf = torch.empty(N, dtype=torch.float64, device=x.device)
            self._est = torch.empty(2 * L.PFB_MAXD, dtype=torch.float64, device=x.device)
        return x

    # ---- predict ----------------------------------------------------------------------------------
    def predict_identity(self, noise_distribution):
        """abstract_particle_filter.py:18-19 (dispatches the identity transition directly)."""
        self.predict_nonlinear(IdentityTransition(), noise_distribution)

    def _predict_params(self, f, noise_distribution, shift_instead_of_add):
        raise NotImplementedError

    def predict_nonlinear(self, f, noise_distribution=None, function_is_vectorized=True, shift_instead_of_add=None):
        """abstract_particle_filter.py:21-54: d <- f(d), then one noise draw per particle."""
        if shift_instead_of_add is None:
            shift_instead_of_add = self._default_shift
        self._flush()
        st = self._filter_state
        assert noise_distribution is None or self._noise_dim_ok(noise_distribution)
        self._pending = self._predict_params(f, noise_distribution, shift_instead_of_add)
        st._est_cache = None
        if not self.lazy_predict:
            self._flush()

    def _noise_dim_ok(self, noise):
        return self.filter_state.dim == noise.dim

    _nonadditive_mode = L.PRED_EUCLID

    def predict_nonlinear_nonadditive(self, f, samples, weights):
        """abstract_particle_filter.py:56-69: one noise sample per particle is drawn from the weighted sample set
        (random.choice(samples, n, p=weights)) and d <- f(d, noise).  f must be a registry
        LinearNoiseTransition(F, G, b): d <- d F^T + noise G^T + b (an arbitrary f(x, noise) is rejected)."""
        if not isinstance(f, LinearNoiseTransition):
            raise NotImplementedError(
                "predict_nonlinear_nonadditive only accepts LinearNoiseTransition(F, G, b); arbitrary Python "
                "callables are not run on the CPU.")
        st = self.filter_state  # materialises a deferred predict
        x = st._records()
        N, D = x.shape
        smp = _as_device_tensor(samples, dtype=torch.float64, device=x.device)
        smp = smp.reshape(smp.shape[0], -1)
        w = _as_device_tensor(weights, dtype=torch.float64, device=x.device).reshape(-1)
        assert smp.shape[0] == w.shape[0], "samples and weights must match in size"
        E = smp.shape[1]
        if f.F.shape[0] != D or f.G.shape[1] != E:
            raise ValueError("LinearNoiseTransition shapes do not match the state / noise dimensions")
        if E > D:
            raise NotImplementedError("noise samples with more components than the state are not supported yet")
        eng = self._eng()
        # 1. resample the noise sample set (multinomial, N draws) with the same CDF + search kernels as the particles
        cdf = torch.empty(w.shape[0], dtype=torch.float64, device=x.device)
        eng.scan(cdf, w=(w / torch.sum(w)).contiguous(), exact=self.scan_exact)
        pad = torch.zeros((smp.shape[0], D), dtype=x.dtype, device=x.device)
        pad[:, :E] = smp.to(x.dtype)
        noise = torch.empty((N, D), dtype=x.dtype, device=x.device)
        src = prandom.source()
        if src.fused:
            eng.resample_rng(cdf, src.seed, src.next_offset(), pad, noise, N, D)
        else:
            eng.resample(cdf, src.uniforms((N,), x.device), pad, noise, N, D)
        # 2. d <- d F^T + noise G^T + b in the predict kernel (A = G^T padded to D x D)
        p = L.PredictParams()
        p.mode, p.dim, p.noise_cols, p.has_F, p.has_A = self._nonadditive_mode, D, D, 1, 1
        p.F[: D * D] = f.F.reshape(-1).tolist()
        A = np.zeros((D, D))
        A[:E, :] = f.G.T
        p.A[: D * D] = A.reshape(-1).tolist()
        if f.b is not None:
            p.has_b = 1
            p.b[:D] = f.b.tolist()
        eng.predict(p, x, x, noise)
        st._d = x.reshape(st._d.shape)
        st._est_cache = None

    def _normals_or_key(self, p, src, shape):
        """standard normals: injected / torch-made array, or the in-kernel Philox key (production)"""
        if src.fused:
            p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_NORMAL, src.seed, src.next_offset()
            return None
        return src.normals(shape, self._filter_state.device, self._filter_state.d.dtype)

    def _fill_transition(self, p, f):
        D = p.dim
        F, b = R.transition_params(f, D)
        if F is not None:
            p.has_F = 1
            p.F[: D * D] = np.asarray(F, dtype=np.float64).reshape(-1).tolist()
        if b is not None:
            p.has_b = 1
            p.b[:D] = np.asarray(b, dtype=np.float64).reshape(-1).tolist()

    # ---- update -----------------------------------------------------------------------------------
    def update_identity(self, meas_noise, measurement, shift_instead_of_add=True):
        """abstract_particle_filter.py:95-109."""
        m = None if measurement is None else R._np(measurement)
        assert m is None or m.shape == (meas_noise.dim,) or m.shape == (getattr(meas_noise, "input_dim", meas_noise.dim),) \
            or meas_noise.dim == 1 and m.shape == ()
        if not shift_instead_of_add:
            raise NotImplementedError()
        likelihood = meas_noise.set_mode(m).pdf
        self.update_nonlinear_using_likelihood(likelihood)

    def update_nonlinear_using_likelihood(self, likelihood, measurement=None):
        """abstract_particle_filter.py:111-125: reweigh, resample N of N, uniform weights."""
        dist = R.as_distribution(likelihood)
        if measurement is not None:
            dist = copy.deepcopy(dist).set_mode(R._np(measurement))
        st = self._filter_state
        x = self._scratch()
        eng = self._eng()
        N, D = x.shape
        lp, keep = R.lik_params(dist, D, images_to_device=R.device_put(x.device))
        if self._pending is not None and st._uniform:
            p, noise = self._pending
            self._pending = None
            gather = self._padded_gather_target(p, x)
            # measured (T^3, 1e8): the von Mises sampler inside the fused kernel costs occupancy and
            # instruction-cache misses -- 5.7 ms fused against 1.9 (predict) + 2.8 (weight) as two kernels
            fuse = p.rng_kind != L.RNG_VONMISES or os.environ.get("PFB_FUSE_VM") == "1"
            if fuse:
                try:
                    # with a padded gather copy the compact predicted particles are dead (the resampling reads the
                    # copy and replaces the particle set): not stored, unless a failed weight check has to leave them
                    dead = gather is not None
                    eng.predict_loglik(p, lp, x, None if dead else x, noise, self._logw, scaled=True)  # fused K(1)+K(2), in place
                    self._predicted_only_in_pad = dead
                except NotImplementedError:  # (mode, likelihood) pair without a fused instantiation
                    fuse = False
            if not fuse:
                eng.predict(p, x, x, noise)
                eng.loglik(lp, x, self._logw, scaled=True)
        else:
            gather = None
            self._flush()
            eng.loglik(lp, x, self._logw, w_prev=st._weights_or_none(), scaled=True)
        self._resample_from_logw(x, keep, gather)

    def _est_kind(self):
        return L.EST_SUM

    def _padded_gather_target(self, p, x):
        """point the predict parameters at the padded copy, if it pays: returns (buffer, components per record)"""
        N, D = x.shape
        Dp = 4 if D <= 4 else 8
        if not (self.padded_gather and self.resampling == "multinomial" and N >= self.padded_gather_min
                and (D * x.element_size()) % 32 != 0 and D not in (1, 2, 4)):
            return None
        if self._pad is None or self._pad.shape != (N, Dp) or self._pad.dtype != x.dtype or self._pad.device != x.device:
            self._pad = torch.empty((N, Dp), dtype=x.dtype, device=x.device)
        p.x_pad_out = self._pad.data_ptr()
        p.x_pad_stride = Dp
        return self._pad, Dp

    def _resample_from_logw(self, x, keep=None, gather=None):
        st = self._filter_state
        eng = self._eng()
        N, D = x.shape
        if self.validate:
            stats = eng.stats.cpu()
            flags = int(stats[L.STAT_FLAGS])
            self.last_log_max = float(stats[L.STAT_MAX])
            if flags & 3 and getattr(self, "_predicted_only_in_pad", False) and gather is not None:
                x.copy_(gather[0][:, :D])  # the update fails: leave the predicted particles behind, like the reference
            self._predicted_only_in_pad = False
            assert not flags & 1, "All weights should be greater than or equal to 0."
            assert not flags & 2, "The sum of all weights should be greater than 0."
        self._predicted_only_in_pad = False
        eng.scan(self._cdf, logw=self._logw, exact=self.scan_exact)
        systematic = self.resampling == "systematic"
        if self.resampling not in ("multinomial", "systematic", "multinomial_sorted"):
            raise ValueError(f"unknown resampling scheme {self.resampling!r}")
        src = prandom.source()
        if self.resampling == "multinomial_sorted":
            # exact multinomial offspring counts, emitted in CDF order (pfb_resample_msort): the draws exist only in the
            # kernel, so there is no uniform stream to inject -- parity runs use "multinomial"
            if not src.fused:
                raise ValueError("resampling='multinomial_sorted' draws its uniforms in the kernel; injected uniforms "
                                 "need resampling='multinomial' (the reference's order)")
            eng.resample_msort(self._cdf, src.seed, src.next_offset(), x, self._alt, N, D, est_kind=self._est_kind(),
                               est_out=self._est)
            return self._finish_resample(x, keep)
        xg, stride = (gather[0], gather[1]) if (gather is not None and not systematic) else (x, 0)
        if src.fused:
            eng.resample_rng(self._cdf, src.seed, src.next_offset(), xg, self._alt, N, D, systematic=systematic,
                             est_kind=self._est_kind(), est_out=self._est, x_in_stride=stride)
        else:
            u = src.uniforms((1,) if systematic else (N,), x.device)
            eng.resample(self._cdf, u, xg, self._alt, N, D, systematic=systematic, est_kind=self._est_kind(),
                         est_out=self._est, x_in_stride=stride)
        return self._finish_resample(x, keep)

    def _finish_resample(self, x, keep=None):
        st = self._filter_state
        N, D = x.shape
        del keep
        new = self._alt
        self._alt = x
        st._d = new.reshape(st._d.shape)
        st._w = torch.full((N,), 1.0 / N, dtype=torch.float64, device=x.device) if not st._uniform else st._w
        st._uniform = True
        st._est_cache = self._estimate_from_sums(self._est, D)

    def _estimate_from_sums(self, est, D):
        return est[:D].clone()

    def association_likelihood(self, likelihood):
        """abstract_particle_filter.py:127-129: sum_i w_i pdf(d_i)."""
        st = self.filter_state
        dist = R.as_distribution(likelihood)
        x = st._records()
        lp, keep = R.lik_params(dist, x.shape[1], images_to_device=R.device_put(x.device))
        # log(w_i pdf_i) from the weighting kernel, then the log-sum-exp reduction of pfb_normalize:
        # sum_i w_i pdf_i = exp(max) * sum exp(l_i - max); no N-sized torch temporaries, nothing on the host
        eng = self._eng()
        N = x.shape[0]
        if self._logw is None or self._logw.shape[0] != N or self._logw.device != x.device:
            logw = torch.empty(N, dtype=torch.float64, device=x.device)
        else:
            logw = self._logw
        stats = torch.empty(L.STAT_COUNT, dtype=torch.float64, device=x.device)
        w_prev = st._weights_or_none()
        eng.loglik(lp, x, logw, w_prev=w_prev, stats=stats)
        eng.normalize(logw, stats=stats)
        del keep
        out = torch.exp(stats[L.STAT_MAX]) * stats[L.STAT_SUMEXP]
        out = torch.where(stats[L.STAT_MAX] == -math.inf, torch.zeros_like(out), out)  # pdf = 0 at every particle
        if w_prev is None:
            out = out / N  # uniform weights 1/N are not materialised
        return out

    # ---- estimate ----------------------------------------------------------------------------------
    def get_point_estimate(self):
        return self.filter_state.mean()


# -------------------------------------------------------------------------------------------------
class EuclideanParticleFilter(AbstractParticleFilter):
    """euclidean_particle_filter.py:18-69."""

    _default_shift = False

    def __init__(self, n_particles, dim, dtype=torch.float64, device=None):
        if not (isinstance(n_particles, int) and n_particles > 0):
            raise ValueError("n_particles must be a positive integer")
        if not (isinstance(dim, int) and dim > 0):
            raise ValueError("dim must be a positive integer")
        device = _default_device() if device is None else torch.device(device)
        initial = LinearDiracDistribution(torch.zeros((n_particles, dim), dtype=dtype, device=device))
        AbstractParticleFilter.__init__(self, initial)

    def _set_filter_state(self, new_state):
        if not isinstance(new_state, LinearDiracDistribution):
            dist_dirac = LinearDiracDistribution.from_distribution(new_state, self._filter_state.d.shape[0])
            dist_dirac._d = dist_dirac._d.to(self._filter_state.d.dtype)
        else:
            dist_dirac = copy.deepcopy(new_state)
        if self._filter_state.d.shape != dist_dirac.d.shape:
            raise ValueError("The shape of new state does not match with the existing state.")
        self._filter_state = dist_dirac

    def _predict_params(self, f, noise, shift):
        D = self._filter_state._stored_dim
        p = L.PredictParams()
        p.mode, p.dim = L.PRED_EUCLID, D
        self._fill_transition(p, f)
        kind, spec = R.predict_noise_spec(noise, D)
        if kind is None:
            return p, None
        if kind != "gauss":
            raise NotImplementedError("Euclidean filters take Gaussian system noise")
        N = self._filter_state.w.shape[0]
        src = prandom.source()
        z = self._normals_or_key(p, src, (N, D))
        p.noise_cols, p.has_A = D, 1
        p.A[: D * D] = spec["A"].reshape(-1).tolist()
        if not shift:  # add-mode keeps the noise mean; shift-mode replaces it by the particle
            p.has_mu_noise = 1
            p.mu_noise[:D] = spec["mu"].tolist()
        return p, z

    def get_point_estimate(self):
        """abstract_euclidean_filter.py:9-11."""
        return self.filter_state.mean()


class HypertoroidalParticleFilter(AbstractParticleFilter):
    """hypertoroidal_particle_filter.py:27-67."""

    _default_shift = True
    _nonadditive_mode = L.PRED_TORUS_ADD  # mod(f(x) + mod(noise G^T)) -- the state stays on the torus

    def __init__(self, n_particles, dim, dtype=torch.float64, device=None):
        device = _default_device() if device is None else torch.device(device)
        if dim == 1:
            points = torch.linspace(0.0, TWO_PI, n_particles + 1, dtype=torch.float64, device=device)[:-1]
        else:
            col = torch.arange(n_particles, dtype=torch.float64, device=device) * (TWO_PI / n_particles)
            points = col[:, None].expand(n_particles, dim)
        filter_state = HypertoroidalDiracDistribution(points.to(dtype).contiguous(), dim=dim)
        AbstractParticleFilter.__init__(self, filter_state)

    def _predict_params(self, f, noise, shift):
        st = self._filter_state
        D = st._stored_dim
        p = L.PredictParams()
        p.mode, p.dim = (L.PRED_TORUS_SHIFT if shift else L.PRED_TORUS_ADD), D
        self._fill_transition(p, f)
        kind, spec = R.predict_noise_spec(noise, D)
        if kind is None:
            return p, None
        N = st.w.shape[0]
        src = prandom.source()
        p.noise_cols = D
        if kind in ("wn", "gauss"):
            z = self._normals_or_key(p, src, (N, D))
            p.has_A = 1
            p.A[: D * D] = spec["A"].reshape(-1).tolist()
            mu = np.mod(spec["mu"], TWO_PI)
        elif kind == "vm":
            kap = spec["kappa"]
            kap = np.broadcast_to(kap, (D,)) if kap.shape[0] in (1, D) else kap
            if src.fused:
                p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VONMISES, src.seed, src.next_offset()
                p.set_von_mises_rng(np.array(kap, dtype=np.float64).tolist(), st.device)
                z = None
            else:
                z = src.vonmises((N, D), torch.as_tensor(np.array(kap, dtype=np.float64)), st.device, st.d.dtype)
            mu = np.mod(np.broadcast_to(spec["mu"], (D,)), TWO_PI)
        else:
            raise NotImplementedError(f"{kind} noise is not defined on the hypertorus")
        if not shift:
            p.has_mu_noise = 1
            p.mu_noise[:D] = np.asarray(mu, dtype=np.float64).tolist()
        return p, z

    def _noise_dim_ok(self, noise):
        return self.filter_state.dim == noise.dim or R.kind_of(noise) == "vm"

    def _est_kind(self):
        return L.EST_TRIG

    def _estimate_from_sums(self, est, D):
        t = est[: 2 * D].reshape(D, 2)
        m = torch.remainder(torch.atan2(t[:, 1], t[:, 0]), TWO_PI)
        return m[0] if self._filter_state._one_d else m

    def get_point_estimate(self):
        """abstract_hypertoroidal_filter.py:17-25."""
        return self.filter_state.mean_direction()


class CircularParticleFilter(HypertoroidalParticleFilter):
    """circular_particle_filter.py:13-37."""

    def __init__(self, n_particles, dtype=torch.float64, device=None):
        device = _default_device() if device is None else torch.device(device)
        pts = torch.linspace(0.0, TWO_PI, n_particles + 1, dtype=torch.float64, device=device)[:-1]
        AbstractParticleFilter.__init__(self, CircularDiracDistribution(pts.to(dtype).contiguous()))

    def compute_association_likelihood(self, likelihood):
        return self.association_likelihood(likelihood)


class ToroidalParticleFilter(HypertoroidalParticleFilter):
    """toroidal_particle_filter.py:9-11."""

    def __init__(self, n_particles, dtype=torch.float64, device=None):
        HypertoroidalParticleFilter.__init__(self, n_particles, 2, dtype=dtype, device=device)


class HypersphericalParticleFilter(AbstractParticleFilter):
    """New (the reference raises "HypersphericalParticleFilter not implemented yet",
    evaluation/configure_for_filter.py:90-100).  State: HypersphericalDiracDistribution on S^dim,
    stored as (N, dim + 1) unit vectors.  Predict follows the sphere pattern of
    hyperhemisphere_cart_prod_particle_filter.py:90-95 (noise.set_mode(f(d)[j]); d[j] = noise.sample(1))
    with von Mises-Fisher noise on S^2, or additive Gaussian noise in the embedding space followed by
    renormalisation.  Estimate = mean direction (hyperspherical_dirac_distribution.py:38-44)."""

    _default_shift = True

    def __init__(self, n_particles, dim, dtype=torch.float64, device=None):
        if not (isinstance(n_particles, int) and n_particles > 0):
            raise ValueError("n_particles must be a positive integer")
        if not (isinstance(dim, int) and dim > 0):
            raise ValueError("dim must be a positive integer")
        device = _default_device() if device is None else torch.device(device)
        d = torch.zeros((n_particles, dim + 1), dtype=dtype, device=device)
        d[:, -1] = 1.0
        AbstractParticleFilter.__init__(self, HypersphericalDiracDistribution(d))

    def _noise_dim_ok(self, noise):
        return noise.dim == self.filter_state.dim or getattr(noise, "dim", None) == self.filter_state.input_dim

    def _predict_params(self, f, noise, shift):
        st = self._filter_state
        D = st._stored_dim
        p = L.PredictParams()
        p.dim = D
        p.mode = L.PRED_EUCLID
        self._fill_transition(p, f)
        kind, spec = R.predict_noise_spec(noise, D)
        if kind is None:
            return p, None
        N = st.w.shape[0]
        src = prandom.source()
        if kind == "vmf":
            if D != 3:
                raise NotImplementedError("von Mises-Fisher system noise is implemented on S^2")
            p.mode, p.noise_cols = L.PRED_SPHERE_VMF, 3
            p.kappa = spec["kappa"]
            p.exp_m2k = float(np.exp(-2 * spec["kappa"]))
            if src.fused:
                p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VMF, src.seed, src.next_offset()
                return p, None
            return p, src.vmf_triples(N, st.device, st.d.dtype)
        if kind == "gauss":
            p.mode, p.noise_cols, p.has_A = L.PRED_SPHERE_ADD, D, 1
            p.A[: D * D] = spec["A"].reshape(-1).tolist()
            if np.any(spec["mu"] != 0):
                p.has_mu_noise = 1
                p.mu_noise[:D] = spec["mu"].tolist()
            return p, self._normals_or_key(p, src, (N, D))
        raise NotImplementedError(f"{kind} noise is not defined on the hypersphere")

    def _estimate_from_sums(self, est, D):
        r = est[:D]
        return r / torch.linalg.norm(r)

    def get_point_estimate(self):
        return self.filter_state.mean_direction()


class HyperhemisphericalParticleFilter(HypersphericalParticleFilter):
    """hyperhemispherical_particle_filter.py:13-46: particles on the upper hemisphere of S^dim (antipodes
    identified).  Same kernels as the hyperspherical filter; every predicted particle is mirrored onto
    x_last >= 0, and the point estimate is the mean axis (principal eigenvector of the scatter matrix)."""

    def __init__(self, n_particles, dim, dtype=torch.float64, device=None):
        HypersphericalParticleFilter.__init__(self, n_particles, dim, dtype=dtype, device=device)
        st = self._filter_state
        self._filter_state = HyperhemisphericalDiracDistribution(st._d)

    def set_state(self, new_state):
        """:35-46  a non-Dirac state is sampled; hyperspherical samples are mirrored onto the hemisphere"""
        if not isinstance(new_state, HyperhemisphericalDiracDistribution):
            samples = new_state.sample(self._filter_state.d.shape[0])
            samples = _as_device_tensor(samples, dtype=self._filter_state.d.dtype, device=self._filter_state.device)
            flip = samples[:, -1] < 0
            samples = torch.where(flip[:, None], -samples, samples)
            new_state = HyperhemisphericalDiracDistribution(samples)
        self.filter_state = new_state

    def _set_filter_state(self, new_state):
        if isinstance(new_state, HyperhemisphericalDiracDistribution):
            assert self._filter_state.d.shape == new_state.d.shape
            self._filter_state = copy.deepcopy(new_state)
        else:
            self.set_state(new_state)

    def _predict_params(self, f, noise, shift):
        p, draws = HypersphericalParticleFilter._predict_params(self, f, noise, shift)
        if p.mode in (L.PRED_SPHERE_VMF, L.PRED_SPHERE_ADD):
            p.renormalize |= 2
        return p, draws

    def _estimate_from_sums(self, est, D):
        return None  # the mean axis needs the scatter matrix, not the resultant: no cached estimate

    def get_point_estimate(self):
        return self.filter_state.mean_axis()


# -------------------------------------------------------------------------------------------------
class WrappedNormalFilter:
    """wrapped_normal_filter.py:10-32: a circular filter whose state is a wrapped normal; the particle-based update
    (update_nonlinear_particle) runs on the device path: samples of the state, CircularDiracDistribution.reweigh,
    moment matching."""

    def __init__(self, wn=None):
        self._filter_state = wn if wn is not None else WrappedNormalDistribution(np.array(0.0), np.array(1.0))

    @property
    def filter_state(self):
        return self._filter_state

    @filter_state.setter
    def filter_state(self, new_state):
        if not hasattr(new_state, "sigma"):
            raise ValueError("the state of a WrappedNormalFilter is a WrappedNormalDistribution")
        self._filter_state = copy.deepcopy(new_state)

    def predict_identity(self, wn_sys):
        """:18-20  convolution of two wrapped normals: means add (mod 2 pi), variances add."""
        a, b = self._filter_state, wn_sys
        mu = np.mod(float(np.asarray(a.mu).reshape(-1)[0]) + float(np.asarray(b.mu).reshape(-1)[0]), 2.0 * math.pi)
        sig = math.sqrt(float(np.asarray(a.sigma).reshape(-1)[0]) ** 2 + float(np.asarray(b.sigma).reshape(-1)[0]) ** 2)
        self._filter_state = WrappedNormalDistribution(np.array(mu), np.array(sig))

    def update_nonlinear_particle(self, likelihood, z, n=100):
        """:27-32  n = 100 samples of the state, reweighed by likelihood(z, .), matched by a wrapped normal.
        `likelihood` is a registry distribution (the reference's callable likelihood(z, x) = its pdf centred at z,
        evaluated at x); anything else is rejected, nothing is evaluated on the CPU."""
        dist = copy.deepcopy(R.as_distribution(likelihood)).set_mode(R._np(z))
        wd = CircularDiracDistribution(self._filter_state.sample(n))
        self._filter_state = wd.reweigh(dist).to_wn()

    def get_point_estimate(self):
        """abstract_circular_filter / abstract_filter: the mean direction of the state."""
        return float(np.asarray(self._filter_state.mu).reshape(-1)[0])


def __getattr__(name):
    """HyperhemisphereCartProdParticleFilter lives in cart_prod.py"""
    if name == "HyperhemisphereCartProdParticleFilter":
        from .cart_prod import HyperhemisphereCartProdParticleFilter

        return HyperhemisphereCartProdParticleFilter
    raise AttributeError(name)
