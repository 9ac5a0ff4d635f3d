"""Distributions of the particle-filter path, same names / constructor signatures / method
semantics as pyrecest.distributions (reference v0.9.1), state on the GPU.

Dirac (weighted-sample) distributions hold `d` as a row-major (N, D) torch CUDA tensor -- the
reference's own layout -- and `w` (N,) float64; every sum, resampling and moment runs in the
libpfb.so kernels.  The parametric kinds (Gaussian, hypertoroidal / circular wrapped normal,
von Mises, von Mises-Fisher) keep their few parameters on the host; their pdf / sample run on the
device.  Reference file:line per method is given in the docstrings (paths under pyrecest/).
"""
from __future__ import annotations

import copy
import math
import warnings

import numpy as np
import torch

from . import _lib as L
from . import random as prandom
from . import registry as R
from ._engine import engine

TWO_PI = 2.0 * math.pi
DEFAULT_DTYPE = torch.float64


def _default_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")


def _as_device_tensor(a, dtype=None, device=None):
    if isinstance(a, torch.Tensor):
        t = a
        if dtype is None:
            dtype = t.dtype if t.dtype in (torch.float32, torch.float64) else DEFAULT_DTYPE
    else:
        t = torch.as_tensor(np.asarray(a))
        if dtype is None:
            dtype = DEFAULT_DTYPE
    if device is None:
        device = t.device if t.is_cuda else _default_device()
    return t.to(device=device, dtype=dtype).contiguous()


def _host(a):
    return R._np(a)


# =================================================================================================
# parametric kinds (noise models / likelihoods / priors)
# =================================================================================================
class _RegistryDistribution:
    """pdf() on the device for every registry kind."""

    dim: int
    input_dim: int

    def _pdf_device(self, xs, stored_dim):
        x = _as_device_tensor(xs)
        x2 = x.reshape(-1, stored_dim).contiguous()
        eng = engine(x2.device)
        lp, keep = R.lik_params(self, stored_dim, images_to_device=R.device_put(x2.device))
        logw = torch.empty(x2.shape[0], dtype=torch.float64, device=x2.device)
        eng.loglik(lp, x2, logw)
        del keep
        return torch.exp(logw)

    def ln_pdf(self, xs):
        return torch.log(self.pdf(xs))


class GaussianDistribution(_RegistryDistribution):
    """nonperiodic/gaussian_distribution.py:13-135."""

    def __init__(self, mu, C, check_validity=True):
        mu = _host(mu)
        C = _host(C)
        assert mu.ndim <= 1, "mu must be 1-dimensional"
        if mu.ndim == 0:
            mu = mu.reshape((1,))
            C = C.reshape((1, 1))
        assert C.ndim == 2, "C must be 2-dimensional"
        assert 1 == mu.shape[0] == C.shape[0] or mu.shape[0] == C.shape[0] == C.shape[1], "Size of C invalid"
        self.dim = mu.shape[0]
        self.input_dim = self.dim
        self.mu = mu
        if check_validity:
            if self.dim == 1:
                assert C > 0, "C must be positive definite"
            elif self.dim == 2:
                assert C[0, 0] > 0.0 and np.linalg.det(C) > 0.0, "C must be positive definite"
            else:
                np.linalg.cholesky(C)
        self.C = C

    def set_mean(self, new_mean):
        new_dist = copy.deepcopy(self)
        new_dist.mu = _host(new_mean).reshape(-1)
        return new_dist

    set_mode = set_mean

    def shift(self, shift_by):
        new = copy.deepcopy(self)
        new.mu = self.mu + _host(shift_by).reshape(-1)
        return new

    def mean(self):
        return self.mu

    def mode(self):
        return self.mu

    def covariance(self):
        return self.C

    def pdf(self, xs):
        xs_t = _as_device_tensor(xs)
        assert self.dim == 1 and xs_t.ndim <= 1 or xs_t.shape[-1] == self.dim, "Dimension incorrect"
        return self._pdf_device(xs_t, self.dim)

    def sample(self, n):
        """random.multivariate_normal(mean, cov, n) (:122-123) == z @ A + mean, z ~ N(0, I)."""
        dev = _default_device()
        z = prandom.source().normals((int(n), self.dim), dev)
        A = torch.as_tensor(R.gaussian_noise_factor(self.C), device=dev)
        x = z @ A
        x += torch.as_tensor(self.mu, device=dev)
        return x


class HypertoroidalWrappedNormalDistribution(_RegistryDistribution):
    """hypertorus/hypertoroidal_wrapped_normal_distribution.py:27-163."""

    def __init__(self, mu, C):
        mu = _host(mu)
        C = _host(C)
        numel_mu = 1 if mu.ndim == 0 else mu.shape[0]
        assert C.ndim == 0 or C.ndim == 2 and C.shape[0] == C.shape[1], "C must be of shape (dim, dim)"
        assert np.allclose(C, C.T, atol=1e-8), "C must be symmetric"
        assert C.ndim == 0 and C > 0.0 or len(np.linalg.cholesky(np.atleast_2d(C))) > 0, "C must be positive definite"
        assert numel_mu == 1 or mu.shape == (C.shape[1],), "mu must be of shape (dim,)"
        self.dim = numel_mu
        self.input_dim = numel_mu
        self.mu = np.mod(mu, TWO_PI)
        self.C = C

    def set_mean(self, mu):
        dist = copy.deepcopy(self)
        dist.mu = np.mod(_host(mu), TWO_PI)
        return dist

    def set_mode(self, m):
        dist = copy.deepcopy(self)
        dist.mu = _host(m)
        return dist

    def mode(self):
        return self.mu

    def shift(self, shift_by):
        shift_by = _host(shift_by)
        assert shift_by.shape == (self.dim,)
        hd = self
        hd.mu = np.mod(self.mu + shift_by, TWO_PI)
        return hd

    def pdf(self, xs, m=3):
        if m != 3:
            raise NotImplementedError("the device likelihood implements the reference default m = 3")
        xs_t = _as_device_tensor(xs)
        assert xs_t.shape[-1] == self.dim or self.dim == 1, "Last dimension of xs must match the distribution dimension"
        return self._pdf_device(xs_t, self.dim)

    def sample(self, n):
        if n <= 0:
            raise ValueError("n must be a positive integer")
        dev = _default_device()
        z = prandom.source().normals((int(n), self.dim), dev)
        A = torch.as_tensor(R.gaussian_noise_factor(np.atleast_2d(self.C)), device=dev)
        s = z @ A
        s += torch.as_tensor(np.atleast_1d(self.mu), device=dev)
        out = torch.empty_like(s)
        engine(dev).wrap_2pi(s, out)
        return out

    def trigonometric_moment(self, n):
        assert isinstance(n, int), "n must be an integer"
        C = np.atleast_2d(self.C)
        mu = np.atleast_1d(self.mu)
        return np.exp(np.array([1j * n * mu[i] - n**2 * C[i, i] / 2 for i in range(self.dim)]))


HypertoroidalWNDistribution = HypertoroidalWrappedNormalDistribution


class WrappedNormalDistribution(HypertoroidalWrappedNormalDistribution):
    """circle/wrapped_normal_distribution.py:34-195 (1-D; pdf is the series summed until it stops changing)."""

    MAX_SIGMA_BEFORE_UNIFORM = 10

    def __init__(self, mu, sigma):
        sigma = _host(sigma)
        HypertoroidalWrappedNormalDistribution.__init__(self, _host(mu), sigma**2)

    @property
    def sigma(self):
        return np.sqrt(self.C)

    def pdf(self, xs):
        if float(np.asarray(self.sigma).reshape(-1)[0]) <= 0:
            raise ValueError(f"sigma must be >0, but received {self.sigma}.")
        xs_t = _as_device_tensor(xs)
        if xs_t.ndim == 0:
            xs_t = xs_t.reshape(1)
        return self._pdf_device(xs_t, 1)

    def sample(self, n):
        dev = _default_device()
        z = prandom.source().normals((int(n),), dev)
        s = float(np.asarray(self.mu).reshape(-1)[0]) + float(np.asarray(self.sigma).reshape(-1)[0]) * z
        out = torch.empty_like(s)
        engine(dev).wrap_2pi(s, out)
        return out

    def trigonometric_moment(self, n):
        return np.exp(1j * n * self.mu - n**2 * self.sigma**2 / 2)

    @staticmethod
    def from_moment(m):
        """:186-190  the wrapped normal with first trigonometric moment m."""
        m = complex(_host(m).reshape(-1)[0]) if not isinstance(m, complex) else m
        mu = np.mod(np.angle(m), 2.0 * np.pi)
        sigma = np.sqrt(-2.0 * np.log(np.abs(m)))
        return WrappedNormalDistribution(np.array(mu), np.array(sigma))


WNDistribution = WrappedNormalDistribution


class VonMisesDistribution(_RegistryDistribution):
    """circle/von_mises_distribution.py:26-165.  sample / set_mean / set_mode are additions: the
    reference's class has none, which is why von Mises system noise fails there."""

    def __init__(self, mu, kappa, norm_const=None):
        kappa = _host(kappa)
        assert np.all(kappa >= 0.0)
        self.dim = 1
        self.input_dim = 1
        self.mu = _host(mu)
        self.kappa = kappa
        self._norm_const = norm_const

    @property
    def norm_const(self):
        if self._norm_const is None:
            self._norm_const = TWO_PI * math.exp(R.log_i0(float(self.kappa)))
        return self._norm_const

    def pdf(self, xs):
        return self._pdf_device(_as_device_tensor(xs), 1)

    def set_mean(self, mu):
        d = copy.deepcopy(self)
        d.mu = np.mod(_host(mu), TWO_PI)
        return d

    def set_mode(self, mu):
        d = copy.deepcopy(self)
        d.mu = _host(mu)
        return d

    def sample(self, n):
        dev = _default_device()
        v = prandom.source().vonmises((int(n),), float(self.kappa), dev)
        s = v + float(np.asarray(self.mu).reshape(-1)[0])
        out = torch.empty_like(s)
        engine(dev).wrap_2pi(s, out)
        return out


VMDistribution = VonMisesDistribution


class VonMisesFisherDistribution(_RegistryDistribution):
    """hypersphere_subset/von_mises_fisher_distribution.py:31-199."""

    def __init__(self, mu, kappa):
        mu = _host(mu)
        kappa = float(_host(kappa))
        assert mu.shape[0] >= 2, "mu must be at least two-dimensional for the circular case"
        assert kappa > 0, "kappa must be a positive scalar"
        assert abs(np.linalg.norm(mu) - 1.0) < 1e-6, "mu must be a normalized"
        self.dim = mu.shape[0] - 1
        self.input_dim = mu.shape[0]
        self.mu = mu
        self.kappa = kappa
        self.C = math.exp(R.vmf_log_norm_const(kappa, mu.shape[0]))

    def pdf(self, xs):
        xs_t = _as_device_tensor(xs)
        assert xs_t.shape[-1] == self.input_dim
        return self._pdf_device(xs_t, self.input_dim)

    def mean_direction(self):
        return self.mu

    def mean(self):
        return self.mu

    def mode(self):
        return self.mu

    def set_mode(self, new_mode):
        new_mode = _host(new_mode)
        assert new_mode.shape == self.mu.shape
        self.mu = copy.deepcopy(new_mode)
        return self

    def set_mean(self, new_mean):
        return self.set_mode(new_mean)

    def sample(self, n):
        """scipy vonmises_fisher(mu, kappa).rvs(n) on S^2 (:59-81), from (u, z1, z2) triples."""
        dev = _default_device()
        n = int(n)
        if self.input_dim != 3:  # other dimensions: Wood's rejection sampler over all draws at once (cart_prod.vmf_rows)
            from .cart_prod import vmf_rows

            modes = torch.as_tensor(self.mu, dtype=torch.float64, device=dev).expand(n, self.input_dim).contiguous()
            return vmf_rows(modes, self.kappa)
        t = prandom.source().vmf_triples(n, dev)
        modes = torch.as_tensor(self.mu, device=dev).expand(n, 3).contiguous()
        out = torch.empty_like(modes)
        p = L.PredictParams()
        p.mode, p.dim, p.noise_cols = L.PRED_SPHERE_VMF, 3, 3
        p.kappa = self.kappa
        p.exp_m2k = float(np.exp(-2 * self.kappa))
        engine(dev).predict(p, modes, out, t)
        return out


VMFDistribution = VonMisesFisherDistribution


class HypertoroidalUniformDistribution:
    """hypertorus/hypertoroidal_uniform_distribution.py:65-72 (prior sampling only)."""

    def __init__(self, dim):
        self.dim = dim
        self.input_dim = dim

    def sample(self, n):
        dev = _default_device()
        return TWO_PI * prandom.source().uniforms((int(n), self.dim), dev)

    def pdf(self, xs):
        xs_t = _as_device_tensor(xs)
        return torch.full((xs_t.reshape(-1, self.dim).shape[0],), 1.0 / TWO_PI**self.dim, dtype=torch.float64, device=xs_t.device)


class CircularUniformDistribution(HypertoroidalUniformDistribution):
    """circle/circular_uniform_distribution.py: the uniform density 1 / (2 pi) on the circle."""

    def __init__(self):
        HypertoroidalUniformDistribution.__init__(self, 1)


class HypersphericalUniformDistribution:
    """hypersphere_subset/hyperspherical_uniform_distribution.py:36-53: normalised standard normals;
    pdf = 1 / surface area."""

    def __init__(self, dim):
        self.dim = dim
        self.input_dim = dim + 1

    def pdf(self, xs):
        xs_t = _as_device_tensor(xs)
        D = self.dim + 1
        area = 2.0 * math.pi ** (0.5 * D) / math.gamma(0.5 * D)
        return torch.full((xs_t.reshape(-1, D).shape[0],), 1.0 / area, dtype=torch.float64, device=xs_t.device)

    def sample(self, n):
        dev = _default_device()
        z = prandom.source().normals((int(n), self.dim + 1), dev)
        return z / torch.linalg.norm(z, dim=1, keepdim=True)


# =================================================================================================
# Dirac distributions
# =================================================================================================
class AbstractDiracDistribution:
    """abstract_dirac_distribution.py:22-139."""

    _manifold = "abstract"

    def __init__(self, d, w=None):
        d_t = _as_device_tensor(d)
        self._one_d = d_t.ndim == 1
        if w is None:
            w_t = torch.full((d_t.shape[0],), 1.0 / d_t.shape[0], dtype=torch.float64, device=d_t.device)
            self._uniform = True
        else:
            w_t = _as_device_tensor(w, dtype=torch.float64, device=d_t.device)
            self._uniform = False
        assert d_t.shape[0] == w_t.shape[0], "Number of Diracs and weights must match."
        self._d = d_t.clone() if isinstance(d, torch.Tensor) and d_t.data_ptr() == d.data_ptr() else d_t
        self._w = w_t.clone() if isinstance(w, torch.Tensor) and w_t.data_ptr() == w.data_ptr() else w_t
        self._est_cache = None
        if not self._uniform:
            self.normalize_in_place()

    # ---- storage ------------------------------------------------------------------------------
    @property
    def d(self):
        """The live particle array.  A filter that owns this distribution predicts in place and ping-pongs two buffers
        when it resamples, so a tensor taken from here is valid until the filter's next predict / update: clone() it to
        keep it (the reference returns fresh arrays; copying 2.4 GB per access would not be a drop-in either)."""
        return self._d

    @d.setter
    def d(self, value):
        t = _as_device_tensor(value, dtype=self._d.dtype if isinstance(value, torch.Tensor) is False else None,
                              device=self._d.device)
        if isinstance(value, torch.Tensor) and t.data_ptr() == value.data_ptr():
            t = t.clone()  # never alias a caller-owned tensor: the filter mutates its particle buffers in place
        self._d = t
        self._est_cache = None

    @property
    def w(self):
        return self._w

    @w.setter
    def w(self, value):
        self._w = _as_device_tensor(value, dtype=torch.float64, device=self._d.device)
        self._uniform = False
        self._est_cache = None

    @property
    def device(self):
        return self._d.device

    @property
    def _stored_dim(self):
        return 1 if self._d.ndim == 1 else self._d.shape[1]

    def _records(self):
        """(N, D) contiguous view of the Dirac locations."""
        d = self._d if self._d.ndim == 2 else self._d.reshape(-1, 1)
        return d if d.is_contiguous() else d.contiguous()

    def _weights_or_none(self):
        return None if self._uniform else self._w

    def _eng(self):
        return engine(self._d.device)

    # ---- reference API --------------------------------------------------------------------------
    def normalize_in_place(self):
        """:41-48  warn + renormalise if |sum w - 1| > 1e-10."""
        s = float(torch.sum(self._w))
        if not math.isclose(s, 1.0, abs_tol=1e-10):
            warnings.warn("Weights are not normalized.", RuntimeWarning)
            self._w = self._w / s

    def normalize(self):
        dist = copy.deepcopy(self)
        dist.normalize_in_place()
        return dist

    def apply_function(self, f, f_supports_multiple=True):
        """:55-67 -- registry transitions only (identity, LinearTransition)."""
        dist = copy.deepcopy(self)
        F, b = R.transition_params(f, self._stored_dim)
        dist._apply_transition(F, b)
        return dist

    def _predict_mode_plain(self):
        return L.PRED_EUCLID

    def _apply_transition(self, F, b):
        x = self._records()
        p = L.PredictParams()
        p.mode, p.dim, p.noise_cols = self._predict_mode_plain(), self._stored_dim, 0
        D = self._stored_dim
        if F is not None:
            p.has_F = 1
            p.F[: D * D] = np.asarray(F, dtype=np.float64).reshape(-1).tolist()
        if b is not None:
            p.has_b = 1
            p.b[:D] = np.asarray(b, dtype=np.float64).reshape(-1).tolist()
        out = torch.empty_like(x)
        self._eng().predict(p, x, out, None)
        self._d = out.reshape(self._d.shape)
        self._est_cache = None

    def _log_weights(self, likelihood):
        """log-likelihood weighting on the device; returns (logw, stats-after-normalise are in engine)."""
        dist_obj = R.as_distribution(likelihood)
        x = self._records()
        lp, keep = R.lik_params(dist_obj, self._stored_dim, images_to_device=R.device_put(x.device))
        logw = torch.empty(x.shape[0], dtype=torch.float64, device=x.device)
        self._eng().loglik(lp, x, logw, w_prev=self._weights_or_none())
        return logw, keep

    def reweigh(self, f):
        """:69-80  w <- f(d) * w / sum, evaluated as max / log-sum-exp normalisation of log f + log w.
        The reference's assertions (f >= 0, sum > 0) map to the device status word."""
        dist = copy.deepcopy(self)
        if isinstance(f, (torch.Tensor, np.ndarray)):
            # likelihood VALUES f(d_i) computed by the caller (a likelihood outside the device registry): (N,) >= 0
            vals = _as_device_tensor(f, dtype=torch.float64, device=self.device).reshape(-1)
            assert vals.shape[0] == self._w.shape[0], "Function returned wrong number of outputs."
            logw = torch.log(vals) + torch.log(self._w)
            keep = None
        else:
            logw, keep = self._log_weights(f)
        eng = self._eng()
        w_new = torch.empty_like(logw)
        eng.normalize(logw, w_new)
        flags = int(eng.stats[L.STAT_FLAGS].item())
        assert not flags & 1, "All weights should be greater than or equal to 0."
        assert not flags & 2, "The sum of all weights should be greater than 0."
        dist._w = w_new
        dist._uniform = False
        dist._est_cache = None
        return dist

    def sample(self, n, exact=True):
        """:82-84  random.choice(d, n, p=w): multinomial draw of n Dirac locations (n may differ from N)."""
        n = int(n)
        x = self._records()
        eng = self._eng()
        N = x.shape[0]
        cdf = torch.empty(N, dtype=torch.float64, device=x.device)
        w = self._w
        eng.scan(cdf, w=w.contiguous(), exact=exact)
        src = prandom.source()
        if x.shape[1] > L.PFB_MAXD:
            # records wider than the gather kernels' PFB_MAXD components (SE(3): 7): the kernel yields the indices,
            # the rows are picked by torch
            idx = torch.empty(n, dtype=torch.int64, device=x.device)
            if src.fused:
                eng.resample_rng(cdf, src.seed, src.next_offset(), None, None, n, 1, idx_out=idx)
            else:
                eng.resample(cdf, src.uniforms((n,), x.device), None, None, n, 1, idx_out=idx)
            return x.index_select(0, idx)
        out = torch.empty((n, x.shape[1]), dtype=x.dtype, device=x.device)
        if src.fused:
            eng.resample_rng(cdf, src.seed, src.next_offset(), x, out, n, x.shape[1])
        else:
            eng.resample(cdf, src.uniforms((n,), x.device), x, out, n, x.shape[1])
        return out.reshape(n) if self._one_d else out

    def entropy(self):
        warnings.warn("Entropy is not defined in a continuous sense")
        return -torch.sum(self._w * torch.log(self._w))

    def integrate(self, left=None, right=None):
        assert left is None and right is None, "Must overwrite in child class to use integral limits"
        return torch.sum(self._w)

    def pdf(self, _):
        raise NotImplementedError("PDF:UNDEFINED, pdf is not defined")

    def mode(self, rel_tol=0.001):
        highest_val, ind = torch.max(self._w, dim=0)
        if float(highest_val) / self._w.numel() < (1 + rel_tol):
            warnings.warn("The samples may be equally weighted, .mode is likely to return a bad result.")
        return self._d[ind]

    @classmethod
    def from_distribution(cls, distribution, n_particles):
        return cls(distribution.sample(n_particles))

    # ---- device moments ---------------------------------------------------------------------------
    def _moment(self, kind, order=1, center=None):
        x = self._records()
        return self._eng().moments(kind, x, x.shape[1], w=self._weights_or_none(), order=order, center=center)


class LinearDiracDistribution(AbstractDiracDistribution):
    """nonperiodic/linear_dirac_distribution.py:13-82."""

    _manifold = "euclid"

    def __init__(self, d, w=None):
        AbstractDiracDistribution.__init__(self, d, w)
        self.dim = self._d.shape[1] if self._d.ndim > 1 else 1
        self.input_dim = self.dim

    def mean(self):
        """:19-21  w @ d."""
        if self._est_cache is not None:
            return self._est_cache
        D = self._stored_dim
        m = self._moment(L.MOM_MEAN)[:D]
        return m[0] if self._one_d else m.clone()

    def set_mean(self, new_mean):
        offset = _as_device_tensor(new_mean, dtype=self._d.dtype, device=self.device) - self.mean()
        self._d = self._d + offset.reshape(1, -1) if self._d.ndim == 2 else self._d + offset
        self._est_cache = None

    def covariance(self):
        """:27-29, 73-82  cov((d - mean).T, aweights=w, bias=True)."""
        D = self._stored_dim
        m = self._moment(L.MOM_MEAN)
        c = self._moment(L.MOM_CENTRAL, center=m)
        return c[: D * D].reshape(D, D).clone()

    @staticmethod
    def from_distribution(distribution, n_particles):
        samples = distribution.sample(n_particles)
        return LinearDiracDistribution(samples)

    @staticmethod
    def weighted_samples_to_mean_and_cov(samples, weights=None):
        ld = LinearDiracDistribution(samples, weights)
        return ld.mean(), ld.covariance()


class HypertoroidalDiracDistribution(AbstractDiracDistribution):
    """hypertorus/hypertoroidal_dirac_distribution.py:29-120."""

    _manifold = "torus"

    def __init__(self, d, w=None, dim=None):
        d_t = _as_device_tensor(d)
        if dim is None:
            if d_t.ndim > 1:
                dim = d_t.shape[-1]
            else:
                raise ValueError("Cannot automatically determine dimension.")
        self.dim = dim
        self.input_dim = dim
        wrapped = torch.empty_like(d_t)
        engine(d_t.device).wrap_2pi(d_t, wrapped)  # mod(d, 2 pi) on construction (:41)
        AbstractDiracDistribution.__init__(self, torch.atleast_1d(wrapped), w=w)

    def _predict_mode_plain(self):
        return L.PRED_TORUS_SHIFT  # noise-free: wraps f(d) like apply_function (:82-85)

    def set_mean(self, mean):
        dist = copy.deepcopy(self)
        shift = _as_device_tensor(mean, dtype=self._d.dtype, device=self.device) - dist.mean_direction()
        return dist.shift(shift)

    def mean_direction(self):
        """:61-70  mod(atan2(Im a, Re a), 2 pi), a = trigonometric_moment(1)."""
        if self._est_cache is not None:
            return self._est_cache
        a = self.trigonometric_moment(1)
        m = torch.remainder(torch.atan2(a.imag, a.real), TWO_PI)
        return m

    def mean(self):
        return self.mean_direction()

    def trigonometric_moment(self, n):
        """:72-80  sum_i w_i exp(1j n d_i) per dimension."""
        D = self._stored_dim
        t = self._moment(L.MOM_TRIG, order=int(n))[: 2 * D].reshape(D, 2)
        a = torch.complex(t[:, 0], t[:, 1])
        return a[0] if (self._one_d and D == 1) else a

    def shift(self, shift_by):
        shift_by = _as_device_tensor(shift_by, dtype=self._d.dtype, device=self.device)
        assert shift_by.shape[-1] == self.dim or self.dim == 1
        hd = copy.copy(self)
        tmp = self._d + (shift_by.reshape(1, -1) if self._d.ndim == 2 else shift_by.reshape(-1)[0])
        out = torch.empty_like(tmp)
        self._eng().wrap_2pi(tmp.contiguous(), out)
        hd._d = out
        hd._est_cache = None
        return hd

    def marginalize_to_1D(self, dimension):
        return CircularDiracDistribution(self._d[:, dimension].contiguous(), self._w)

    def to_toroidal_wd(self):
        assert self.dim == 2, "The dimension must be 2"
        return ToroidalDiracDistribution(self._d, self._w)


class CircularDiracDistribution(HypertoroidalDiracDistribution):
    """circle/circular_dirac_distribution.py:5-26."""

    def __init__(self, d, w=None):
        d_t = _as_device_tensor(d)
        HypertoroidalDiracDistribution.__init__(self, d_t, w, dim=1)
        assert w is None or tuple(d_t.squeeze().shape) == tuple(self._w.shape), "The shapes of d and w should match."

    def to_wn(self):
        """circle/abstract_circular_distribution.py:58-68: wrapped normal with the same first trigonometric moment
        (the moment is one launch of the moment kernel; two doubles come back)."""
        m = self.trigonometric_moment(1)
        return WrappedNormalDistribution.from_moment(complex(m.reshape(-1)[0].item()))


class ToroidalDiracDistribution(HypertoroidalDiracDistribution):
    """hypertorus/toroidal_dirac_distribution.py:9-55."""

    def __init__(self, d, w=None):
        HypertoroidalDiracDistribution.__init__(self, d, w, dim=2)

class LinBoundedCartProdDiracDistribution(AbstractDiracDistribution):
    """cart_prod/lin_bounded_cart_prod_dirac_distribution.py:14-52 -- bounded (periodic) dimensions first, then linear.
    All moments are taken on the whole (N, D) record with the kernels of the torus and the linear Diracs and the wanted
    components picked from the result: no marginal copies of the particle set are made."""

    _manifold = "cart_prod"

    def __init__(self, bound_dim, d, w=None):
        if not int(bound_dim) >= 1:
            raise ValueError("bound_dim must be a positive integer")
        AbstractDiracDistribution.__init__(self, d, w)
        assert self._d.ndim == 2 and self._d.shape[1] > int(bound_dim), "need at least one linear dimension"
        self.bound_dim = int(bound_dim)
        self.lin_dim = self._d.shape[1] - self.bound_dim
        self.dim = self._d.shape[1]
        self.input_dim = self.dim

    def marginalize_periodic(self):
        """:17-20"""
        return LinearDiracDistribution(self._d[:, self.bound_dim:].contiguous(), self._w.clone())

    def marginalize_linear(self):
        raise NotImplementedError

    def linear_mean(self):
        """:22-23"""
        return self._moment(L.MOM_MEAN)[self.bound_dim:self.dim].clone()

    def linear_covariance(self, approximate_mean=False):
        """:25-32"""
        if approximate_mean:
            warnings.warn("Formula for mean is trivial, so approximate_mean is ignored.")
        D, b = self.dim, self.bound_dim
        m = self._moment(L.MOM_MEAN)
        c = self._moment(L.MOM_CENTRAL, center=m)[: D * D].reshape(D, D)
        return c[b:, b:].clone()

    def hybrid_mean(self):
        """:37-41  concatenate(periodic.mean_direction(), linear.mean())"""
        D, b = self.dim, self.bound_dim
        t = self._moment(L.MOM_TRIG, order=1)[: 2 * D].reshape(D, 2).clone()
        per = torch.remainder(torch.atan2(t[:b, 1], t[:b, 0]), TWO_PI)
        return torch.cat((per, self.linear_mean()))

    def mean(self):
        """abstract_lin_bounded_cart_prod_distribution.py:35-43"""
        return self.hybrid_mean()

    @classmethod
    def from_distribution(cls, distribution, n_particles):
        """:43-52"""
        return cls(distribution.bound_dim, distribution.sample(n_particles))


class HypercylindricalDiracDistribution(LinBoundedCartProdDiracDistribution):
    """cart_prod/hypercylindrical_dirac_distribution.py:18-49"""

    _manifold = "cylinder"

    def marginalize_linear(self):
        """:33-36"""
        return HypertoroidalDiracDistribution(self._d[:, : self.bound_dim].contiguous(), self._w.clone())

    def hybrid_moment(self):
        """:38-49  w @ [cos(periodic) | sin(periodic) | linear]"""
        D, b = self.dim, self.bound_dim
        t = self._moment(L.MOM_TRIG, order=1)[: 2 * D].reshape(D, 2).clone()
        return torch.cat((t[:b, 0], t[:b, 1], self.linear_mean()))


class AbstractHypersphereSubsetDiracDistribution(AbstractDiracDistribution):
    """hypersphere_subset/abstract_hypersphere_subset_dirac_distribution.py:11-33."""

    _manifold = "sphere"

    def __init__(self, d, w=None):
        AbstractDiracDistribution.__init__(self, d, w=w)
        self.dim = self._d.shape[-1] - 1
        self.input_dim = self._d.shape[-1]

    def _predict_mode_plain(self):
        return L.PRED_EUCLID

    def moment(self):
        """:18-26  sum_i w_i d_i d_i^T (a Python loop over particles in the reference)."""
        D = self._stored_dim
        return self._moment(L.MOM_SCATTER)[: D * D].reshape(D, D).clone()

    def entropy(self):
        return -torch.sum(self._w * torch.log(self._w))

    def mean_axis(self):
        """abstract_hypersphere_subset_distribution.py:207-225: principal eigenvector of the scatter matrix
        (the moment() kernel), signed so that its last component is >= 0."""
        mom = self.moment()
        evals, evecs = torch.linalg.eigh(0.5 * (mom + mom.T))  # ascending
        m = evecs[:, -1]
        return m if float(m[-1]) >= 0 else -m


class HyperhemisphericalDiracDistribution(AbstractHypersphereSubsetDiracDistribution):
    """hypersphere_subset/hyperhemispherical_dirac_distribution.py: Diracs on the upper hemisphere, antipodes
    identified; mean() is the mean axis (abstract_hyperhemispherical_distribution.py:31-39)."""

    def mean(self):
        return self.mean_axis()


class HypersphericalDiracDistribution(AbstractHypersphereSubsetDiracDistribution):
    """hypersphere_subset/hyperspherical_dirac_distribution.py:14-44."""

    def mean_resultant_vector(self):
        """:43-44  sum_i w_i d_i."""
        D = self._stored_dim
        return self._moment(L.MOM_MEAN)[:D].clone()

    def mean_direction(self):
        """:38-41  r / |r|."""
        if self._est_cache is not None:
            return self._est_cache
        r = self.mean_resultant_vector()
        return r / torch.linalg.norm(r)

    def mean(self):
        """abstract_hyperspherical_distribution.py:37-45."""
        return self.mean_direction()


class LinHypersphereCartProdDiracDistribution(LinBoundedCartProdDiracDistribution):
    """cart_prod/lin_hypersphere_cart_prod_dirac_distribution.py:11-25 -- a unit vector of bound_dim + 1 components
    first, then linear components.  The records can be wider than the moment / gather kernels' PFB_MAXD (SE(3): 7), so
    every moment is taken on a marginal copy, as the reference does."""

    _manifold = "lin_hypersphere"

    def __init__(self, bound_dim, d, w=None):
        d_t = _as_device_tensor(d)
        nrm = torch.linalg.norm(d_t[:, : int(bound_dim) + 1], dim=-1)
        assert float((nrm - 1.0).abs().max()) < 1e-5, "The hypersphere ssubset part of d must be normalized"
        AbstractDiracDistribution.__init__(self, d, w)
        self.bound_dim = int(bound_dim)
        self.lin_dim = self._d.shape[1] - self.bound_dim - 1
        self.dim = self.bound_dim + self.lin_dim
        self.input_dim = self.dim + 1

    def marginalize_periodic(self):
        """lin_bounded_cart_prod_dirac_distribution.py:17-20: d[:, bound_dim:] -- for a hypersphere part of
        bound_dim + 1 components this keeps the unit vector's LAST component next to the linear ones; the reference
        slices there, and so do linear_mean() / linear_covariance() built on it."""
        return LinearDiracDistribution(self._d[:, self.bound_dim:].contiguous(), None if self._uniform else self._w.clone())

    def linear_mean(self):
        return self.marginalize_periodic().mean()

    def linear_covariance(self, approximate_mean=False):
        if approximate_mean:
            warnings.warn("Formula for mean is trivial, so approximate_mean is ignored.")
        return self.marginalize_periodic().covariance()

    def marginalize_linear(self):
        return HyperhemisphericalDiracDistribution(self._d[:, : self.bound_dim + 1].contiguous(),
                                                   None if self._uniform else self._w.clone())

    def hybrid_mean(self):
        """mean axis of the unit-vector part, then the mean of the linear components proper.  PARITY UNPINNED: the
        reference's hybrid_mean() raises ValueError for these classes (its hemispherical mean_direction is broken)."""
        lin = LinearDiracDistribution(self._d[:, self.bound_dim + 1:].contiguous(), None if self._uniform else self._w.clone())
        return torch.cat((self.marginalize_linear().mean(), lin.mean()))


class SE3DiracDistribution(LinHypersphereCartProdDiracDistribution):
    """se3_dirac_distribution.py:13-48: (N, 7) records, unit quaternion (last component >= 0 by convention) first,
    translation last."""

    def __init__(self, d, w=None):
        LinHypersphereCartProdDiracDistribution.__init__(self, 3, d, w)

    def mean(self):
        return self.hybrid_mean()

    @staticmethod
    def from_distribution(distribution, n_particles):
        assert isinstance(n_particles, int) and n_particles > 0, "n_particles must be a positive integer"
        return SE3DiracDistribution(distribution.sample(n_particles))


def __getattr__(name):
    """the Watson / hemisphere Cartesian-product classes live in cart_prod.py (which imports this module)"""
    if name in ("WatsonDistribution", "HyperhemisphericalWatsonDistribution", "HyperhemisphereCartProdDiracDistribution"):
        from . import cart_prod

        return getattr(cart_prod, name)
    raise AttributeError(name)
