"""the hemisphere Cartesian-product particle filter with von Mises-Fisher / Watson system noise
(pyrecest/filters/hyperhemisphere_cart_prod_particle_filter.py:23-118,
pyrecest/distributions/cart_prod/hyperhemisphere_cart_prod_dirac_distribution.py,
pyrecest/distributions/hypersphere_subset/{watson_distribution,hyperhemispherical_watson_distribution}.py).

A state is a tuple of n_hemispheres unit vectors of dim_hemisphere + 1 components each (e.g. a dual quaternion:
two points of the upper hemisphere of S^3), stored as one (N, (dim_hemisphere + 1) n_hemispheres) record.  The
reference predicts part by part in a Python loop over particles (`noise.set_mode(f(d)[j]); d[j] = noise.sample(1)`).
Here every part is predicted for all particles at once on the device:

  * von Mises-Fisher noise on S^2 parts: the predict kernel of the hyperspherical filter (scipy's `_rvs_3d`, parity
    pinned through the reference, tests/golden/vmf_sampler.npz);
  * von Mises-Fisher noise on S^d, d != 2: Wood's rejection sampler (the algorithm scipy uses there), and
  * Watson noise (any d): the angular-central-Gaussian rejection sampler of Kent, Ganeiber and Mardia,
    both as device tensor operations over all particles (draws from pyrecest_b200.random: normals and uniforms only),
    rotated from the canonical axis to every particle's own mode by a Householder reflection.

PARITY UNPINNED for the two rejection samplers: the reference draws through scipy's generator (VMF) and a
Metropolis-Hastings chain (Watson), whose streams cannot be injected; they are checked against the distributions'
moments instead (tests/test_gpu_filters.py).  Records are wider than the gather kernels' PFB_MAXD for S^3 x 2, so
resampling picks rows by the indices the resampling kernel returns.  Likelihoods: a tuple of registry distributions (one
per part, their log-likelihoods are added), or likelihood values computed by the caller; Python callables are
rejected like everywhere else."""
from __future__ import annotations

import copy
import math

import numpy as np
import torch

from . import _lib as L
from . import random as prandom
from . import registry as R
from .distributions import (
    AbstractDiracDistribution,
    HyperhemisphericalDiracDistribution,
    VonMisesFisherDistribution,
    _as_device_tensor,
    _default_device,
    _host,
)
from ._engine import engine


# ---------------------------------------------------------------------------------------------------------------
# row-wise samplers: one draw per mode, modes (N, q) unit vectors on the device
# ---------------------------------------------------------------------------------------------------------------
def _reflect_from_last_axis(x, modes):
    """x is expressed with the symmetry axis at e_q; the Householder reflection that takes e_q to modes[j] is applied
    to row j (both distributions are rotationally symmetric about their axis, so any orthogonal map e_q -> mode does)."""
    v = -modes.clone()
    v[:, -1] += 1.0                                   # v = e_q - mode
    vv = (v * v).sum(dim=1, keepdim=True)
    coef = torch.where(vv > 1e-30, 2.0 * (v * x).sum(dim=1, keepdim=True) / vv.clamp_min(1e-300), torch.zeros_like(vv))
    return x - coef * v


def _gamma_half_integer(a2, n, src, dev):
    """Gamma(a2 / 2, 1) draws (a2 a positive integer) from uniforms and normals only: a sum of floor(a2 / 2)
    exponentials, plus z^2 / 2 for odd a2."""
    k = a2 // 2
    g = torch.zeros(n, dtype=torch.float64, device=dev)
    if k > 0:
        u = 1.0 - src.uniforms((n, k), dev)            # (0, 1]
        g = -torch.log(u).sum(dim=1)
    if a2 % 2:
        z = src.normals((n,), dev)
        g = g + 0.5 * z * z
    return g


def vmf_rows(modes, kappa):
    """One von Mises-Fisher(modes[j], kappa) draw per row, any dimension (Wood 1994, as scipy's
    vonmises_fisher._rejection_sampling): w = mode . x from a Beta-based envelope, the rest uniform on S^(q-2)."""
    n, q = modes.shape
    dev = modes.device
    src = prandom.source()
    m = q - 1
    b = m / (2.0 * kappa + math.sqrt(4.0 * kappa * kappa + m * m))
    x0 = (1.0 - b) / (1.0 + b)
    c = kappa * x0 + m * math.log(1.0 - x0 * x0)
    w = torch.empty(n, dtype=torch.float64, device=dev)
    todo = torch.arange(n, device=dev)
    for _ in range(200):
        k = todo.numel()
        if k == 0:
            break
        g1, g2 = _gamma_half_integer(m, k, src, dev), _gamma_half_integer(m, k, src, dev)
        z = g1 / (g1 + g2)                             # Beta(m / 2, m / 2)
        cand = (1.0 - (1.0 + b) * z) / (1.0 - (1.0 - b) * z)
        u = 1.0 - src.uniforms((k,), dev)
        ok = kappa * cand + m * torch.log1p(-x0 * cand) - c >= torch.log(u)
        w[todo[ok]] = cand[ok]
        todo = todo[~ok]
    else:
        raise RuntimeError("von Mises-Fisher rejection sampler did not terminate")
    v = src.normals((n, q - 1), dev)
    v = v / torch.linalg.norm(v, dim=1, keepdim=True)
    x = torch.cat((torch.sqrt(torch.clamp(1.0 - w * w, min=0.0))[:, None] * v, w[:, None]), dim=1)
    return _reflect_from_last_axis(x, modes)


def watson_rows(modes, kappa):
    """One Watson(modes[j], kappa) draw per row, density prop. to exp(kappa (mode . x)^2), any dimension and sign of
    kappa: angular-central-Gaussian envelope (Kent, Ganeiber, Mardia 2013) for the Bingham form exp(-x^T A x) with
    A = kappa (I - mode mode^T) (kappa > 0) or A = |kappa| mode mode^T (kappa < 0)."""
    n, q = modes.shape
    dev = modes.device
    src = prandom.source()
    if kappa == 0.0:
        x = src.normals((n, q), dev)
        return x / torch.linalg.norm(x, dim=1, keepdim=True)
    ak = abs(kappa)
    mult = (q - 1) if kappa > 0 else 1                 # multiplicity of the eigenvalue |kappa| of A
    # b solves sum_i 1 / (b + 2 lambda_i) = 1:  (q - mult) / b + mult / (b + 2 ak) = 1
    lo, hi = 1e-12, float(q)
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if (q - mult) / mid + mult / (mid + 2.0 * ak) > 1.0:
            lo = mid
        else:
            hi = mid
    b = 0.5 * (lo + hi)
    om = 1.0 + 2.0 * ak / b                            # eigenvalue of Omega = I + 2 A / b on A's |kappa| eigenspace
    log_m = -0.5 * (q - b) + 0.5 * q * math.log(q / b)
    out = torch.empty((n, q), dtype=torch.float64, device=dev)
    todo = torch.arange(n, device=dev)
    for _ in range(500):
        k = todo.numel()
        if k == 0:
            break
        y = src.normals((k, q), dev)
        if kappa > 0:
            y[:, :-1] *= 1.0 / math.sqrt(om)           # the orthogonal complement of the axis carries the eigenvalue
        else:
            y[:, -1] *= 1.0 / math.sqrt(om)
        x = y / torch.linalg.norm(y, dim=1, keepdim=True)
        heavy = (x[:, :-1] ** 2).sum(dim=1) if kappa > 0 else x[:, -1] ** 2     # x^T A x / |kappa|
        xax = ak * heavy
        xox = 1.0 + (om - 1.0) * heavy
        u = 1.0 - src.uniforms((k,), dev)
        ok = torch.log(u) < -xax - log_m + 0.5 * q * torch.log(xox)
        out[todo[ok]] = x[ok]
        todo = todo[~ok]
    else:
        raise RuntimeError("Watson rejection sampler did not terminate")
    return _reflect_from_last_axis(out, modes)


# ---------------------------------------------------------------------------------------------------------------
# distributions
# ---------------------------------------------------------------------------------------------------------------
class WatsonDistribution:
    """hypersphere_subset/watson_distribution.py:27-130 (pdf, mode, set_mode, sample)."""

    EPSILON = 1e-6

    def __init__(self, mu, kappa):
        mu = _host(mu)
        assert mu.ndim == 1, "mu must be a 1-D vector"
        assert abs(np.linalg.norm(mu) - 1.0) < self.EPSILON, "mu is unnormalized"
        self.mu = mu
        self.kappa = float(_host(kappa))
        self.dim = mu.shape[0] - 1
        self.input_dim = mu.shape[0]

    @property
    def ln_norm_const(self):
        """:57-68  log Gamma(q / 2) - log(2 pi^(q / 2)) - log 1F1(1/2; q / 2; kappa)"""
        from scipy.special import hyp1f1

        q = self.input_dim
        return math.lgamma(0.5 * q) - math.log(2.0 * math.pi ** (0.5 * q)) - math.log(float(hyp1f1(0.5, 0.5 * q, self.kappa)))

    @property
    def norm_const(self):
        return math.exp(self.ln_norm_const)

    def ln_pdf(self, xs):
        xs_t = _as_device_tensor(xs)
        assert xs_t.shape[-1] == self.input_dim, "Last dimension of xs must be dim + 1"
        mu = torch.as_tensor(self.mu, dtype=xs_t.dtype, device=xs_t.device)
        return self.ln_norm_const + self.kappa * (xs_t @ mu) ** 2

    def pdf(self, xs):
        return torch.exp(self.ln_pdf(xs))

    def mode(self):
        if self.kappa >= 0:
            return self.mu
        raise NotImplementedError("the mode of a girdle (kappa < 0) Watson distribution is a great circle")

    def set_mode(self, new_mode):
        new_mode = _host(new_mode)
        assert new_mode.shape == self.mu.shape
        self.mu = copy.deepcopy(new_mode)
        return self

    def sample(self, n):
        dev = _default_device()
        modes = torch.as_tensor(self.mu, dtype=torch.float64, device=dev).expand(int(n), self.input_dim).contiguous()
        return watson_rows(modes, self.kappa)


class HyperhemisphericalWatsonDistribution:
    """hypersphere_subset/hyperhemispherical_watson_distribution.py:14-64: twice the Watson density on the upper
    hemisphere (last coordinate >= 0); samples of the full-sphere distribution are mirrored."""

    def __init__(self, mu, kappa):
        mu = _host(mu)
        assert mu[-1] >= 0
        self.dist_full_sphere = WatsonDistribution(mu, kappa)
        self.dim = self.dist_full_sphere.dim
        self.input_dim = self.dist_full_sphere.input_dim

    mu = property(lambda self: self.dist_full_sphere.mu, lambda self, v: setattr(self.dist_full_sphere, "mu", _host(v)))
    kappa = property(lambda self: self.dist_full_sphere.kappa, lambda self, v: setattr(self.dist_full_sphere, "kappa", float(v)))

    def pdf(self, xs):
        return 2.0 * self.dist_full_sphere.pdf(xs)

    def set_mode(self, mu):
        self.mu = mu
        return self

    def mode(self):
        return self.mu

    def sample(self, n):
        s = self.dist_full_sphere.sample(n)
        return torch.where((s[:, -1] < 0)[:, None], -s, s)


class HyperhemisphereCartProdDiracDistribution(AbstractDiracDistribution):
    """cart_prod/hyperhemisphere_cart_prod_dirac_distribution.py:7-41."""

    _manifold = "hemisphere_cart_prod"

    def __init__(self, d, w, dim_hemisphere, n_hemispheres):
        AbstractDiracDistribution.__init__(self, d, w)
        self.dim_hemisphere = int(dim_hemisphere)
        self.n_hemispheres = int(n_hemispheres)
        assert self._d.shape[-1] == (1 + self.dim_hemisphere) * self.n_hemispheres, "Dimension is not correct."
        self.dim = self.dim_hemisphere * self.n_hemispheres
        self.input_dim = self._d.shape[-1]

    def part(self, i):
        """the i-th hemisphere's unit vectors as a Dirac distribution of its own (a contiguous copy)"""
        q = self.dim_hemisphere + 1
        return HyperhemisphericalDiracDistribution(self._d[:, i * q:(i + 1) * q].contiguous(), None if self._uniform else self._w.clone())

    def mean_axes(self):
        """the mean axis of every part, concatenated (the reference defines no mean for this class)"""
        return torch.cat([self.part(i).mean_axis() for i in range(self.n_hemispheres)])


# ---------------------------------------------------------------------------------------------------------------
# the filter
# ---------------------------------------------------------------------------------------------------------------
class HyperhemisphereCartProdParticleFilter:
    """hyperhemisphere_cart_prod_particle_filter.py:23-118."""

    def __init__(self, n_particles, dim_hemisphere, n_hemispheres, dtype=torch.float64, device=None):
        if not (isinstance(n_particles, (int, np.integer)) and n_particles > 0):
            raise ValueError("n_particles must be a positive integer")
        dev = _default_device() if device is None else torch.device(device)
        q = dim_hemisphere + 1
        d = torch.zeros((int(n_particles), q * n_hemispheres), dtype=dtype, device=dev)
        d[:, q - 1::q] = 1.0  # (the reference leaves the array uninitialised; here every part starts at the pole)
        self._filter_state = HyperhemisphereCartProdDiracDistribution(d, None, dim_hemisphere, n_hemispheres)

    # -- state ---------------------------------------------------------------------------------------------
    @property
    def filter_state(self):
        return self._filter_state

    @filter_state.setter
    def filter_state(self, new_state):
        """:97-118  a hyper(hemi)spherical distribution of the parts' dimension is sampled n_particles x n_hemispheres
        times (spherical samples mirrored onto the upper hemisphere) and reshaped; a Dirac state of equal shape is
        deep-copied."""
        st = self._filter_state
        if isinstance(new_state, HyperhemisphereCartProdDiracDistribution):
            assert new_state.d.shape == st.d.shape
            self._filter_state = copy.deepcopy(new_state)
            return
        if not hasattr(new_state, "sample"):
            raise ValueError("the state must be a distribution")
        assert new_state.dim == st.dim_hemisphere
        n = st.d.shape[0]
        samples = _as_device_tensor(new_state.sample(n * st.n_hemispheres), dtype=st.d.dtype, device=st.device)
        samples = torch.where((samples[:, -1] < 0)[:, None], -samples, samples)
        st._d = samples.reshape(n, st.input_dim).contiguous()
        st._w = torch.full((n,), 1.0 / n, dtype=torch.float64, device=st.device)
        st._uniform = True

    def set_state(self, new_state):
        """:42-57"""
        self.filter_state = new_state

    # -- predict -------------------------------------------------------------------------------------------
    def predict_nonlinear_each_part(self, f, noise_distribution=None, function_is_vectorized=True, shift_instead_of_add=True):
        """:59-95  every hemisphere's part is moved by f (a registry transition; identity functions are recognised)
        and then replaced by a draw of the noise distribution centred at it."""
        assert function_is_vectorized, "Only vectorized functions are supported"
        st = self._filter_state
        assert noise_distribution is None or noise_distribution.dim == st.dim_hemisphere, \
            "Noise dimension must match state dimension in Cartesian product"
        assert shift_instead_of_add, "Only shifting is supported"
        q = st.dim_hemisphere + 1
        F, bvec = R.transition_params(f, q)
        kind = None
        if noise_distribution is not None:
            name = type(noise_distribution).__name__
            if R.kind_of(noise_distribution) == "vmf":
                kind = "vmf"
            elif "Watson" in name:
                kind = "watson"
            else:
                raise NotImplementedError("noise must be a VonMisesFisherDistribution or a HyperhemisphericalWatsonDistribution")
        for i in range(st.n_hemispheres):
            cur = st._d[:, i * q:(i + 1) * q].contiguous()
            if F is not None:
                cur = cur @ torch.as_tensor(np.asarray(F, dtype=np.float64), dtype=cur.dtype, device=cur.device).T
            if bvec is not None:
                cur = cur + torch.as_tensor(np.asarray(bvec, dtype=np.float64), dtype=cur.dtype, device=cur.device)
            if kind == "vmf":
                kappa = float(_host(noise_distribution.kappa))
                if q == 3:  # S^2: the predict kernel (scipy's sampler, parity-pinned)
                    p = L.PredictParams()
                    p.mode, p.dim, p.noise_cols = L.PRED_SPHERE_VMF, 3, 3
                    p.kappa, p.exp_m2k = kappa, float(np.exp(-2.0 * kappa))
                    cur = cur.contiguous()
                    out = torch.empty_like(cur)
                    engine(cur.device).predict(p, cur, out, prandom.source().vmf_triples(cur.shape[0], cur.device, cur.dtype))
                    cur = out
                else:
                    cur = vmf_rows(cur.to(torch.float64), kappa).to(cur.dtype)
            elif kind == "watson":
                cur = watson_rows(cur.to(torch.float64), float(_host(noise_distribution.kappa))).to(cur.dtype)
                if "Hyperhemispherical" in type(noise_distribution).__name__:
                    cur = torch.where((cur[:, -1] < 0)[:, None], -cur, cur)
            st._d[:, i * q:(i + 1) * q] = cur
        st._est_cache = None

    # -- update --------------------------------------------------------------------------------------------
    def update_nonlinear_using_likelihood(self, likelihood, measurement=None):
        """abstract_particle_filter.py:111-125: reweigh, resample n of n, uniform weights.  `likelihood`: a tuple of
        registry distributions, one per hemisphere (independent parts: the log-likelihoods add), or the likelihood
        VALUES at the particles (a tensor / array of n_particles non-negative numbers)."""
        st = self._filter_state
        n = st.d.shape[0]
        q = st.dim_hemisphere + 1
        eng = engine(st.device)
        if isinstance(likelihood, (torch.Tensor, np.ndarray)):
            vals = _as_device_tensor(likelihood, dtype=torch.float64, device=st.device).reshape(-1)
            assert vals.shape[0] == n, "Function returned wrong number of outputs."
            logw = torch.log(vals)
        elif isinstance(likelihood, (tuple, list)):
            if len(likelihood) != st.n_hemispheres:
                raise ValueError("one likelihood per hemisphere is needed")
            logw = torch.zeros(n, dtype=torch.float64, device=st.device)
            tmp = torch.empty(n, dtype=torch.float64, device=st.device)
            for i, lk in enumerate(likelihood):
                dist = R.as_distribution(lk)
                lp, keep = R.lik_params(dist, q, images_to_device=R.device_put(st.device))
                eng.loglik(lp, st._d[:, i * q:(i + 1) * q].contiguous(), tmp)
                logw += tmp
                del keep
        else:
            raise NotImplementedError("likelihood must be a tuple of registry distributions (one per hemisphere) or an array "
                                      "of likelihood values; arbitrary Python callables are rejected, not run on the CPU")
        if not st._uniform:
            logw = logw + torch.log(st._w)
        w = torch.empty_like(logw)
        eng.normalize(logw, w)
        flags = int(eng.stats[L.STAT_FLAGS].item())
        assert not flags & 1, "All weights should be greater than or equal to 0."
        assert not flags & 2, "The sum of all weights should be greater than 0."
        cdf = torch.empty(n, dtype=torch.float64, device=st.device)
        eng.scan(cdf, w=w, exact=True)
        idx = torch.empty(n, dtype=torch.int64, device=st.device)
        src = prandom.source()
        if src.fused:
            eng.resample_rng(cdf, src.seed, src.next_offset(), None, None, n, 1, idx_out=idx)
        else:
            eng.resample(cdf, src.uniforms((n,), st.device), None, None, n, 1, idx_out=idx)
        st._d = st._d.index_select(0, idx)  # (records of up to 8+ components: rows picked by the kernel's indices)
        st._w = torch.full((n,), 1.0 / n, dtype=torch.float64, device=st.device)
        st._uniform = True
        st._est_cache = None
        del measurement

    def get_point_estimate(self):
        """the mean axis of every hemisphere, concatenated (the reference's generic estimate, filter_state.mean(), does
        not exist for this state class)"""
        return self._filter_state.mean_axes()
