"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the particle-filter hot path.

A vectorised numpy restatement of pyRecEst v0.9.1's predict / update / resample / estimate
arithmetic (reference tree /root/reference, citations below are relative to it).  It exists
to CHECK the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

Parity pins (see tests/test_oracle_golden.py and tests/golden/make_golden.py):
  * every function below is compared with the real reference, run in the build container
    through oracle/ref_harness.py with injected draws; the inputs/outputs of those runs are
    committed under tests/golden/*.npz;
  * the reference's own known-answer values are asserted directly.
PARITY UNPINNED (absent or broken in the reference, defined here -- SURVEY.md 8c):
  von Mises noise on the torus, VMF noise through a HypersphericalParticleFilter,
  additive+renormalise noise on the sphere, systematic resampling.  Their restatements
  follow the reference's add/shift-then-wrap pattern and scipy's S^2 VMF sampler.

All randomness is injected: standard normals Z, uniforms U, VMF triples (u, z1, z2).
"""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi  # the reference writes `2.0 * pi` with math.pi everywhere


# --------------------------------------------------------------------------------------
# noise helpers
# --------------------------------------------------------------------------------------
def gaussian_noise_factor(C):
    """A such that a legacy numpy multivariate_normal draw is Z @ A + mean.

    numpy RandomState.multivariate_normal: (u, s, v) = svd(cov); x = dot(Z, sqrt(s)[:,None]*v).
    Called from pyrecest/distributions/nonperiodic/gaussian_distribution.py:122-123 and
    pyrecest/distributions/hypertorus/hypertoroidal_wrapped_normal_distribution.py:113.
    """
    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    (_, s, v) = np.linalg.svd(C)
    return np.sqrt(s)[:, None] * v


def gaussian_rows(Z, C, mean=None):
    """Z (N, d) standard normals -> N draws of N(mean, C), numpy-legacy arithmetic."""
    Z = np.asarray(Z, dtype=np.float64)
    x = np.dot(Z.reshape(-1, Z.shape[-1]), gaussian_noise_factor(C))
    if mean is not None:
        x += np.asarray(mean, dtype=np.float64)
    return x


def apply_transition(d, F=None, b=None):
    """Registry transitions: identity (F is None) or linear x @ F.T + b."""
    if F is None:
        out = np.array(d, dtype=np.float64, copy=True)
    else:
        out = np.asarray(d, dtype=np.float64) @ np.asarray(F, dtype=np.float64).T
    if b is not None:
        out = out + np.asarray(b, dtype=np.float64)
    return out


# --------------------------------------------------------------------------------------
# predict  (pyrecest/filters/abstract_particle_filter.py:21-54)
# --------------------------------------------------------------------------------------
def predict_euclid(d, noise_rows, F=None, b=None):
    """EuclideanParticleFilter.predict_nonlinear, add-mode
    (pyrecest/filters/euclidean_particle_filter.py:59-69; abstract_particle_filter.py:44-47):
    d[i] <- f(d)[i] + noise.sample(1);  noise_rows are the N samples (mean included)."""
    fd = apply_transition(d, F, b)
    if noise_rows is None:
        return fd
    return fd + np.asarray(noise_rows, dtype=np.float64).reshape(fd.shape)


def predict_torus(d, noise_rows, F=None, b=None, shift=True, noise_mean=None):
    """HypertoroidalParticleFilter.predict_nonlinear
    (pyrecest/filters/hypertoroidal_particle_filter.py:43-56).

    noise_rows: zero-mean draws in angle space (Z @ A for a wrapped normal, von Mises draws
    in [-pi, pi] for von Mises noise), shape of d.
    shift-mode (default, abstract_particle_filter.py:49-50): the noise is re-centred at the
    particle: set_mean wraps the particle (hypertoroidal_wrapped_normal_distribution.py:62),
    sample() adds the mean to the draw and wraps (:113-115), the filter wraps once more (:56).
    add-mode (:44-47): d[i] + sample(1) with sample = mod(draw + noise_mean), then the final mod.
    """
    fd = apply_transition(d, F, b)
    if noise_rows is None:
        return np.mod(fd, TWO_PI)
    n = np.asarray(noise_rows, dtype=np.float64).reshape(fd.shape)
    if shift:
        mean = np.mod(fd, TWO_PI)
        s = np.mod(n + mean, TWO_PI)
        return np.mod(s, TWO_PI)
    mu = 0.0 if noise_mean is None else np.mod(np.asarray(noise_mean, dtype=np.float64), TWO_PI)
    s = np.mod(n + mu, TWO_PI)
    return np.mod(fd + s, TWO_PI)


def vmf_canonical_s2(triples, kappa):
    """scipy.stats.vonmises_fisher._rvs_3d (scipy/stats/_multivariate.py): a VMF(e1, kappa)
    draw on S^2 from (u, z1, z2): x = 1 + log(u + (1-u) e^{-2k})/k, (y, z) on the circle of
    radius sqrt(1-x^2) in direction (z1, z2)/|(z1, z2)| (_sample_uniform_direction)."""
    t = np.asarray(triples, dtype=np.float64).reshape(-1, 3)
    u = t[:, 0]
    x = 1.0 + np.log(u + (1.0 - u) * np.exp(-2 * kappa)) / kappa
    temp = np.sqrt(1.0 - np.square(x))
    zz = t[:, 1:3]
    nrm = np.sqrt(np.sum(np.square(zz), axis=1))  # np.linalg.norm(axis=-1) for float64
    circ = zz / nrm[:, None]
    return np.stack([x, temp * circ[:, 0], temp * circ[:, 1]], axis=-1)


def householder_from_e1(mu):
    """The orthogonal map scipy's _rotate_samples applies: Q (times rotsign) from the QR of the
    matrix whose first column is mu.  Closed form (LAPACK dgeqrf/dlarfg convention):
    beta = -sign(mu0)|mu|, tau = (beta - mu0)/beta, v = [1, mu[1:]/(mu0 - beta)],
    H = I - tau v v^T, Q[:,0] = H e1 = mu/beta (so Q e1 = -sign(mu0) mu/|mu|) and
    rotsign = +1 iff Q e1 ~ mu.  The product rotsign * H is returned as a dense (D, D)."""
    mu = np.asarray(mu, dtype=np.float64)
    D = mu.shape[0]
    alpha = mu[0]
    xnorm = np.sqrt(np.sum(np.square(mu[1:])))
    if xnorm == 0.0:
        H = np.eye(D)
        q0 = H[:, 0]
    else:
        beta = -np.copysign(np.sqrt(alpha * alpha + xnorm * xnorm), alpha)
        tau = (beta - alpha) / beta
        v = np.concatenate([[1.0], mu[1:] / (alpha - beta)])
        H = np.eye(D) - tau * np.outer(v, v)
        q0 = H[:, 0]
    rotsign = 1.0 if np.allclose(q0, mu) else -1.0
    return rotsign * H


def rotate_from_e1(samples, mu):
    """Row-wise: rotsign * (I - tau v v^T) s, evaluated per particle like the device does:
    s - tau (v.s) v.  mu may be (D,) or (N, D) (one mode per particle)."""
    s = np.asarray(samples, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    if mu.ndim == 1:
        mu = np.broadcast_to(mu, s.shape)
    alpha = mu[:, 0]
    xn2 = np.sum(np.square(mu[:, 1:]), axis=1)
    nrm = np.sqrt(alpha * alpha + xn2)
    beta = -np.copysign(nrm, alpha)
    degenerate = xn2 == 0.0
    safe_beta = np.where(degenerate, 1.0, beta)
    tau = np.where(degenerate, 0.0, (beta - alpha) / safe_beta)
    denom = np.where(degenerate, 1.0, alpha - beta)
    v = np.concatenate([np.ones((s.shape[0], 1)), mu[:, 1:] / denom[:, None]], axis=1)
    vs = np.sum(v * s, axis=1)
    out = s - (tau * vs)[:, None] * v
    # rotsign: H e1 = mu/beta -> equals mu only if beta ~ +1 i.e. mu0 < 0 (or degenerate, mu0>0)
    q0_is_mu = np.where(degenerate, alpha > 0.0, beta > 0.0)
    return np.where(q0_is_mu[:, None], out, -out)


def predict_sphere_vmf(d, triples, kappa, renormalize=False):
    """Sphere pattern of pyrecest/filters/hyperhemisphere_cart_prod_particle_filter.py:90-95:
    noise.set_mode(f(d)[j]); d[j] = noise.sample(1), with VonMisesFisherDistribution.sample ->
    scipy vonmises_fisher(mu, kappa).rvs(1)
    (pyrecest/distributions/hypersphere_subset/von_mises_fisher_distribution.py:59-81).
    S^2 only (D = 3).  PARITY UNPINNED as a filter; the sampler itself is pinned against scipy."""
    d = np.asarray(d, dtype=np.float64)
    out = rotate_from_e1(vmf_canonical_s2(triples, kappa), d)
    if renormalize:
        out = out / np.sqrt(np.sum(out * out, axis=1))[:, None]
    return out


def predict_sphere_additive(d, noise_rows):
    """x <- (x + n)/|x + n|  (additive noise in the embedding space, projected back onto S^d).
    PARITY UNPINNED (no reference equivalent)."""
    y = np.asarray(d, dtype=np.float64) + np.asarray(noise_rows, dtype=np.float64)
    return y / np.sqrt(np.sum(y * y, axis=1))[:, None]


# --------------------------------------------------------------------------------------
# likelihoods
# --------------------------------------------------------------------------------------
_LOG_2PI = np.log(2 * np.pi)


def gaussian_whitening(C):
    """(W, log_pdet) as scipy.stats.multivariate_normal uses them (_PSD): eigh(C),
    W = V / sqrt(lambda), log_pdet = sum log lambda.  Used by GaussianDistribution.pdf
    (pyrecest/distributions/nonperiodic/gaussian_distribution.py:43-50)."""
    import scipy.linalg

    C = np.atleast_2d(np.asarray(C, dtype=np.float64))
    s, u = scipy.linalg.eigh(C, lower=True, check_finite=False)  # scipy/stats/_multivariate.py _PSD
    return np.multiply(u, np.sqrt(1.0 / s)), float(np.sum(np.log(s)))


def logpdf_gauss(x, mu, C):
    """scipy multivariate_normal._logpdf: -0.5 (d log 2pi + log|C| + |(x - mu) W|^2)."""
    x = np.asarray(x, dtype=np.float64)
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    if x.ndim == 1:
        x = x.reshape(-1, mu.shape[0])
    W, logdet = gaussian_whitening(C)
    dev = x - mu
    maha = np.sum(np.square(np.dot(dev, W)), axis=-1)
    return -0.5 * (mu.shape[0] * _LOG_2PI + logdet + maha)


def pdf_gauss(x, mu, C):
    return np.exp(logpdf_gauss(x, mu, C))


def pdf_wn_multi(x, mu, C, m=3):
    """HypertoroidalWrappedNormalDistribution.pdf
    (pyrecest/distributions/hypertorus/hypertoroidal_wrapped_normal_distribution.py:65-93):
    xs <- (xs + pi) % 2pi - pi, then the sum over all (2m+1)^d offsets of mvn.pdf(xs + 2 pi k)."""
    mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
    d = mu.shape[0]
    xs = np.atleast_2d(np.asarray(x, dtype=np.float64))
    xs = (xs + np.pi) % (2 * np.pi) - np.pi
    offs = [np.arange(-m, m + 1) * 2.0 * np.pi for _ in range(d)]
    combos = np.stack(np.meshgrid(*offs, indexing="ij"), -1).reshape(-1, d)
    out = np.zeros(xs.shape[0])
    for off in combos:
        out += pdf_gauss(xs + off[None, :], mu, C)
    return out


def pdf_wn_1d(x, mu, sigma, max_iterations=1000):
    """WrappedNormalDistribution.pdf (pyrecest/distributions/circle/wrapped_normal_distribution.py:59-127):
    per element, add exp(tp)+exp(tm) for k = 1, 2, ... until the fp sum stops changing."""
    xs = np.atleast_1d(np.asarray(x, dtype=np.float64))
    if sigma > 10:
        return 1.0 / (2.0 * np.pi) * np.ones(xs.shape[0])
    xx = np.mod(xs, 2.0 * np.pi)
    xx = np.where(xx < 0, xx + 2.0 * np.pi, xx)
    xx = xx - mu
    tmp = -1.0 / (2.0 * sigma**2)
    nc = 1.0 / np.sqrt(2.0 * np.pi) / sigma
    result = np.exp(xx * xx * tmp)
    active = np.ones(xs.shape[0], dtype=bool)
    for k in range(1, max_iterations + 1):
        xp = xx + 2 * np.pi * k
        xm = xx - 2 * np.pi * k
        new = result + (np.exp(xp * xp * tmp) + np.exp(xm * xm * tmp))
        changed = active & (new != result)
        result = np.where(active, new, result)
        active = changed
        if not active.any():
            break
    return result * nc


def pdf_vm(x, mu, kappa):
    """VonMisesDistribution.pdf (pyrecest/distributions/circle/von_mises_distribution.py:42-50)."""
    from scipy.special import iv

    x = np.asarray(x, dtype=np.float64)
    return np.exp(kappa * np.cos(x - mu)) / (2.0 * np.pi * iv(0, kappa))


def vmf_norm_const(kappa, D):
    """VonMisesFisherDistribution.__init__ (von_mises_fisher_distribution.py:41-49)."""
    from scipy.special import iv

    dim = D - 1
    if dim == 2:
        return kappa / (4 * np.pi * np.sinh(kappa))
    return kappa ** ((dim + 1) / 2.0 - 1) / ((2.0 * np.pi) ** ((dim + 1) / 2.0) * iv((dim + 1) / 2 - 1, kappa))


def pdf_vmf(x, mu, kappa):
    """VonMisesFisherDistribution.pdf (:51-54): C exp(kappa mu^T x)."""
    x = np.asarray(x, dtype=np.float64)
    mu = np.asarray(mu, dtype=np.float64)
    return vmf_norm_const(kappa, mu.shape[0]) * np.exp(kappa * mu.T @ x.T)


LIKELIHOODS = {
    "gauss": lambda x, p: pdf_gauss(x, p["mu"], p["C"]),
    "wn_multi": lambda x, p: pdf_wn_multi(x, p["mu"], p["C"], p.get("m", 3)),
    "wn_1d": lambda x, p: pdf_wn_1d(x, p["mu"], p["sigma"]),
    "vm": lambda x, p: pdf_vm(x, p["mu"], p["kappa"]),
    "vmf": lambda x, p: pdf_vmf(x, p["mu"], p["kappa"]),
}


def likelihood(kind, x, params):
    return LIKELIHOODS[kind](x, params)


# --------------------------------------------------------------------------------------
# update: reweigh + resample
# --------------------------------------------------------------------------------------
def reweigh(w, lik):
    """AbstractDiracDistribution.reweigh (pyrecest/distributions/abstract_dirac_distribution.py:69-80)."""
    w = np.asarray(w, dtype=np.float64)
    lik = np.asarray(lik, dtype=np.float64)
    assert lik.shape == w.shape, "Function returned wrong output dimensions."
    assert np.all(lik >= 0), "All weights should be greater than or equal to 0."
    assert np.sum(lik) > 0, "The sum of all weights should be greater than 0."
    wn = lik * w
    return wn / np.sum(wn)


def cdf_of(w):
    """numpy Generator.choice: cdf = p.cumsum(); cdf /= cdf[-1]  (numpy/random/_generator.pyx),
    reached through pyrecest/_backend/_shared_numpy/random.py:28-29.  np.cumsum is strictly
    sequential: c_i = fl(c_{i-1} + p_i)."""
    cdf = np.cumsum(np.asarray(w, dtype=np.float64))
    cdf /= cdf[-1]
    return cdf


def multinomial_indices(w, u):
    """idx_j = #{i : cdf_i <= u_j} (searchsorted side='right'), int64."""
    return np.searchsorted(cdf_of(w), np.asarray(u, dtype=np.float64), side="right").astype(np.int64)


def systematic_uniforms(u0, n):
    """u_j = (j + u0)/n in fp64 (one add, one divide).  PARITY UNPINNED (not in the reference)."""
    return (np.arange(n, dtype=np.float64) + np.float64(u0)) / np.float64(n)


def systematic_indices(w, u0, n=None):
    """Systematic resampling on the reference's CDF; indices clamped to N-1."""
    w = np.asarray(w, dtype=np.float64)
    n = w.shape[0] if n is None else n
    idx = np.searchsorted(cdf_of(w), systematic_uniforms(u0, n), side="right").astype(np.int64)
    return np.minimum(idx, w.shape[0] - 1)


def sample_dirac(d, w, u):
    """AbstractDiracDistribution.sample (abstract_dirac_distribution.py:82-84)."""
    return np.asarray(d)[multinomial_indices(w, u)]


def update_resample(d, w, lik, u):
    """AbstractParticleFilter.update_nonlinear_using_likelihood
    (pyrecest/filters/abstract_particle_filter.py:111-125): reweigh, sample N of N, uniform weights.
    Returns (d_new, w_new, idx, w_posterior)."""
    wp = reweigh(w, lik)
    idx = multinomial_indices(wp, u)
    N = wp.shape[0]
    return np.asarray(d)[idx], (1 / N) * np.ones_like(wp), idx, wp


# --------------------------------------------------------------------------------------
# sharded filter -- PARITY UNPINNED: the reference has no multi-device path.
# Definition: every shard g (a contiguous range of particles) forms its own sequential cumsum c^g of
# p = exp(logw - M), M the global max; shard offsets are added sequentially in rank order,
# O_0 = 0, O_{g+1} = fl(O_g + c^g[-1]); the global CDF is v_i = fl(fl(O_g + c^g_i) / O_G).  With one shard
# this is exactly cdf_of().  Systematic resampling then runs on that global CDF.
# --------------------------------------------------------------------------------------
def sharded_cdf(p_shards):
    cs = [np.cumsum(np.asarray(p, dtype=np.float64)) for p in p_shards]
    offs = [0.0]
    for c in cs:
        offs.append(offs[-1] + (c[-1] if c.size else 0.0))
    total = offs[-1]
    v = np.concatenate([(offs[g] + cs[g]) / total for g in range(len(cs))])
    return v, np.array(offs), cs


def sharded_systematic_indices(p_shards, u0, n_out=None):
    v, _, _ = sharded_cdf(p_shards)
    n_out = v.shape[0] if n_out is None else n_out
    idx = np.searchsorted(v, systematic_uniforms(u0, n_out), side="right").astype(np.int64)
    return np.minimum(idx, v.shape[0] - 1)


def offspring_boundaries(offs, u0, n_out):
    """B_g = #{j : u_j < fl(O_g / O_G)}: offspring [B_g, B_{g+1}) of the comb come from shard g."""
    total = offs[-1]
    B = []
    for g in range(len(offs)):
        E = offs[g] / total
        j = int(np.floor(E * n_out - u0))
        j = min(max(j, 0), n_out)
        while j < n_out and (np.float64(j) + np.float64(u0)) / np.float64(n_out) < E:
            j += 1
        while j > 0 and (np.float64(j - 1) + np.float64(u0)) / np.float64(n_out) >= E:
            j -= 1
        B.append(j)
    B[-1] = n_out
    return B


# --------------------------------------------------------------------------------------
# moments / estimates
# --------------------------------------------------------------------------------------
def uniform_w(N):
    return np.ones(N) / N


def mean_linear(d, w):
    """LinearDiracDistribution.mean (pyrecest/distributions/nonperiodic/linear_dirac_distribution.py:19-21)."""
    return np.asarray(w) @ np.asarray(d)


def covariance_linear(d, w):
    """weighted_samples_to_mean_and_cov (:73-82): cov((d - m).T, aweights=w, bias=True)."""
    d = np.asarray(d, dtype=np.float64)
    if d.ndim == 1:
        d = d[:, None]
    m = w @ d
    return np.atleast_2d(np.cov((d - m).T, aweights=w, bias=True))


def trig_moment(d, w, n=1):
    """HypertoroidalDiracDistribution.trigonometric_moment
    (pyrecest/distributions/hypertorus/hypertoroidal_dirac_distribution.py:72-80)."""
    d = np.asarray(d, dtype=np.float64)
    if d.ndim == 1:
        d = d[:, None]
    dim = d.shape[1]
    return np.sum(np.exp(1j * n * d.T) * np.tile(w, (dim, 1)), axis=1)


def mean_direction_torus(d, w):
    """HypertoroidalDiracDistribution.mean_direction (:61-70)."""
    a = trig_moment(d, w, 1)
    return np.mod(np.arctan2(np.imag(a), np.real(a)), TWO_PI)


def hybrid_moment(d, w, bound_dim):
    """HypercylindricalDiracDistribution.hybrid_moment
    (pyrecest/distributions/cart_prod/hypercylindrical_dirac_distribution.py:39-50): w @ [cos(periodic) | sin(periodic) | linear]."""
    d = np.asarray(d, dtype=np.float64)
    S = np.column_stack((np.cos(d[:, :bound_dim]), np.sin(d[:, :bound_dim]), d[:, bound_dim:]))
    return np.asarray(w) @ S


def hybrid_mean(d, w, bound_dim):
    """LinBoundedCartProdDiracDistribution.hybrid_mean
    (pyrecest/distributions/cart_prod/lin_bounded_cart_prod_dirac_distribution.py:37-41): circular mean direction of the
    periodic part (the part is a HypertoroidalDiracDistribution, whose constructor wraps it, :41 there) and weighted mean
    of the linear part."""
    d = np.asarray(d, dtype=np.float64)
    return np.concatenate((mean_direction_torus(np.mod(d[:, :bound_dim], TWO_PI), w), mean_linear(d[:, bound_dim:], w)))


def angular_error(alpha, beta):
    """AbstractHypertoroidalDistribution.angular_error
    (pyrecest/distributions/hypertorus/abstract_hypertoroidal_distribution.py:146-167)."""
    diff = np.abs(np.mod(alpha, TWO_PI) - np.mod(beta, TWO_PI))
    return np.minimum(diff, TWO_PI - diff)


def distance(manifold_name, est, true):
    """pyrecest/evaluation/get_distance_function.py:14-56 for one (estimate, groundtruth) pair."""
    if "circle" in manifold_name or "hypertorus" in manifold_name:
        return np.linalg.norm(angular_error(est, true))
    if "hypersphere" in manifold_name:
        return np.arccos(np.dot(est, true))
    if "uclidean" in manifold_name:
        return np.linalg.norm(est - true)
    raise NotImplementedError(manifold_name)


def summarize(manifold_name, last_estimates, groundtruths, runtimes, run_failed):
    """determine_all_deviations (pyrecest/evaluation/determine_all_deviations.py:9-52) + summarize_filter_results
    (pyrecest/evaluation/summarize_filter_results.py:40-58): rows (error_mean, error_std, time_mean, failure_rate)."""
    dev = np.array([[distance(manifold_name, e, g[-1]) for e, g in zip(est_c, groundtruths)] for est_c in last_estimates])
    return dev, np.column_stack([dev.mean(axis=1), dev.std(axis=1), np.mean(runtimes, axis=1),
                                 np.sum(run_failed, axis=1) / run_failed.shape[1]])


def mean_resultant(d, w):
    """HypersphericalDiracDistribution.mean_resultant_vector
    (pyrecest/distributions/hypersphere_subset/hyperspherical_dirac_distribution.py:43-44)."""
    return np.sum(np.asarray(d) * np.reshape(w, (-1, 1)), axis=0)


def mean_direction_sphere(d, w):
    """HypersphericalDiracDistribution.mean_direction (:38-41)."""
    r = mean_resultant(d, w)
    return r / np.linalg.norm(r)


def scatter(d, w):
    """AbstractHypersphereSubsetDiracDistribution.moment
    (pyrecest/distributions/hypersphere_subset/abstract_hypersphere_subset_dirac_distribution.py:18-26):
    sum_i w_i d_i d_i^T (the reference loops over i)."""
    d = np.asarray(d, dtype=np.float64)
    return (d * np.asarray(w)[:, None]).T @ d


def mean_axis(d, w):
    """AbstractHypersphereSubsetDistribution._compute_mean_axis_from_moment
    (pyrecest/distributions/hypersphere_subset/abstract_hypersphere_subset_distribution.py:207-225): eigenvector
    of the largest eigenvalue of the scatter matrix, signed so that its last component is >= 0."""
    D, V = np.linalg.eig(scatter(d, w))
    D, V = np.real(D), np.real(V)
    m = V[:, np.argsort(D)][:, -1]
    return m if m[-1] >= 0 else -m


def predict_nonadditive(d, samples, weights, u, F, G, b=None, torus=False):
    """AbstractParticleFilter.predict_nonlinear_nonadditive (pyrecest/filters/abstract_particle_filter.py:56-69):
    noise_samples = choice(samples, n, p=weights / sum(weights)); d = f(d, noise) with the registry form
    f(x, n) = x F^T + n G^T + b.  torus=True wraps like the hypertoroidal filters do (PARITY UNPINNED: the
    reference's hypertoroidal override, hypertoroidal_particle_filter.py:58-67, does not run)."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / np.sum(w)
    noise = np.asarray(samples, dtype=np.float64).reshape(w.shape[0], -1)[multinomial_indices(w, u)]
    fd = np.asarray(d, dtype=np.float64) @ np.asarray(F).T
    if b is not None:
        fd = fd + np.asarray(b)
    gn = noise @ np.asarray(G).T
    if torus:
        return np.mod(fd + np.mod(gn, TWO_PI), TWO_PI)
    return fd + gn


def association_likelihood(lik, w):
    """AbstractParticleFilter.association_likelihood (abstract_particle_filter.py:127-129)."""
    return np.sum(lik * w)


# --------------------------------------------------------------------------------------
# whole step, as the bench / smoke use it
# --------------------------------------------------------------------------------------
def step(manifold, d, noise, lik_kind, lik_params, u, F=None, b=None, kappa_noise=None,
         resampling="multinomial"):
    """predict -> update -> resample -> estimate on one manifold ("euclid" | "torus" | "sphere").
    `noise`: additive rows (euclid), zero-mean angle rows (torus), VMF triples (sphere).
    `u`: (N,) uniforms for multinomial, scalar u0 for systematic."""
    if manifold == "euclid":
        dp = predict_euclid(d, noise, F, b)
    elif manifold == "torus":
        dp = predict_torus(d, noise, F, b, shift=True)
    elif manifold == "sphere":
        dp = predict_sphere_vmf(d, noise, kappa_noise)
    else:
        raise ValueError(manifold)
    N = dp.shape[0]
    lik = likelihood(lik_kind, dp, lik_params)
    wp = reweigh(uniform_w(N), lik)
    if resampling == "multinomial":
        idx = multinomial_indices(wp, u)
    else:
        idx = systematic_indices(wp, u)
    dn = dp[idx]
    wn = uniform_w(N)
    if manifold == "euclid":
        est = mean_linear(dn, wn)
    elif manifold == "torus":
        est = mean_direction_torus(dn, wn)
    else:
        est = mean_direction_sphere(dn, wn)
    return {"d_pred": dp, "lik": lik, "w_post": wp, "idx": idx, "d": dn, "est": est}
