"""resampling="multinomial_sorted" (pfb_resample_msort): exact multinomial offspring counts emitted in CDF order.
The draws are a pure function of (seed, offset, n_out) -- oracle/philox.py restates the popcount tree and the
in-leaf draws in numpy -- so the test is bit-exact: uniforms, indices (searchsorted of numpy.cumsum, side='right',
the reference's arithmetic: pyrecest/_backend/_shared_numpy/random.py:28-29) and gathered particles."""
import numpy as np
import pytest
import torch

from oracle import pf_oracle as O
from oracle import philox as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from pyrecest_b200._engine import engine

    return engine("cuda:0")


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda:0").contiguous()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("n_out", [1, 5, 32, 33, 64, 65, 513, 1000, 4097, 100_000, 1_000_003])
def test_uniforms_match_the_oracle_tree(eng, n_out):
    out = torch.empty(n_out, dtype=torch.float64, device="cuda:0")
    eng.msort_uniforms(out, 0xDEADBEEF12345, 7)
    u, leaf = P.msort_uniforms(0xDEADBEEF12345, 7, n_out)
    assert np.array_equal(host(out), u)
    K = 1 << P.msort_levels(n_out)
    assert np.array_equal(np.floor(u * K).astype(np.int64), leaf)  # slot order = leaf order
    assert np.all((u >= 0) & (u < 1))


def _weights(kind, n, rng):
    if kind == "lik":
        x = rng.standard_normal(n) * 2.5
        return np.exp(-0.5 * (x - 0.7) ** 2)
    if kind == "uniform":
        return np.full(n, 1.0 / n)
    if kind == "zeros":
        w = rng.random(n) * (rng.random(n) < 0.3)
        w[0] = 0.0
        w[-1] = 0.0
        w[n // 2] = 1.0
        return w
    if kind == "heavy":  # a few particles hold almost all the weight: they span many leaves and buckets
        w = rng.random(n) * 1e-9
        w[rng.integers(0, n, size=3)] = 1.0
        return w
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["lik", "uniform", "zeros", "heavy"])
@pytest.mark.parametrize("n,n_out", [(1, 1), (3, 1000), (1000, 3), (2049, 2049), (50_000, 50_000), (300_001, 200_000)])
def test_indices_and_particles_bitexact(eng, kind, n, n_out):
    from pyrecest_b200 import _lib as L

    rng = np.random.default_rng(n * 7 + n_out)
    w = _weights(kind, n, rng)
    if not w.sum() > 0:
        w[0] = 1.0
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    c = np.cumsum(w)
    assert np.array_equal(host(cdf), c)
    x = rng.random((n, 3)) * 2 * np.pi
    xd = dev(x)
    out = torch.empty((n_out, 3), dtype=torch.float64, device="cuda:0")
    idx = torch.empty(n_out, dtype=torch.int64, device="cuda:0")
    est = torch.zeros(2 * L.PFB_MAXD, dtype=torch.float64, device="cuda:0")
    seed, offset = 991, 12
    eng.resample_msort(cdf, seed, offset, xd, out, n_out, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est)
    u, _ = P.msort_uniforms(seed, offset, n_out)
    ref_idx = np.minimum(np.searchsorted(c / c[-1], u, side="right"), n - 1)
    assert np.array_equal(host(idx), ref_idx)
    assert np.array_equal(host(out), x[ref_idx])
    e = host(est)[:6].reshape(3, 2)
    np.testing.assert_allclose(e[:, 0], np.cos(x[ref_idx]).mean(axis=0), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(e[:, 1], np.sin(x[ref_idx]).mean(axis=0), rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 6])
def test_every_record_size_and_sum_estimate(eng, dtype, dim):
    from pyrecest_b200 import _lib as L

    rng = np.random.default_rng(dim)
    n = 70_001
    w = _weights("lik", n, rng)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    x = rng.standard_normal((n, dim))
    xd = dev(x, dtype)
    out = torch.empty((n, dim), dtype=dtype, device="cuda:0")
    est = torch.zeros(2 * L.PFB_MAXD, dtype=torch.float64, device="cuda:0")
    eng.resample_msort(cdf, 5, 1, xd, out, n, dim, est_kind=L.EST_SUM, est_out=est)
    u, _ = P.msort_uniforms(5, 1, n)
    c = np.cumsum(w)
    ref_idx = np.minimum(np.searchsorted(c / c[-1], u, side="right"), n - 1)
    xs = host(xd).astype(np.float64)
    assert np.array_equal(host(out).astype(np.float64), xs[ref_idx])
    np.testing.assert_allclose(host(est)[:dim], xs[ref_idx].mean(axis=0), rtol=1e-11, atol=1e-13)


def test_is_a_permutation_of_the_reference_order_draw(eng):
    """The sorted-order call equals pfb_resample (the reference-order search) fed with the same uniforms."""
    rng = np.random.default_rng(3)
    n = 123_457
    w = _weights("lik", n, rng)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    u = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.msort_uniforms(u, 77, 3)
    a = torch.empty(n, dtype=torch.int64, device="cuda:0")
    b = torch.empty(n, dtype=torch.int64, device="cuda:0")
    eng.resample_msort(cdf, 77, 3, None, None, n, 1, idx_out=a)
    eng.resample(cdf, u, None, None, n, 1, idx_out=b)
    assert torch.equal(a, b)
    # offspring counts follow the weights: chi-square of the counts over 64 blocks of sources
    counts = np.bincount(host(a) // (n // 64 + 1), minlength=64).astype(np.float64)
    p = np.add.reduceat(w, np.arange(0, n, n // 64 + 1)) / w.sum()
    chi2 = ((counts - n * p) ** 2 / (n * p)).sum()
    assert chi2 < 64 + 6 * np.sqrt(2 * 64), chi2


def test_filter_mode_multinomial_sorted(eng):
    """public API: the filter step with resampling='multinomial_sorted' reproduces the oracle step fed with the
    oracle's restatement of the in-kernel uniforms (and refuses injected uniforms)."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter

    N = 40_000
    rng = np.random.default_rng(8)
    d0 = rng.standard_normal((N, 2))
    pf = EuclideanParticleFilter(N, 2)
    pf.resampling = "multinomial_sorted"
    pf.filter_state.d = d0
    prandom.seed(4242)
    src = prandom.source()
    z = np.array([0.3, -0.2])
    pf.update_identity(GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)), z)
    used_offset = src._offset
    u, _ = P.msort_uniforms(src.seed, used_offset, N)
    lik = O.likelihood("gauss", d0, {"mu": z, "C": 0.5 * np.eye(2)})
    wp = O.reweigh(O.uniform_w(N), lik)
    idx = O.multinomial_indices(wp, u)
    got = host(pf.filter_state.d)
    differ = int(np.any(got != d0[idx], axis=1).sum())
    assert differ <= 2, differ  # (the filter scans exp(l - max), the oracle w / sum(w): see test_gpu_parity_scale.py)
    np.testing.assert_allclose(host(pf.get_point_estimate()), d0[idx].mean(axis=0), rtol=1e-9)
    with prandom.inject(uniforms=rng.random(N)):
        with pytest.raises(ValueError):
            pf.update_identity(GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)), z)
