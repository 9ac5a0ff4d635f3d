"""Parity of the CUDA kernels, called through the C ABI (libpfb.so via pyrecest_b200._engine),
against the oracle and the golden fixtures made from the real reference.
Bar: bit-exact for indices, the pure add / numpy.mod predict paths and the sequential-exact CDF;
rtol 1e-12 (fp64) / 1e-5 (fp32 storage) for everything that goes through exp / log / cos."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import pf_oracle as O

pytestmark = pytest.mark.gpu

RTOL64 = 1e-12
RTOL32 = 1e-5


@pytest.fixture(scope="module")
def eng():
    from pyrecest_b200._engine import engine

    return engine("cuda:0")


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda:0").contiguous()


def host(t):
    return t.detach().cpu().numpy()


def pparams(mode, dim, noise_cols=None, F=None, b=None, A=None, mu=None, kappa=0.0, renorm=0):
    from pyrecest_b200 import _lib as L

    p = L.PredictParams()
    p.mode, p.dim = mode, dim
    p.noise_cols = dim if noise_cols is None else noise_cols
    if F is not None:
        p.has_F = 1
        p.F[: dim * dim] = np.asarray(F, dtype=np.float64).reshape(-1).tolist()
    if b is not None:
        p.has_b = 1
        p.b[:dim] = np.asarray(b, dtype=np.float64).tolist()
    if A is not None:
        p.has_A = 1
        p.A[: dim * dim] = np.asarray(A, dtype=np.float64).reshape(-1).tolist()
    if mu is not None:
        p.has_mu_noise = 1
        p.mu_noise[:dim] = np.asarray(mu, dtype=np.float64).tolist()
    p.kappa = kappa
    p.exp_m2k = float(np.exp(-2 * kappa))
    p.renormalize = renorm
    return p


def test_abi_version():
    from pyrecest_b200 import _lib as L

    assert L.lib().pfb_abi_version() == L.ABI_VERSION


def test_wrap_bitexact(eng, golden):
    g = golden("mod_edges")
    x = dev(g["x"])
    y = torch.empty_like(x)
    eng.wrap_2pi(x, y)
    assert np.array_equal(host(y), g["y"])
    rng = np.random.default_rng(0)
    big = np.concatenate([rng.standard_normal(100_003) * 50, rng.random(1000) * 2 * math.pi])
    x = dev(big)
    y = torch.empty_like(x)
    eng.wrap_2pi(x, y)
    assert np.array_equal(host(y), np.mod(big, 2.0 * np.pi))


def test_predict_euclid(eng, golden):
    from pyrecest_b200 import _lib as L

    g = golden("euclid_cv")
    d = g["d0"]
    for t in range(g["Z"].shape[0]):
        rows = O.gaussian_rows(g["Z"][t], g["Q"], np.zeros(2))
        x = dev(d)
        out = torch.empty_like(x)
        # linear transition on the device + host-made noise rows
        eng.predict(pparams(L.PRED_EUCLID, 2, F=g["F"]), x, out, dev(rows))
        np.testing.assert_allclose(host(out), g["d_pred"][t], rtol=1e-14, atol=1e-15)
        # identity transition: pure adds, bit-exact
        fd = d @ g["F"].T
        eng.predict(pparams(L.PRED_EUCLID, 2), dev(fd), out, dev(rows))
        assert np.array_equal(host(out), O.predict_euclid(fd, rows))
        # standard normals + A on the device
        eng.predict(pparams(L.PRED_EUCLID, 2, F=g["F"], A=O.gaussian_noise_factor(g["Q"])), x, out, dev(g["Z"][t]))
        np.testing.assert_allclose(host(out), g["d_pred"][t], rtol=RTOL64, atol=1e-14)
        d = g["d_post"][t]


def test_predict_torus_shift_bitexact(eng, golden):
    from pyrecest_b200 import _lib as L

    g = golden("torus_t3")
    d = g["d0"]
    for t in range(g["Z"].shape[0]):
        rows = O.gaussian_rows(g["Z"][t], g["Cn"])
        x = dev(d)
        out = torch.empty_like(x)
        eng.predict(pparams(L.PRED_TORUS_SHIFT, 3), x, out, dev(rows))
        assert np.array_equal(host(out), g["d_pred"][t])
        eng.predict(pparams(L.PRED_TORUS_SHIFT, 3, A=O.gaussian_noise_factor(g["Cn"])), x, out, dev(g["Z"][t]))
        diff = np.abs(host(out) - g["d_pred"][t])
        diff = np.minimum(diff, 2 * np.pi - diff)
        assert diff.max() < 1e-12
        d = g["d_post"][t]
    # in place, circle
    g = golden("circle_case")
    rows = O.gaussian_rows(g["Z"][0][:, None], np.array([[float(g["sig_n"]) ** 2]]))
    x = dev(g["d0"][:, None])
    eng.predict(pparams(L.PRED_TORUS_SHIFT, 1), x, x, dev(rows))
    assert np.array_equal(host(x)[:, 0], g["d_pred"][0])


def test_predict_torus_addmode_bitexact(eng, golden):
    from pyrecest_b200 import _lib as L

    g = golden("torus_addmode")
    rows = O.gaussian_rows(g["Z"], g["C"])
    x = dev(np.mod(g["d0"], 2 * np.pi))
    out = torch.empty_like(x)
    eng.predict(pparams(L.PRED_TORUS_ADD, 2, mu=np.mod(g["mu_n"], 2 * np.pi)), x, out, dev(rows))
    assert np.array_equal(host(out), g["d_pred"])


def test_predict_sphere_vmf(eng, golden):
    from pyrecest_b200 import _lib as L

    g = golden("vmf_sampler")
    x = dev(g["modes"])
    out = torch.empty_like(x)
    for kappa in g["kappas"]:
        eng.predict(pparams(L.PRED_SPHERE_VMF, 3, kappa=float(kappa)), x, out, dev(g["triples"]))
        np.testing.assert_allclose(host(out), g[f"out_k{kappa}"], rtol=0, atol=2e-14)
        np.testing.assert_allclose(host(out), O.predict_sphere_vmf(g["modes"], g["triples"], float(kappa)), rtol=0, atol=2e-14)
    rng = np.random.default_rng(3)
    n = rng.standard_normal(g["modes"].shape) * 0.1
    eng.predict(pparams(L.PRED_SPHERE_ADD, 3), x, out, dev(n))
    np.testing.assert_allclose(host(out), O.predict_sphere_additive(g["modes"], n), rtol=RTOL64)


class _D:  # minimal stand-ins recognised by registry.kind_of through their class names
    pass


def _dist(name, **kw):
    cls = type(name, (), {})
    o = cls()
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def _loglik(eng, dist, x, dim, dtype=torch.float64, w_prev=None):
    from pyrecest_b200 import registry as R

    lp, keep = R.lik_params(dist, dim, images_to_device=lambda t: dev(t))
    xd = dev(x.reshape(-1, dim), dtype)
    logw = torch.empty(xd.shape[0], dtype=torch.float64, device="cuda:0")
    eng.loglik(lp, xd, logw, w_prev=None if w_prev is None else dev(w_prev))
    return host(logw), float(host(eng.stats)[0])


@pytest.mark.parametrize("dtype,rtol", [(torch.float64, RTOL64), (torch.float32, RTOL32)])
def test_loglik_tables(eng, golden, dtype, rtol):
    g = golden("likelihoods")
    f32 = dtype == torch.float32

    def cmp(lw, ref, x, fn):
        if f32:  # compare against the oracle evaluated at the fp32-rounded particle
            ref = fn(x.astype(np.float32).astype(np.float64))
        np.testing.assert_allclose(np.exp(lw), ref, rtol=rtol, atol=1e-300)

    for d in (1, 2, 3, 4):
        x, mu, Cc = g[f"g{d}_x"], g[f"g{d}_mu"], g[f"g{d}_C"]
        lw, mx = _loglik(eng, _dist("GaussianDistribution", mu=mu, C=Cc), x, d, dtype)
        cmp(lw, g[f"g{d}_pdf"], x, lambda xx: O.pdf_gauss(xx, mu, Cc))
        assert mx == lw.max()
    for tag in ("wn2", "wn3", "wn3b"):
        x, mu, Cc = g[f"{tag}_x"], g[f"{tag}_mu"], g[f"{tag}_C"]
        lw, _ = _loglik(eng, _dist("HypertoroidalWrappedNormalDistribution", mu=mu, C=Cc), x, x.shape[1], dtype)
        cmp(lw, g[f"{tag}_pdf"], x, lambda xx: O.pdf_wn_multi(xx, mu, Cc))
    x = g["wn1_x"]
    for i, s in enumerate(g["wn1_sigma"]):
        lw, _ = _loglik(eng, _dist("WrappedNormalDistribution", mu=g["wn1_mu"], C=np.array(s**2)), x, 1, dtype)
        cmp(lw, g["wn1_pdf"][i], x, lambda xx: O.pdf_wn_1d(xx, float(g["wn1_mu"]), float(s)))
    for i, k in enumerate(g["vm_kappa"]):
        lw, _ = _loglik(eng, _dist("VonMisesDistribution", mu=g["vm_mu"], kappa=k), x, 1, dtype)
        cmp(lw, g["vm_pdf"][i], x, lambda xx: O.pdf_vm(xx, float(g["vm_mu"]), float(k)))
    for D in (3, 4):
        x, mu = g[f"vmf{D}_x"], g[f"vmf{D}_mu"]
        for i, k in enumerate(g[f"vmf{D}_kappa"]):
            lw, _ = _loglik(eng, _dist("VonMisesFisherDistribution", mu=mu, kappa=k), x, D, dtype)
            cmp(lw, g[f"vmf{D}_pdf"][i], x, lambda xx: O.pdf_vmf(xx, mu, float(k)))


def test_wn_multi_matches_full_343_sum_on_random_inputs(eng):
    rng = np.random.default_rng(11)
    C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
    for scale, mu in ((0.5, [1.0, 2.0, 3.0]), (0.05, [6.2, 0.01, 3.1]), (3.0, [0.0, 5.0, 1.0])):
        x = rng.random((4000, 3)) * 2 * np.pi
        x[:8] = [[0, 0, 0], [2 * np.pi] * 3, [np.pi] * 3, [np.pi, 0, 2 * np.pi], mu, [0.0, np.pi, np.pi], [1e-9] * 3, [6.28] * 3]
        lw, _ = _loglik(eng, _dist("HypertoroidalWrappedNormalDistribution", mu=np.array(mu), C=scale * C3), x, 3)
        ref = O.pdf_wn_multi(x, np.array(mu), scale * C3)
        ok = ref > 1e-280
        np.testing.assert_allclose(np.exp(lw[ok]), ref[ok], rtol=RTOL64)


def test_normalize_and_reweigh(eng, golden):
    g = golden("dirac_moments")
    w, lik = g["w"], g["lin_lik"]
    logw = dev(np.log(lik) + np.log(w))
    wout = torch.empty_like(logw)
    eng.normalize(logw, wout)
    np.testing.assert_allclose(host(wout), g["lin_reweighed"], rtol=RTOL64)
    st = host(eng.stats)
    p = host(wout)
    np.testing.assert_allclose(st[1] ** 2 / st[2], 1.0 / np.sum(p * p), rtol=1e-12)
    assert st[3] == 0.0
    # -inf and NaN handling
    bad = np.log(lik)
    bad[:] = -np.inf
    eng.normalize(dev(bad), None)
    assert int(host(eng.stats)[3]) & 2
    bad = np.log(lik)
    bad[5] = np.nan
    eng.normalize(dev(bad), None)
    assert int(host(eng.stats)[3]) & 1


CASES = ["lik", "wide", "ties", "zeros", "uniform", "single"]


@pytest.mark.parametrize("name", CASES)
def test_scan_seqexact_and_indices_bitexact(eng, golden, name):
    g = golden("resample_cases")
    w, u, idx = g[f"{name}_w"], g[f"{name}_u"], g[f"{name}_idx"]
    wd = dev(w)
    cdf = torch.empty_like(wd)
    eng.scan(cdf, w=wd, exact=True)
    ref = np.cumsum(w)
    assert np.array_equal(host(cdf), ref), f"max ulp-ish diff {np.max(np.abs(host(cdf) - ref) / ref[-1])}"
    assert eng.scan_error_flags() == 0
    assert host(eng.stats)[4] == ref[-1]
    out_idx = torch.empty(u.shape[0], dtype=torch.int64, device="cuda:0")
    eng.resample(cdf, dev(u), None, None, u.shape[0], 1, idx_out=out_idx)
    assert np.array_equal(host(out_idx), idx)
    # the fast scan is only tolerance-equal, its indices may differ at exact ties
    eng.scan(cdf, w=wd, exact=False)
    # (the "ties" case is where a blocked sum is MORE accurate than the sequential one: 1 + 2^-53 + ... )
    np.testing.assert_allclose(host(cdf), ref, rtol=1e-13 if name != "ties" else 1e-11)


@pytest.mark.parametrize("n", [1, 2, 2047, 2048, 2049, 300_001, 3_000_000])
def test_scan_seqexact_sizes(eng, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) * 2.5
    w = np.exp(-0.5 * (x - 0.7) ** 2) * (rng.random(n) < 0.97)
    wd = dev(w)
    cdf = torch.empty_like(wd)
    eng.scan(cdf, w=wd, exact=True)
    assert np.array_equal(host(cdf), np.cumsum(w))
    assert eng.scan_error_flags() == 0
    # from log-weights: p = exp(l - max) is formed on the device; the scan of those p is exact
    lw = -0.5 * (x - 0.7) ** 2
    lwd = dev(lw)
    eng.normalize(lwd, None)
    eng.scan(cdf, logw=lwd, exact=True)
    c = host(cdf)
    p = np.diff(np.concatenate([[0.0], c]))
    np.testing.assert_allclose(p, np.exp(lw - lw.max()), rtol=1e-9, atol=1e-14 * c[-1])  # diff of a running sum
    u = rng.random(min(n, 100_000))
    idx = torch.empty(u.shape[0], dtype=torch.int64, device="cuda:0")
    eng.resample(cdf, dev(u), None, None, u.shape[0], 1, idx_out=idx)
    cn = c / c[-1]
    assert np.array_equal(host(idx), np.minimum(np.searchsorted(cn, u, side="right"), n - 1))


def test_scan_adversarial_dynamic_range(eng):
    rng = np.random.default_rng(5)
    n = 200_000
    # every binade from 2^-1060 (subnormal) up to 2^40, zeros in between, huge jumps
    e = np.sort(rng.integers(-1060, 40, size=n))
    w = np.ldexp(1.0 + rng.random(n), e) * (rng.random(n) < 0.9)
    w[0] = 0.0
    w[1] = 5e-324
    wd = dev(w)
    cdf = torch.empty_like(wd)
    eng.scan(cdf, w=wd, exact=True)
    assert np.array_equal(host(cdf), np.cumsum(w))
    w2 = np.concatenate([np.full(70_000, 2.0 ** -53), [1.0], np.full(70_000, 2.0 ** -53), [2.0 ** -52] * 5000])
    wd = dev(w2)
    cdf = torch.empty_like(wd)
    eng.scan(cdf, w=wd, exact=True)
    assert np.array_equal(host(cdf), np.cumsum(w2))
    assert eng.scan_error_flags() == 0


@pytest.fixture(params=[("1", ""), ("1", "2pass"), ("0", "")], ids=["bucket_records", "bucket_records_2pass_build", "guide_window"])
def lookup_path(request, monkeypatch):
    """multinomial resampling has two index-lookup structures and two builds of the first (pfb_resample.cu);
    PFB_BUCKET_RECORDS / PFB_RECORD_BUILD are read per call"""
    monkeypatch.setenv("PFB_BUCKET_RECORDS", request.param[0])
    monkeypatch.setenv("PFB_RECORD_BUILD", request.param[1])
    return request.param


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_resample_gather_and_estimates(eng, golden, dtype, lookup_path):
    from pyrecest_b200 import _lib as L

    g = golden("torus_t3")
    N = g["d0"].shape[0]
    for t in range(2):
        wpost = O.reweigh(O.uniform_w(N), g["lik"][t])
        cdf = torch.empty(N, dtype=torch.float64, device="cuda:0")
        eng.scan(cdf, w=dev(wpost), exact=True)
        x = dev(g["d_pred"][t], dtype)
        out = torch.empty_like(x)
        idx = torch.empty(N, dtype=torch.int64, device="cuda:0")
        est = torch.empty(6, dtype=torch.float64, device="cuda:0")
        eng.resample(cdf, dev(g["U"][t]), x, out, N, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est)
        assert np.array_equal(host(idx), g["idx"][t])
        assert np.array_equal(host(out), host(x)[g["idx"][t]])
        if dtype == torch.float64:
            assert np.array_equal(host(out), g["d_post"][t])
        e = host(est).reshape(3, 2)
        md = np.mod(np.arctan2(e[:, 1], e[:, 0]), 2 * np.pi)
        np.testing.assert_allclose(md, g["est"][t], rtol=RTOL64 if dtype == torch.float64 else RTOL32)
    g = golden("euclid_cv")
    N = g["d0"].shape[0]
    lik = O.pdf_gauss(g["d_pred"][0], g["zs"][0], g["R"])
    cdf = torch.empty(N, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(O.reweigh(O.uniform_w(N), lik)), exact=True)
    x = dev(g["d_pred"][0], dtype)
    out = torch.empty_like(x)
    est = torch.empty(2, dtype=torch.float64, device="cuda:0")
    idx = torch.empty(N, dtype=torch.int64, device="cuda:0")
    eng.resample(cdf, dev(g["U"][0]), x, out, N, 2, idx_out=idx, est_kind=L.EST_SUM, est_out=est)
    assert np.array_equal(host(idx), g["idx"][0])
    np.testing.assert_allclose(host(est), g["est"][0], rtol=RTOL64 if dtype == torch.float64 else RTOL32)


def test_sample_n_not_equal_N_and_systematic(eng, golden, lookup_path):
    g = golden("dirac_moments")
    w, u = g["w"], g["tor_sample_u"]
    N = w.shape[0]
    cdf = torch.empty(N, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    x = dev(np.mod(g["tor_d"], 2 * np.pi))
    out = torch.empty((u.shape[0], 3), dtype=torch.float64, device="cuda:0")
    idx = torch.empty(u.shape[0], dtype=torch.int64, device="cuda:0")
    eng.resample(cdf, dev(u), x, out, u.shape[0], 3, idx_out=idx)
    assert np.array_equal(host(idx), g["tor_sample_idx"])
    assert np.array_equal(host(out), g["tor_sample"])
    for u0, n_out in ((0.37, N), (0.0, 2 * N + 1), (np.nextafter(1.0, 0), 17)):
        idx = torch.empty(n_out, dtype=torch.int64, device="cuda:0")
        eng.resample(cdf, dev(np.array([u0])), None, None, n_out, 1, systematic=True, idx_out=idx)
        assert np.array_equal(host(idx), O.systematic_indices(w, u0, n_out))


@pytest.mark.parametrize("dtype,rtol", [(torch.float64, RTOL64), (torch.float32, RTOL32)])
def test_moments(eng, golden, dtype, rtol):
    from pyrecest_b200 import _lib as L

    g = golden("dirac_moments")
    w = dev(g["w"])
    r32 = (lambda a: a.astype(np.float32).astype(np.float64)) if dtype == torch.float32 else (lambda a: a)
    x = dev(g["lin_d"], dtype)
    m = eng.moments(L.MOM_MEAN, x, 4, w=w)
    np.testing.assert_allclose(host(m)[:4], O.mean_linear(r32(g["lin_d"]), g["w"]), rtol=rtol)
    np.testing.assert_allclose(host(m)[4], 1.0, rtol=1e-14)
    cov = eng.moments(L.MOM_CENTRAL, x, 4, w=w, center=m)
    np.testing.assert_allclose(host(cov)[:16].reshape(4, 4), O.covariance_linear(r32(g["lin_d"]), g["w"]), rtol=max(rtol, 1e-11))
    x = dev(g["tor_d"], dtype)
    for order, key in ((1, "tor_m1"), (2, "tor_m2")):
        t = host(eng.moments(L.MOM_TRIG, x, 3, w=w, order=order))[:6].reshape(3, 2)
        ref = O.trig_moment(r32(g["tor_d"]), g["w"], order)
        np.testing.assert_allclose(t[:, 0] + 1j * t[:, 1], ref, rtol=rtol * 10)
    x = dev(g["sph_d"], dtype)
    np.testing.assert_allclose(host(eng.moments(L.MOM_MEAN, x, 3, w=w))[:3], O.mean_resultant(r32(g["sph_d"]), g["w"]), rtol=rtol)
    np.testing.assert_allclose(host(eng.moments(L.MOM_SCATTER, x, 3, w=w))[:9].reshape(3, 3), O.scatter(r32(g["sph_d"]), g["w"]), rtol=rtol * 10, atol=1e-15)
    # uniform weights (w = NULL)
    mu = host(eng.moments(L.MOM_MEAN, x, 3))[:3]
    np.testing.assert_allclose(mu, r32(g["sph_d"]).mean(axis=0), rtol=rtol * 10)


def test_fused_predict_loglik_equals_separate_calls(eng, golden):
    from pyrecest_b200 import _lib as L
    from pyrecest_b200 import registry as R

    g = golden("torus_t3")
    rows = O.gaussian_rows(g["Z"][0], g["Cn"])
    x = dev(g["d0"])
    lp, keep = R.lik_params(_dist("HypertoroidalWrappedNormalDistribution", mu=g["zs"][0], C=g["Cl"]), 3, images_to_device=dev)
    out1, out2 = torch.empty_like(x), torch.empty_like(x)
    l1 = torch.empty(x.shape[0], dtype=torch.float64, device="cuda:0")
    l2 = torch.empty_like(l1)
    pp = pparams(L.PRED_TORUS_SHIFT, 3)
    eng.predict(pp, x, out1, dev(rows))
    eng.loglik(lp, out1, l1)
    m1 = float(host(eng.stats)[0])
    eng.predict_loglik(pp, lp, x, out2, dev(rows), l2)
    assert np.array_equal(host(out1), host(out2)) and np.array_equal(host(l1), host(l2))
    assert m1 == float(host(eng.stats)[0])
    np.testing.assert_allclose(np.exp(host(l2)), g["lik"][0], rtol=RTOL64)


def test_rejects_bad_arguments(eng):
    from pyrecest_b200 import _lib as L

    x = torch.zeros((4, 3), dtype=torch.float64, device="cuda:0")
    with pytest.raises(NotImplementedError):
        eng.predict(pparams(L.PRED_SPHERE_VMF, 2, kappa=1.0), x[:, :2].contiguous(), x[:, :2].contiguous(), x)
    with pytest.raises(ValueError):
        eng.predict(pparams(L.PRED_EUCLID, 9), x, x, x)
    lp = L.LikParams()
    lp.kind, lp.dim = 17, 3
    with pytest.raises(NotImplementedError):
        eng.loglik(lp, x, torch.zeros(4, dtype=torch.float64, device="cuda:0"))


# ---- device RNG (production path) -------------------------------------------------------------------
def test_rng_uniforms_match_philox_restatement(eng):
    from oracle import philox as P

    for seed, offset, n in ((0, 0, 1000), (0x1234567890ABCDEF, 7, 100_003), (2**64 - 1, 2**40 + 5, 4097)):
        out = torch.empty(n, dtype=torch.float64, device="cuda:0")
        eng.rng_uniform(out, seed, offset)
        ref = P.resample_uniforms(seed, offset, n)
        assert np.array_equal(host(out), ref)
        assert ref.min() >= 0.0 and ref.max() < 1.0


def test_rng_resample_equals_injecting_the_same_uniforms(eng, golden):
    from oracle import philox as P

    g = golden("resample_cases")
    w = g["lik_w"]
    n = w.shape[0]
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    idx = torch.empty(n, dtype=torch.int64, device="cuda:0")
    eng.resample_rng(cdf, 99, 3, None, None, n, 1, idx_out=idx)
    assert np.array_equal(host(idx), O.multinomial_indices(w, P.resample_uniforms(99, 3, n)))
    eng.resample_rng(cdf, 99, 4, None, None, n, 1, systematic=True, idx_out=idx)
    assert np.array_equal(host(idx), O.systematic_indices(w, P.systematic_u0(99, 4), n))


def test_rng_normals_and_von_mises_and_vmf(eng):
    from oracle import philox as P
    from pyrecest_b200 import _lib as L
    from scipy.special import iv

    n = 400_000
    x = torch.zeros((n, 3), dtype=torch.float64, device="cuda:0")
    out = torch.empty_like(x)
    p = pparams(L.PRED_EUCLID, 3)
    p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_NORMAL, 5, 11
    eng.predict(p, x, out, None)
    z = host(out)
    np.testing.assert_allclose(z, P.normals(5, 11, n, 3), rtol=1e-12, atol=1e-14)
    assert abs(z.mean()) < 5 / np.sqrt(3 * n) and abs(z.var() - 1) < 0.01
    assert np.abs(np.corrcoef(z.T) - np.eye(3)).max() < 0.01
    # von Mises per dimension: E cos = I1/I0, E sin = 0
    p = pparams(L.PRED_TORUS_SHIFT, 3)
    p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VONMISES, 5, 12
    kap = [0.0, 2.0, 50.0]
    p.set_von_mises_rng(kap)
    eng.predict(p, x, out, None)
    v = host(out)
    assert v.min() >= 0 and v.max() <= 2 * np.pi
    # bit-exact against the numpy restatement of the strip sampler (same Philox blocks, fp64 acceptance test)
    ref = P.vonmises(5, 12, n, kap, {k: L.vm_strip_table(k) for k in kap if k >= 1e-8})
    assert np.array_equal(v, np.mod(ref, 2 * np.pi))
    from scipy import stats

    for c, k in enumerate(kap):
        if k >= 1e-8:
            centred = np.mod(v[:, c] + np.pi, 2 * np.pi) - np.pi
            assert stats.kstest(centred, stats.vonmises(k).cdf).pvalue > 1e-4
    for c, k in enumerate(kap):
        a = iv(1, k) / iv(0, k)
        assert abs(np.cos(v[:, c]).mean() - a) < 5 / np.sqrt(n), (k, np.cos(v[:, c]).mean(), a)
        assert abs(np.sin(v[:, c]).mean()) < 5 / np.sqrt(n)
    # VMF on S^2 around e3: resultant length coth k - 1/k
    modes = torch.zeros((n, 3), dtype=torch.float64, device="cuda:0")
    modes[:, 2] = 1.0
    p = pparams(L.PRED_SPHERE_VMF, 3, kappa=8.0)
    p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VMF, 5, 13
    eng.predict(p, modes, out, None)
    s = host(out)
    np.testing.assert_allclose(np.linalg.norm(s, axis=1), 1.0, atol=1e-14)
    m = s.mean(axis=0)
    assert abs(m[2] - (1 / np.tanh(8.0) - 1 / 8.0)) < 5 / np.sqrt(n) and np.abs(m[:2]).max() < 5 / np.sqrt(n)


def test_resample_sharded_kernel_matches_sharded_oracle(eng):
    """pfb_resample_sharded for every shard of a 3-way split (emulated on one GPU) == oracle.sharded_*"""
    from pyrecest_b200.sharded import offspring_boundaries, shard_offsets

    rng = np.random.default_rng(21)
    G, n_loc, D = 3, 50_000, 4
    N = G * n_loc
    x = rng.standard_normal((N, D))
    logw = -0.5 * np.sum((x - 0.2) ** 2, axis=1) - 3.0 * (np.arange(N) >= n_loc)
    M = logw.max()
    u0 = 0.123456789
    cdfs, totals = [], []
    for g in range(G):
        lw = dev(logw[g * n_loc:(g + 1) * n_loc])
        eng.normalize(lw, None)                       # fills stats; then impose the global shift
        eng.stats[0:1].copy_(torch.tensor([M], dtype=torch.float64))
        c = torch.empty(n_loc, dtype=torch.float64, device="cuda:0")
        eng.scan(c, logw=lw, exact=True)
        cdfs.append(c)
        totals.append(float(host(eng.stats)[4]))
    offs = shard_offsets(totals)
    B = offspring_boundaries(offs, u0, N)
    pieces = []
    for g in range(G):
        cnt = B[g + 1] - B[g]
        out = torch.empty((cnt, D), dtype=torch.float64, device="cuda:0")
        idx = torch.empty(cnt, dtype=torch.int64, device="cuda:0")
        eng.resample_sharded(cdfs[g], offs[g], offs[-1], u0, B[g], cnt, N, dev(x[g * n_loc:(g + 1) * n_loc]), out, D, idx_out=idx)
        pieces.append(host(out))
        assert host(idx).min() >= 0 and host(idx).max() <= n_loc - 1
    got = np.concatenate(pieces)
    # oracle on the device-made p (exp differs from numpy's by an ulp, so rebuild p from the device CDFs)
    shards = [np.diff(np.concatenate([[0.0], host(c)])) for c in cdfs]
    shards_exact = [np.exp(logw[g * n_loc:(g + 1) * n_loc] - M) for g in range(G)]
    idx_ref = O.sharded_systematic_indices(shards_exact, u0)
    mism = np.any(got != x[idx_ref], axis=1).sum()
    assert mism <= 3, mism
    assert B == O.offspring_boundaries(np.array(offs), u0, N)


@pytest.mark.parametrize("n", [1, 300, 519, 2047, 2049])
def test_fresh_engine_workspace_fits_every_call(n):
    """A new engine's workspace (sized by pfb_workspace_bytes(n)) must satisfy every kernel of the cycle for the
    same n -- in particular the weighting kernels, which work on whole 2048-particle tiles."""
    from pyrecest_b200 import registry as R
    from pyrecest_b200._engine import Engine

    e = Engine("cuda:0")
    rng = np.random.default_rng(n)
    x = dev(rng.normal(size=(n, 2)))
    lp, _keep = R.lik_params(_dist("GaussianDistribution", mu=np.zeros(2), C=np.eye(2)), 2, images_to_device=dev)
    logw = torch.empty(n, dtype=torch.float64, device="cuda:0")
    e.loglik(lp, x, logw)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    e.scan(cdf, logw=logw, exact=True)
    w = np.exp(host(logw) - host(logw).max())
    np.testing.assert_allclose(host(cdf), np.cumsum(w), rtol=1e-13)


def test_resample_lookup_adversarial(eng, lookup_path):
    """crowded buckets (more entries than a bucket record holds, more than its 16-bit count), heavy particles
    spanning many buckets, and uniforms that sit exactly on / next to normalised CDF values (searchsorted ties)"""
    rng = np.random.default_rng(11)
    parts = [rng.random(5000) + 0.5, np.full(200, 1e-9), [800.0], np.full(70_000, 1e-13), rng.random(3000),
             np.full(40, 2.0 ** -30), [0.0, 0.0, 0.0], rng.random(20_000) * 1e-3, [2500.0], rng.random(1000)]
    w = np.concatenate([np.asarray(p, dtype=np.float64) for p in parts])
    n = w.shape[0]
    c = np.cumsum(w)
    v = c / c[-1]
    pick = rng.integers(0, n, 20_000)
    u = np.concatenate([rng.random(100_000), v[pick], np.nextafter(v[pick], 0.0), np.nextafter(v[pick], 2.0),
                        [0.0, np.nextafter(1.0, 0.0)], (np.arange(4096) + 0.0) / 4096.0])
    u = u[u < 1.0]
    ref = np.minimum(np.searchsorted(v, u, side="right"), n - 1)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    assert np.array_equal(host(cdf), c)
    idx = torch.empty(u.shape[0], dtype=torch.int64, device="cuda:0")
    eng.resample(cdf, dev(u), None, None, u.shape[0], 1, idx_out=idx)
    assert np.array_equal(host(idx), ref)
    x = dev(rng.normal(size=(n, 3)))
    out = torch.empty((u.shape[0], 3), dtype=torch.float64, device="cuda:0")
    eng.resample(cdf, dev(u), x, out, u.shape[0], 3)
    assert np.array_equal(host(out), host(x)[ref])


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("dim", [1, 2, 3, 4, 5, 6])
def test_resample_gather_every_record_size(eng, dtype, dim):
    """the gather reads sector-aligned spans and picks the record out of them: every (dtype, dim) alignment"""
    rng = np.random.default_rng(dim)
    n = 10_007
    w = rng.random(n)
    u = rng.random(30_011)
    c = np.cumsum(w)
    ref = np.minimum(np.searchsorted(c / c[-1], u, side="right"), n - 1)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    xs = dev(rng.normal(size=(n + 3, dim)), dtype)
    for shift in (0, 1, 3):  # views that start at odd record offsets inside the allocation
        x = xs[shift:shift + n]
        out = torch.empty((u.shape[0], dim), dtype=dtype, device="cuda:0")
        eng.resample(cdf, dev(u), x, out, u.shape[0], dim)
        assert np.array_equal(host(out), host(x)[ref])


def test_resample_heavy_particles_span_many_buckets(eng, lookup_path):
    """a few particles holding most of the weight: runs of 1e4..1e5 empty buckets (the grid-wide fill of the record
    build), next to crowded buckets of weightless particles"""
    rng = np.random.default_rng(5)
    n = 2_000_003
    w = rng.random(n) * 1e-3
    w[[0, 777, 1_000_000, n - 1]] = [300.0, 40.0, 500.0, 90.0]
    w[1_200_000:1_300_000] = 1e-12
    c = np.cumsum(w)
    v = c / c[-1]
    u = np.concatenate([rng.random(400_000), v[[0, 776, 777, 999_999, 1_000_000, n - 2]], [0.0]])
    u = u[u < 1.0]
    ref = np.minimum(np.searchsorted(v, u, side="right"), n - 1)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    assert np.array_equal(host(cdf), c)
    idx = torch.empty(u.shape[0], dtype=torch.int64, device="cuda:0")
    for _ in range(2):  # twice: the job counter of the fill kernel cleans itself
        eng.resample(cdf, dev(u), None, None, u.shape[0], 1, idx_out=idx)
        assert np.array_equal(host(idx), ref)


def test_wn_cell_masks_do_not_change_the_likelihood(eng, golden):
    """summing only the cell's candidate images around the cell's reference image (lik.cell_grid > 0) gives the
    log-likelihood of the gridless path (best image found per particle, all images visited) to rounding"""
    from pyrecest_b200 import registry as R

    g = golden("likelihoods")
    rng = np.random.default_rng(3)
    for tag in ("wn2", "wn3", "wn3b"):
        mu, Cc = g[f"{tag}_mu"], g[f"{tag}_C"]
        D = mu.shape[0]
        x = dev(rng.random((200_000, D)) * 2 * np.pi)
        out = []
        for grid in (None, 0):
            lp, keep = R.lik_params(_dist("HypertoroidalWrappedNormalDistribution", mu=mu, C=Cc), D, images_to_device=dev)
            assert lp.cell_grid == (R.WN_CELL_GRID_FINE if lp.n_images <= 32 else (R.WN_CELL_GRID if lp.n_images <= 64 else 0))
            assert lp.cell_span == (2 if lp.n_images <= 32 else 0)  # (32-bit masks on the fine grid, else uint64)
            if grid is not None:
                lp.cell_grid, lp.cell_span = grid, 0
            logw = torch.empty(x.shape[0], dtype=torch.float64, device="cuda:0")
            eng.loglik(lp, x, logw)
            out.append(host(logw))
        np.testing.assert_allclose(out[0], out[1], rtol=2e-14, atol=2e-14)


def test_systematic_with_degenerate_weights(eng):
    """one or two particles hold almost all the weight: their runs of offspring (> 8192) go through the job list and
    are copied by the whole grid; indices and particles equal the oracle's comb"""
    rng = np.random.default_rng(21)
    n, n_out = 200_000, 1_500_001
    w = rng.random(n) * 1e-4
    w[1234] = 30.0
    w[150_000] = 12.0
    x = dev(rng.normal(size=(n, 3)))
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, w=dev(w), exact=True)
    ref = O.systematic_indices(w, 0.3141, n_out)
    for _ in range(2):  # twice: the job counter cleans itself
        idx = torch.empty(n_out, dtype=torch.int64, device="cuda:0")
        out = torch.empty((n_out, 3), dtype=torch.float64, device="cuda:0")
        eng.resample(cdf, dev(np.array([0.3141])), x, out, n_out, 3, systematic=True, idx_out=idx)
        assert np.array_equal(host(idx), ref)
        assert np.array_equal(host(out), host(x)[ref])


def test_resample_small_and_odd_sizes(eng, lookup_path):
    """n from 1 to a few thousand, n_out unrelated to n, zero weights at both ends: every size takes the same
    kernels as 1e8 particles (one bucket, partial tiles, look-ahead past the end of the data)"""
    rng = np.random.default_rng(99)
    for n in list(range(1, 40)) + [63, 64, 65, 255, 256, 257, 2047, 2048, 2049, 2080, 4095, 4097, 6143]:
        w = rng.random(n) ** 3
        if n > 4:
            w[0] = 0.0
            w[-1] = 0.0
            w[n // 2] = 0.0
        if w.sum() == 0:
            w[0] = 1.0
        n_out = int(rng.integers(1, 3 * n + 50))
        u = rng.random(n_out)
        c = np.cumsum(w)
        ref = np.minimum(np.searchsorted(c / c[-1], u, side="right"), n - 1)
        cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
        eng.scan(cdf, w=dev(w), exact=True)
        assert np.array_equal(host(cdf), c), n
        x = dev(rng.normal(size=(n, 2)))
        out = torch.empty((n_out, 2), dtype=torch.float64, device="cuda:0")
        idx = torch.empty(n_out, dtype=torch.int64, device="cuda:0")
        eng.resample(cdf, dev(u), x, out, n_out, 2, idx_out=idx)
        assert np.array_equal(host(idx), ref), n
        assert np.array_equal(host(out), host(x)[ref]), n
        u0 = float(rng.random())
        eng.resample(cdf, dev(np.array([u0])), x, out, n_out, 2, systematic=True, idx_out=idx)
        assert np.array_equal(host(idx), O.systematic_indices(w, u0, n_out)), n


@pytest.mark.parametrize("n", [1, 255, 257, 5000, 300_001])
def test_slice_scaled_weights_scan_is_the_exact_cumsum_of_its_addends(eng, n):
    """PFB_WEIGHTS_SCALED: p_i = pdf_i 2^(-e_s) per slice of 256 particles, e_s an integer slice exponent; the scan's
    addends are p_i 2^(e_s - E) (exact scalings) and its CDF is numpy.cumsum of those addends, bit for bit."""
    from pyrecest_b200 import registry as R
    from pyrecest_b200.distributions import GaussianDistribution

    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, 2)) * np.array([3.0, 40.0])  # log-likelihoods spread over ~1000 nats
    xd = dev(x)
    dist = GaussianDistribution(np.array([0.3, -0.2]), np.diag([0.5, 0.8]))
    lp, keep = R.lik_params(dist, 2, images_to_device=R.device_put(xd.device))
    logw = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.loglik(lp, xd, logw)
    l = host(logw)
    p = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.loglik(lp, xd, p, scaled=True)
    cdf = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf, logw=p, exact=True)
    assert eng.scan_error_flags() == 0
    ph = host(p)
    nsl = (n + 255) // 256
    lpad = np.full(nsl * 256, -np.inf)
    lpad[:n] = l
    # the slice exponents (integers; taken from the first 32 weighted particles of a slice, moved up when a later
    # particle lies more than e^600 above): recovered from the slice's largest weight
    ppad = np.zeros(nsl * 256)
    ppad[:n] = ph
    top = lpad.reshape(nsl, 256).argmax(axis=1)
    rows = np.arange(nsl)
    e = np.round((lpad.reshape(nsl, 256)[rows, top] - np.log(ppad.reshape(nsl, 256)[rows, top])) * 1.4426950408889634)
    assert np.all(ppad.reshape(nsl, 256).max(axis=1) < np.exp(601.0))
    np.testing.assert_allclose(ph, np.exp(l - np.repeat(e, 256)[:n] * 0.6931471805599453), rtol=2e-12, atol=1e-300)
    a = ph * np.exp2(np.repeat(e, 256)[:n] - e.max())
    assert np.array_equal(host(cdf), np.cumsum(a))
    # the same weights as the log format, up to rounding: normalised CDFs agree
    cdf2 = torch.empty(n, dtype=torch.float64, device="cuda:0")
    eng.scan(cdf2, logw=logw, exact=True)
    np.testing.assert_allclose(host(cdf) / host(cdf)[-1], host(cdf2) / host(cdf2)[-1], rtol=1e-12, atol=1e-15)
    with pytest.raises(Exception):  # the scaled format cannot be re-read without its slice exponents
        eng._tile_stats_key = None
        eng.lib  # noqa: B018
        from pyrecest_b200 import _lib as L
        import ctypes as C
        L.check(eng.lib.pfb_scan_cdf(L.SCAN_SEQEXACT | L.SCAN_SLICE_SCALED, n, None, C.c_void_p(p.data_ptr()),
                                     C.c_void_p(cdf.data_ptr()), C.c_void_p(eng.stats.data_ptr()),
                                     C.c_void_p(eng._ws.data_ptr()), eng._ws.numel(), None))


@pytest.mark.parametrize("world,skew,heavy", [(1, 0.0, False), (2, 0.0, False), (4, 6.0, False), (3, 0.0, True), (8, 3.0, False)])
def test_sharded_device_plan_kernel_equals_sharded_oracle(eng, world, skew, heavy):
    """pfb_resample_sharded_dev (one launch per rank, plan made on the device from the all-gathered totals): the G
    ranks are emulated on one GPU by G launches that store into G output buffers through the peer table."""
    from pyrecest_b200 import _lib as L

    rng = np.random.default_rng(world * 7 + int(skew))
    n_loc, D = 20_011, 4
    N = n_loc * world
    x = rng.standard_normal((N, D))
    logw = -0.5 * np.sum((x - 0.3) ** 2, axis=1) - skew * (np.arange(N) // n_loc)
    if heavy:
        logw[n_loc + 5] += 40.0  # one particle of shard 1 takes nearly everything: offspring cross into every rank
    M = logw.max()
    u0 = 0.6180339887
    shards = [np.exp(logw[g * n_loc:(g + 1) * n_loc] - M) for g in range(world)]
    idx = O.sharded_systematic_indices(shards, u0)
    cdfs, totals = [], []
    for g in range(world):
        c = torch.empty(n_loc, dtype=torch.float64, device="cuda:0")
        eng.scan(c, w=dev(shards[g]), exact=True)
        cdfs.append(c)
        totals.append(float(np.cumsum(shards[g])[-1]))
    totals_dev = dev(np.array(totals))
    outs = [torch.full((n_loc, D), np.nan, dtype=torch.float64, device="cuda:0") for _ in range(world)]
    ptrs = [o.data_ptr() for o in outs]
    est = torch.zeros((world, 2 * L.PFB_MAXD), dtype=torch.float64, device="cuda:0")
    status = torch.ones(1, dtype=torch.float64, device="cuda:0")
    for g in range(world):
        eng.resample_sharded_dev(cdfs[g], totals_dev, world, g, u0, ptrs, dev(x[g * n_loc:(g + 1) * n_loc]), D, L.EST_SUM,
                                 est[g], status)
    got = np.concatenate([host(o) for o in outs])
    assert float(status) == 0.0
    assert not np.isnan(got).any()  # every slot written exactly by its owner shard
    assert np.array_equal(got, x[idx])
    np.testing.assert_allclose(host(est)[:, :D].sum(axis=0), x[idx].sum(axis=0), rtol=1e-11, atol=1e-9)
    # a zero total is reported through the status word and nothing is written
    bad = torch.zeros(world, dtype=torch.float64, device="cuda:0")
    eng.resample_sharded_dev(cdfs[0], bad, world, 0, u0, ptrs, dev(x[:n_loc]), D, L.EST_SUM, est[0], status)
    assert float(status) == 1.0


@pytest.mark.parametrize("skew", [False, True])
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_resampling_pull_and_push_with_all_shards_on_one_device(eng, world, skew):
    """pfb_resample_sharded_pull / pfb_resample_sharded_dev with the `world` shards of one filter living on the same
    GPU (the peer tables simply point at local buffers): every rank's output equals the oracle's global systematic
    resampling of the sharded CDF, row for row -- the NCCL / NVLink run of the same kernels is
    tests/multi_gpu/run_sharded.py."""
    from pyrecest_b200 import _lib as L

    rng = np.random.default_rng(100 * world + skew)
    n, D = 70_001, 4
    x = [rng.standard_normal((n, D)) for _ in range(world)]
    p = [rng.random(n) * (np.exp(-3.0 * g) if skew else 1.0) for g in range(world)]
    if skew:
        p[0][1234] = 40_000.0   # a heavy particle: its offspring span ranks and go through the job list
    u0 = 0.6180339887
    idx = O.sharded_systematic_indices(p, u0)
    want = np.concatenate(x)[idx].reshape(world, n, D)
    xd = [dev(a) for a in x]
    cdfs = []
    totals = torch.zeros(world, dtype=torch.float64, device="cuda:0")
    for g in range(world):
        c = torch.empty(n, dtype=torch.float64, device="cuda:0")
        eng.scan(c, w=dev(p[g]), exact=True)
        totals[g] = c[-1]
        cdfs.append(c)
    status = torch.zeros(1, dtype=torch.float64, device="cuda:0")
    est = torch.zeros(2 * L.PFB_MAXD, dtype=torch.float64, device="cuda:0")
    # pull: one call per rank, local output
    sums = np.zeros(D)
    for r in range(world):
        out = torch.empty((n, D), dtype=torch.float64, device="cuda:0")
        eng.resample_sharded_pull([c.data_ptr() for c in cdfs], [a.data_ptr() for a in xd], totals, world, r, u0, out, n, D,
                                  L.EST_SUM, est, status)
        assert float(status[0]) == 0.0
        assert np.array_equal(host(out), want[r]), (world, skew, r)
        sums += host(est)[:D]
    np.testing.assert_allclose(sums, want.reshape(-1, D).sum(axis=0), rtol=1e-11, atol=1e-9)
    # push: one call per rank, every rank stores into the owners' buffers
    outs = [torch.empty((n, D), dtype=torch.float64, device="cuda:0") for _ in range(world)]
    for r in range(world):
        eng.resample_sharded_dev(cdfs[r], totals, world, r, u0, [o.data_ptr() for o in outs], xd[r], D, L.EST_SUM, est, status)
    for r in range(world):
        assert np.array_equal(host(outs[r]), want[r]), (world, skew, r)
    # an invalid global total: status 1, nothing written
    totals[0] = float("nan")
    out = torch.full((n, D), 7.0, dtype=torch.float64, device="cuda:0")
    eng.resample_sharded_pull([c.data_ptr() for c in cdfs], [a.data_ptr() for a in xd], totals, world, 0, u0, out, n, D,
                              L.EST_SUM, est, status)
    assert float(status[0]) == 1.0 and bool((out == 7.0).all())
