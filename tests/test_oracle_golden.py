"""Pins oracle/pf_oracle.py against (1) the reference's own known-answer values and
(2) outputs of the real reference run with injected draws (tests/golden/*.npz, made by
tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import pf_oracle as O

TWO_PI = 2.0 * np.pi


def test_mod_edges(golden):
    g = golden("mod_edges")
    assert np.array_equal(np.mod(g["x"], TWO_PI), g["y"])


def test_euclid_cv_matches_reference(golden):
    g = golden("euclid_cv")
    d = O.gaussian_rows(g["Z0"], g["prior_C"], g["prior_mu"])
    assert np.array_equal(d, g["d0"])
    N = d.shape[0]
    for t in range(g["Z"].shape[0]):
        noise = O.gaussian_rows(g["Z"][t], g["Q"], np.zeros(2))
        dp = O.predict_euclid(d, noise, F=g["F"])
        np.testing.assert_allclose(dp, g["d_pred"][t], rtol=1e-14, atol=1e-15)
        lik = O.pdf_gauss(dp, g["zs"][t], g["R"])
        dn, wn, idx, _ = O.update_resample(g["d_pred"][t], O.uniform_w(N), lik, g["U"][t])
        assert np.array_equal(idx, g["idx"][t])
        assert np.array_equal(dn, g["d_post"][t])
        np.testing.assert_allclose(O.mean_linear(dn, wn), g["est"][t], rtol=1e-13)
        d = g["d_post"][t]


def test_config0_full_size_matches_reference(golden, config0_inputs):
    """BASELINE.json configs[0] as specified -- R^2 constant velocity, 10^4 particles, 100 steps -- run through the
    real reference (tests/golden/make_golden.py config0): the oracle reproduces every resampled index, the particle
    sets bit for bit and the estimates."""
    g, c = golden("config0"), config0_inputs
    N, T = int(g["N"]), int(g["T"])
    assert np.array_equal(c["zs"], g["zs"])
    d = O.gaussian_rows(c["Z0"], np.eye(2), np.zeros(2))
    ar = np.arange(1, N + 1)
    for t in range(T):
        dp = O.predict_euclid(d, O.gaussian_rows(c["Z"][t], c["Q"], np.zeros(2)), F=c["F"])
        lik = O.pdf_gauss(dp, c["zs"][t], c["R"])
        d, wn, idx, _ = O.update_resample(dp, O.uniform_w(N), lik, c["U"][t])
        assert [int(idx.sum()), int((idx * ar).sum() % (2 ** 61 - 1))] == g["idx_sum"][t].tolist(), t
        if f"idx_{t}" in g:
            assert np.array_equal(idx, g[f"idx_{t}"])
            np.testing.assert_allclose(d, g[f"d_{t}"], rtol=1e-13, atol=1e-14)
        np.testing.assert_allclose(O.mean_linear(d, wn), g["est"][t], rtol=1e-12, atol=1e-13)


def test_torus_t3_matches_reference(golden):
    g = golden("torus_t3")
    d = np.mod(O.gaussian_rows(g["Z0"], g["C"], g["mu0"]), TWO_PI)
    assert np.array_equal(d, g["d0"])
    N = d.shape[0]
    for t in range(g["Z"].shape[0]):
        n = O.gaussian_rows(g["Z"][t], g["Cn"])
        dp = O.predict_torus(d, n, shift=True)
        assert np.array_equal(dp, g["d_pred"][t])
        lik = O.pdf_wn_multi(dp, g["zs"][t], g["Cl"])
        np.testing.assert_allclose(lik, g["lik"][t], rtol=1e-13)
        dn, wn, idx, _ = O.update_resample(dp, O.uniform_w(N), g["lik"][t], g["U"][t])
        assert np.array_equal(idx, g["idx"][t])
        assert np.array_equal(dn, g["d_post"][t])
        np.testing.assert_allclose(O.mean_direction_torus(dn, wn), g["est"][t], rtol=1e-13)
        d = dn


def test_torus_addmode_matches_reference(golden):
    g = golden("torus_addmode")
    n = O.gaussian_rows(g["Z"], g["C"])
    dp = O.predict_torus(np.mod(g["d0"], TWO_PI), n, shift=False, noise_mean=g["mu_n"])
    assert np.array_equal(dp, g["d_pred"])


def test_circle_matches_reference(golden):
    g = golden("circle_case")
    d = g["d0"]
    N = d.shape[0]
    liks = [lambda x: O.pdf_vm(x, float(g["vm_mu"]), float(g["vm_kappa"])),
            lambda x: O.pdf_wn_1d(x, float(g["wn_mu"]), float(g["wn_sigma"]))]
    for t in range(2):
        n = O.gaussian_rows(g["Z"][t][:, None], np.array([[float(g["sig_n"]) ** 2]]))[:, 0]
        dp = O.predict_torus(d, n, shift=True)
        assert np.array_equal(dp, g["d_pred"][t])
        lik = liks[t](dp)
        np.testing.assert_allclose(lik, g["lik"][t], rtol=1e-13)
        dn, wn, idx, _ = O.update_resample(dp, O.uniform_w(N), g["lik"][t], g["U"][t])
        assert np.array_equal(idx, g["idx"][t])
        assert np.array_equal(dn, g["d_post"][t])
        np.testing.assert_allclose(O.mean_direction_torus(dn, wn), g["est"][t], rtol=1e-13)
        d = dn


def test_likelihood_tables(golden):
    g = golden("likelihoods")
    for d in (1, 2, 3, 4):
        np.testing.assert_allclose(O.pdf_gauss(g[f"g{d}_x"], g[f"g{d}_mu"], g[f"g{d}_C"]), g[f"g{d}_pdf"], rtol=1e-13)
    for tag in ("wn2", "wn3", "wn3b"):
        np.testing.assert_allclose(O.pdf_wn_multi(g[f"{tag}_x"], g[f"{tag}_mu"], g[f"{tag}_C"]), g[f"{tag}_pdf"], rtol=1e-13)
    for i, s in enumerate(g["wn1_sigma"]):
        np.testing.assert_allclose(O.pdf_wn_1d(g["wn1_x"], float(g["wn1_mu"]), float(s)), g["wn1_pdf"][i], rtol=1e-14)
    for i, k in enumerate(g["vm_kappa"]):
        np.testing.assert_allclose(O.pdf_vm(g["wn1_x"], float(g["vm_mu"]), float(k)), g["vm_pdf"][i], rtol=1e-14)
    for D in (3, 4):
        for i, k in enumerate(g[f"vmf{D}_kappa"]):
            np.testing.assert_allclose(O.pdf_vmf(g[f"vmf{D}_x"], g[f"vmf{D}_mu"], float(k)), g[f"vmf{D}_pdf"][i], rtol=1e-13)


def test_vmf_sampler_matches_scipy_through_reference(golden):
    g = golden("vmf_sampler")
    for kappa in g["kappas"]:
        out = O.predict_sphere_vmf(g["modes"], g["triples"], float(kappa))
        np.testing.assert_allclose(out, g[f"out_k{kappa}"], rtol=0, atol=5e-15)
    # dense Householder form agrees with the per-row form
    for j in range(10):
        Q = O.householder_from_e1(g["modes"][j])
        s = O.vmf_canonical_s2(g["triples"][j], 50.0)
        np.testing.assert_allclose(Q @ s[0], g["out_k50.0"][j], atol=5e-15)


def test_dirac_moments(golden):
    g = golden("dirac_moments")
    w = g["w"]
    np.testing.assert_allclose(O.mean_linear(g["lin_d"], w), g["lin_mean"], rtol=1e-14)
    np.testing.assert_allclose(O.covariance_linear(g["lin_d"], w), g["lin_cov"], rtol=1e-13)
    np.testing.assert_allclose(O.trig_moment(g["tor_d"], w, 1), g["tor_m1"], rtol=1e-14)
    np.testing.assert_allclose(O.trig_moment(g["tor_d"], w, 2), g["tor_m2"], rtol=1e-14)
    np.testing.assert_allclose(O.mean_direction_torus(g["tor_d"], w), g["tor_meandir"], rtol=1e-14)
    np.testing.assert_allclose(O.mean_direction_torus(g["circ_d"], w), g["circ_meandir"], rtol=1e-14)
    np.testing.assert_allclose(O.mean_resultant(g["sph_d"], w), g["sph_res"], rtol=1e-14)
    np.testing.assert_allclose(O.mean_direction_sphere(g["sph_d"], w), g["sph_meandir"], rtol=1e-14)
    np.testing.assert_allclose(O.scatter(g["sph_d"], w), g["sph_scatter"], rtol=1e-12)
    assert np.array_equal(O.multinomial_indices(w, g["tor_sample_u"]), g["tor_sample_idx"])
    assert np.array_equal(O.sample_dirac(np.mod(g["tor_d"], TWO_PI), w, g["tor_sample_u"]), g["tor_sample"])
    np.testing.assert_allclose(O.reweigh(w, g["lin_lik"]), g["lin_reweighed"], rtol=1e-15)


@pytest.mark.parametrize("name", ["lik", "wide", "ties", "zeros", "uniform", "single"])
def test_resample_indices(golden, name):
    g = golden("resample_cases")
    assert np.array_equal(O.multinomial_indices(g[f"{name}_w"], g[f"{name}_u"]), g[f"{name}_idx"])


# ---- the reference's own known-answer values ---------------------------------
def test_known_answers_vmf():
    # pyrecest/tests/distributions/test_von_mises_fisher_distribution.py:99-171
    assert abs(O.vmf_norm_const(2.0, 3) - 2.0 / (4 * np.pi * np.sinh(2.0))) < 1e-15
    mu = np.array([0.0, 0.0, 1.0])
    x = np.array([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    p = O.pdf_vmf(x, mu, 2.0)
    np.testing.assert_allclose(p[0] / p[1], np.exp(2.0), rtol=1e-14)


def test_known_answers_hypertoroidal_wn():
    # pyrecest/tests/distributions/test_hypertoroidal_wrapped_normal_distribution.py:11-23
    mu = np.array([1.0, 2.0])
    C = np.array([[0.5, 0.1], [0.1, 0.3]])
    xs = np.array([[0.0, 1.0, 2.0], [1.0, 2.0, 3.0]]).T
    ref = np.array([0.0499028191873498, 0.425359477472412, 0.0499028191873498])
    np.testing.assert_allclose(O.pdf_wn_multi(xs, mu, C), ref, rtol=2e-6)


def test_known_answers_wn1d_series():
    # pyrecest/tests/distributions/test_wrapped_normal_distribution.py:17-50 (+-20-term series, rtol 1e-7)
    mu, sigma = 3.0, 1.5
    xs = np.linspace(0, 2 * np.pi, 50)
    k = np.arange(-20, 21)
    ser = np.sum(np.exp(-((xs[:, None] - mu + 2 * np.pi * k) ** 2) / (2 * sigma**2)), axis=1) / (np.sqrt(2 * np.pi) * sigma)
    np.testing.assert_allclose(O.pdf_wn_1d(xs, mu, sigma), ser, rtol=1e-7)
    np.testing.assert_allclose(O.pdf_wn_1d(xs, mu, 100.0), 1 / (2 * np.pi))


def test_systematic_definition():
    rng = np.random.default_rng(0)
    w = rng.random(1000)
    w /= w.sum()
    idx = O.systematic_indices(w, 0.37)
    assert np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() <= 999
    counts = np.bincount(idx, minlength=1000)
    assert np.all(np.abs(counts - 1000 * w) < 1.0 + 1e-9)


@pytest.mark.parametrize("tag,bd", [("c21", 2), ("c13", 1), ("c22", 2)])
def test_hypercylindrical_moments_against_reference_run(golden, tag, bd):
    """hybrid_mean / hybrid_moment / linear moments of HypercylindricalDiracDistribution,
    reweigh and the sample round trip, all from a run of the real reference (tests/golden/make_golden.py)"""
    g = golden("hypercylindrical")
    d, w = g[f"{tag}_d"], g[f"{tag}_w"]
    np.testing.assert_allclose(O.hybrid_moment(d, w, bd), g[f"{tag}_hybrid_moment"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(O.hybrid_mean(d, w, bd), g[f"{tag}_hybrid_mean"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(O.mean_linear(d[:, bd:], w), g[f"{tag}_linear_mean"], rtol=1e-13)
    np.testing.assert_allclose(O.covariance_linear(d[:, bd:], w), g[f"{tag}_linear_cov"], rtol=1e-12)
    w2 = O.reweigh(w, g[f"{tag}_lik"])
    np.testing.assert_allclose(w2, g[f"{tag}_reweighed_w"], rtol=1e-13)
    idx = O.multinomial_indices(g[f"{tag}_reweighed_w"], g[f"{tag}_sample_u"])
    assert np.array_equal(idx, g[f"{tag}_sample_idx"])
    assert np.array_equal(d[idx], g[f"{tag}_sample"])


@pytest.mark.parametrize("tag,manifold", [("torus", "hypertorus"), ("circle", "circle"), ("sphere", "hypersphere"), ("euclid", "Euclidean")])
def test_evaluation_summary_against_reference_run(golden, tag, manifold):
    """deviations and summary rows equal the reference's determine_all_deviations /
    summarize_filter_results on the same estimates"""
    g = golden("evaluation_summary")
    dev, rows = O.summarize(manifold, g[f"{tag}_est"], g[f"{tag}_gt"], g[f"{tag}_runtimes"], g[f"{tag}_failed"])
    np.testing.assert_allclose(dev, g[f"{tag}_dev"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(rows, g[f"{tag}_summary"], rtol=1e-12)


def test_se3_dirac_and_wn_filter_update_against_reference_run(golden):
    """the oracle's generic Dirac pieces on the reference's SE(3) run (7-component records) and its
    WrappedNormalFilter.update_nonlinear_particle run"""
    g = golden("se3_and_wn_filter")
    d, w = g["se3_d"], g["se3_w"]
    np.testing.assert_allclose(O.mean_linear(d[:, 3:], w), g["se3_linear_mean"], rtol=1e-13)   # the reference's slice
    np.testing.assert_allclose(O.covariance_linear(d[:, 3:], w), g["se3_linear_cov"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(O.scatter(d[:, :4], w), g["se3_quat_moment"], rtol=1e-13, atol=1e-15)
    wp = O.reweigh(w, g["se3_lik"])
    np.testing.assert_allclose(wp, g["se3_reweighed_w"], rtol=1e-14)
    idx = O.multinomial_indices(g["se3_reweighed_w"], g["se3_sample_u"])
    assert np.array_equal(idx, g["se3_sample_idx"]) and np.array_equal(d[idx], g["se3_sample"])
    for tag in ("wnf_a", "wnf_b"):
        mu0, sig0, z, kap = g[f"{tag}_in"]
        x = g[f"{tag}_samples"]
        wn = O.reweigh(O.uniform_w(x.shape[0]), O.pdf_vm(x, z, kap))
        m = O.trig_moment(x[:, None], wn)[0]
        out = [np.mod(np.angle(m), TWO_PI), np.sqrt(-2.0 * np.log(np.abs(m)))]
        np.testing.assert_allclose(out, g[f"{tag}_out"], rtol=1e-12)
