"""The public API (pyrecest_b200.filters / .distributions) against the golden runs of the real
reference with the same injected draws, plus the reference's own statistical filter tests."""
import math

import numpy as np
import pytest
import torch

from oracle import pf_oracle as O

pytestmark = pytest.mark.gpu


def host(t):
    return t.detach().cpu().numpy()


def test_euclid_cv_against_reference_run(golden):
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter, LinearTransition

    g = golden("euclid_cv")
    N, T = g["d0"].shape[0], g["Z"].shape[0]
    pf = EuclideanParticleFilter(N, 2)
    with prandom.inject(normals=np.concatenate([g["Z0"].ravel(), g["Z"].ravel()]), uniforms=g["U"]):
        pf.filter_state = GaussianDistribution(g["prior_mu"], g["prior_C"])
        np.testing.assert_allclose(host(pf.filter_state.d), g["d0"], rtol=1e-13, atol=1e-15)
        for t in range(T):
            pf.predict_nonlinear(LinearTransition(g["F"]), GaussianDistribution(np.zeros(2), g["Q"]))
            np.testing.assert_allclose(host(pf.filter_state.d), g["d_pred"][t], rtol=1e-12, atol=1e-13)
            pf.update_identity(GaussianDistribution(np.zeros(2), g["R"]), g["zs"][t])
            np.testing.assert_allclose(host(pf.filter_state.d), g["d_post"][t], rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(host(pf.filter_state.w), 1.0 / N)
            np.testing.assert_allclose(host(pf.get_point_estimate()), g["est"][t], rtol=1e-12)


def test_config0_full_size_against_reference_run(golden, config0_inputs):
    """BASELINE.json configs[0] as specified (EuclideanParticleFilter(10^4, 2), constant velocity, Gaussian noise and
    likelihood, 100 steps) through the public API with the draws the real reference consumed
    (tests/golden/config0.npz): every resampled index over the 100 steps (checksums per step, full index vectors at
    steps 0 / 49 / 99 through the particle sets), particles and estimates."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter, LinearTransition

    g, c = golden("config0"), config0_inputs
    N, T = int(g["N"]), int(g["T"])
    pf = EuclideanParticleFilter(N, 2)
    with prandom.inject(normals=np.concatenate([c["Z0"].ravel(), c["Z"].ravel()]), uniforms=c["U"]):
        pf.filter_state = GaussianDistribution(np.zeros(2), np.eye(2))
        for t in range(T):
            pf.predict_nonlinear(LinearTransition(c["F"]), GaussianDistribution(np.zeros(2), c["Q"]))
            pf.update_identity(GaussianDistribution(np.zeros(2), c["R"]), c["zs"][t])
            np.testing.assert_allclose(host(pf.get_point_estimate()), g["est"][t], rtol=1e-12, atol=1e-13, err_msg=str(t))
            if f"d_{t}" in g:
                # equal particle sets <=> equal index vectors (the predicted particles are distinct)
                np.testing.assert_allclose(host(pf.filter_state.d), g[f"d_{t}"], rtol=1e-12, atol=1e-13, err_msg=str(t))


def test_torus_t3_against_reference_run(golden):
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN
    from pyrecest_b200.filters import HypertoroidalParticleFilter

    g = golden("torus_t3")
    N, T = g["d0"].shape[0], g["Z"].shape[0]
    pf = HypertoroidalParticleFilter(N, 3)
    assert pf.filter_state.d.shape == (N, 3)
    with prandom.inject(normals=np.concatenate([g["Z0"].ravel(), g["Z"].ravel()]), uniforms=g["U"]):
        pf.filter_state = HWN(g["mu0"], g["C"])
        assert np.abs(host(pf.filter_state.d) - g["d0"]).max() < 1e-13
        for t in range(T):
            pf.predict_identity(HWN(np.zeros(3), g["Cn"]))
            assert np.abs(host(pf.filter_state.d) - g["d_pred"][t]).max() < 1e-12
            pf.update_identity(HWN(np.zeros(3), g["Cl"]), g["zs"][t])
            assert np.abs(host(pf.filter_state.d) - g["d_post"][t]).max() < 1e-12
            np.testing.assert_allclose(host(pf.get_point_estimate()), g["est"][t], rtol=1e-12)
            # cached estimate == recomputed moment
            pf.filter_state._est_cache = None
            np.testing.assert_allclose(host(pf.get_point_estimate()), g["est"][t], rtol=1e-12)


def test_circle_against_reference_run(golden):
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import (
        HypertoroidalWrappedNormalDistribution as HWN,
        VonMisesDistribution,
        WrappedNormalDistribution,
    )
    from pyrecest_b200.filters import CircularParticleFilter

    g = golden("circle_case")
    N = g["d0"].shape[0]
    pf = CircularParticleFilter(N)
    assert pf.filter_state.d.shape == (N,)
    np.testing.assert_allclose(host(pf.filter_state.d), g["d0"], rtol=1e-15, atol=1e-15)
    pf.filter_state.d = g["d0"]
    liks = [VonMisesDistribution(np.array(2.0), 1.5), WrappedNormalDistribution(np.array(4.0), np.array(0.7))]
    with prandom.inject(normals=g["Z"], uniforms=g["U"]):
        for t in range(2):
            pf.predict_identity(HWN(np.zeros(1), np.array([[float(g["sig_n"]) ** 2]])))
            assert pf.filter_state.d.shape == (N,)
            assert np.abs(host(pf.filter_state.d) - g["d_pred"][t]).max() < 1e-13
            np.testing.assert_allclose(host(liks[t].pdf(pf.filter_state.d)), g["lik"][t], rtol=1e-12)
            pf.update_nonlinear_using_likelihood(liks[t])
            assert np.abs(host(pf.filter_state.d) - g["d_post"][t]).max() < 1e-13
            np.testing.assert_allclose(host(pf.get_point_estimate()), g["est"][t][0], rtol=1e-12)


def test_sphere_filter_against_oracle():
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import VonMisesFisherDistribution as VMF
    from pyrecest_b200.filters import HypersphericalParticleFilter

    rng = np.random.default_rng(12)
    N = 5000
    trip0 = np.column_stack([rng.random(N), rng.standard_normal((N, 2))])
    trip = np.column_stack([rng.random(2 * N), rng.standard_normal((2 * N, 2))]).reshape(2, N, 3)
    U = rng.random((2, N))
    z = np.array([1.0, 1.0, 1.0]) / math.sqrt(3)
    pf = HypersphericalParticleFilter(N, 2)
    with prandom.inject(vmf_triples=np.concatenate([trip0.ravel(), trip.ravel()]), uniforms=U):
        pf.filter_state = VMF(np.array([0.0, 0.0, 1.0]), 5.0)
        d = O.predict_sphere_vmf(np.tile([0.0, 0.0, 1.0], (N, 1)), trip0, 5.0)
        np.testing.assert_allclose(host(pf.filter_state.d), d, atol=2e-14)
        for t in range(2):
            pf.predict_identity(VMF(np.array([0.0, 0.0, 1.0]), 50.0))
            pf.update_nonlinear_using_likelihood(VMF(z, 10.0))
            ref = O.step("sphere", d, trip[t], "vmf", {"mu": z, "kappa": 10.0}, U[t], kappa_noise=50.0)
            np.testing.assert_allclose(host(pf.filter_state.d), ref["d"], atol=1e-13)
            np.testing.assert_allclose(host(pf.get_point_estimate()), ref["est"], rtol=1e-12)
            d = ref["d"]
    # additive Gaussian noise + renormalise
    from pyrecest_b200.distributions import GaussianDistribution

    Zs = rng.standard_normal((N, 3))
    with prandom.inject(normals=Zs):
        pf.predict_identity(GaussianDistribution(np.zeros(3), 0.01 * np.eye(3)))
    ref = O.predict_sphere_additive(d, O.gaussian_rows(Zs, 0.01 * np.eye(3)))
    np.testing.assert_allclose(host(pf.filter_state.d), ref, rtol=1e-12, atol=1e-14)


def test_von_mises_noise_on_torus_and_systematic_resampling():
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN, VonMisesDistribution
    from pyrecest_b200.filters import HypertoroidalParticleFilter

    rng = np.random.default_rng(13)
    N = 4000
    C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
    d0 = rng.random((N, 3)) * 2 * np.pi
    vm = rng.vonmises(0.0, 10.0, size=(N, 3))
    z = np.array([1.0, 2.0, 3.0])
    pf = HypertoroidalParticleFilter(N, 3)
    pf.filter_state.d = d0
    pf.resampling = "systematic"
    with prandom.inject(vonmises=vm, uniforms=np.array([0.4321])):
        pf.predict_identity(VonMisesDistribution(np.zeros(3), np.array([10.0, 10.0, 10.0])))
        dp = O.predict_torus(d0, vm, shift=True)
        assert np.array_equal(host(pf.filter_state.d), dp)
        pf.update_identity(HWN(np.zeros(3), 0.5 * C3), z)
    lik = O.pdf_wn_multi(dp, z, 0.5 * C3)
    idx = O.systematic_indices(O.reweigh(O.uniform_w(N), lik), 0.4321)
    got = host(pf.filter_state.d)
    mism = np.any(got != dp[idx], axis=1).sum()
    assert mism <= 2, f"{mism} particles differ from the oracle's systematic resampling"
    assert np.all(np.diff(idx) >= 0)


def test_statistical_convergence_like_reference_tests():
    """pyrecest/tests/filters/test_euclidean_particle_filter.py:24-34 and
    test_hypertoroidal_particle_filter.py:50-58 (device RNG, statistical tolerance)."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution, HypertoroidalWrappedNormalDistribution as HWN
    from pyrecest_b200.filters import EuclideanParticleFilter, HypertoroidalParticleFilter

    prandom.seed(42)
    C = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
    mu = np.array([5.0, 6.0, 7.0])
    pf = EuclideanParticleFilter(1000, 3)
    pf.filter_state = GaussianDistribution(mu, C)
    forced_mean = np.array([1.0, 2.0, 3.0])
    sys_noise = GaussianDistribution(np.zeros(3), 0.5 * C)
    for _ in range(50):
        pf.predict_identity(sys_noise)
        for _ in range(3):
            pf.update_identity(sys_noise, forced_mean)
    np.testing.assert_allclose(host(pf.get_point_estimate()), forced_mean, atol=0.15)

    hpf = HypertoroidalParticleFilter(500, 3)
    hpf.filter_state = HWN(np.array([1.0, 1.0, 1.0]) + math.pi / 2, C)
    forced = np.array([1.0, 2.0, 3.0])
    for _ in range(5):
        hpf.predict_identity(HWN(np.zeros(3), 0.5 * C))
        for _ in range(3):
            hpf.update_identity(HWN(np.zeros(3), 0.5 * C), forced)
    est = host(hpf.get_point_estimate())
    assert est.shape == (3,)
    np.testing.assert_allclose(est, forced, atol=0.15)


def test_api_shapes_errors_and_rejections():
    from pyrecest_b200.distributions import (
        GaussianDistribution,
        HypertoroidalDiracDistribution,
        LinearDiracDistribution,
    )
    from pyrecest_b200.filters import EuclideanParticleFilter, HypertoroidalParticleFilter, LinearTransition

    with pytest.raises(ValueError):
        EuclideanParticleFilter(0, 2)
    with pytest.raises(ValueError):
        EuclideanParticleFilter(10, 0)
    pf = EuclideanParticleFilter(100, 1)
    assert pf.filter_state.d.shape == (100, 1)
    with pytest.raises(ValueError):
        pf.filter_state = LinearDiracDistribution(np.zeros((50, 1)))
    pf = EuclideanParticleFilter(64, 3)
    with pytest.raises(NotImplementedError):
        pf.predict_nonlinear(lambda x: x**2, None)
    with pytest.raises(NotImplementedError):
        pf.update_nonlinear_using_likelihood(lambda x: np.ones(64))
    with pytest.raises(ValueError):
        pf.predict_nonlinear(LinearTransition(np.eye(2)), None)
    # all particles infinitely far -> the reference's "sum of all weights > 0" assertion
    pf.filter_state.d = np.full((64, 3), 1e200)
    with pytest.raises(AssertionError):
        pf.update_identity(GaussianDistribution(np.zeros(3), np.eye(3)), np.zeros(3))
    # noise-free predict equals apply_function (test_circular_particle_filter.py:71-83 pattern)
    hp = HypertoroidalParticleFilter(200, 2)
    rng = np.random.default_rng(0)
    d0 = rng.random((200, 2)) * 2 * np.pi
    hp.filter_state = HypertoroidalDiracDistribution(d0)
    F = np.array([[1.0, 0.5], [0.0, 1.0]])
    expect = hp.filter_state.apply_function(LinearTransition(F))
    hp.predict_nonlinear(LinearTransition(F), None)
    assert np.array_equal(host(hp.filter_state.d), host(expect.d))
    np.testing.assert_allclose(host(expect.d), np.mod(d0 @ F.T, 2 * np.pi), atol=1e-12)
    # Dirac sample with n != N, weights renormalisation warning
    with pytest.warns(RuntimeWarning):
        dd = HypertoroidalDiracDistribution(d0, np.ones(200))
    s = dd.sample(77)
    assert s.shape == (77, 2)
    np.testing.assert_allclose(host(dd.w).sum(), 1.0)


def test_dirac_moments_api(golden):
    from pyrecest_b200.distributions import (
        CircularDiracDistribution,
        HypersphericalDiracDistribution,
        HypertoroidalDiracDistribution,
        LinearDiracDistribution,
    )

    g = golden("dirac_moments")
    w = g["w"]
    ld = LinearDiracDistribution(g["lin_d"], w)
    np.testing.assert_allclose(host(ld.mean()), g["lin_mean"], rtol=1e-12)
    np.testing.assert_allclose(host(ld.covariance()), g["lin_cov"], rtol=1e-11)
    td = HypertoroidalDiracDistribution(g["tor_d"], w)
    np.testing.assert_allclose(host(td.trigonometric_moment(1)), g["tor_m1"], rtol=1e-12)
    np.testing.assert_allclose(host(td.trigonometric_moment(2)), g["tor_m2"], rtol=1e-11)
    np.testing.assert_allclose(host(td.mean_direction()), g["tor_meandir"], rtol=1e-12)
    cd = CircularDiracDistribution(g["circ_d"], w)
    np.testing.assert_allclose(host(cd.mean_direction()), g["circ_meandir"][0], rtol=1e-12)
    sd = HypersphericalDiracDistribution(g["sph_d"], w)
    np.testing.assert_allclose(host(sd.mean_resultant_vector()), g["sph_res"], rtol=1e-12)
    np.testing.assert_allclose(host(sd.mean_direction()), g["sph_meandir"], rtol=1e-12)
    np.testing.assert_allclose(host(sd.moment()), g["sph_scatter"], rtol=1e-11, atol=1e-15)
    from pyrecest_b200 import random as prandom

    with prandom.inject(uniforms=g["tor_sample_u"]):
        smp = td.sample(1234)
    assert np.array_equal(host(smp), g["tor_sample"])
    from pyrecest_b200.distributions import GaussianDistribution

    rw = ld.reweigh(GaussianDistribution(np.ones(4), np.eye(4)))
    np.testing.assert_allclose(host(rw.w), g["lin_reweighed"], rtol=1e-12)


def test_fused_rng_path_is_reproducible_and_seed_dependent():
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN, VonMisesDistribution
    from pyrecest_b200.filters import HypertoroidalParticleFilter

    C = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])

    def run(seed):
        prandom.seed(seed)
        pf = HypertoroidalParticleFilter(20_000, 3)
        pf.filter_state = HWN(np.array([1.0, 1.0, 1.0]) + math.pi / 2, C)
        for _ in range(5):
            pf.predict_identity(VonMisesDistribution(np.zeros(3), np.array([10.0, 10.0, 10.0])))
            pf.update_identity(HWN(np.zeros(3), 0.5 * C), np.array([1.0, 2.0, 3.0]))
        return host(pf.filter_state.d), host(pf.get_point_estimate())

    d1, e1 = run(7)
    d2, e2 = run(7)
    d3, e3 = run(8)
    assert np.array_equal(d1, d2) and np.array_equal(e1, e2)
    assert not np.array_equal(d1, d3)
    np.testing.assert_allclose(e1, [1.0, 2.0, 3.0], atol=0.1)
    np.testing.assert_allclose(e3, [1.0, 2.0, 3.0], atol=0.1)


def test_fp32_storage_option_matches_oracle_to_1e5():
    """dtype=float32 stores particles (and noise) in fp32, arithmetic / weights / CDF stay fp64: states and
    estimates within 1e-5 relative of the fp64 oracle (BASELINE.json north_star)."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter, LinearTransition

    rng = np.random.default_rng(77)
    N, D = 6000, 4
    F = np.eye(4)
    F[0, 2] = F[1, 3] = 1.0
    Q, Rm = 0.1 * np.eye(4), 0.5 * np.eye(4)
    d0 = rng.standard_normal((N, D))
    Z = rng.standard_normal((N, D))
    U = rng.random(N)
    z = np.array([0.3, -0.2, 0.1, 0.0])
    pf = EuclideanParticleFilter(N, D, dtype=torch.float32)
    pf.filter_state.d = torch.as_tensor(d0, dtype=torch.float32)
    assert pf.filter_state.d.dtype == torch.float32
    with prandom.inject(normals=Z, uniforms=U):
        pf.predict_nonlinear(LinearTransition(F), GaussianDistribution(np.zeros(4), Q))
        dp = host(pf.filter_state.d).astype(np.float64)
        pf.update_identity(GaussianDistribution(np.zeros(4), Rm), z)
    ref_pred = O.predict_euclid(d0, O.gaussian_rows(Z, Q), F=F)
    np.testing.assert_allclose(dp, ref_pred, rtol=1e-5, atol=1e-5)
    # resampling of the fp32 particles: compare with the oracle run on exactly those particles
    lik = O.pdf_gauss(dp, z, Rm)
    dn, wn, idx, _ = O.update_resample(dp, O.uniform_w(N), lik, U)
    got = host(pf.filter_state.d).astype(np.float64)
    assert (np.abs(got - dn).max(axis=1) > 0).sum() <= 1
    np.testing.assert_allclose(host(pf.get_point_estimate()), O.mean_linear(dn, wn), rtol=1e-5, atol=1e-6)
    assert pf.filter_state.d.dtype == torch.float32 and pf.filter_state.w.dtype == torch.float64


def test_nonadditive_predict_and_hemispherical_filter(golden):
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HyperhemisphericalDiracDistribution, VonMisesFisherDistribution as VMF
    from pyrecest_b200.filters import (EuclideanParticleFilter, HyperhemisphericalParticleFilter,
                                       HypertoroidalParticleFilter, LinearNoiseTransition)

    rng = np.random.default_rng(31)
    N = 4000
    d0 = rng.standard_normal((N, 3))
    samples = rng.standard_normal((50, 2))
    weights = rng.random(50)
    u = rng.random(N)
    F = np.array([[1.0, 0.1, 0.0], [0.0, 1.0, 0.1], [0.0, 0.0, 1.0]])
    G = rng.standard_normal((3, 2))
    b = np.array([0.1, 0.2, 0.3])
    pf = EuclideanParticleFilter(N, 3)
    pf.filter_state.d = d0
    with prandom.inject(uniforms=u):
        pf.predict_nonlinear_nonadditive(LinearNoiseTransition(F, G, b), samples, weights)
    np.testing.assert_allclose(host(pf.filter_state.d), O.predict_nonadditive(d0, samples, weights, u, F, G, b), rtol=1e-12, atol=1e-13)
    with pytest.raises(NotImplementedError):
        pf.predict_nonlinear_nonadditive(lambda x, n: x + n, samples, weights)
    t0 = rng.random((N, 2)) * 2 * np.pi
    hp = HypertoroidalParticleFilter(N, 2)
    hp.filter_state.d = t0
    with prandom.inject(uniforms=u):
        hp.predict_nonlinear_nonadditive(LinearNoiseTransition(np.eye(2), np.eye(2)), samples, weights)
    ref = O.predict_nonadditive(t0, samples, weights, u, np.eye(2), np.eye(2), torus=True)
    diff = np.abs(host(hp.filter_state.d) - ref)
    assert np.minimum(diff, 2 * np.pi - diff).max() < 1e-12
    # hemispherical filter (pyrecest/tests/filters/test_hyperhemispherical_particle_filter.py:38-50 pattern)
    prandom.seed(1)
    hh = HyperhemisphericalParticleFilter(2000, 2)
    assert hh.filter_state.d.shape == (2000, 3) and isinstance(hh.filter_state, HyperhemisphericalDiracDistribution)
    mode = np.array([1.0, 0.0, 0.0])
    hh.set_state(VMF(mode, 20.0))
    assert float(hh.filter_state.d[:, -1].min()) >= 0.0
    est = host(hh.get_point_estimate())
    assert min(np.linalg.norm(est - mode), np.linalg.norm(est + mode)) < 0.1
    for _ in range(3):
        hh.predict_identity(VMF(np.array([0.0, 0.0, 1.0]), 100.0))
        assert float(hh.filter_state.d[:, -1].min()) >= 0.0
        hh.update_nonlinear_using_likelihood(VMF(np.array([0.0, 0.6, 0.8]), 30.0))
    est = host(hh.get_point_estimate())
    assert np.linalg.norm(est - np.array([0.0, 0.6, 0.8])) < 0.3 and est[-1] >= 0
    np.testing.assert_allclose(est, O.mean_axis(host(hh.filter_state.d), host(hh.filter_state.w)), atol=1e-10)


@pytest.mark.parametrize("tag,bd", [("c21", 2), ("c13", 1), ("c22", 2)])
def test_hypercylindrical_dirac_against_reference_run(golden, tag, bd):
    """HypercylindricalDiracDistribution on the device vs the reference's own outputs"""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HypercylindricalDiracDistribution

    g = golden("hypercylindrical")
    d, w = g[f"{tag}_d"], g[f"{tag}_w"]
    hd = HypercylindricalDiracDistribution(bd, d, w)
    host = lambda t: t.detach().cpu().numpy()  # noqa: E731
    np.testing.assert_allclose(host(hd.hybrid_moment()), g[f"{tag}_hybrid_moment"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(host(hd.hybrid_mean()), g[f"{tag}_hybrid_mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(host(hd.mean()), g[f"{tag}_hybrid_mean"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(host(hd.linear_mean()), g[f"{tag}_linear_mean"], rtol=1e-12)
    np.testing.assert_allclose(host(hd.linear_covariance()), g[f"{tag}_linear_cov"], rtol=1e-11)
    np.testing.assert_allclose(host(hd.marginalize_linear().trigonometric_moment(1)), g[f"{tag}_per_m1"], rtol=1e-12)
    np.testing.assert_allclose(host(hd.marginalize_periodic().mean()), g[f"{tag}_linear_mean"], rtol=1e-12)
    rw = hd.reweigh(g[f"{tag}_lik"])  # likelihood values evaluated by the caller
    np.testing.assert_allclose(host(rw.w), g[f"{tag}_reweighed_w"], rtol=1e-12)
    np.testing.assert_allclose(host(rw.hybrid_mean()), g[f"{tag}_reweighed_mean"], rtol=1e-11)
    rw.w = g[f"{tag}_reweighed_w"]  # the reference's weights bit for bit: the draw below must then pick the same indices
    with prandom.inject(uniforms=g[f"{tag}_sample_u"]):
        smp = rw.sample(500)
    assert np.array_equal(host(smp), g[f"{tag}_sample"])
    with pytest.raises(ValueError):
        HypercylindricalDiracDistribution(0, d, w)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("D", [3, 5, 6])
def test_padded_gather_copy_gives_identical_particles(D, dtype):
    """the gather of the multinomial resampling may read the padded copy the predict kernel left behind
    (pfb_predict_params.x_pad_out): same particles, same estimate, bit for bit"""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter, LinearTransition

    rng = np.random.default_rng(D)
    A = rng.normal(size=(D, D))
    Q = A @ A.T / D + 0.1 * np.eye(D)
    F = np.eye(D) + 0.05 * rng.normal(size=(D, D))
    outs = []
    for padded in (True, False):
        prandom.seed(2024)
        pf = EuclideanParticleFilter(60_000, D, dtype=dtype)
        pf.padded_gather, pf.padded_gather_min = padded, 0
        pf.filter_state = GaussianDistribution(np.zeros(D), np.eye(D))
        for t in range(3):
            pf.predict_nonlinear(LinearTransition(F), GaussianDistribution(np.zeros(D), 0.2 * Q))
            pf.update_identity(GaussianDistribution(np.zeros(D), Q), 0.3 * np.ones(D))
            est = pf.get_point_estimate()
        assert (pf._pad is not None) == padded
        outs.append((pf.filter_state.d.cpu().numpy().copy(), est.cpu().numpy().copy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


def test_failed_update_leaves_predicted_particles_with_padded_gather():
    """when the weight check fails (here: a NaN measurement) the reference has already applied the predict; with the
    padded gather copy the compact predicted particles are not stored by the fused kernel, so they are restored"""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import EuclideanParticleFilter, IdentityTransition

    states = []
    for padded in (True, False):
        prandom.seed(5)
        pf = EuclideanParticleFilter(50_000, 3)
        pf.padded_gather, pf.padded_gather_min = padded, 0
        pf.filter_state = GaussianDistribution(np.zeros(3), np.eye(3))
        pf.predict_nonlinear(IdentityTransition(), GaussianDistribution(np.zeros(3), 0.1 * np.eye(3)))
        with pytest.raises(AssertionError):
            pf.update_identity(GaussianDistribution(np.zeros(3), np.eye(3)), np.array([np.nan, 0.0, 0.0]))
        states.append(pf.filter_state.d.cpu().numpy().copy())
    assert np.isfinite(states[0]).all()
    assert np.array_equal(states[0], states[1])


def test_association_likelihood_reference_known_answers():
    """pyrecest/tests/filters/test_circular_particle_filter.py:115-130: sum_i w_i pdf(d_i) with a uniform likelihood is
    1 / (2 pi) to 1e-10, and a von Mises likelihood centred on the particles gives more; computed by the weighting
    kernel + the log-sum-exp reduction (no host arithmetic), also for non-uniform weights and on S^2."""
    from pyrecest_b200.distributions import (
        CircularDiracDistribution,
        CircularUniformDistribution,
        HypersphericalDiracDistribution,
        HypersphericalUniformDistribution,
        VonMisesDistribution,
        VonMisesFisherDistribution,
    )
    from pyrecest_b200.filters import CircularParticleFilter, HypersphericalParticleFilter

    dist = CircularDiracDistribution(np.array([1.0, 2.0, 3.0]), np.array([1 / 3, 1 / 3, 1 / 3]))
    pf = CircularParticleFilter(3)
    pf.filter_state = dist
    assert abs(float(pf.association_likelihood(CircularUniformDistribution())) - 1.0 / (2.0 * math.pi)) < 1e-10
    assert float(pf.association_likelihood(VonMisesDistribution(np.array(2.0), np.array(1.0)))) > 1.0 / (2.0 * math.pi)
    assert float(pf.compute_association_likelihood(CircularUniformDistribution())) == float(
        pf.association_likelihood(CircularUniformDistribution()))
    # weighted particles against the oracle's restatement: sum w_i pdf(x_i)
    rng = np.random.default_rng(5)
    d = rng.random(5000) * 2 * np.pi
    w = rng.random(5000)
    w /= w.sum()
    pf = CircularParticleFilter(5000)
    pf.filter_state = CircularDiracDistribution(d, w)
    vm = VonMisesDistribution(np.array(2.0), np.array(1.5))
    ref = float(np.sum(w * O.likelihood("vm", d, {"mu": 2.0, "kappa": 1.5})))
    np.testing.assert_allclose(float(pf.association_likelihood(vm)), ref, rtol=1e-12)
    # sphere: uniform likelihood = 1 / (4 pi); VMF against the oracle
    x = rng.standard_normal((4000, 3))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    sp = HypersphericalParticleFilter(4000, 2)
    sp.filter_state = HypersphericalDiracDistribution(x)
    np.testing.assert_allclose(float(sp.association_likelihood(HypersphericalUniformDistribution(2))), 1 / (4 * math.pi), rtol=1e-12)
    mu = np.array([0.0, 0.6, 0.8])
    ref = float(np.mean(O.likelihood("vmf", x, {"mu": mu, "kappa": 3.0})))
    np.testing.assert_allclose(float(sp.association_likelihood(VonMisesFisherDistribution(mu, 3.0))), ref, rtol=1e-12)


def test_se3_dirac_against_reference_run(golden):
    """SE3DiracDistribution (7-component records: unit quaternion, translation): the marginals, linear_mean /
    linear_covariance (on the reference's d[:, bound_dim:] slice), reweigh with caller-computed likelihood values,
    and sample with the injected uniforms -- indices bit for bit (tests/golden/se3_and_wn_filter.npz)."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HyperhemisphericalDiracDistribution, LinearDiracDistribution, SE3DiracDistribution

    g = golden("se3_and_wn_filter")
    s3 = SE3DiracDistribution(g["se3_d"], g["se3_w"])
    assert (s3.dim, s3.bound_dim, s3.lin_dim, s3.input_dim) == (6, 3, 3, 7)
    np.testing.assert_allclose(host(s3.linear_mean()), g["se3_linear_mean"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(host(s3.linear_covariance()), g["se3_linear_cov"], rtol=1e-11, atol=1e-13)
    ml = s3.marginalize_linear()
    assert isinstance(ml, HyperhemisphericalDiracDistribution) and ml.d.shape == (700, 4)
    np.testing.assert_allclose(host(ml.moment()), g["se3_quat_moment"], rtol=1e-12, atol=1e-14)
    mp = s3.marginalize_periodic()
    assert isinstance(mp, LinearDiracDistribution)
    np.testing.assert_allclose(host(mp.mean()), g["se3_periodic_marginal_mean"], rtol=1e-12, atol=1e-14)
    rw = s3.reweigh(g["se3_lik"])
    np.testing.assert_allclose(host(rw.w), g["se3_reweighed_w"], rtol=1e-12)
    with prandom.inject(uniforms=g["se3_sample_u"]):
        smp = host(rw.sample(400))
    assert np.array_equal(smp, g["se3_d"][g["se3_sample_idx"]])
    assert np.array_equal(smp, g["se3_sample"])
    # hybrid_mean / mean: mean axis of the quaternion part + mean translation (the reference raises here)
    hm = host(s3.mean())
    assert hm.shape == (7,) and abs(np.linalg.norm(hm[:4]) - 1.0) < 1e-12 and hm[3] >= 0
    np.testing.assert_allclose(hm[4:], g["se3_w"] @ g["se3_d"][:, 4:], rtol=1e-12)
    with pytest.raises(AssertionError):
        SE3DiracDistribution(np.ones((4, 7)))


@pytest.mark.parametrize("tag", ["wnf_a", "wnf_b"])
def test_wrapped_normal_filter_particle_update_against_reference_run(golden, tag):
    """WrappedNormalFilter.update_nonlinear_particle (wrapped_normal_filter.py:27-32) with the reference's 100
    state samples injected: the moment-matched wrapped normal equals the reference's."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import VonMisesDistribution, WrappedNormalDistribution
    from pyrecest_b200.filters import WrappedNormalFilter

    g = golden("se3_and_wn_filter")
    mu0, sig0, z, kap = g[f"{tag}_in"]
    f = WrappedNormalFilter(WrappedNormalDistribution(np.array(mu0), np.array(sig0)))
    smp = g[f"{tag}_samples"]
    f.filter_state.sample = lambda n: torch.as_tensor(smp[:n], device="cuda:0")
    f.update_nonlinear_particle(VonMisesDistribution(np.array(0.0), kap), z)
    np.testing.assert_allclose([float(f.filter_state.mu), float(np.asarray(f.filter_state.sigma).reshape(-1)[0])], g[f"{tag}_out"], rtol=1e-11)
    assert f.get_point_estimate() == pytest.approx(g[f"{tag}_out"][0], rel=1e-11)
    with pytest.raises(NotImplementedError):
        f.update_nonlinear_particle(lambda zz, x: x, z)
    f.predict_identity(WrappedNormalDistribution(np.array(6.0), np.array(0.3)))
    assert float(f.filter_state.mu) == pytest.approx((g[f"{tag}_out"][0] + 6.0) % (2 * np.pi), rel=1e-12)
    assert float(np.asarray(f.filter_state.sigma).reshape(-1)[0]) == pytest.approx(math.hypot(g[f"{tag}_out"][1], 0.3), rel=1e-12)


def test_hemisphere_cart_prod_filter_like_the_reference_tests():
    """pyrecest/tests/filters/test_hyperhemisphere_cart_prod_particle_filter.py: init / set_state / predict / update on
    two S^3 hemispheres (8-component records), with the reference's own parameters; the likelihood of test_update
    (abs(sum(x))) is passed as values, a Python callable is rejected."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import VonMisesFisherDistribution
    from pyrecest_b200.filters import HyperhemisphereCartProdParticleFilter, IdentityTransition

    prandom.seed(2)
    n = 1000
    pf = HyperhemisphereCartProdParticleFilter(n, 3, 2)
    assert pf.filter_state.d.shape == (n, 8)
    pf.filter_state = VonMisesFisherDistribution(np.array([0.0, 0.0, 0.0, 1.0]), 1.0)
    d = host(pf.filter_state.d)
    assert d.shape == (n, 8) and np.all(d[:, 3] >= 0) and np.all(d[:, 7] >= 0)
    np.testing.assert_allclose(np.linalg.norm(d[:, :4], axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(np.linalg.norm(d[:, 4:], axis=1), 1.0, atol=1e-12)
    pf.predict_nonlinear_each_part(IdentityTransition(), VonMisesFisherDistribution(np.array([0.0, 0.0, 0.0, 1.0]), 1.0))
    d = host(pf.filter_state.d)
    np.testing.assert_allclose(np.linalg.norm(d[:, 4:], axis=1), 1.0, atol=1e-12)
    pf.update_nonlinear_using_likelihood(np.abs(d.sum(axis=1)))
    assert pf.filter_state.d.shape == (n, 8)
    np.testing.assert_allclose(host(pf.filter_state.w), 1.0 / n)
    with pytest.raises(NotImplementedError):
        pf.update_nonlinear_using_likelihood(lambda x: np.abs(x.sum(axis=1)))
    with pytest.raises(NotImplementedError):
        pf.predict_nonlinear_each_part(lambda x: x, None)


def test_hemisphere_cart_prod_filter_tracks_with_watson_and_vmf_noise():
    """two S^2 hemispheres, Watson system noise (hemispherical) on the device, per-part von Mises-Fisher likelihoods:
    the mean axes move to the measured axes; injected uniforms give the oracle's resampling indices."""
    from oracle import pf_oracle as O
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import HyperhemisphericalWatsonDistribution, VonMisesFisherDistribution
    from pyrecest_b200.filters import HyperhemisphereCartProdParticleFilter, IdentityTransition

    prandom.seed(4)
    n = 20_000
    pf = HyperhemisphereCartProdParticleFilter(n, 2, 2)
    pf.filter_state = VonMisesFisherDistribution(np.array([0.0, 0.0, 1.0]), 2.0)
    z1 = np.array([0.6, 0.0, 0.8])
    z2 = np.array([0.0, -0.6, 0.8])
    for _ in range(6):
        pf.predict_nonlinear_each_part(IdentityTransition(), HyperhemisphericalWatsonDistribution(np.array([0.0, 0.0, 1.0]), 60.0))
        pf.update_nonlinear_using_likelihood((VonMisesFisherDistribution(z1, 40.0), VonMisesFisherDistribution(z2, 40.0)))
    est = host(pf.get_point_estimate())
    assert est.shape == (6,)
    assert abs(est[:3] @ z1) > 0.99 and abs(est[3:] @ z2) > 0.99, est
    # resampling indices with injected uniforms == the oracle's (weights from the caller's likelihood values)
    d = host(pf.filter_state.d)
    lik = np.exp(3.0 * d[:, 2]) * (1.0 + d[:, 5] ** 2)
    u = np.random.default_rng(0).random(n)
    with prandom.inject(uniforms=u):
        pf.update_nonlinear_using_likelihood(lik)
    idx = O.multinomial_indices(O.reweigh(O.uniform_w(n), lik), u)
    assert np.array_equal(host(pf.filter_state.d), d[idx])


@pytest.mark.parametrize("q,kappa", [(3, 8.0), (4, 2.5), (4, -6.0), (6, 15.0)])
def test_watson_sampler_moments(q, kappa):
    """cart_prod.watson_rows (angular-central-Gaussian rejection): E[(mu . x)^2] equals d/dkappa log 1F1(1/2; q/2; kappa)
    = 1F1(3/2; q/2 + 1; kappa) / (q 1F1(1/2; q/2; kappa)), and the draws are symmetric about the axis."""
    from scipy.special import hyp1f1

    from pyrecest_b200 import random as prandom
    from pyrecest_b200.cart_prod import watson_rows

    prandom.seed(q)
    n = 400_000
    rng = np.random.default_rng(q)
    mu = rng.standard_normal(q)
    mu /= np.linalg.norm(mu)
    modes = torch.as_tensor(mu, device="cuda:0").expand(n, q).contiguous()
    x = host(watson_rows(modes, kappa))
    np.testing.assert_allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-12)
    t = x @ mu
    want = hyp1f1(1.5, 0.5 * q + 1.0, kappa) / (q * hyp1f1(0.5, 0.5 * q, kappa))
    assert abs((t * t).mean() - want) < 5.0 * (t * t).std() / math.sqrt(n)
    assert abs(t.mean()) < 5.0 / math.sqrt(n)
    perp = x - np.outer(t, mu)
    assert np.abs(perp.mean(axis=0)).max() < 5.0 / math.sqrt(n)


@pytest.mark.parametrize("q,kappa", [(2, 3.0), (4, 1.0), (4, 25.0), (5, 6.0)])
def test_vmf_rejection_sampler_moments(q, kappa):
    """cart_prod.vmf_rows (Wood's sampler, S^d with d != 2): mean resultant length A_q(kappa) = I_{q/2}(kappa) / I_{q/2-1}(kappa)
    along the mode, nothing across it; the pdf of the samples' own distribution integrates consistently (VMF.sample)."""
    from scipy.special import ive

    from pyrecest_b200 import random as prandom
    from pyrecest_b200.cart_prod import vmf_rows

    prandom.seed(10 + q)
    n = 400_000
    rng = np.random.default_rng(q)
    mu = rng.standard_normal(q)
    mu /= np.linalg.norm(mu)
    modes = torch.as_tensor(mu, device="cuda:0").expand(n, q).contiguous()
    x = host(vmf_rows(modes, kappa))
    np.testing.assert_allclose(np.linalg.norm(x, axis=1), 1.0, atol=1e-12)
    t = x @ mu
    want = ive(0.5 * q, kappa) / ive(0.5 * q - 1.0, kappa)
    assert abs(t.mean() - want) < 5.0 * t.std() / math.sqrt(n)
    perp = x - np.outer(t, mu)
    assert np.abs(perp.mean(axis=0)).max() < 5.0 / math.sqrt(n)


def test_watson_pdf_against_the_reference():
    """pdf of WatsonDistribution / HyperhemisphericalWatsonDistribution at the mode and across it (values from a run of the
    reference, watson_distribution.py:70-84, hyperhemispherical_watson_distribution.py:21-22)"""
    from pyrecest_b200.distributions import HyperhemisphericalWatsonDistribution, WatsonDistribution

    mu = np.array([0.6, 0.0, 0.0, 0.8])
    xs = np.array([mu, np.roll(mu, 1)])
    np.testing.assert_allclose(host(WatsonDistribution(mu, 7.5).pdf(xs)), [1.608853460435332, 0.005009204238815092], rtol=1e-12)
    np.testing.assert_allclose(host(HyperhemisphericalWatsonDistribution(mu, 7.5).pdf(xs)), [3.217706920870664, 0.010018408477630183], rtol=1e-12)
