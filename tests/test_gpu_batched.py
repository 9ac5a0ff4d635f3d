"""Batched Monte-Carlo runner: every run of the batch must equal a single filter run with the same draws
(oracle per run), including run lengths that are not a multiple of the scan tile."""
import numpy as np
import pytest
import torch

from oracle import pf_oracle as O

pytestmark = pytest.mark.gpu


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("records", ["1", "0"], ids=["bucket_records", "guide_window"])
@pytest.mark.parametrize("n", [700, 2048, 5000])
def test_batched_torus_runs_equal_single_runs(n, records, monkeypatch):
    from pyrecest_b200 import random as prandom

    monkeypatch.setenv("PFB_BUCKET_RECORDS", records)  # both lookup structures of the multinomial resampling
    from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN
    from pyrecest_b200.evaluation import BatchedParticleFilter

    rng = np.random.default_rng(n)
    R, D, T = 6, 2, 3
    C = np.array([[0.5, 0.1], [0.1, 0.3]])
    d0 = rng.random((R, n, D)) * 2 * np.pi
    Z = rng.standard_normal((T, R * n, D))
    U = rng.random((T, R * n))
    zs = rng.random((T, R, D)) * 2 * np.pi
    bf = BatchedParticleFilter("torus", R, n, D)
    bf.set_particles(d0)
    ref = d0.copy()
    for t in range(T):
        with prandom.inject(normals=Z[t], uniforms=U[t]):
            bf.predict_identity(HWN(np.zeros(D), 0.2 * C))
            bf.update_identity(HWN(np.zeros(D), 0.5 * C), zs[t])
        got = host(bf.d)
        est = host(bf.get_point_estimates())
        for r in range(R):
            noise = O.gaussian_rows(Z[t].reshape(R, n, D)[r], 0.2 * C)
            o = O.step("torus", ref[r], noise, "wn_multi", {"mu": zs[t, r], "C": 0.5 * C}, U[t].reshape(R, n)[r])
            diff = np.abs(got[r] - o["d"])
            diff = np.minimum(diff, 2 * np.pi - diff)
            bad = (diff.max(axis=1) > 1e-11).sum()
            assert bad <= 1, f"run {r} step {t}: {bad} particles differ"
            if bad == 0:
                np.testing.assert_allclose(est[r], o["est"], rtol=1e-10)
        ref = got
    assert not host(bf.failed).any()


def test_batched_euclid_and_failure_flag():
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.evaluation import BatchedParticleFilter
    from pyrecest_b200.filters import LinearTransition

    rng = np.random.default_rng(1)
    R, n, D = 4, 3000, 2
    F = np.array([[1.0, 1.0], [0.0, 1.0]])
    Q, Rm = 0.1 * np.eye(2), 0.5 * np.eye(2)
    d0 = rng.standard_normal((R, n, D))
    Z = rng.standard_normal((R * n, D))
    U = rng.random(R * n)
    zs = rng.standard_normal((R, D))
    zs[2] = np.nan  # this run's weights are all NaN -> failed, the others unaffected
    bf = BatchedParticleFilter("euclid", R, n, D)
    bf.set_particles(d0)
    with prandom.inject(normals=Z, uniforms=U):
        bf.predict_nonlinear(LinearTransition(F), GaussianDistribution(np.zeros(2), Q))
        bf.update_identity(GaussianDistribution(np.zeros(2), Rm), zs)
    got, est, failed = host(bf.d), host(bf.get_point_estimates()), host(bf.failed)
    assert failed.tolist() == [False, False, True, False]
    assert np.isnan(est[2]).all()
    for r in (0, 1, 3):
        o = O.step("euclid", d0[r], O.gaussian_rows(Z.reshape(R, n, D)[r], Q), "gauss", {"mu": zs[r], "C": Rm},
                   U.reshape(R, n)[r], F=F)
        bad = (np.abs(got[r] - o["d"]).max(axis=1) > 1e-11).sum()
        assert bad <= 1
        if bad == 0:
            np.testing.assert_allclose(est[r], o["est"], rtol=1e-10)


def test_iterate_configs_and_runs_like_reference_evaluation():
    """R2randomWalk-shaped scenario (pyrecest/evaluation/simulation_database.py:29-35), device RNG: estimates track
    the truth within the tolerance of the reference's own evaluation test."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.evaluation import iterate_configs_and_runs

    prandom.seed(3)
    rng = np.random.default_rng(3)
    n_runs, T = 40, 10
    scen = {"manifold": "Euclidean", "n_timesteps": T,
            "initial_prior": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)),
            "meas_noise": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)),
            "sys_noise": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2))}
    gt = np.cumsum(rng.standard_normal((n_runs, T, 2)) * np.sqrt(0.5), axis=1)
    meas = gt + rng.standard_normal((n_runs, T, 2)) * np.sqrt(0.5)
    states, times, failed, _, _ = iterate_configs_and_runs(gt, meas, scen, [{"name": "pf", "parameter": [500, 2000]}],
                                                           {"tolerate_failure": True})
    assert states.shape == (2,) and times.shape == (2, n_runs) and not failed.any()
    est = iterate_configs_and_runs.last_estimates
    err = np.linalg.norm(est - gt[None, :, -1, :], axis=2).mean(axis=1)
    assert err[1] < 1.0 and err[0] < 1.2, err
    assert states[1].d.shape == (n_runs, 2000, 2)
