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


@pytest.mark.parametrize("tag,manifold", [("torus", "hypertorus"), ("circle", "circle"), ("sphere", "hypersphere"), ("euclid", "Euclidean")])
def test_device_error_summary_against_reference_run(tag, manifold):
    """evaluation.determine_all_deviations / summarize_filter_results on the device == the reference's (golden)"""
    from conftest import load_golden
    from pyrecest_b200 import evaluation as E

    g = load_golden("evaluation_summary")
    gt = torch.as_tensor(g[f"{tag}_gt"], device="cuda:0")
    est = torch.as_tensor(g[f"{tag}_est"], device="cuda:0")
    dev = E.determine_all_deviations(est, gt, manifold)
    np.testing.assert_allclose(host(dev), g[f"{tag}_dev"], rtol=1e-12, atol=1e-14)
    cfgs = [{"name": "pf", "parameter": 10}, {"name": "pf", "parameter": 20}]
    rows = E.summarize_filter_results({"manifold": manifold}, cfgs, g[f"{tag}_runtimes"], gt, g[f"{tag}_failed"], est)
    got = np.array([[r["error_mean"], r["error_std"], r["time_mean"], r["failure_rate"]] for r in rows])
    np.testing.assert_allclose(got, g[f"{tag}_summary"], rtol=1e-11)
    assert "error_mean" not in cfgs[0]  # the caller's configs are not modified


def test_device_scenario_generation_statistics():
    """generate_groundtruths / generate_measurements: random walk increments and measurement residuals have the
    moments of the configured noises; the reference's t-1 indexing is reproduced; a full study runs end to end"""
    from pyrecest_b200 import evaluation as E
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution, HypertoroidalWrappedNormalDistribution as HWN
    from pyrecest_b200.distributions import VonMisesFisherDistribution

    prandom.seed(77)
    R_, T = 20_000, 4
    C = np.array([[0.05, 0.01], [0.01, 0.03]])
    sc = {"n_timesteps": T, "manifold": "hypertorus", "initial_prior": HWN(np.array([1.0, 2.0]), 4 * C),
          "sys_noise": HWN(np.zeros(2), C), "meas_noise": HWN(np.zeros(2), 0.5 * C)}
    gt = E.generate_groundtruths(sc, R_)
    assert gt.shape == (R_, T, 2)
    inc = host(gt[:, 1:] - gt[:, :-1]).reshape(-1, 2)
    inc = np.mod(inc + np.pi, 2 * np.pi) - np.pi
    np.testing.assert_allclose(np.cov(inc.T), C, atol=2e-3)
    z = E.generate_measurements(gt, sc)
    assert float(z.min()) >= 0 and float(z.max()) < 2 * np.pi
    res = host(z[:, 2] - gt[:, 1])  # reference indexing: the measurement of step t belongs to groundtruth[t - 1]
    res = np.mod(res + np.pi, 2 * np.pi) - np.pi
    np.testing.assert_allclose(np.cov(res.T), 0.5 * C, atol=2e-3)
    z2 = E.generate_measurements(gt, sc, reference_indexing=False)
    res2 = np.mod(host(z2[:, 2] - gt[:, 2]) + np.pi, 2 * np.pi) - np.pi
    np.testing.assert_allclose(np.cov(res2.T), 0.5 * C, atol=2e-3)
    # Euclidean + sphere measurement models
    sce = {"n_timesteps": 3, "manifold": "Euclidean", "initial_prior": GaussianDistribution(np.zeros(2), np.eye(2)),
           "sys_noise": GaussianDistribution(np.zeros(2), C), "meas_noise": GaussianDistribution(np.zeros(2), C)}
    ge = E.generate_groundtruths(sce, 5000)
    ze = E.generate_measurements(ge, sce, reference_indexing=False)
    np.testing.assert_allclose(np.cov(host(ze - ge).reshape(-1, 2).T), C, atol=3e-3)
    modes = torch.zeros((5000, 1, 3), dtype=torch.float64, device="cuda:0")
    modes[..., 2] = 1.0
    zs = E.generate_measurements(modes, {"meas_noise": VonMisesFisherDistribution(np.array([0.0, 0.0, 1.0]), 20.0)})
    np.testing.assert_allclose(host(torch.linalg.norm(zs, dim=-1)), 1.0, atol=1e-12)
    assert abs(float(zs[:, 0, 2].mean()) - (1 / np.tanh(20.0) - 1 / 20.0)) < 5e-3
    # the whole study on the device: scenario -> batched filters -> summary
    small = dict(sc, n_timesteps=3)
    gts = E.generate_groundtruths(small, 64)
    meas = E.generate_measurements(gts, small, reference_indexing=False)
    cfgs = [{"name": "pf", "parameter": [500, 2000]}]
    states, times, failed, _, _ = E.iterate_configs_and_runs(gts, meas, small, cfgs, {"tolerate_failure": True})
    rows = E.summarize_filter_results(small, [{"name": "pf", "parameter": 500}, {"name": "pf", "parameter": 2000}], times,
                                      gts, failed, E.iterate_configs_and_runs.last_estimates)
    assert all(0 < r["error_mean"] < 0.5 for r in rows)


def test_result_file_schema_and_reload(tmp_path):
    """evaluate_for_variables writes the reference's .npy result file (evaluate_for_variables.py:51-70: the same eight
    keys) and evaluate_for_file reads it back and re-runs the filters on its ground truths and measurements
    (evaluate_for_file.py:31-71); two measurements per time step are two updates
    (perform_predict_update_cycles.py:57-65)."""
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200 import evaluation as E

    rng = np.random.default_rng(8)
    n_runs, T = 12, 6

    def scenario():
        return {"name": "r2walk", "manifold": "Euclidean", "n_timesteps": T,
                "initial_prior": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)),
                "meas_noise": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)),
                "sys_noise": GaussianDistribution(np.zeros(2), 0.5 * np.eye(2))}

    gt = np.cumsum(rng.standard_normal((n_runs, T, 2)) * np.sqrt(0.5), axis=1)
    meas2 = gt[:, :, None, :] + rng.standard_normal((n_runs, T, 2, 2)) * np.sqrt(0.5)   # two measurements per step
    prandom.seed(5)
    out = E.evaluate_for_variables(gt, meas2, [{"name": "pf", "parameter": [400, 1500]}], scenario(), save_folder=str(tmp_path),
                                   tolerate_failure=True)
    assert len(out) == 8
    est2 = E.iterate_configs_and_runs.last_estimates.copy()
    data = np.load(E.evaluate_for_variables.last_file, allow_pickle=True).item()
    assert tuple(data.keys()) == E.RESULT_KEYS
    assert data["groundtruths"].shape == (n_runs, T, 2) and data["measurements"].shape == (n_runs, T, 2, 2)
    assert data["runtimes"].shape == (2, n_runs) and data["run_failed"].shape == (2, n_runs) and not data["run_failed"].any()
    assert data["last_filter_states"][1].shape == (n_runs, 1500, 2)
    assert data["filter_configs"] == [{"name": "pf", "parameter": 400}, {"name": "pf", "parameter": 1500}]
    assert data["scenario_config"]["name"] == "r2walk" and "initial_prior" not in data["scenario_config"]
    # reload: the same study from the file alone (scenario_config supplies only the models), same seed -> same estimates
    prandom.seed(5)
    sc = scenario()
    del sc["n_timesteps"], sc["name"]
    again = E.evaluate_for_file(E.evaluate_for_variables.last_file, [{"name": "pf", "parameter": [400, 1500]}], sc,
                                save_folder=str(tmp_path), tolerate_failure=True)
    assert sc["n_timesteps"] == T and list(sc["n_meas_at_individual_time_step"]) == [2] * T
    assert list(sc["apply_sys_noise_times"]) == [True] * (T - 1) + [False]
    np.testing.assert_array_equal(E.iterate_configs_and_runs.last_estimates, est2)
    np.testing.assert_array_equal(again[0][1], out[0][1])
    # two measurements per step pull the estimate closer to the truth than one
    prandom.seed(5)
    E.iterate_configs_and_runs(gt, meas2[:, :, 0, :], scenario(), [{"name": "pf", "parameter": [1500]}], {"tolerate_failure": True})
    err1 = np.linalg.norm(E.iterate_configs_and_runs.last_estimates[0] - gt[:, -1, :], axis=1).mean()
    err2 = np.linalg.norm(est2[1] - gt[:, -1, :], axis=1).mean()
    assert err2 < err1 * 1.05, (err1, err2)
