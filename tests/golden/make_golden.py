"""Generates tests/golden/*.npz by running the REAL reference (pyRecEst v0.9.1 from
/root/reference, numpy backend) through oracle/ref_harness.py with injected draws.

Run in the build container only:   python tests/golden/make_golden.py
The fixtures hold the inputs (particles, injected normals / uniforms / VMF triples,
parameters) and the reference's outputs; tests compare oracle/pf_oracle.py and the CUDA path
with them.  Nothing here is imported at test time.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_harness import LAST, DrawQueue, import_reference, injected_draws  # noqa: E402
sys.path.insert(0, HERE)
from config0_inputs import config0_draws  # noqa: E402

import_reference()
from pyrecest.distributions import (  # noqa: E402
    CircularDiracDistribution,
    GaussianDistribution,
    HypertoroidalDiracDistribution,
    HypertoroidalWrappedNormalDistribution,
    LinearDiracDistribution,
    VonMisesDistribution,
    VonMisesFisherDistribution,
    WrappedNormalDistribution,
)
from pyrecest.distributions.hypersphere_subset.hyperspherical_dirac_distribution import (  # noqa: E402
    HypersphericalDiracDistribution,
)
from pyrecest.filters import EuclideanParticleFilter, HypertoroidalParticleFilter  # noqa: E402
from pyrecest.filters.circular_particle_filter import CircularParticleFilter  # noqa: E402

C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: np.shape(v) for k, v in arrs.items()})


def euclid_cv(N=2000, T=3, seed=1):
    """Config 1 shape: R^2 constant velocity, Gaussian noise + Gaussian likelihood."""
    rng = np.random.default_rng(seed)
    F = np.array([[1.0, 1.0], [0.0, 1.0]])
    Q = np.array([[0.1, 0.02], [0.02, 0.1]])
    R = 0.5 * np.eye(2)
    Z0 = rng.standard_normal((N, 2))
    Z = rng.standard_normal((T, N, 2))
    U = rng.random((T, N))
    zs = rng.standard_normal((T, 2))
    pf = EuclideanParticleFilter(N, 2)
    q = DrawQueue(normals=np.concatenate([Z0.ravel(), Z.ravel()]), uniforms=U)
    out = {"d_pred": [], "idx": [], "d_post": [], "est": []}
    with injected_draws(q):
        pf.filter_state = GaussianDistribution(np.array([0.5, -0.25]), np.eye(2))
        d0 = pf.filter_state.d.copy()
        for t in range(T):
            pf.predict_nonlinear(lambda x: x @ F.T, GaussianDistribution(np.zeros(2), Q))
            out["d_pred"].append(pf.filter_state.d.copy())
            pf.update_identity(GaussianDistribution(np.zeros(2), R), zs[t])
            out["idx"].append(LAST["choice_idx"].copy())
            out["d_post"].append(pf.filter_state.d.copy())
            out["est"].append(pf.get_point_estimate().copy())
    save("euclid_cv", F=F, Q=Q, R=R, Z0=Z0, Z=Z, U=U, zs=zs, d0=d0,
         prior_mu=np.array([0.5, -0.25]), prior_C=np.eye(2),
         **{k: np.stack(v) for k, v in out.items()})


def torus_t3(N=1000, T=2, seed=2):
    """Config 2 shape at test size: T^3, wrapped-normal noise (shift-mode) + WN likelihood."""
    rng = np.random.default_rng(seed)
    Z0 = rng.standard_normal((N, 3))
    Z = rng.standard_normal((T, N, 3))
    U = rng.random((T, N))
    zs = np.array([[1.0, 2.0, 3.0], [1.2, 2.1, 2.9]])
    mu0 = np.array([1.0, 1.0, 1.0]) + np.pi / 2
    Cn = 0.1 * C3
    Cl = 0.5 * C3
    pf = HypertoroidalParticleFilter(N, 3)
    q = DrawQueue(normals=np.concatenate([Z0.ravel(), Z.ravel()]), uniforms=U)
    out = {"d_pred": [], "idx": [], "d_post": [], "est": [], "lik": []}
    with injected_draws(q):
        pf.filter_state = HypertoroidalWrappedNormalDistribution(mu0, C3)
        d0 = pf.filter_state.d.copy()
        for t in range(T):
            pf.predict_identity(HypertoroidalWrappedNormalDistribution(np.zeros(3), Cn))
            out["d_pred"].append(pf.filter_state.d.copy())
            lik = HypertoroidalWrappedNormalDistribution(np.zeros(3), Cl).set_mode(zs[t])
            out["lik"].append(lik.pdf(pf.filter_state.d))
            pf.update_identity(HypertoroidalWrappedNormalDistribution(np.zeros(3), Cl), zs[t])
            out["idx"].append(LAST["choice_idx"].copy())
            out["d_post"].append(pf.filter_state.d.copy())
            out["est"].append(pf.get_point_estimate().copy())
    save("torus_t3", C=C3, Cn=Cn, Cl=Cl, mu0=mu0, Z0=Z0, Z=Z, U=U, zs=zs, d0=d0,
         **{k: np.stack(v) for k, v in out.items()})


def torus_addmode(N=400, seed=3):
    """T^2 add-mode predict (shift_instead_of_add=False) with a non-zero noise mean, plus edge
    particles exactly at 0, 2pi-ulp and tiny negatives produced by f."""
    rng = np.random.default_rng(seed)
    C = np.array([[0.3, 0.1], [0.1, 0.2]])
    mu_n = np.array([0.25, 6.0])
    Z = rng.standard_normal((N, 2))
    d0 = rng.random((N, 2)) * 2 * np.pi
    d0[0] = [0.0, np.nextafter(2 * np.pi, 0)]
    d0[1] = [1e-300, 5e-324]
    pf = HypertoroidalParticleFilter(N, 2)
    pf.filter_state = HypertoroidalDiracDistribution(d0)
    q = DrawQueue(normals=Z)
    with injected_draws(q):
        pf.predict_nonlinear(lambda x: x, HypertoroidalWrappedNormalDistribution(mu_n, C),
                             True, False)
    save("torus_addmode", C=C, mu_n=mu_n, Z=Z, d0=d0, d_pred=pf.filter_state.d.copy())


def mod_edges():
    """numpy.mod(x, 2pi) on edge values, as applied by the torus filter/dirac."""
    tp = 2.0 * np.pi
    x = np.array([0.0, -0.0, tp, -tp, 2 * tp, -1e-20, 1e-20, -1e-300, np.nextafter(tp, 0),
                  np.nextafter(tp, 10), -np.nextafter(tp, 0), 7.0, -7.0, 1e6, -1e6, 1e15, -1e15,
                  3 * tp, 100 * tp, -5e-324, 5e-324, 12.566370614359172, 6.283185307179587,
                  -3.0, 3.0, 1e300, -1e300])
    save("mod_edges", x=x, y=np.mod(x, tp))


def circle_case(N=500, T=2, seed=4):
    """CircularParticleFilter: 1-D wrapped-normal noise (shift-mode), von Mises and 1-D WN likelihoods."""
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((T, N))
    U = rng.random((T, N))
    pf = CircularParticleFilter(N)
    d0 = pf.filter_state.d.copy()
    out = {"d_pred": [], "idx": [], "d_post": [], "est": [], "lik": []}
    sig_n = 0.4
    q = DrawQueue(normals=Z, uniforms=U)
    liks = [VonMisesDistribution(np.array(2.0), 1.5), WrappedNormalDistribution(np.array(4.0), np.array(0.7))]
    with injected_draws(q):
        for t in range(T):
            # HypertoroidalWN with 1x1 covariance goes through multivariate_normal
            pf.predict_identity(HypertoroidalWrappedNormalDistribution(np.zeros(1), np.array([[sig_n**2]])))
            out["d_pred"].append(pf.filter_state.d.copy())
            out["lik"].append(liks[t].pdf(pf.filter_state.d))
            pf.update_nonlinear_using_likelihood(liks[t])
            out["idx"].append(LAST["choice_idx"].copy())
            out["d_post"].append(pf.filter_state.d.copy())
            out["est"].append(np.atleast_1d(pf.get_point_estimate()).copy())
    save("circle_case", Z=Z, U=U, d0=d0, sig_n=sig_n, vm_mu=2.0, vm_kappa=1.5, wn_mu=4.0, wn_sigma=0.7,
         **{k: np.stack(v) for k, v in out.items()})


def likelihood_tables(seed=5):
    rng = np.random.default_rng(seed)
    arrs = {}
    for d in (1, 2, 3, 4):
        A = rng.standard_normal((d, d))
        C = A @ A.T + 0.3 * np.eye(d)
        mu = rng.standard_normal(d)
        x = rng.standard_normal((300, d)) * 2
        arrs[f"g{d}_C"], arrs[f"g{d}_mu"], arrs[f"g{d}_x"] = C, mu, x
        arrs[f"g{d}_pdf"] = GaussianDistribution(mu, C).pdf(x if d > 1 else x[:, 0])
    # hypertoroidal WN, d = 2, 3 (+ broad covariance)
    for tag, d, scale in (("wn2", 2, 1.0), ("wn3", 3, 0.5), ("wn3b", 3, 4.0)):
        C = scale * (C3[:d, :d])
        mu = np.array([1.0, 2.0, 3.0])[:d]
        x = rng.random((200, d)) * 2 * np.pi
        arrs[f"{tag}_C"], arrs[f"{tag}_mu"], arrs[f"{tag}_x"] = C, mu, x
        arrs[f"{tag}_pdf"] = HypertoroidalWrappedNormalDistribution(mu, C).pdf(x)
    # 1-D WN series for several sigma (incl. > 10 -> uniform)
    x1 = np.concatenate([rng.random(200) * 2 * np.pi, [0.0, 2 * np.pi, np.pi]])
    arrs["wn1_x"] = x1
    sig = np.array([0.05, 0.3, 1.0, 2.5, 6.0, 9.99, 11.0])
    arrs["wn1_sigma"] = sig
    arrs["wn1_mu"] = np.array(1.3)
    arrs["wn1_pdf"] = np.stack([WrappedNormalDistribution(np.array(1.3), np.array(s)).pdf(x1) for s in sig])
    # von Mises
    kap = np.array([0.0, 0.1, 1.0, 10.0, 200.0])
    arrs["vm_kappa"] = kap
    arrs["vm_mu"] = np.array(0.7)
    arrs["vm_pdf"] = np.stack([VonMisesDistribution(np.array(0.7), k).pdf(x1) for k in kap])
    # VMF on S^2 and S^3
    for D in (3, 4):
        mu = rng.standard_normal(D)
        mu /= np.linalg.norm(mu)
        x = rng.standard_normal((200, D))
        x /= np.linalg.norm(x, axis=1)[:, None]
        arrs[f"vmf{D}_mu"], arrs[f"vmf{D}_x"] = mu, x
        kk = np.array([0.5, 10.0, 50.0])
        arrs[f"vmf{D}_kappa"] = kk
        arrs[f"vmf{D}_pdf"] = np.stack([VonMisesFisherDistribution(mu, k).pdf(x) for k in kk])
    save("likelihoods", **arrs)


def vmf_sampler(seed=6):
    """VonMisesFisherDistribution.sample on S^2 from injected (u, z1, z2), one mode per row
    (the set_mode + sample(1) loop of hyperhemisphere_cart_prod_particle_filter.py:90-95)."""
    rng = np.random.default_rng(seed)
    N = 300
    modes = rng.standard_normal((N, 3))
    modes /= np.linalg.norm(modes, axis=1)[:, None]
    modes[0] = [1.0, 0.0, 0.0]
    modes[1] = [-1.0, 0.0, 0.0]
    modes[2] = [0.0, 0.0, 1.0]
    modes[3] = [0.0, -1.0, 0.0]
    modes[4] = [1e-9, 1.0, 0.0] / np.linalg.norm([1e-9, 1.0, 0.0])
    modes[5] = [-1e-9, 0.0, 1.0] / np.linalg.norm([-1e-9, 0.0, 1.0])
    trip = np.column_stack([rng.random(N), rng.standard_normal((N, 2))])
    outs = {}
    for kappa in (0.5, 50.0):
        res = np.empty((N, 3))
        q = DrawQueue(vmf_triples=trip)
        with injected_draws(q):
            for j in range(N):
                res[j] = VonMisesFisherDistribution(modes[j], kappa).sample(1)
        outs[f"out_k{kappa}"] = res
    save("vmf_sampler", modes=modes, triples=trip, kappas=np.array([0.5, 50.0]), **outs)


def dirac_moments(seed=7):
    rng = np.random.default_rng(seed)
    N = 700
    arrs = {}
    w = rng.random(N)
    w /= w.sum()
    arrs["w"] = w
    # linear
    d = rng.standard_normal((N, 4)) * np.array([1, 2, 3, 0.5]) + np.array([0.1, -2, 5, 0])
    ld = LinearDiracDistribution(d, w)
    arrs["lin_d"], arrs["lin_mean"], arrs["lin_cov"] = d, ld.mean(), ld.covariance()
    # torus
    t = rng.random((N, 3)) * 2 * np.pi
    td = HypertoroidalDiracDistribution(t, w)
    arrs["tor_d"] = t
    arrs["tor_m1"], arrs["tor_m2"] = td.trigonometric_moment(1), td.trigonometric_moment(2)
    arrs["tor_meandir"] = td.mean_direction()
    c = rng.random(N) * 2 * np.pi
    cd = CircularDiracDistribution(c, w)
    arrs["circ_d"], arrs["circ_m1"], arrs["circ_meandir"] = c, np.atleast_1d(cd.trigonometric_moment(1)), np.atleast_1d(cd.mean_direction())
    # sphere
    s = rng.standard_normal((N, 3)) + np.array([0.5, 0.2, 1.0])
    s /= np.linalg.norm(s, axis=1)[:, None]
    sd = HypersphericalDiracDistribution(s, w)
    arrs["sph_d"], arrs["sph_res"], arrs["sph_meandir"], arrs["sph_scatter"] = s, sd.mean_resultant_vector(), sd.mean_direction(), sd.moment()
    # sample with n != N, and reweigh
    u = rng.random(1234)
    with injected_draws(DrawQueue(uniforms=u)):
        smp = td.sample(1234)
        arrs["tor_sample_idx"] = LAST["choice_idx"].copy()
    arrs["tor_sample_u"], arrs["tor_sample"] = u, smp
    lik = np.exp(-0.5 * np.sum((d - 1.0) ** 2, axis=1))
    arrs["lin_lik"] = lik
    arrs["lin_reweighed"] = ld.reweigh(lambda x: np.exp(-0.5 * np.sum((x - 1.0) ** 2, axis=1))).w
    save("dirac_moments", **arrs)


def resample_cases(seed=8):
    """choice(d, n, p=w) index vectors for weight vectors that stress the sequential cumsum:
    likelihood-shaped, wide dynamic range, tie-heavy (exact half-ulp addends), zeros."""
    rng = np.random.default_rng(seed)
    arrs = {}
    cases = {}
    N = 200_000
    x = rng.standard_normal(N) * 3
    w = np.exp(-0.5 * (x - 1.0) ** 2)
    cases["lik"] = w / w.sum()
    w = 2.0 ** rng.integers(-60, 1, size=50_000).astype(np.float64) * (1 + rng.random(50_000))
    cases["wide"] = w / w.sum()
    w = np.full(30_000, 2.0 ** -53)
    w[0] = 1.0
    w[1::7] = 2.0 ** -52
    cases["ties"] = w  # deliberately unnormalised: choice() normalises the CDF only
    w = rng.random(10_000)
    w[rng.random(10_000) < 0.5] = 0.0
    cases["zeros"] = w / w.sum()
    cases["uniform"] = np.ones(65_537) / 65_537
    cases["single"] = np.array([1.0])
    for name, w in cases.items():
        n = max(1000, min(w.shape[0], 50_000))
        u = rng.random(n)
        u[:3] = [0.0, np.nextafter(1.0, 0), 0.5]
        cdf = np.cumsum(w)
        cdf /= cdf[-1]
        # u values sitting exactly on CDF entries (side='right' semantics)
        take = rng.integers(0, w.shape[0], size=min(200, n - 3))
        u[3:3 + take.shape[0]] = np.minimum(cdf[take], np.nextafter(1.0, 0))
        with injected_draws(DrawQueue(uniforms=u)):
            import pyrecest.backend as B

            # kahan-sum check inside numpy's choice needs sum(p)==1 within sqrt(eps): normalise for the call
            B.random.choice(np.arange(w.shape[0]), n, p=w)
        arrs[f"{name}_w"], arrs[f"{name}_u"], arrs[f"{name}_idx"] = w, u, LAST["choice_idx"].copy()
        arrs[f"{name}_cumsum_tail"] = np.cumsum(w)[-4:]
    save("resample_cases", **arrs)


def hypercylindrical(seed=9):
    """HypercylindricalDiracDistribution (periodic dims first, then linear) -- hybrid_mean,
    hybrid_moment, linear_mean / linear_covariance, the marginals, and a reweigh + sample round trip"""
    from pyrecest.distributions.cart_prod.hypercylindrical_dirac_distribution import HypercylindricalDiracDistribution

    rng = np.random.default_rng(seed)
    N = 900
    arrs = {}
    for tag, bd, ld in (("c21", 2, 1), ("c13", 1, 3), ("c22", 2, 2)):
        d = np.column_stack([rng.random((N, bd)) * 2 * np.pi, rng.standard_normal((N, ld)) * 2.0 + np.arange(ld)])
        w = rng.random(N)
        w /= w.sum()
        hd = HypercylindricalDiracDistribution(bd, d, w)
        arrs[f"{tag}_d"], arrs[f"{tag}_w"] = d, w
        arrs[f"{tag}_hybrid_mean"] = np.asarray(hd.hybrid_mean())
        arrs[f"{tag}_hybrid_moment"] = np.asarray(hd.hybrid_moment())
        arrs[f"{tag}_linear_mean"] = np.asarray(hd.linear_mean())
        arrs[f"{tag}_linear_cov"] = np.atleast_2d(np.asarray(hd.linear_covariance()))
        arrs[f"{tag}_per_m1"] = np.atleast_1d(hd.marginalize_linear().trigonometric_moment(1))
        lik = np.exp(-0.5 * np.sum((d[:, bd:] - 1.0) ** 2, axis=1)) * (1.2 + np.cos(d[:, 0] - 1.0))
        arrs[f"{tag}_lik"] = lik
        rw = hd.reweigh(lambda x, bd=bd: np.exp(-0.5 * np.sum((x[:, bd:] - 1.0) ** 2, axis=1)) * (1.2 + np.cos(x[:, 0] - 1.0)))
        arrs[f"{tag}_reweighed_w"] = rw.w
        arrs[f"{tag}_reweighed_mean"] = np.asarray(rw.hybrid_mean())
        u = rng.random(500)
        with injected_draws(DrawQueue(uniforms=u)):
            arrs[f"{tag}_sample"] = rw.sample(500)
            arrs[f"{tag}_sample_idx"] = LAST["choice_idx"].copy()
        arrs[f"{tag}_sample_u"] = u
    save("hypercylindrical", **arrs)


def evaluation_summary(seed=10):
    """the error summary either side of the batched Monte-Carlo runner -- the reference's
    get_distance_function / determine_all_deviations / summarize_filter_results on given final estimates"""
    from pyrecest.evaluation.determine_all_deviations import determine_all_deviations
    from pyrecest.evaluation.get_distance_function import get_distance_function
    from pyrecest.evaluation.summarize_filter_results import summarize_filter_results

    rng = np.random.default_rng(seed)
    R, T, C = 1200, 4, 2
    arrs = {}
    for tag, manifold, d in (("torus", "hypertorus", 3), ("circle", "circle", 1), ("sphere", "hypersphere", 3), ("euclid", "Euclidean", 4)):
        if tag == "sphere":
            gt = rng.standard_normal((R, T, d))
            gt /= np.linalg.norm(gt, axis=-1, keepdims=True)
            est = gt[None, :, -1, :] + 0.2 * rng.standard_normal((C, R, d))
            est /= np.linalg.norm(est, axis=-1, keepdims=True)
        elif tag == "euclid":
            gt = rng.standard_normal((R, T, d)) * 3
            est = gt[None, :, -1, :] + rng.standard_normal((C, R, d))
        else:
            gt = rng.random((R, T, d)) * 4 * np.pi - 2 * np.pi          # groundtruth walks are not wrapped
            est = np.mod(gt[None, :, -1, :] + 0.3 * rng.standard_normal((C, R, d)), 2 * np.pi)
        gts = np.empty((R, T), dtype=object)
        for r in range(R):
            for t in range(T):
                gts[r, t] = gt[r, t]
        results = [[est[c, r] for r in range(R)] for c in range(C)]
        dist = get_distance_function(manifold)
        dev = determine_all_deviations(results, None, dist, gts)
        runtimes = rng.random((C, R)) * 1e-3
        failed = rng.random((C, R)) < 0.01
        summ = summarize_filter_results({"manifold": manifold, "mtt": False}, [{"name": "pf", "parameter": 10}, {"name": "pf", "parameter": 20}],
                                        runtimes, gts, failed, last_filter_states=results)  # (last_estimates alone raises in the reference)
        arrs[f"{tag}_gt"], arrs[f"{tag}_est"], arrs[f"{tag}_dev"] = gt, est, dev
        arrs[f"{tag}_runtimes"], arrs[f"{tag}_failed"] = runtimes, failed
        arrs[f"{tag}_summary"] = np.array([[s_["error_mean"], s_["error_std"], s_["time_mean"], s_["failure_rate"]] for s_ in summ])
    save("evaluation_summary", **arrs)


def config0(N=10_000, T=100):
    """BASELINE.json configs[0] at full size through the real reference (about a minute): EuclideanParticleFilter
    (10^4, 2), constant velocity, Gaussian noise and likelihood, 100 steps.  Outputs kept: the estimate and an index
    checksum of every step, the resampled indices of steps 0 / 49 / 99 and the final particle set."""
    import time

    c = config0_draws(N, T)
    pf = EuclideanParticleFilter(N, 2)
    q = DrawQueue(normals=np.concatenate([c["Z0"].ravel(), c["Z"].ravel()]), uniforms=c["U"])
    est, idx_sum, idx_keep, d_keep = [], [], {}, {}
    t0 = time.perf_counter()
    with injected_draws(q):
        pf.filter_state = GaussianDistribution(np.zeros(2), np.eye(2))
        for t in range(T):
            pf.predict_nonlinear(lambda x: x @ c["F"].T, GaussianDistribution(np.zeros(2), c["Q"]))
            pf.update_identity(GaussianDistribution(np.zeros(2), c["R"]), c["zs"][t])
            idx = LAST["choice_idx"]
            idx_sum.append([int(idx.sum()), int((idx * np.arange(1, N + 1)).sum() % (2 ** 61 - 1))])
            if t in (0, T // 2 - 1, T - 1):
                idx_keep[f"idx_{t}"] = idx.astype(np.int32)
                d_keep[f"d_{t}"] = pf.filter_state.d.copy()
            est.append(pf.get_point_estimate().copy())
    secs = time.perf_counter() - t0
    print(f"reference: {N} particles x {T} steps in {secs:.1f} s = {N * T / secs:.3g} particle-steps/s")
    save("config0", N=N, T=T, zs=c["zs"], est=np.stack(est), idx_sum=np.array(idx_sum, dtype=np.int64),
         reference_seconds=secs, **idx_keep, **d_keep)


def se3_and_wn_filter(seed=11):
    """SURVEY 8(f) rank 4, the rest: SE3DiracDistribution (unit quaternion part first, then the translation) and
    WrappedNormalFilter.update_nonlinear_particle.  What the reference can do is recorded: the marginals,
    linear_mean / linear_covariance (taken on d[:, bound_dim:] = the LAST FOUR columns -- the reference slices at
    bound_dim = 3 although the quaternion has four components; a drop-in has to return the same), reweigh and sample
    of the 7-component records.  hybrid_mean() / mean() raise ValueError in the reference (the hemispherical mean
    direction is broken there): parity unpinned."""
    from pyrecest.distributions.se3_dirac_distribution import SE3DiracDistribution
    from pyrecest.filters.wrapped_normal_filter import WrappedNormalFilter

    rng = np.random.default_rng(seed)
    N = 700
    q = rng.standard_normal((N, 4)) * np.array([0.3, 0.3, 0.3, 1.0])
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1.0
    d = np.hstack([q, rng.standard_normal((N, 3)) * 2.0 + np.array([1.0, -2.0, 0.5])])
    w = rng.random(N)
    w /= w.sum()
    s3 = SE3DiracDistribution(d, w)
    arrs = {"se3_d": d, "se3_w": w, "se3_linear_mean": np.asarray(s3.linear_mean()),
            "se3_linear_cov": np.asarray(s3.linear_covariance()),
            "se3_quat_moment": np.asarray(s3.marginalize_linear().moment()),
            "se3_periodic_marginal_mean": np.asarray(s3.marginalize_periodic().mean())}
    lik = np.exp(-0.5 * np.sum((d[:, 4:] - 1.0) ** 2, axis=1)) * (1.5 + d[:, 3])
    arrs["se3_lik"] = lik
    rw = s3.reweigh(lambda x: np.exp(-0.5 * np.sum((x[:, 4:] - 1.0) ** 2, axis=1)) * (1.5 + x[:, 3]))
    arrs["se3_reweighed_w"] = rw.w
    u = rng.random(400)
    with injected_draws(DrawQueue(uniforms=u)):
        arrs["se3_sample"] = rw.sample(400)
        arrs["se3_sample_idx"] = LAST["choice_idx"].copy()
    arrs["se3_sample_u"] = u
    # WrappedNormalFilter.update_nonlinear_particle (wrapped_normal_filter.py:27-32): 100 samples of the state,
    # reweigh by likelihood(z, .), moment-match a wrapped normal.  The samples are injected (the method draws them
    # from numpy's global stream).
    for tag, (mu0, sig0, z, kap) in {"wnf_a": (1.0, 0.8, 2.0, 2.0), "wnf_b": (5.9, 0.4, 0.3, 6.0)}.items():
        f = WrappedNormalFilter(WrappedNormalDistribution(np.array(mu0), np.array(sig0)))
        smp = np.mod(mu0 + sig0 * rng.standard_normal(100), 2 * np.pi)
        f.filter_state.sample = lambda n, smp=smp: smp[:n]
        f.update_nonlinear_particle(lambda zz, x, kap=kap: VonMisesDistribution(np.array(zz), kap).pdf(x), z)
        arrs[f"{tag}_in"] = np.array([mu0, sig0, z, kap])
        arrs[f"{tag}_samples"] = smp
        arrs[f"{tag}_out"] = np.array([float(f.filter_state.mu), float(f.filter_state.sigma)])
    save("se3_and_wn_filter", **arrs)


if __name__ == "__main__":
    if len(sys.argv) > 1:  # regenerate only the named fixtures
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    hypercylindrical()
    euclid_cv()
    torus_t3()
    torus_addmode()
    mod_edges()
    circle_case()
    likelihood_tables()
    vmf_sampler()
    dirac_moments()
    resample_cases()
