"""Out-of-bounds WRITE check for every output-producing entry point of libpfb (compute-sanitizer is closed on this
pool): every output array sits inside a larger allocation between two guard bands filled with a sentinel; after the
call the bands must be untouched.  Sizes are deliberately awkward (not multiples of 32 / 256 / 2048, n_out != n).
Reads are covered indirectly: results are compared with the oracle elsewhere on the same kinds of sizes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
BAND = 4096  # elements per guard band
SENT = {torch.float64: -7.25e77, torch.float32: -7.25e30, torch.int64: -0x5A5A5A5A5A5A5A5A}


class Guarded:
    """tensors carved out of guarded allocations; check() verifies every band"""

    def __init__(self):
        self.items = []

    def make(self, shape, dtype=torch.float64):
        n = int(np.prod(shape))
        raw = torch.full((n + 2 * BAND,), SENT[dtype], dtype=dtype, device=DEV)
        view = raw[BAND:BAND + n].view(shape)
        self.items.append((raw, n, dtype))
        return view

    def check(self, what):
        torch.cuda.synchronize()
        for raw, n, dtype in self.items:
            lo, hi = raw[:BAND], raw[BAND + n:]
            assert bool((lo == SENT[dtype]).all()) and bool((hi == SENT[dtype]).all()), f"{what}: write outside an output array"


@pytest.fixture(scope="module")
def eng():
    from pyrecest_b200._engine import engine

    return engine(DEV)


@pytest.mark.parametrize("n", [1, 33, 255, 2049, 70_001])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_guard_bands_of_the_single_filter_entry_points(eng, n, dtype):
    from pyrecest_b200 import _lib as L
    from pyrecest_b200 import registry as R
    from pyrecest_b200.distributions import GaussianDistribution, HypertoroidalWrappedNormalDistribution as HWN

    g = Guarded()
    rng = np.random.default_rng(n)
    C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
    x = torch.as_tensor(rng.random((n, 3)) * 6.28, device=DEV).to(dtype)
    noise = torch.as_tensor(rng.vonmises(0.0, 10.0, size=(n, 3)), device=DEV).to(dtype)
    # predict (array-fed and in-kernel draws), with the padded gather copy
    p = L.PredictParams()
    p.mode, p.dim, p.noise_cols = L.PRED_TORUS_SHIFT, 3, 3
    out, pad = g.make((n, 3), dtype), g.make((n, 4), dtype)
    p.x_pad_out, p.x_pad_stride = pad.data_ptr(), 4
    eng.predict(p, x, out, noise)
    q = L.PredictParams()
    q.mode, q.dim, q.noise_cols = L.PRED_TORUS_SHIFT, 3, 3
    q.rng_kind, q.rng_seed, q.rng_offset = L.RNG_VONMISES, 5, 1
    q.set_von_mises_rng([10.0, 3.0, 0.0], device=DEV)
    out2 = g.make((n, 3), dtype)
    eng.predict(q, x, out2, None)
    g.check("predict")
    # weighting: log and slice-scaled formats, fused and stand-alone
    lp, keep = R.lik_params(HWN(np.array([1.0, 2.0, 3.0]), 0.5 * C3), 3, images_to_device=R.device_put(torch.device(DEV)))
    for scaled in (False, True):
        logw, xo = g.make((n,)), g.make((n, 3), dtype)
        eng.loglik(lp, out, logw, scaled=scaled)
        eng.predict_loglik(p, lp, x, xo, noise, logw, scaled=scaled)
        cdf = g.make((n,))
        eng.scan(cdf, logw=logw, exact=True)
    g.check("weighting + scan")
    w = g.make((n,))
    l2 = torch.as_tensor(-rng.random(n) * 30.0, device=DEV)
    eng.normalize(l2, w)
    cdf_w, cdf_f = g.make((n,)), g.make((n,))
    eng.scan(cdf_w, w=w.contiguous(), exact=True)
    eng.scan(cdf_f, w=w.contiguous(), exact=False)
    g.check("normalize + scan")
    # resampling: every scheme, n_out != n, indices and estimates
    for n_out in (n, max(1, n // 3), 2 * n + 5):
        u = torch.as_tensor(rng.random(n_out), device=DEV)
        u0 = torch.as_tensor(rng.random(1), device=DEV)
        est = g.make((2 * L.PFB_MAXD,))
        for kind in ("multinomial", "systematic", "rng", "rng_systematic", "msort", "padded"):
            y, idx = g.make((n_out, 3), dtype), g.make((n_out,), torch.int64)
            if kind == "multinomial":
                eng.resample(cdf_w, u, out, y, n_out, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est)
            elif kind == "systematic":
                eng.resample(cdf_w, u0, out, y, n_out, 3, systematic=True, idx_out=idx, est_kind=L.EST_SUM, est_out=est)
            elif kind == "rng":
                eng.resample_rng(cdf_w, 3, 9, out, y, n_out, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est)
            elif kind == "rng_systematic":
                eng.resample_rng(cdf_w, 3, 9, out, y, n_out, 3, systematic=True, idx_out=idx)
            elif kind == "msort":
                eng.resample_msort(cdf_w, 3, 9, out, y, n_out, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est)
            else:
                eng.resample(cdf_w, u, pad, y, n_out, 3, idx_out=idx, est_kind=L.EST_TRIG, est_out=est, x_in_stride=4)
            g.check(f"resample {kind} n_out={n_out}")
            assert int(idx.min()) >= 0 and int(idx.max()) <= n - 1
    # moments, wrap, uniforms
    for kind in (L.MOM_MEAN, L.MOM_TRIG, L.MOM_SCATTER):
        m = g.make((3 * 3 + 2 * 3 + 2,))
        eng.moments(kind, out, 3, w=w.contiguous(), out=m)
    wr = g.make((n, 3), dtype)
    eng.wrap_2pi(x * 3.0 - 5.0, wr)
    uu, um = g.make((n,)), g.make((n,))
    eng.rng_uniform(uu, 1, 2)
    eng.msort_uniforms(um, 1, 2)
    g.check("moments / wrap / uniforms")
    del keep
    # Euclidean fused kernel with a linear transition (general, non-PLAIN path), R^4
    F = np.eye(4)
    F[0, 2] = F[1, 3] = 1.0
    pe = L.PredictParams()
    pe.mode, pe.dim, pe.noise_cols, pe.has_F, pe.has_A = L.PRED_EUCLID, 4, 4, 1, 1
    pe.F[:16] = F.reshape(-1).tolist()
    pe.A[:16] = (0.3 * np.eye(4)).reshape(-1).tolist()
    le, keep = R.lik_params(GaussianDistribution(np.full(4, 0.3), 0.5 * np.eye(4)), 4)
    x4 = torch.as_tensor(rng.standard_normal((n, 4)), device=DEV).to(dtype)
    z4 = torch.as_tensor(rng.standard_normal((n, 4)), device=DEV).to(dtype)
    o4, l4 = g.make((n, 4), dtype), g.make((n,))
    eng.predict_loglik(pe, le, x4, o4, z4, l4, scaled=True)
    g.check("euclid fused")


@pytest.mark.parametrize("n,runs", [(700, 5), (2049, 3)])
def test_guard_bands_of_the_batched_and_sharded_entry_points(eng, n, runs):
    from pyrecest_b200 import _lib as L
    from pyrecest_b200 import registry as R
    from pyrecest_b200.distributions import GaussianDistribution

    g = Guarded()
    rng = np.random.default_rng(n)
    pad = eng.run_pad(n)
    x = torch.as_tensor(rng.standard_normal((runs * n, 2)), device=DEV)
    lp, keep = R.lik_params(GaussianDistribution(np.zeros(2), 0.5 * np.eye(2)), 2)
    mu = torch.as_tensor(rng.standard_normal((runs, 2)), device=DEV).contiguous()
    lp.mu_batch_dev = mu.data_ptr()
    for scaled in (False, True):
        logw, cdf = g.make((runs * pad,)), g.make((runs * pad,))
        eng.loglik_batched(lp, x, runs, n, logw, scaled=scaled)
        eng.scan_batched(logw, cdf, runs, n, exact=True, scaled=scaled)
        for systematic in (False, True):
            y = g.make((runs * n, 2))
            eng.resample_batched(cdf, x, y, runs, n, 2, seed=4, offset=7, systematic=systematic)
        g.check(f"batched scaled={scaled}")
    m = g.make((runs, 2 * 2 + 2 * 2 + 2))
    eng.moments_batched(L.MOM_MEAN, x, runs, n, 2, out=m)
    g.check("moments_batched")
    # sharded pull / push with three shards on this device
    world = 3
    xs = [torch.as_tensor(rng.standard_normal((n, 4)), device=DEV) for _ in range(world)]
    cdfs, totals = [], torch.zeros(world, dtype=torch.float64, device=DEV)
    for r in range(world):
        c = torch.empty(n, dtype=torch.float64, device=DEV)
        eng.scan(c, w=torch.as_tensor(rng.random(n) * (10.0 ** -r), device=DEV), exact=True)
        totals[r] = c[-1]
        cdfs.append(c)
    status, est = g.make((1,)), g.make((2 * L.PFB_MAXD,))
    outs = [g.make((n, 4)) for _ in range(world)]
    for r in range(world):
        eng.resample_sharded_pull([c.data_ptr() for c in cdfs], [a.data_ptr() for a in xs], totals, world, r, 0.37, outs[r], n, 4,
                                  L.EST_SUM, est, status)
    g.check("sharded pull")
    outs = [g.make((n, 4)) for _ in range(world)]
    for r in range(world):
        eng.resample_sharded_dev(cdfs[r], totals, world, r, 0.37, [o.data_ptr() for o in outs], xs[r], 4, L.EST_SUM, est, status)
    g.check("sharded push")
    del keep
