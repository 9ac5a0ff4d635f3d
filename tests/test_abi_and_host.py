"""CPU-only checks: the C-ABI library loads and exports every symbol include/pfb.h declares, the
ctypes structs mirror the header, the host-side registry logic, and the product's refusal to run
without its CUDA library."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import pf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "pfb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from pyrecest_b200 import _lib as L

    names = header_functions()
    assert set(names) == set(L.SIGNATURES), (names, sorted(L.SIGNATURES))
    handle = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), n
    lib = L.lib()
    assert lib.pfb_abi_version() == L.ABI_VERSION == 8
    assert lib.pfb_workspace_bytes(0) >= 256
    assert lib.pfb_workspace_bytes(10**8) > lib.pfb_workspace_bytes(0)


def test_struct_sizes_match_header():
    from pyrecest_b200 import _lib as L

    assert ctypes.sizeof(L.PredictParams) == 8 * 4 + 8 * (36 + 6 + 36 + 6 + 2) + 8 + 16 + 48 + 48 + 8 + 24 + 8 + 8
    assert ctypes.sizeof(L.LikParams) == 4 * 4 + 8 * (6 + 36 + 3) + 8 + 8 + 8


def test_argument_validation_without_gpu():
    """host-side validation returns error codes before any CUDA call"""
    from pyrecest_b200 import _lib as L

    lib = L.lib()
    p = L.PredictParams()
    p.dim = 99
    assert lib.pfb_predict(ctypes.byref(p), 0, 10, None, None, None, None) == 1
    assert b"dim" in lib.pfb_last_error()
    lp = L.LikParams()
    lp.kind, lp.dim = 42, 3
    assert lib.pfb_loglik(ctypes.byref(lp), 0, 10, None, None, None, None, None, 0, None) == 4
    assert lib.pfb_scan_cdf(0, 10, None, None, None, None, None, 0, None) == 1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pyrecest_b200 import _lib as L

    monkeypatch.setattr(L, "_LIB", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.PfbError):
        L.lib()


def test_no_gpu_means_no_engine():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pyrecest_b200._engine import Engine

    with pytest.raises(RuntimeError):
        Engine("cuda:0")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyrecest_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_registry_rejects_callables_and_maps_kinds():
    from pyrecest_b200 import registry as R
    from pyrecest_b200.distributions import GaussianDistribution, VonMisesFisherDistribution

    g = GaussianDistribution(np.zeros(2), np.eye(2))
    assert R.kind_of(g) == "gauss"
    assert R.as_distribution(g.pdf) is g
    with pytest.raises(NotImplementedError):
        R.as_distribution(lambda x: x)
    with pytest.raises(NotImplementedError):
        R.transition_params(lambda x: x, 2)
    assert R.transition_params(R.IdentityTransition(), 2) == (None, None)
    with pytest.raises(ValueError):
        R.transition_params(R.LinearTransition(np.eye(3)), 2)
    v = VonMisesFisherDistribution(np.array([0.0, 0.0, 1.0]), 2.0)
    np.testing.assert_allclose(v.C, 2.0 / (4 * np.pi * np.sinh(2.0)), rtol=1e-14)
    for D, k in ((3, 0.5), (3, 700.0), (4, 10.0), (5, 3.0)):
        ref = O.vmf_norm_const(k, D) if k < 600 else None
        if ref is not None:
            np.testing.assert_allclose(np.exp(R.vmf_log_norm_const(k, D)), ref, rtol=1e-13)
        assert np.isfinite(R.vmf_log_norm_const(k, D))


def test_registry_likelihood_parameters_follow_scipy():
    from pyrecest_b200 import registry as R

    rng = np.random.default_rng(3)
    A = rng.standard_normal((3, 3))
    C = A @ A.T + 0.2 * np.eye(3)
    W, logdet = R.gaussian_whitening(C)
    Wo, lo = O.gaussian_whitening(C)
    assert np.array_equal(W, Wo) and logdet == lo
    np.testing.assert_allclose(W @ W.T, np.linalg.inv(C), rtol=1e-10)
    assert np.array_equal(R.gaussian_noise_factor(C), O.gaussian_noise_factor(C))
    np.testing.assert_allclose(R.log_i0(3.0), np.log(np.i0(3.0)), rtol=1e-13)


@pytest.mark.parametrize("scale,mu", [(0.5, [1.0, 2.0, 3.0]), (0.05, [6.2, 0.01, 3.1]), (1.0, [0.0, 0.0, 0.0]), (4.0, [3.0, 5.0, 1.0])])
def test_wn_image_pruning_reproduces_the_full_343_term_sum(scale, mu):
    """The pruned image table (host logic of the wn_multi kernel) evaluated in numpy equals the
    reference's full (2m+1)^d sum to rounding, for points all over [0, 2pi]^3 incl. the edges."""
    from pyrecest_b200 import registry as R

    C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
    C = scale * C3
    mu = np.array(mu)
    W, logdet = R.gaussian_whitening(C)
    table, K = R.wn_image_table(mu, W @ W.T)
    assert 1 <= K <= 343 and table.shape == (K, 7)
    rng = np.random.default_rng(0)
    x = rng.random((3000, 3)) * 2 * np.pi
    x[:4] = [[0, 0, 0], [2 * np.pi] * 3, [np.pi] * 3, [np.pi, 0, 2 * np.pi]]
    xs = (x + np.pi) % (2 * np.pi) - np.pi
    dev = xs - mu
    logc = -0.5 * (3 * np.log(2 * np.pi) + logdet)
    tot = np.zeros(x.shape[0])
    for row in table:
        y = (dev + row[4:7]) @ W
        tot += np.exp(logc - 0.5 * np.sum(y * y, axis=1))
    ref = O.pdf_wn_multi(x, mu, C)
    ok = ref > 1e-290
    np.testing.assert_allclose(tot[ok], ref[ok], rtol=2e-12)  # |maha| up to ~1400: rounding of the exponent, not pruning
    if scale <= 0.5:
        assert K < 100  # pruning is effective for peaked likelihoods


def test_bench_reference_arm_runs_on_cpu():
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-particles", "2000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"


@pytest.mark.parametrize("kappa", [1e-8, 0.3, 10.0, 777.0, 1e6])
def test_vm_strip_table_is_a_valid_equal_area_hat(kappa):
    """pfb_vm_strip_table (host code of libpfb, no GPU): contiguous strips covering [0, pi], every hat A / w_k above
    the density on its strip, area equal to the numpy restatement's, rejection probability ~1e-3"""
    from scipy.special import ive

    from oracle import philox as P
    from pyrecest_b200 import _lib as L

    table, A = L.vm_strip_table(kappa)
    a, w = table[:, 0], table[:, 1]
    assert table.shape == (L.VM_STRIPS, 2) and a[0] == 0.0 and np.all(w >= 0)
    np.testing.assert_allclose(a[1:], a[:-1] + w[:-1], rtol=0, atol=1e-15)
    assert abs(a[-1] + w[-1] - np.pi) < 1e-15
    g = np.exp(-2.0 * kappa * np.sin(0.5 * a) ** 2)  # density at the left (highest) end of each strip
    assert np.all(A >= g * w * (1 - 1e-12))
    t_ref, A_ref = P.vm_strip_table(kappa)
    np.testing.assert_allclose(A, A_ref, rtol=1e-12)
    np.testing.assert_allclose(table, t_ref, rtol=1e-9, atol=1e-13)
    efficiency = np.pi * ive(0, kappa) / (L.VM_STRIPS * A)
    assert 0.998 < efficiency <= 1.0


def test_vm_strip_sampler_restatement_is_von_mises():
    from scipy import stats

    from oracle import philox as P
    from pyrecest_b200 import _lib as L

    kap = [0.0, 0.7, 25.0]
    x = P.vonmises(3, 9, 100_000, kap, {k: L.vm_strip_table(k) for k in kap if k >= 1e-8})
    assert np.abs(x).max() < np.pi
    assert stats.kstest(x[:, 0], stats.uniform(-np.pi, 2 * np.pi).cdf).pvalue > 1e-4
    for c in (1, 2):
        assert stats.kstest(x[:, c], stats.vonmises(kap[c]).cdf).pvalue > 1e-4


@pytest.mark.parametrize("any_mean", [False, True])
@pytest.mark.parametrize("G", [8, 16])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_wn_cell_masks_cover_the_full_screen(d, G, any_mean):
    """registry.wn_cell_masks / wn_cell_table: at every point of the dev box, the images that can matter
    (r_k <= min r + cut) are inside the candidate mask of the point's cell, and so is the minimiser; the cell's
    reference image is one of its candidates.  Both grids (8 and 16 per 2 pi) and both boxes: one measurement, and the
    measurement-independent box dev in [-3 pi, pi)^d of the batched calls (2 G cells per dimension)."""
    from pyrecest_b200 import registry as R

    if any_mean and G == 16 and d == 3:
        pytest.skip("32^3 cells: not a configuration the host builds")
    rng = np.random.default_rng(d)
    A = rng.normal(size=(d, d))
    C = A @ A.T / d + 0.2 * np.eye(d)
    mu = rng.random(d) * 2 * np.pi
    W, _ = R.gaussian_whitening(C)
    if any_mean:  # the united table of evaluation._wn_params_any_mean
        rows = {}
        for corner in np.array(np.meshgrid(*[[0.0, 2 * np.pi]] * d)).T.reshape(-1, d):
            t, _ = R.wn_image_table(corner, W @ W.T)
            for row in t:
                rows[tuple(np.round(row[d + 1:] / (2 * np.pi)).astype(int))] = row
        table = np.array([rows[k] for k in sorted(rows, key=lambda kk: rows[kk][d])])
        K = table.shape[0]
    else:
        table, K = R.wn_image_table(mu, W @ W.T)
    masks = R.wn_cell_masks(mu, table, G=G, any_mean=any_mean)
    if masks is None:
        assert K > 64
        return
    n_cells = 2 * G if any_mean else G
    assert masks.shape == (n_cells**d,) and masks.dtype == np.uint64
    g, h = table[:, :d], table[:, d]
    y = rng.random((20_000, d)) * 2 * np.pi - np.pi  # wrapped particle coordinate
    mus = rng.random((20_000, d)) * 2 * np.pi if any_mean else mu
    dev = y - mus
    r = 2.0 * dev @ g.T + h
    need = r <= r.min(axis=1, keepdims=True) + R.WN_SCREEN_CUT
    lo = -3 * np.pi if any_mean else -(np.pi + mu)
    q = np.clip(((dev - lo) * (G / (2 * np.pi))).astype(int), 0, n_cells - 1)
    cell = (q * (n_cells ** np.arange(d))).sum(axis=1)
    have = ((masks[cell][:, None] >> np.arange(K, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    assert np.all(have[need])
    assert have.sum(axis=1).mean() < K  # and it prunes
    for narrow in ([False, True] if K <= 32 else [False]):
        tab = R.wn_cell_table(mu, table, masks, G=G, any_mean=any_mean, narrow=narrow, limit=500.0)
        assert tab is not None
        raw = tab.view(np.uint8)
        nc = n_cells**d
        refs = raw[nc * 4: nc * 5] if narrow else raw[nc * 8: nc * 9]
        m2 = raw[: nc * 4].view(np.uint32).astype(np.uint64) if narrow else raw[: nc * 8].view(np.uint64)
        assert np.array_equal(m2, masks)
        assert np.all((masks >> refs.astype(np.uint64)) & np.uint64(1))  # the reference image is a candidate of its cell
        # and it is the best image at the cell's centre among the candidates
        t_ref = 0.5 * r[np.arange(r.shape[0]), refs[cell]]
        assert np.all(t_ref - 0.5 * r.min(axis=1) <= 500.0)


def test_graft_entry_build_runs():
    """the driver's build check: compiles (make is a no-op when up to date), loads and imports"""
    import __graft_entry__ as g

    g.build()


def test_reference_arm_line_has_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: one JSON line with the
    same metric / config keys, impl = reference, a cpu_baseline describing the run and an e2e that repeats the value"""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-particles", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_derived_draw_sources_differ_per_shard_and_are_reproducible():
    """shards seeded alike must not draw identical noise; the per-rank child stream is a pure function of
    (seed, rank)."""
    from pyrecest_b200 import random as prandom

    a = prandom.DeviceDraws(seed=7)
    b = prandom.DeviceDraws(seed=7)
    assert a.derive(0).seed == b.derive(0).seed and a.derive(3).seed == b.derive(3).seed
    seeds = {a.derive(r).seed for r in range(8)} | {a.seed}
    assert len(seeds) == 9
    assert a.derive(2) is a.derive(2)  # one child per rank: its call counter keeps advancing


def test_watson_normalising_constant_against_the_reference():
    """cart_prod.WatsonDistribution.norm_const (host math: scipy 1F1) equals the reference's (mpmath,
    watson_distribution.py:44-55), values recorded from a run of the reference: S^2, S^3 and the circle, bipolar and girdle"""
    from pyrecest_b200.cart_prod import WatsonDistribution

    for mu, kappa, want in [(np.array([0.0, 0.0, 1.0]), 2.0, 0.03365575103329092),
                            (np.array([0.6, 0.0, 0.0, 0.8]), 7.5, 0.0008898317028250381),
                            (np.array([0.0, 1.0]), -1.5, 0.29409677016225244)]:
        w = WatsonDistribution(mu, kappa)
        assert w.dim == mu.shape[0] - 1 and w.input_dim == mu.shape[0]
        np.testing.assert_allclose(w.norm_const, want, rtol=1e-12)
    with pytest.raises(AssertionError):
        WatsonDistribution(np.array([0.0, 2.0]), 1.0)
