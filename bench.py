#!/usr/bin/env python
"""bench.py -- particle-steps/s of the fused predict + update + resample + estimate step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload t3|s2|r2|r4|r4s|t2mc] [--particles P]
                  [--resampling multinomial_sorted|multinomial|systematic] [--storage f64|f32]
  python bench.py --impl reference ...      # the CPU restatement of the reference (oracle/) on host cores

Workload at N=1 (BASELINE.json configs[1]): HypertoroidalParticleFilter on T^3, 1e8 particles, von Mises
system noise (kappa = 10 per dim), wrapped-normal likelihood (cov 0.5 C, z = [1, 2, 3], m = 3, i.e. the
reference's 343-image sum), exact multinomial resampling of N of N (offspring in CDF order; --resampling multinomial is
the reference's order), mean-direction estimate.  fp64.
N > 1 (torchrun): every rank runs an independent filter of the same size (independent Monte-Carlo runs shard
with no communication -- north_star), "scaling": "weak"; the line's `sharded` block then reports configs[4]: ONE R^4
filter sharded over all ranks (1.25e8 particles per GPU, global systematic resampling over NVLink).

value : steps timed with CUDA events, noise draws + uniforms already resident in HBM (a ring of
        pre-generated buffers larger than L2), 3 kernel launches per step through the C ABI.
e2e   : the same step through the public API (pf.predict_identity / pf.update_identity /
        pf.get_point_estimate), draws generated on the device inside the timed region, the measurement
        coming from pinned host memory and the estimate + status word read back to the host every step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
METRIC = "particle-steps/s (predict+update+resample)"

WORKLOADS = {
    # name: (manifold, D, noise columns E, default particles, description)
    "t3": ("torus", 3, 3, 10**8, "HypertoroidalParticleFilter T^3, von Mises noise + wrapped-normal likelihood"),
    "s2": ("sphere", 3, 3, 10**8, "HypersphericalParticleFilter S^2, VMF noise + VMF likelihood, mean direction"),
    "r2": ("euclid", 2, 2, 10**4, "EuclideanParticleFilter R^2 constant velocity, Gaussian noise + likelihood"),
    "r4": ("euclid", 4, 4, 10**8, "EuclideanParticleFilter R^4 constant velocity, Gaussian noise + likelihood"),
    # BASELINE.json configs[4]: ONE filter sharded over the GPUs, global systematic resampling (weak: 1.25e8 per GPU)
    # BASELINE.json configs[3]: independent Monte-Carlo runs, batched (1250 runs x 1e5 particles per GPU = 10k runs on 8)
    "t2mc": ("torus", 2, 2, 100_000, "BatchedParticleFilter: 1250 independent T^2 filter runs x 1e5 particles per GPU, WN noise + WN likelihood"),
    "r4s": ("euclid", 4, 4, 125_000_000, "ShardedParticleFilter R^4 constant velocity, particles sharded over the GPUs, global systematic resampling"),
}


def algorithmic_bytes(D, E, sb=8):
    """B = 8 (4D + E + 5) per particle-step, split per kernel (sb: bytes per stored particle / noise
    component: 4 with --storage f32; weights, CDF and uniforms stay fp64)."""
    return {"predict_loglik": sb * (2 * D + E) + 8, "scan": 16, "resample": sb * 2 * D + 16}


def random_access_roofline(kernel, resampling, n, D, ms, padded=False):
    """Multinomial resampling is bound by DRAM random accesses, not bytes: B200 serves ~3.7e10 random reads/s from
    footprints >= 1 GiB whatever their size up to 128 B (tools/micro/sector_fetch.cu, profiles/r01g_sector_fetch_micro.log).
    Per particle: one 64-byte bucket record + the particle record (1 + (8 D - 8)/64 64-byte fetches; exactly 1
    when the gather reads the padded copy the predict kernel left behind)."""
    if kernel != "resample" or resampling != "multinomial":
        return None
    R_ = 8 * D  # bytes per record; it straddles a 64-byte fetch when (i R mod 64) + R > 64
    straddle = 0.0 if padded else sum(((i * R_) % 64) + R_ > 64 for i in range(64)) / 64.0
    per_particle = 2.0 + straddle  # (the padded gather copy never straddles)
    peak = 3.7e10
    achieved = per_particle * n / (ms * 1e-3)
    return {"reads_per_particle": per_particle, "achieved_reads_per_s": achieved, "peak_reads_per_s": peak,
            "frac": achieved / peak, "peak_source": "measured: profiles/r01g_sector_fetch_micro.log (footprint >= 1 GiB)",
            "note": "time includes the 0.45 ms record build, which streams"}


def ncu_traffic(name, n, resampling, exact, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this exact configuration, else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = f"{name}:{n}:f64:{resampling}:{'exact' if exact else 'fast'}"
        return t[key][kernel], t[key]["source"]
    except Exception:
        return None, None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# workload definitions shared by both arms
# ---------------------------------------------------------------------------------------------
def cv_matrix(D):
    F = np.eye(D)
    for i in range(D // 2):
        F[i, D // 2 + i] = 1.0
    return F


def workload_params(name):
    manifold, D, E, _, _ = WORKLOADS[name]
    if name == "t3":
        return {"z": np.array([1.0, 2.0, 3.0]), "lik_C": 0.5 * C3, "kappa_noise": np.array([10.0, 10.0, 10.0]),
                "prior_mu": np.array([1.0, 1.0, 1.0]) + math.pi / 2, "prior_C": C3}
    if name == "s2":
        return {"z": np.array([1.0, 1.0, 1.0]) / math.sqrt(3), "kappa_lik": 10.0, "kappa_noise": 50.0,
                "prior_mu": np.array([0.0, 0.0, 1.0]), "prior_kappa": 5.0}
    return {"z": np.full(D, 0.3), "F": cv_matrix(D), "Q": 0.1 * np.eye(D), "R": 0.5 * np.eye(D),
            "prior_mu": np.zeros(D), "prior_C": np.eye(D)}


def make_dists(name, P):
    """(prior, transition, noise, likelihood-with-mode) as public-API objects."""
    from pyrecest_b200 import distributions as Dn
    from pyrecest_b200.filters import IdentityTransition, LinearTransition

    if name == "t3":
        return (Dn.HypertoroidalWrappedNormalDistribution(P["prior_mu"], P["prior_C"]), IdentityTransition(),
                Dn.VonMisesDistribution(np.zeros(3), P["kappa_noise"]),
                Dn.HypertoroidalWrappedNormalDistribution(np.zeros(3), P["lik_C"]))
    if name == "s2":
        return (Dn.VonMisesFisherDistribution(P["prior_mu"], P["prior_kappa"]), IdentityTransition(),
                Dn.VonMisesFisherDistribution(np.array([0.0, 0.0, 1.0]), P["kappa_noise"]),
                Dn.VonMisesFisherDistribution(P["z"], P["kappa_lik"]))
    D = P["F"].shape[0]
    return (Dn.GaussianDistribution(P["prior_mu"], P["prior_C"]), LinearTransition(P["F"]),
            Dn.GaussianDistribution(np.zeros(D), P["Q"]), Dn.GaussianDistribution(np.zeros(D), P["R"]))


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores
# ---------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_oracle_step(O, pool, nthreads, manifold, d, noise, lik_kind, lik_params, u, **kw):
    """oracle.step with the per-particle parts (predict, likelihood) split over host threads -- numpy releases the
    GIL inside its kernels; reweigh / cumsum / searchsorted / estimate are inherently serial and stay so."""
    n = d.shape[0]
    if pool is None or n < 4 * nthreads:
        return O.step(manifold, d, noise, lik_kind, lik_params, u, **kw)
    bounds = np.linspace(0, n, nthreads + 1).astype(int)

    def part(i):
        a, b = bounds[i], bounds[i + 1]
        if manifold == "euclid":
            dp = O.predict_euclid(d[a:b], noise[a:b], kw.get("F"), kw.get("b"))
        elif manifold == "torus":
            dp = O.predict_torus(d[a:b], noise[a:b], kw.get("F"), kw.get("b"), shift=True)
        else:
            dp = O.predict_sphere_vmf(d[a:b], noise[a:b], kw["kappa_noise"])
        return dp, O.likelihood(lik_kind, dp, lik_params)

    parts = list(pool.map(part, range(nthreads)))
    dp = np.concatenate([p[0] for p in parts])
    lik = np.concatenate([p[1] for p in parts])
    wp = O.reweigh(O.uniform_w(n), lik)
    idx = O.multinomial_indices(wp, u)
    dn = dp[idx]
    wn = O.uniform_w(n)
    est = O.mean_linear(dn, wn) if manifold == "euclid" else (O.mean_direction_torus(dn, wn) if manifold == "torus" else O.mean_direction_sphere(dn, wn))
    return {"d": dn, "est": est}


def cpu_step_fn(name, n, seed=0):
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pf_oracle as O

    nthreads = cpu_threads()
    pool = ThreadPoolExecutor(nthreads) if nthreads > 1 else None

    P = workload_params(name)
    rng = np.random.default_rng(seed)
    manifold, D, E, _, _ = WORKLOADS[name]
    if name == "t3":
        d = np.mod(O.gaussian_rows(rng.standard_normal((n, 3)), P["prior_C"], P["prior_mu"]), 2 * np.pi)
        lik = ("wn_multi", {"mu": P["z"], "C": P["lik_C"]})
        draw = lambda: rng.vonmises(0.0, 10.0, size=(n, 3))  # noqa: E731
        kw = {}
    elif name == "s2":
        t0 = np.column_stack([rng.random(n), rng.standard_normal((n, 2))])
        d = O.predict_sphere_vmf(np.tile(P["prior_mu"], (n, 1)), t0, P["prior_kappa"])
        lik = ("vmf", {"mu": P["z"], "kappa": P["kappa_lik"]})
        draw = lambda: np.column_stack([rng.random(n), rng.standard_normal((n, 2))])  # noqa: E731
        kw = {"kappa_noise": P["kappa_noise"]}
    else:
        d = O.gaussian_rows(rng.standard_normal((n, D)), P["prior_C"], P["prior_mu"])
        lik = ("gauss", {"mu": P["z"], "C": P["R"]})
        draw = lambda: O.gaussian_rows(rng.standard_normal((n, D)), P["Q"])  # noqa: E731
        kw = {"F": P["F"]}
    state = {"d": d}

    def step():
        noise = draw()
        u = rng.random(n)
        t0 = time.perf_counter()
        r = cpu_oracle_step(O, pool, nthreads, manifold, state["d"], noise, lik[0], lik[1], u, **kw)
        dt = time.perf_counter() - t0
        state["d"] = r["d"]
        return dt

    return step


def run_cpu(name, n, steps, warmup):
    step = cpu_step_fn(name, n)
    try:  # the particle ranges are split over our own threads: numpy's BLAS / OpenMP pools stay single-threaded
        from threadpoolctl import threadpool_limits  # (nested pools over-subscribe the cores: 32-core boxes ran slower)

        limit = threadpool_limits(limits=1)
    except Exception:
        limit = None
    try:
        for _ in range(warmup):
            step()
        total = sum(step() for _ in range(steps))
    finally:
        if limit is not None:
            limit.restore_original_limits()
    return n * steps / total, total / steps


# ---------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (ts, r) in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for (_, r) in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pyrecest_b200 import _lib as L
    from pyrecest_b200 import random as prandom
    from pyrecest_b200 import registry as R
    from pyrecest_b200._engine import engine
    from pyrecest_b200 import filters as Fl

    name = args.workload
    manifold, D, E, n_default, desc = WORKLOADS[name]
    n = args.particles or n_default
    P = workload_params(name)
    prior, trans, noise_dist, lik_dist = make_dists(name, P)
    eng = engine(dev)
    prandom.seed(1234 + rank)

    # ---- public-API filter (used for e2e and to build the initial particle set)
    cls = {"torus": Fl.HypertoroidalParticleFilter, "sphere": Fl.HypersphericalParticleFilter,
           "euclid": Fl.EuclideanParticleFilter}[manifold]
    sdt = torch.float32 if args.storage == "f32" else torch.float64
    pf = cls(n, D if manifold != "sphere" else D - 1, dtype=sdt)
    pf.resampling = args.resampling  # 'multinomial' (the reference's scheme, default) or 'systematic'
    pf.filter_state = prior

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value leg: resident inputs, C-ABI calls ------------------------------------------------
    nbuf = max(2, min(4, int(math.ceil(512e6 / (n * 8.0 * (E + 1))))))  # ring > L2 (126 MB)
    src = prandom.source()
    noise_ring, u_ring = [], []
    for _ in range(nbuf):
        if name == "t3":
            noise_ring.append(src.vonmises((n, 3), torch.as_tensor(P["kappa_noise"]), dev, sdt))
        elif name == "s2":
            noise_ring.append(src.vmf_triples(n, dev, sdt))
        else:
            noise_ring.append(src.normals((n, D), dev, sdt))
        u_ring.append(src.uniforms((n,), dev))
    with prandom.null():  # parameter block for array-fed draws (rng_kind = 0): the value leg reads the ring
        pp, _ = pf._predict_params(trans, noise_dist, pf._default_shift)
    lp, keep = R.lik_params(lik_dist.set_mode(P["z"]), D, images_to_device=R.device_put(dev))
    x = pf.filter_state._records().clone()
    alt = torch.empty_like(x)
    logw = torch.empty(n, dtype=torch.float64, device=dev)
    cdf = torch.empty(n, dtype=torch.float64, device=dev)
    est = torch.empty(2 * L.PFB_MAXD, dtype=torch.float64, device=dev)
    est_kind = L.EST_TRIG if manifold == "torus" else L.EST_SUM
    exact = not args.fast_scan
    # the filters' padded gather copy (pfb_predict_params.x_pad_out), as the public API uses it
    gather = pf._padded_gather_target(pp, x)
    xg_stride = gather[1] if gather is not None else 0

    def step(i, ev=None):
        nonlocal x, alt
        nb = noise_ring[i % nbuf]
        ub = u_ring[i % nbuf]
        if ev:
            ev[0].record()
        # (with the padded gather copy the compact predicted particles are dead: not stored, as in filters.py)
        eng.predict_loglik(pp, lp, x, None if gather is not None else x, nb, logw, scaled=not args.log_weights)
        if ev:
            ev[1].record()
        eng.scan(cdf, logw=logw, exact=exact)
        if ev:
            ev[2].record()
        if args.resampling == "multinomial_sorted":  # in-kernel draws (no uniform array exists for this mode)
            eng.resample_msort(cdf, 1234 + rank, i + 1, x, alt, n, D, est_kind=est_kind, est_out=est)
        else:
            eng.resample(cdf, ub, gather[0] if gather is not None else x, alt, n, D, systematic=args.resampling == "systematic",
                         est_kind=est_kind, est_out=est, x_in_stride=xg_stride)
        if ev:
            ev[3].record()
        x, alt = alt, x

    for i in range(args.warmup):
        step(i)
    barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    launches0 = eng.launches
    t_wall0 = time.time()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        step(args.warmup + i, evs[i])
    end.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_total = start.elapsed_time(end)
    gpu_launches = eng.launches - launches0
    kt = np.array([[e[k].elapsed_time(e[k + 1]) for k in range(3)] for e in evs]).mean(axis=0)  # ms per kernel
    flags = int(eng.stats[L.STAT_FLAGS].item())
    est_host = est.cpu().numpy()
    assert flags == 0 and np.all(np.isfinite(est_host[:D])), "bench step produced invalid weights"
    assert eng.scan_error_flags() == 0

    # ---- e2e leg: public API ------------------------------------------------------------------------
    z_pinned = torch.as_tensor(P["z"]).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def api_step():
        z = z_pinned.numpy()  # measurement arrives in pinned host memory; parameters ride in the launch
        pf.predict_nonlinear(trans, noise_dist)
        pf.update_identity(lik_dist, z)
        return pf.get_point_estimate().cpu()  # D2H of the estimate (the status word is read inside update)

    if args.e2e_steps > 0:
        for _ in range(min(args.warmup, 3)):
            api_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            api_step()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    else:
        e2e_ms = float("nan")

    # ---- reduce over ranks ---------------------------------------------------------------------------
    vals = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(vals[0]), float(vals[1])
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    e2e_value = world * n / (e2e_ms * 1e-3)

    sharded = None
    if world > 1 and not args.no_sharded:
        del noise_ring, u_ring, x, alt, logw, cdf, pf
        torch.cuda.empty_cache()
        sharded = sharded_block(args, world, rank, dev, args.steps, args.warmup, n=args.sharded_particles or None)
    if rank == 0:
        peak, peak_src = peaks()
        ab = algorithmic_bytes(D, E, 4 if args.storage == "f32" else 8)
        names = ["predict_loglik", "scan", "resample"]
        k = int(np.argmax(kt))
        achieved = ab[names[k]] * n / (kt[k] * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(name, n, args.resampling, exact, names[k])
        step_gbs = sum(ab.values()) * n / (ms_step * 1e-3) / 1e9
        cpu_n = min(args.cpu_particles, n)  # (a filter smaller than the CPU sample, configs[0]: the CPU arm at the same N)
        cpu_steps = 2 if cpu_n >= 100_000 else 50
        cpu_val, cpu_s = run_cpu(name, cpu_n, cpu_steps, 1) if not args.no_cpu else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {desc}", "particles_per_gpu": n, "resampling": args.resampling,
                       "storage": args.storage + (" (particles and noise draws stored in fp32, arithmetic fp64)" if args.storage == "f32" else ""),
                       "scan": "sequential-exact" if exact else "fast", "likelihood_images": int(lp.n_images) if name == "t3" else None,
                       "l2_policy": f"inputs larger than L2: ring of {nbuf} pre-generated noise/uniform buffers "
                                    f"({n * 8 * (E + 1) / 1e6:.0f} MB each) + {n * 8 * D / 1e6:.0f} MB state",
                       "parallelism": "1 filter per GPU, no collective" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "h2d_bytes_per_step": int(8 * len(P["z"])), "d2h_bytes_per_step": int(8 * D + 8 * L.STAT_COUNT),
                    "path": "public API: predict_nonlinear + update_identity + get_point_estimate; draws generated on device"},
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": names[k], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel_ms": {nm: float(t) for nm, t in zip(names, kt)},
                         "step_algorithmic_gbs": step_gbs, "step_frac": step_gbs / peak,
                         "bytes_per_particle": ab,
                         "random_access": random_access_roofline(names[k], args.resampling, n, D, kt[k], padded=gather is not None)},
            "cpu_baseline": None if args.no_cpu else {
                "value": cpu_val, "unit": "particle-steps/s", "cores": cpu_threads(), "kind": "port",
                "sample": f"oracle/pf_oracle.py (vectorised numpy restatement, predict + likelihood split over {cpu_threads()} host threads) "
                          f"on {cpu_n} particles x {cpu_steps} steps of the same workload, {cpu_s:.4f} s/step"},
        }
        if sharded is not None:
            line["sharded"] = sharded
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sharded_block(args, world, rank, dev, steps, warmup, n=None):
    """BASELINE.json configs[4] inside an N-GPU run: ONE EuclideanParticleFilter (R^4 constant velocity) sharded over
    all ranks, 1.25e8 particles per GPU, global systematic resampling whose offspring are stored straight into the
    owner rank's buffer over NVLink (pyrecest_b200.sharded, device-side plan).  Two scenarios -- `even` (all shards
    statistically alike: almost every offspring stays on its rank) and `skewed` (shard g's particles sit 1.5 g sigma
    away from the measurement, so most offspring are produced by shard 0 and cross to the other ranks; the particle
    set is re-installed before every timed step, untimed, because one resampling step evens the shards out) -- and
    the 1-GPU anchor: the same filter code with a one-rank group on every GPU at once.  Returns the dict rank 0
    prints as the line's `sharded` block."""
    import torch
    import torch.distributed as dist

    from pyrecest_b200 import random as prandom
    from pyrecest_b200.sharded import ShardedParticleFilter

    name = "r4s"
    manifold, D, E, n_default, desc = WORKLOADS[name]
    n = n or n_default
    P = workload_params(name)
    prior, trans, noise_dist, lik_dist = make_dists(name, P)
    lik = lik_dist.set_mode(P["z"])
    prandom.seed(4321)  # the same seed on every rank: the shards derive their own streams from it
    u_rng = np.random.default_rng(11)
    u0s = u_rng.random(4096)
    solo_groups = [dist.new_group([r]) for r in range(world)]  # (every rank must create every group)

    def timed(pf, k_steps, k_warm, reset=None):
        """per-step CUDA events, every step entered through a barrier; returns mean ms per step (max over ranks)"""
        it = [0]

        def step():
            pf.predict_nonlinear(trans, noise_dist)
            pf.update_nonlinear_using_likelihood(lik, u0=float(u0s[it[0] % 4096]))
            it[0] += 1
            return pf.get_point_estimate()

        for _ in range(k_warm):
            if reset is not None:
                reset()
            step()
        evs = []
        for _ in range(k_steps):
            if reset is not None:
                reset()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            est = step()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        pf.check()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs) / k_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert bool(torch.isfinite(est).all())
        return float(ms)

    out = {"workload": f"{name}: {desc}", "particles_per_gpu": n, "particles_total": n * world, "resampling": "systematic (global)",
           "scan": "sequential-exact per shard against the global shift", "plan": "device-side (no host read of the shard totals)",
           "exchange": "pull: every rank produces its own slots and reads remote sources (CDF shard + particles, peer-mapped "
                       "symmetric memory) over NVLink inside the resampling kernel; all stores local"}
    # ---- the sharded filter, even scenario
    pf = ShardedParticleFilter(manifold, n * world, D)
    pf.filter_state = prior
    pf.time_scatter = True
    even_ms = timed(pf, steps, warmup)
    made, remote = pf.offspring_traffic()
    sc_ms = pf.scatter_ms()
    tr = torch.tensor([made, remote], dtype=torch.float64, device=dev)
    dist.all_reduce(tr)
    out["even"] = {"ms_per_step": even_ms, "value": world * n / (even_ms * 1e-3),
                   "offspring_crossing_nvlink_frac": float(tr[1] / tr[0]),
                   "remote_offspring_bytes_per_step": int(tr[1]) * 8 * D, "resample_kernel_ms": sc_ms}
    # ---- skewed scenario: shard g sits 1.5 g sigma from the measurement; re-installed before every step (untimed)
    x_even = pf.filter_state.d.clone()
    x_skew = x_even.clone()
    x_skew[:, :2] += 1.5 * rank
    del x_even
    skew_steps = max(3, min(steps, 8))
    skew_ms = timed(pf, skew_steps, 2, reset=lambda: pf.set_local_particles(x_skew))
    made, remote = pf.offspring_traffic()
    sc_ms = pf.scatter_ms()
    tr = torch.tensor([made, remote, remote], dtype=torch.float64, device=dev)
    mx = tr[2:].clone()
    dist.all_reduce(tr[:2])
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    scat = torch.tensor([sc_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(scat, op=dist.ReduceOp.MAX)
    out["skewed"] = {"ms_per_step": skew_ms, "value": world * n / (skew_ms * 1e-3),
                     "offspring_crossing_nvlink_frac": float(tr[1] / tr[0]),
                     "remote_offspring_bytes_per_step": int(tr[1]) * 8 * D,
                     "busiest_rank_remote_offspring_bytes": int(mx[0]) * 8 * D, "resample_kernel_ms": float(scat),
                     # (an upper bound of the bytes read over NVLink: a remote source with k offspring is read once)
                     "busiest_rank_nvlink_read_gbs_upper": int(mx[0]) * 8 * D / (float(scat) * 1e-3) / 1e9 if float(scat) > 0 else None,
                     "nvlink_peak_gbs_per_direction": 770.0,
                     "note": "particle set re-installed (untimed) before every timed step; " + str(skew_steps) + " steps"}
    # ---- time in the step's three small collectives, back to back on the stream (not overlapped with anything)
    a = torch.zeros(1, dtype=torch.float64, device=dev)
    g = torch.zeros(world, dtype=torch.float64, device=dev)
    e = torch.zeros(D, dtype=torch.float64, device=dev)
    for k in range(13):
        if k == 3:
            torch.cuda.synchronize()
            dist.barrier()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
        dist.all_reduce(a, op=dist.ReduceOp.MAX)
        dist.all_gather_into_tensor(g, a)
        dist.all_reduce(e)
    t1.record()
    torch.cuda.synchronize()
    coll = torch.tensor([t0.elapsed_time(t1) / 10], dtype=torch.float64, device=dev)
    dist.all_reduce(coll, op=dist.ReduceOp.MAX)
    out["collectives_ms_per_step"] = float(coll)
    out["collectives"] = "all_reduce(MAX, 1 double) + all_gather(1 double per rank) + all_reduce(SUM, D doubles); particles move by peer loads inside the resampling kernel, no all_to_all"
    del pf, x_skew
    torch.cuda.empty_cache()
    # ---- anchor: the same filter with a one-rank group, on every GPU at once
    solo = ShardedParticleFilter(manifold, n, D, group=solo_groups[rank])
    solo.filter_state = prior
    anchor_ms = timed(solo, max(3, min(steps, 10)), warmup)
    del solo
    torch.cuda.empty_cache()
    out["anchor_1gpu_ms_per_step"] = anchor_ms
    out["even"]["efficiency"] = anchor_ms / even_ms
    out["skewed"]["efficiency"] = anchor_ms / skew_ms
    return out


def run_mc(args):
    """configs[3]: independent Monte-Carlo runs packed into one launch per step (pyrecest_b200.evaluation),
    runs sharded over the GPUs with no communication."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from pyrecest_b200 import random as prandom
    from pyrecest_b200._engine import engine
    from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN
    from pyrecest_b200.evaluation import BatchedParticleFilter

    n = args.particles or 100_000
    runs = args.runs
    D = 2
    prandom.seed(99 + rank)
    eng = engine(dev)
    C = 0.5 * np.eye(2)
    bf = BatchedParticleFilter("torus", runs, n, D)
    bf.set_prior(HWN(np.array([1.0, 2.0]), C))
    rng = np.random.default_rng(rank)
    T = args.steps + args.warmup + args.e2e_steps + 1
    meas = np.mod(np.array([1.0, 2.0]) + np.cumsum(rng.standard_normal((T, runs, D)) * 0.3, axis=0), 2 * np.pi)
    meas_pinned = torch.as_tensor(meas).pin_memory()
    sys_noise, meas_noise = HWN(np.zeros(2), C), HWN(np.zeros(2), C)
    state = {"t": 0}

    def step():
        z = meas_pinned[state["t"]].numpy()
        state["t"] += 1
        bf.update_identity(meas_noise, z)       # update -> estimate -> predict, the reference's loop order
        est = bf.get_point_estimates()
        bf.predict_identity(sys_noise)
        return est

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    l0 = eng.launches
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    start.record()
    for _ in range(args.steps):
        step()
    end.record()
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    gpu_launches = eng.launches - l0
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
    barrier()  # rank 0 has just spent ~0.2 s collecting the clock samples: start the e2e clocks together
    t0 = time.perf_counter()
    for _ in range(max(1, args.e2e_steps)):
        est_host = step().cpu()
    barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) * 1e3 / max(1, args.e2e_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    assert bool(torch.isfinite(est_host).all()) and not bool(bf.failed.any())
    if rank == 0:
        ms_step = float(ms) / args.steps
        per_gpu = runs * n
        peak, peak_src = peaks()
        ab = sum(algorithmic_bytes(D, 0).values()) - 8
        gbs = ab * per_gpu / (ms_step * 1e-3) / 1e9
        print(json.dumps({
            "metric": METRIC, "value": world * per_gpu / (ms_step * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"t2mc: {WORKLOADS['t2mc'][4]}", "runs_per_gpu": runs, "particles_per_run": n,
                       "particles_per_gpu": per_gpu, "resampling": "multinomial", "scan": "sequential-exact per run",
                       "l2_policy": f"state {per_gpu * 8 * D / 1e6:.0f} MB per GPU, far larger than L2",
                       "parallelism": "runs sharded over GPUs, no collective"},
            "e2e": {"value": world * per_gpu / (float(e2e) * 1e-3), "unit": "particle-steps/s", "ms_per_step": float(e2e),
                    "h2d_bytes_per_step": int(runs * D * 8), "d2h_bytes_per_step": int(runs * D * 8),
                    "path": "public API: BatchedParticleFilter.update_identity + get_point_estimates + predict_identity"},
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step (per GPU)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                         "frac": gbs / peak, "traffic": None, "peak_source": peak_src, "bytes_per_particle": ab},
            "cpu_baseline": None,
        }))
    if world > 1:
        dist.destroy_process_group()


def run_sharded(args):
    """configs[4]: one EuclideanParticleFilter over all ranks (pyrecest_b200.sharded), in-kernel Philox noise,
    global systematic resampling (all_reduce max, all_gather totals, pull-form resampling kernel reading peer memory
    over NVLink, all_reduce of the estimate sums)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pyrecest_b200 import random as prandom
    from pyrecest_b200._engine import engine
    from pyrecest_b200.sharded import ShardedParticleFilter

    name = args.workload
    manifold, D, E, n_default, desc = WORKLOADS[name]
    n = args.particles or n_default
    P = workload_params(name)
    prior, trans, noise_dist, lik_dist = make_dists(name, P)
    prandom.seed(1234 + rank)
    eng = engine(dev)
    pf = ShardedParticleFilter(manifold, n * world, D)
    pf.filter_state = prior
    z_pinned = torch.as_tensor(P["z"]).pin_memory()
    rng = np.random.default_rng(7)

    def step():
        pf.predict_nonlinear(trans, noise_dist)
        pf.update_nonlinear_using_likelihood(lik_dist.set_mode(z_pinned.numpy()), u0=float(rng.random()))
        return pf.get_point_estimate()

    for _ in range(args.warmup):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    dist.barrier()  # every rank enters the timed region together (rank 0 just slept)
    torch.cuda.synchronize()
    l0 = eng.launches
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    start.record()
    for _ in range(args.steps):
        est = step()
    end.record()
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    gpu_launches = eng.launches - l0
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # e2e: the same call sequence with the estimate read back to the host every step
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(max(1, args.e2e_steps)):
        est_host = step().cpu()
    dist.barrier()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / max(1, args.e2e_steps)], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    assert bool(torch.isfinite(est_host).all())
    if rank == 0:
        ms_step = float(ms) / args.steps
        peak, peak_src = peaks()
        ab = sum(algorithmic_bytes(D, 0).values()) - 8  # in-kernel noise, one comb offset: no noise / uniform arrays
        gbs = ab * n / (ms_step * 1e-3) / 1e9
        cpu_val, cpu_s = run_cpu("r4", args.cpu_particles, 2, 1) if not args.no_cpu else (None, None)
        print(json.dumps({
            "metric": METRIC, "value": world * n / (ms_step * 1e-3), "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{name}: {desc}", "particles_per_gpu": n, "particles_total": n * world,
                       "resampling": "systematic (global)", "scan": "sequential-exact per shard",
                       "l2_policy": f"state {n * 8 * D / 1e6:.0f} MB per GPU, far larger than L2",
                       "parallelism": f"particles sharded over {world} GPU(s); all_reduce(max) + all_gather(totals) per step, every rank produces its own output slots and reads remote sources from the peers' buffers (NVLink loads) inside the resampling kernel, all_reduce(sum) of the estimate"},
            "e2e": {"value": world * n / (float(e2e_ms) * 1e-3), "unit": "particle-steps/s", "ms_per_step": float(e2e_ms),
                    "h2d_bytes_per_step": int(8 * D), "d2h_bytes_per_step": int(8 * D + 8 * world),
                    "path": "public API: ShardedParticleFilter.predict_nonlinear + update + get_point_estimate"},
            "gpu_launches": int(gpu_launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole step (per GPU)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                         "frac": gbs / peak, "traffic": None, "peak_source": peak_src, "bytes_per_particle": ab},
            "cpu_baseline": None if args.no_cpu else {"value": cpu_val, "unit": "particle-steps/s", "cores": cpu_threads(), "kind": "port",
                                                      "sample": f"oracle port, unsharded R^4 step on {args.cpu_particles} particles"},
        }))
    dist.destroy_process_group()


def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    manifold, D, E, n_default, desc = WORKLOADS[name]
    n = args.cpu_particles
    val, s_per_step = run_cpu(name, n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "particle-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {desc}", "particles_per_gpu": args.particles or n_default,
                   "note": "reference arm = the numpy restatement of pyRecEst's algorithm (oracle/pf_oracle.py); the real "
                           "reference is pure Python, cannot travel to the GPU box and runs ~100x slower "
                           "(1.05e4 particle-steps/s on T^3, BASELINE.md)"},
        "cpu_baseline": {"value": val, "unit": "particle-steps/s", "cores": cpu_threads(), "kind": "port",
                         "sample": f"{n} particles per step (bounded sample of the workload); predict + likelihood split over {cpu_threads()} host threads, cumsum / searchsorted serial"},
        "e2e": {"value": val, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="t3", choices=sorted(WORKLOADS))
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--cpu-particles", type=int, default=500_000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--fast-scan", action="store_true", help="plain blocked scan instead of the sequential-exact one")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the sharded-filter block (configs[4])")
    ap.add_argument("--sharded-particles", type=int, default=0, help="particles per GPU of the sharded block (default 1.25e8)")
    ap.add_argument("--log-weights", action="store_true", help="value leg: log-weights between the weighting kernel and the scan "
                    "(PFB_WEIGHTS_LOG) instead of the slice-scaled linear weights")
    ap.add_argument("--runs", type=int, default=1250, help="t2mc: independent runs per GPU")
    ap.add_argument("--resampling", default="multinomial_sorted", choices=["multinomial", "multinomial_sorted", "systematic"])
    ap.add_argument("--storage", default="f64", choices=["f64", "f32"], help="storage type of particles and noise draws (north_star's "
                    "stated fp32 option; the arithmetic is fp64 either way and the line's dtype says f64)")
    args = ap.parse_args()
    # stdout carries ONE JSON line: libraries that print to file descriptor 1 from native code (NCCL's version banner at
    # communicator init) are sent to stderr; the line itself is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.gpus > 1 or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        args.no_cpu = True  # the CPU baseline is reported at N = 1 only
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if args.workload == "r4s":
            args.workload = "r4"
        if args.workload == "t2mc":
            args.workload = "t3"
        run_reference(args)
    elif args.workload == "r4s":
        run_sharded(args)
    elif args.workload == "t2mc":
        run_mc(args)
    else:
        run_gpu(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
