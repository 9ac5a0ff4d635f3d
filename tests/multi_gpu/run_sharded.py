"""Multi-GPU check of pyrecest_b200.sharded (run on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/multi_gpu/run_sharded.py

Every rank runs a ShardedParticleFilter over NCCL with injected noise and comb offsets; rank 0 gathers the
particles and compares them with the oracle's sharded definition (and the estimate with the oracle's mean)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pf_oracle as O  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyrecest_b200 import random as prandom
    from pyrecest_b200.distributions import GaussianDistribution
    from pyrecest_b200.filters import LinearTransition
    from pyrecest_b200.sharded import ShardedParticleFilter

    n_loc, D, T = 200_000, 4, 3
    N = n_loc * world
    rng = np.random.default_rng(5)
    x0 = rng.standard_normal((N, D))
    if os.environ.get("SKEW", "1") == "1":
        # very unequal shards: the particles of shard g sit g standard deviations away from the measurements, so almost
        # all offspring come from shard 0 and large pieces cross to every other rank
        x0[:, 0] += 1.5 * (np.arange(N) // n_loc)
    Z = rng.standard_normal((T, N, D))
    u0s = rng.random(T)
    zs = rng.standard_normal((T, D)) * 0.3
    F = np.eye(D)
    F[0, 2] = F[1, 3] = 1.0
    Q, R = 0.1 * np.eye(D), 0.5 * np.eye(D)
    pf = ShardedParticleFilter("euclid", N, D)
    pf.pull = os.environ.get("PULL", "1") == "1"  # pull form (default) or push form of the exchange
    sl = slice(rank * n_loc, (rank + 1) * n_loc)
    pf.set_local_particles(x0[sl])
    ref = x0
    ok = True
    for t in range(T):
        with prandom.inject(normals=Z[t][sl]):
            pf.predict_nonlinear(LinearTransition(F), GaussianDistribution(np.zeros(D), Q))
            pf.update_nonlinear_using_likelihood(GaussianDistribution(zs[t], R), u0=float(u0s[t]))
        est = pf.get_point_estimate().cpu().numpy()
        got = pf.gather_particles().cpu().numpy()
        # oracle, sharded definition
        dp = O.predict_euclid(ref, O.gaussian_rows(Z[t], Q), F=F)
        lw = O.logpdf_gauss(dp, zs[t], R)
        M = lw.max()
        shards = [np.exp(lw[g * n_loc:(g + 1) * n_loc] - M) for g in range(world)]
        idx = O.sharded_systematic_indices(shards, u0s[t])
        want = dp[idx]
        if rank == 0:
            close = np.isclose(got, want, rtol=1e-12, atol=1e-13).all(axis=1)
            print(f"[{'pull' if pf.pull else 'push'}] step {t}: particles differing from the sharded oracle: {(~close).sum()} of {N}; "
                  f"estimate err {np.abs(est - want.mean(axis=0)).max():.2e}; "
                  f"(offspring produced here, stored on other ranks) {pf.offspring_traffic()}")
            ok &= (~close).sum() <= 4 and np.allclose(est, want.mean(axis=0), rtol=1e-9, atol=1e-12)
        ref = got
    if rank == 0:
        print("SHARDED_OK" if ok else "SHARDED_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
