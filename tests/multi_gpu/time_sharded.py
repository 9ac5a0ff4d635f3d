"""phase timing of the sharded update (diagnostic): torchrun --nproc-per-node N tests/multi_gpu/time_sharded.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import bench
from pyrecest_b200 import sharded as S, random as prandom
n = int(os.environ.get("NPART", 125_000_000))
P = bench.workload_params("r4s"); prior, trans, noise, lik = bench.make_dists("r4s", P)
prandom.seed(1 + rank)
pf = S.ShardedParticleFilter("euclid", n * world, 4)
pf.filter_state = prior
orig = {k: getattr(dist, k) for k in ("all_reduce", "all_gather", "all_to_all_single", "broadcast")}
acc = {}
def wrap(name, f):
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
        acc[name] = acc.get(name, 0) + time.perf_counter() - t; return r
    return g
for k, f in orig.items(): setattr(dist, k, wrap(k, f))
be = pf.backend
be.scan = wrap("scan", be.scan); be.offspring = wrap("offspring", be.offspring)
for it in range(4):
    acc.clear()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    pf.predict_nonlinear(trans, noise)
    pf.update_nonlinear_using_likelihood(lik.set_mode(P["z"]), u0=0.37)
    est = pf.get_point_estimate()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    for r in range(world):
        if rank == r: print(f"rank {r} iter {it}: total {1e3*(t1-t0):.1f} ms  " + "  ".join(f"{k} {1e3*v:.1f}" for k, v in acc.items()), pf.last_plan["send"], flush=True)
        dist.barrier()
# and the same loop without any instrumentation
for k, f in orig.items(): setattr(dist, k, f)
pf.backend = S.CudaShardBackend(pf.local._filter_state.device)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for it in range(5):
        pf.predict_nonlinear(trans, noise)
        pf.update_nonlinear_using_likelihood(lik.set_mode(P["z"]), u0=0.37)
        est = pf.get_point_estimate()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0: print(f"iter plain loop: {1e3*(t1-t0)/5:.1f} ms/step")
dist.destroy_process_group()
