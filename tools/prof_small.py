"""host-side overhead of one step of a small filter (R^2, 1e4 particles): cProfile of the public-API loop"""
import cProfile, pstats, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, bench
from pyrecest_b200 import filters as Fl
P = bench.workload_params("r2")
prior, trans, noise, lik = bench.make_dists("r2", P)
pf = Fl.EuclideanParticleFilter(10**4, 2)
pf.filter_state = prior
z = P["z"]
def step():
    pf.predict_nonlinear(trans, noise)
    pf.update_identity(lik, z)
    return pf.get_point_estimate().cpu()
for _ in range(20): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t0) * 5)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
