import sys, time, cProfile, pstats
sys.path.insert(0, "/root/repo")
import numpy as np, torch, bench
from pyrecest_b200 import random as prandom, filters as Fl
P = bench.workload_params("t3")
prior, trans, noise, lik = bench.make_dists("t3", P)
n = 10**8
pf = Fl.HypertoroidalParticleFilter(n, 3)
pf.filter_state = prior
import os
if os.environ.get('LAZY') == '0': pf.lazy_predict = False
z = P["z"]
def step():
    pf.predict_nonlinear(trans, noise)
    pf.update_identity(lik, z)
    return pf.get_point_estimate().cpu()
for _ in range(3): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t0)*200)
# time sections
def timed(f,*a):
    torch.cuda.synchronize(); t=time.perf_counter(); r=f(*a); torch.cuda.synchronize(); return (time.perf_counter()-t)*1e3, r
for _ in range(2):
    a,_=timed(pf.predict_nonlinear, trans, noise)
    b,_=timed(pf.update_identity, lik, z)
    c,_=timed(lambda: pf.get_point_estimate().cpu())
    print("predict %.2f update %.2f est %.2f"%(a,b,c))
pr=cProfile.Profile(); pr.enable()
for _ in range(3): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
