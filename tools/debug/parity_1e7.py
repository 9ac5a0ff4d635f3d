import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import pf_oracle as O
from pyrecest_b200 import random as prandom
from pyrecest_b200.distributions import HypertoroidalWrappedNormalDistribution as HWN, VonMisesDistribution
from pyrecest_b200.filters import HypertoroidalParticleFilter
import test_gpu_fullsize as T
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(n)
C = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]]); z = np.array([1.0, 2.0, 3.0])
d0 = rng.random((n, 3)) * 2 * np.pi; vm = rng.vonmises(0.0, 10.0, size=(n, 3)); u = rng.random(n)
pf = HypertoroidalParticleFilter(n, 3, device="cuda:0"); pf.filter_state.d = d0
with prandom.inject(vonmises=vm, uniforms=u):
    pf.predict_identity(VonMisesDistribution(np.zeros(3), np.array([10.0, 10.0, 10.0])))
    pf.update_identity(HWN(np.zeros(3), 0.5 * C), z)
got = pf.filter_state.d.cpu().numpy()
dp = O.predict_torus(d0, vm, shift=True)
lik = T._wn_pdf_pruned(dp, z, 0.5 * C)
wp = O.reweigh(O.uniform_w(n), lik); idx = O.multinomial_indices(wp, u)
cdf = np.cumsum(wp); cdf /= cdf[-1]
bad = np.nonzero(np.any(got != dp[idx], axis=1))[0]
print("bad", bad.size)
# device-side pieces for comparison
eng = pf._eng()
for j in bad[:10]:
    i = idx[j]
    win = np.arange(max(i - 2000, 0), min(i + 2000, n))
    hit = win[np.all(dp[win] == got[j], axis=1)]
    print("j", j, "ref idx", i, "gpu idx", hit, "u", u[j], "cdf[i-1..i+1]", cdf[i-1:i+2], "u-cdf[i-1]", u[j]-cdf[i-1], "cdf[i]-u", cdf[i]-u[j], "w", wp[i-1:i+2]*n)
