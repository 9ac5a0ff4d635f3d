"""One small invocation of every libpfb entry point, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py      (separate gpurun call)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pyrecest_b200 import random as prandom  # noqa: E402
from pyrecest_b200.distributions import (  # noqa: E402
    GaussianDistribution, HypertoroidalDiracDistribution, HypertoroidalWrappedNormalDistribution as HWN,
    LinearDiracDistribution, VonMisesDistribution, VonMisesFisherDistribution as VMF)
from pyrecest_b200.evaluation import BatchedParticleFilter  # noqa: E402
from pyrecest_b200.filters import (  # noqa: E402
    CircularParticleFilter, EuclideanParticleFilter, HypersphericalParticleFilter, HypertoroidalParticleFilter,
    LinearTransition)

prandom.seed(1)
rng = np.random.default_rng(0)
C3 = np.array([[0.7, 0.4, 0.2], [0.4, 0.6, 0.1], [0.2, 0.1, 1.0]])
N = 5000
pf = HypertoroidalParticleFilter(N, 3)
pf.filter_state = HWN(np.array([1.0, 2.0, 3.0]), C3)
for res in ("multinomial", "systematic"):
    pf.resampling = res
    pf.predict_identity(VonMisesDistribution(np.zeros(3), np.array([10.0, 5.0, 0.0])))
    pf.update_identity(HWN(np.zeros(3), 0.5 * C3), np.array([1.0, 2.0, 3.0]))
    pf.get_point_estimate()
pf.scan_exact = False
pf.predict_identity(HWN(np.zeros(3), 0.1 * C3))
pf.update_identity(HWN(np.zeros(3), 0.5 * C3), np.array([1.0, 2.0, 3.0]))
ep = EuclideanParticleFilter(N + 37, 4)
ep.filter_state = GaussianDistribution(np.zeros(4), np.eye(4))
ep.predict_nonlinear(LinearTransition(np.eye(4), np.ones(4)), GaussianDistribution(np.zeros(4), 0.1 * np.eye(4)))
ep.update_identity(GaussianDistribution(np.zeros(4), np.eye(4)), np.zeros(4))
ep.filter_state.covariance()
ep.association_likelihood(GaussianDistribution(np.zeros(4), np.eye(4)))
sp = HypersphericalParticleFilter(N, 2)
sp.filter_state = VMF(np.array([0.0, 0.0, 1.0]), 5.0)
sp.predict_identity(VMF(np.array([0.0, 0.0, 1.0]), 50.0))
sp.update_nonlinear_using_likelihood(VMF(np.array([1.0, 0.0, 0.0]), 10.0))
sp.filter_state.moment()
cp = CircularParticleFilter(777)
cp.predict_identity(HWN(np.zeros(1), np.array([[0.2]])))
cp.update_nonlinear_using_likelihood(VonMisesDistribution(np.array(2.0), 1.5))
td = HypertoroidalDiracDistribution(rng.random((3000, 2)) * 7, rng.random(3000) / 1500)
td.sample(4321)
td.trigonometric_moment(2)
LinearDiracDistribution(rng.standard_normal((100, 1)), None).mean()
with prandom.inject(normals=rng.standard_normal((2 * 3 * 2500, 2)), uniforms=rng.random(3 * 2500)):
    bf = BatchedParticleFilter("torus", 3, 2500, 2)
    bf.set_prior(HWN(np.array([1.0, 2.0]), 0.5 * np.eye(2)))
    bf.predict_identity(HWN(np.zeros(2), 0.2 * np.eye(2)))
    bf.update_identity(HWN(np.zeros(2), 0.5 * np.eye(2)), rng.random((3, 2)) * 6)
    bf.get_point_estimates()
w = np.concatenate([2.0 ** rng.integers(-60, 1, size=20000).astype(float), np.zeros(5), [1e-300, 5e-324]])
from pyrecest_b200._engine import engine  # noqa: E402

eng = engine("cuda:0")
wd = torch.as_tensor(w, device="cuda:0")
cdf = torch.empty_like(wd)
eng.scan(cdf, w=wd, exact=True)
assert np.array_equal(cdf.cpu().numpy(), np.cumsum(w)) and eng.scan_error_flags() == 0
torch.cuda.synchronize()
print("sanitize_small ok")
