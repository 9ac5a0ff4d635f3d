"""soak: a few hundred steps of the public-API filters (self-cleaning counters, job lists, ping-pong buffers)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pyrecest_b200 import filters as Fl, random as prandom
prandom.seed(7)
for name, cls, n, steps in (("t3", Fl.HypertoroidalParticleFilter, 5_000_000, 300), ("r4", Fl.EuclideanParticleFilter, 3_000_000, 200)):
    P = bench.workload_params(name)
    prior, trans, noise, lik = bench.make_dists(name, P)
    for scheme in ("multinomial", "systematic"):
        pf = cls(n, 3 if name == "t3" else 4)
        pf.resampling = scheme
        pf.filter_state = prior
        z = np.array(P["z"], dtype=float)
        worst = 0.0
        for t in range(steps):
            zt = z + 0.05 * np.sin(0.1 * t)
            pf.predict_nonlinear(trans, noise)
            pf.update_identity(lik, zt)
            est = pf.get_point_estimate().cpu().numpy()
            err = np.abs(est - zt)
            if name == "t3":
                err = np.minimum(err, 2 * np.pi - err)
            worst = max(worst, float(err.max()))
        assert np.isfinite(est).all()
        print(name, scheme, "steps", steps, "worst |estimate - measurement|", round(worst, 4), flush=True)
print("soak ok")
