"""Stand-alone launches of the kernels that are not part of the bench step (profiled by tools/profile_all.sh):
moments (mean / trig / scatter / central), pfb_normalize (lse_reduce + weights), the array-fed multinomial
resampling in the reference's order, and the in-kernel draw generators, at 1e8 particles."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyrecest_b200 import _lib as L  # noqa: E402
from pyrecest_b200._engine import engine  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
dev = torch.device("cuda", 0)
eng = engine(dev)
g = torch.Generator(device=dev).manual_seed(1)
x = torch.rand((n, 3), device=dev, dtype=torch.float64, generator=g) * 6.28
w = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
w /= w.sum()
for rep in range(2):
    eng.moments(L.MOM_MEAN, x, 3, w=w)
    eng.moments(L.MOM_TRIG, x, 3, w=w, order=1)
    eng.moments(L.MOM_SCATTER, x, 3, w=w)
    m = eng.moments(L.MOM_MEAN, x, 3, w=w)[:3].clone()
    eng.moments(L.MOM_CENTRAL, x, 3, w=w, center=m)
    logw = torch.log(w)
    out = torch.empty_like(logw)
    eng.normalize(logw, out)
    cdf = torch.empty(n, dtype=torch.float64, device=dev)
    eng.scan(cdf, w=w, exact=True)
    u = torch.rand(n, device=dev, dtype=torch.float64, generator=g)
    y = torch.empty_like(x)
    est = torch.empty(2 * L.PFB_MAXD, dtype=torch.float64, device=dev)
    eng.resample(cdf, u, x, y, n, 3, est_kind=L.EST_TRIG, est_out=est)
torch.cuda.synchronize()
print("standalone kernels ok", float(est[0]))
