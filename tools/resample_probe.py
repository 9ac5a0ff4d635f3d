"""Probe of the multinomial resampling kernels on synthetic weights (run under ncu to read sectors / bytes per
particle).  usage: python tools/resample_probe.py [n] [case ...]   cases: uniform, uniform_idx, expo, lognormal"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from pyrecest_b200 import _lib as L  # noqa: E402
from pyrecest_b200._engine import engine  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
cases = sys.argv[2:] or ["uniform_idx", "uniform", "expo", "lognormal"]
eng = engine("cuda:0")
g = torch.Generator(device="cuda:0").manual_seed(1)
x = torch.rand((n, 3), dtype=torch.float64, device="cuda:0", generator=g)
xo = torch.empty_like(x)
u = torch.rand(n, dtype=torch.float64, device="cuda:0", generator=g)
idx = torch.empty(n, dtype=torch.int64, device="cuda:0")
for case in cases:
    if case.startswith("uniform"):
        w = torch.ones(n, dtype=torch.float64, device="cuda:0")
    elif case == "expo":
        w = -torch.log(torch.rand(n, dtype=torch.float64, device="cuda:0", generator=g))
    else:
        w = torch.exp(2.0 * torch.randn(n, dtype=torch.float64, device="cuda:0", generator=g))
    cdf = torch.cumsum(w, 0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        if rep == 2:
            ev[0].record()
        if case.endswith("_idx"):
            eng.resample(cdf, u, None, None, n, 3, idx_out=idx)
        else:
            eng.resample(cdf, u, x, xo, n, 3)
    ev[1].record()
    torch.cuda.synchronize()
    print(case, n, "ms", ev[0].elapsed_time(ev[1]), flush=True)
