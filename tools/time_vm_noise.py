"""how long does a stand-alone von Mises noise pass take (predict kernel with in-kernel RNG on resident particles)?"""
import sys
import torch
sys.path.insert(0, ".")
from pyrecest_b200 import _lib as L
from pyrecest_b200._engine import engine
n = 10**8
eng = engine("cuda:0")
p = L.PredictParams()
p.mode, p.dim, p.noise_cols = L.PRED_TORUS_SHIFT, 3, 3
p.rng_kind, p.rng_seed, p.rng_offset = L.RNG_VONMISES, 11, 1
p.set_von_mises_rng([10.0, 10.0, 10.0])
x = torch.zeros((n, 3), dtype=torch.float64, device="cuda:0")
y = torch.empty_like(x)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for rep in range(4):
    if rep == 1: ev[0].record()
    eng.predict(p, x, y, None)
ev[1].record(); torch.cuda.synchronize()
print("predict kernel with VM RNG:", ev[0].elapsed_time(ev[1]) / 3, "ms")
