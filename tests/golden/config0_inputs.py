"""Injected inputs of BASELINE.json configs[0] exactly as SURVEY.md 8(d) C1 states them: truth / measurements from
default_rng(1), standard normals (prior N x 2, then T x N x 2) and uniforms (T x N) from default_rng(0).
100 steps of 10^4 particles are 24 MB of draws -- not committed: tests/golden/make_golden.py (which ran the real
reference on them) and the tests both call this function; numpy's PCG64 streams are stable across versions.
No dependency on the reference or on the oracle."""
import numpy as np


def config0_draws(N=10_000, T=100):
    rng = np.random.default_rng(0)
    Z0 = rng.standard_normal((N, 2))
    Z = rng.standard_normal((T, N, 2))
    U = rng.random((T, N))
    F = np.array([[1.0, 1.0], [0.0, 1.0]])
    tr = np.random.default_rng(1)
    x = tr.standard_normal(2)
    zs = np.empty((T, 2))
    for t in range(T):
        x = F @ x + float(np.sqrt(0.1)) * tr.standard_normal(2)
        zs[t] = x + float(np.sqrt(0.5)) * tr.standard_normal(2)
    return {"F": F, "Q": 0.1 * np.eye(2), "R": 0.5 * np.eye(2), "Z0": Z0, "Z": Z, "U": U, "zs": zs}
