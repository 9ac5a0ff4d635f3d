"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Reference harness: imports the real pyRecEst from /root/reference (numpy backend) with
stand-in modules for the third-party packages that are not installed in the build
container, and patches the three RNG entry points of the particle-filter path so that the
reference consumes draws injected by the caller instead of fresh random numbers.

Only usable where /root/reference exists (the build container).  It is used by
tests/golden/make_golden.py to produce the committed fixtures and by the "not gpu" tests
that compare oracle/pf_oracle.py with the real reference when the reference is present.

Patched entry points (reference file:line):
  * numpy.random.multivariate_normal -- called by GaussianDistribution.sample
    (pyrecest/distributions/nonperiodic/gaussian_distribution.py:122-123) and
    HypertoroidalWrappedNormalDistribution.sample
    (pyrecest/distributions/hypertorus/hypertoroidal_wrapped_normal_distribution.py:109-115)
  * pyrecest._backend._shared_numpy.random.choice (:28-29) -- a fresh default_rng() per call
  * scipy.stats.vonmises_fisher(...).rvs -- called by VonMisesFisherDistribution.sample
    (pyrecest/distributions/hypersphere_subset/von_mises_fisher_distribution.py:59-81)
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"

_STUB_MODULES = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
    "matplotlib.collections", "matplotlib.cm", "matplotlib.figure", "matplotlib.axes",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
    "quaternion", "pyshtools", "pyshtools.expand", "pyshtools.shio",
    "shapely", "shapely.geometry", "shapely.affinity", "shapely.plotting", "shapely.ops",
    "filterpy", "filterpy.kalman", "filterpy.common", "healpy", "parameterized",
]


class _Anything:
    """Attribute sink: any attribute access returns another sink; calling it too."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        # classes are needed where the reference subclasses / isinstance-checks
        cls = _StubMeta(name, (), {"__slots__": (), "__init__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pyrecest"))


def import_reference():
    """Import pyrecest (numpy backend) from /root/reference with stubs. Returns the package."""
    if not reference_available():
        raise RuntimeError("reference tree not present: " + REFERENCE_ROOT)
    os.environ.setdefault("PYRECEST_BACKEND", "numpy")
    for name in _STUB_MODULES:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:  # not installed -> stub
                mod = _StubModule(name)
                mod.__path__ = []  # behave like a package
                sys.modules[name] = mod
                if "." in name:
                    parent, child = name.rsplit(".", 1)
                    setattr(sys.modules[parent], child, mod)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import pyrecest  # noqa: F401
    import pyrecest.distributions  # noqa: F401
    import pyrecest.filters  # noqa: F401

    return sys.modules["pyrecest"]


class DrawQueue:
    """FIFO of injected draws, consumed in the order the reference asks for them."""

    def __init__(self, normals=None, uniforms=None, vmf_triples=None):
        self.normals = None if normals is None else np.asarray(normals, dtype=np.float64)
        self.uniforms = None if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        self.vmf = None if vmf_triples is None else np.asarray(vmf_triples, dtype=np.float64)
        self.n_pos = 0
        self.u_pos = 0
        self.v_pos = 0

    def take_normals(self, count):
        flat = self.normals.reshape(-1)
        out = flat[self.n_pos:self.n_pos + count]
        assert out.size == count, "injected normal draws exhausted"
        self.n_pos += count
        return out.copy()

    def take_uniforms(self, count):
        flat = self.uniforms.reshape(-1)
        out = flat[self.u_pos:self.u_pos + count]
        assert out.size == count, "injected uniforms exhausted"
        self.u_pos += count
        return out.copy()

    def take_vmf(self, count):
        rows = self.vmf.reshape(-1, 3)
        out = rows[self.v_pos:self.v_pos + count]
        assert out.shape[0] == count, "injected VMF triples exhausted"
        self.v_pos += count
        return out.copy()


class _InjectedGenerator:
    """Stands in for numpy.random.default_rng(): .choice runs numpy's own algorithm
    (Generator.choice with p) but takes the uniforms from the queue."""

    def __init__(self, queue):
        self.q = queue

    def choice(self, a, size=None, replace=True, p=None):
        assert replace and p is not None
        a = np.asarray(a)
        p = np.asarray(p, dtype=np.float64)
        n = int(size)
        # numpy/random/_generator.pyx, Generator.choice, replace=True, p given:
        #   cdf = p.cumsum(); cdf /= cdf[-1]; uniform_samples = self.random(shape)
        #   idx = cdf.searchsorted(uniform_samples, side='right')
        cdf = p.cumsum()
        cdf /= cdf[-1]
        u = self.q.take_uniforms(n)
        idx = cdf.searchsorted(u, side="right").astype(np.int64)
        self.last_idx = idx
        LAST["choice_idx"] = idx
        return a[idx]


LAST: dict = {}


@contextlib.contextmanager
def injected_draws(queue: DrawQueue):
    """Patch the reference's RNG entry points to consume `queue`."""
    import_reference()
    import scipy.stats as sst

    shared_random = importlib.import_module("pyrecest._backend._shared_numpy.random")
    backend_random = sys.modules["pyrecest.backend"].random

    def mvn(mean, cov, size=None, **kw):
        # numpy legacy RandomState.multivariate_normal (numpy/random/mtrand.pyx):
        #   x = standard_normal(final_shape).reshape(-1, d); (u, s, v) = svd(cov)
        #   x = dot(x, sqrt(s)[:, None] * v); x += mean
        mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
        cov = np.atleast_2d(np.asarray(cov, dtype=np.float64))
        d = mean.shape[0]
        if size is None:
            shape = ()
        elif isinstance(size, (int, np.integer)):
            shape = (int(size),)
        else:
            shape = tuple(size)
        count = int(np.prod(shape)) if shape else 1
        x = queue.take_normals(count * d).reshape(-1, d)
        (_, s, v) = np.linalg.svd(cov)
        x = np.dot(x, np.sqrt(s)[:, None] * v)
        x += mean
        return x.reshape(shape + (d,))

    def choice(*args, **kwargs):
        return _InjectedGenerator(queue).choice(*args, **kwargs)

    orig_vmf = sst.vonmises_fisher

    class _InjectedRandomState(np.random.RandomState):
        # scipy's _rvs_3d draws random(n) then _sample_uniform_direction ->
        # standard_normal((n, 2)) (scipy/stats/_multivariate.py, vonmises_fisher_gen)
        def random(self, n):
            self._rows = queue.take_vmf(int(np.prod(n)))
            return self._rows[:, 0].copy()

        def standard_normal(self, shape):
            return self._rows[:, 1:3].reshape(shape).copy()

    def vmf_factory(mu=None, kappa=1, seed=None):
        frozen = orig_vmf(mu, kappa)
        real_rvs = frozen.rvs

        def rvs(size=1, random_state=None):
            return real_rvs(size, random_state=_InjectedRandomState())

        frozen.rvs = rvs
        return frozen

    saved = (np.random.multivariate_normal, shared_random.choice, backend_random.choice,
             backend_random.multivariate_normal, sst.vonmises_fisher)
    np.random.multivariate_normal = mvn
    shared_random.choice = choice
    backend_random.choice = choice
    backend_random.multivariate_normal = mvn
    sst.vonmises_fisher = vmf_factory
    try:
        yield queue
    finally:
        (np.random.multivariate_normal, shared_random.choice, backend_random.choice,
         backend_random.multivariate_normal, sst.vonmises_fisher) = saved
