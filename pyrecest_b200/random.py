"""Draw source of the filters: where standard normals, uniforms, von Mises draws and VMF triples
come from.  Default = the torch CUDA generator (draws are produced on the device and never touch
the host).  `inject(...)` replaces it by host-supplied arrays consumed in order, which is how the
parity tests feed the CUDA path the same draws the reference consumed
(oracle/ref_harness.py patches the reference's three RNG entry points the same way)."""
from __future__ import annotations

import contextlib

import torch


class DeviceDraws:
    """Production source.  The per-step draws of the filters (system noise, resampling uniforms) are
    generated INSIDE the predict / resample kernels by a counter-based Philox4x32-10 keyed with
    (seed, call counter, particle index): `fused = True` tells the filters to pass the key instead of
    arrays.  Prior sampling (`distribution.sample(n)`) uses the torch CUDA generator below."""

    fused = True

    def __init__(self, seed=None):
        import os

        self._gens = {}
        self._seed = seed
        self.seed = int(seed) & (2**64 - 1) if seed is not None else int.from_bytes(os.urandom(8), "little")
        self._offset = 0

    def next_offset(self):
        self._offset += 1
        return self._offset

    def derive(self, k):
        """child source number k (e.g. the rank of a shard): an independent stream that is still a function of this
        source's seed, so that SPMD code calling seed(s) with the same s on every rank stays reproducible without the
        shards drawing identical noise"""
        if not hasattr(self, "_children"):
            self._children = {}
        if k not in self._children:
            z = (self.seed + 0x9E3779B97F4A7C15 * (int(k) + 1)) & (2**64 - 1)  # splitmix64 finaliser
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2**64 - 1)
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2**64 - 1)
            child = DeviceDraws(seed=(z ^ (z >> 31)) & (2**63 - 1))
            self._children[k] = child
        return self._children[k]

    def _gen(self, device):
        key = str(device)
        if key not in self._gens:
            g = torch.Generator(device=device)
            if self._seed is not None:
                g.manual_seed(self._seed)
            else:
                g.seed()
            self._gens[key] = g
        return self._gens[key]

    def normals(self, shape, device, dtype=torch.float64):
        return torch.randn(shape, device=device, dtype=dtype, generator=self._gen(device))

    def uniforms(self, shape, device):
        return torch.rand(shape, device=device, dtype=torch.float64, generator=self._gen(device))

    def vonmises(self, shape, kappa, device, dtype=torch.float64):
        # zero-mean von Mises draws in [-pi, pi) (Best-Fisher rejection, as torch implements it)
        k = torch.as_tensor(kappa, dtype=torch.float64, device=device).expand(shape)
        loc = torch.zeros(shape, dtype=torch.float64, device=device)
        return torch.distributions.VonMises(loc, k).sample().to(dtype)

    def vmf_triples(self, n, device, dtype=torch.float64):
        t = torch.empty((n, 3), device=device, dtype=dtype)
        g = self._gen(device)
        t[:, 0] = torch.rand(n, device=device, dtype=dtype, generator=g)
        t[:, 1:] = torch.randn((n, 2), device=device, dtype=dtype, generator=g)
        return t


class InjectedDraws:
    """FIFO of host-provided draws (numpy arrays / tensors)."""

    fused = False

    def __init__(self, normals=None, uniforms=None, vonmises=None, vmf_triples=None):
        self._q = {"normals": normals, "uniforms": uniforms, "vonmises": vonmises, "vmf": vmf_triples}
        self._pos = {k: 0 for k in self._q}

    def _take(self, key, count, device, dtype):
        src = self._q[key]
        if src is None:
            raise RuntimeError(f"no injected {key} available")
        flat = torch.as_tensor(src).reshape(-1)
        pos = self._pos[key]
        out = flat[pos:pos + count]
        if out.numel() != count:
            raise RuntimeError(f"injected {key} exhausted")
        self._pos[key] = pos + count
        return out.to(device=device, dtype=dtype)

    def normals(self, shape, device, dtype=torch.float64):
        n = int(torch.Size(shape).numel())
        return self._take("normals", n, device, dtype).reshape(shape)

    def uniforms(self, shape, device):
        n = int(torch.Size(shape).numel())
        return self._take("uniforms", n, device, torch.float64).reshape(shape)

    def vonmises(self, shape, kappa, device, dtype=torch.float64):
        n = int(torch.Size(shape).numel())
        return self._take("vonmises", n, device, dtype).reshape(shape)

    def vmf_triples(self, n, device, dtype=torch.float64):
        return self._take("vmf", 3 * n, device, dtype).reshape(n, 3)


class NullDraws:
    """Hands out no draws at all (None): used to build a parameter block whose noise array the caller supplies."""

    fused = False

    def normals(self, *a, **k):
        return None

    uniforms = vonmises = vmf_triples = normals


_SOURCE = DeviceDraws()


def source():
    return _SOURCE


def seed(s):
    """Seed the device draw source (the analogue of pyrecest.backend.random.seed)."""
    global _SOURCE
    _SOURCE = DeviceDraws(seed=s)


@contextlib.contextmanager
def using(src):
    """temporarily make `src` the draw source"""
    global _SOURCE
    saved = _SOURCE
    _SOURCE = src
    try:
        yield src
    finally:
        _SOURCE = saved


@contextlib.contextmanager
def null():
    global _SOURCE
    saved = _SOURCE
    _SOURCE = NullDraws()
    try:
        yield _SOURCE
    finally:
        _SOURCE = saved


@contextlib.contextmanager
def inject(normals=None, uniforms=None, vonmises=None, vmf_triples=None):
    global _SOURCE
    saved = _SOURCE
    _SOURCE = InjectedDraws(normals, uniforms, vonmises, vmf_triples)
    try:
        yield _SOURCE
    finally:
        _SOURCE = saved
