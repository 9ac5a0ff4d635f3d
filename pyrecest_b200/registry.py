"""Host-side reduction of the reference's objects to the POD parameter blocks of include/pfb.h.

Transitions and likelihoods are a REGISTRY: identity / linear transitions, and the built-in
distribution kinds (Gaussian, hypertoroidal wrapped normal, circular wrapped normal, von Mises,
von Mises-Fisher).  An arbitrary Python callable is rejected -- it is never run on the CPU
(BASELINE.json north_star).  Objects are recognised by class name and attributes, so both the
classes of pyrecest_b200.distributions and the reference's own pyrecest.distributions instances
(numpy backend) are accepted.

Only tiny (<= 6 x 6) host arrays are touched here; numpy (and scipy.linalg.eigh when present,
to follow scipy.stats.multivariate_normal's factorisation) do that arithmetic.
"""
from __future__ import annotations

import itertools
import math

import numpy as np

from . import _lib as L

TWO_PI = 2.0 * math.pi


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# transitions
# --------------------------------------------------------------------------------------------
class IdentityTransition:
    """f(x) = x.  What predict_identity dispatches (abstract_particle_filter.py:18-19 passes an
    identity lambda, which cannot be recognised, so the filters override predict_identity)."""

    def __call__(self, x):
        return x


class LinearTransition:
    """f(x) = x @ F.T + b, the registry form of a linear system model."""

    def __init__(self, F, b=None):
        self.F = _np(F)
        assert self.F.ndim == 2 and self.F.shape[0] == self.F.shape[1], "F must be square"
        self.b = None if b is None else _np(b).reshape(-1)

    def __call__(self, x):
        import torch

        F = torch.as_tensor(self.F, dtype=x.dtype, device=x.device)
        y = x @ F.T
        if self.b is not None:
            y = y + torch.as_tensor(self.b, dtype=x.dtype, device=x.device)
        return y


class LinearNoiseTransition:
    """f(x, n) = x @ F.T + n @ G.T + b: the registry form of a non-additive system model for
    predict_nonlinear_nonadditive (abstract_particle_filter.py:56-69 takes an arbitrary f(x, noise))."""

    def __init__(self, F, G, b=None):
        self.F = _np(F)
        self.G = _np(G)
        assert self.F.ndim == 2 and self.F.shape[0] == self.F.shape[1], "F must be square"
        assert self.G.ndim == 2 and self.G.shape[0] == self.F.shape[0], "G must have one row per state component"
        self.b = None if b is None else _np(b).reshape(-1)


def transition_params(f, dim):
    """(has_F, F, has_b, b) for a registry transition; raises for anything else."""
    if f is None or isinstance(f, IdentityTransition):
        return None, None
    if isinstance(f, LinearTransition):
        if f.F.shape[0] != dim:
            raise ValueError(f"LinearTransition is {f.F.shape[0]}-D, filter state is {dim}-D")
        return f.F, f.b
    raise NotImplementedError(
        "predict_nonlinear only accepts registry transitions (IdentityTransition, LinearTransition(F, b)); "
        f"got {type(f).__name__}. Arbitrary Python callables are not run on the CPU."
    )


# --------------------------------------------------------------------------------------------
# distribution kinds
# --------------------------------------------------------------------------------------------
_KIND_BY_NAME = {
    "GaussianDistribution": "gauss",
    "HypertoroidalWrappedNormalDistribution": "wn_multi",
    "HypertoroidalWNDistribution": "wn_multi",
    "ToroidalWrappedNormalDistribution": "wn_multi",
    "WrappedNormalDistribution": "wn_1d",
    "WNDistribution": "wn_1d",
    "VonMisesDistribution": "vm",
    "VMDistribution": "vm",
    "VonMisesFisherDistribution": "vmf",
    "VMFDistribution": "vmf",
    # uniform densities (constant likelihood; association_likelihood of the reference's tests uses one)
    "CircularUniformDistribution": "uniform",
    "HypertoroidalUniformDistribution": "uniform",
    "HypersphericalUniformDistribution": "uniform",
    "HyperhemisphericalUniformDistribution": "uniform",
}


def kind_of(obj):
    for cls in type(obj).__mro__:
        k = _KIND_BY_NAME.get(cls.__name__)
        if k is not None:
            return k
    return None


def as_distribution(likelihood):
    """Accept a registry distribution or its bound .pdf (what update_identity builds,
    abstract_particle_filter.py:108); reject other callables."""
    if kind_of(likelihood) is not None:
        return likelihood
    owner = getattr(likelihood, "__self__", None)
    if owner is not None and getattr(likelihood, "__name__", "") == "pdf" and kind_of(owner) is not None:
        return owner
    raise NotImplementedError(
        "likelihood must be one of the built-in distributions (Gaussian, HypertoroidalWrappedNormal, "
        "WrappedNormal, VonMises, VonMisesFisher) or its bound .pdf; arbitrary Python callables are "
        f"rejected, not run on the CPU (got {type(likelihood).__name__})."
    )


_FACTOR_CACHE: dict = {}


def _memo(tag, C, make):
    """a filter passes the same covariance every step: the svd / eigh of a d x d matrix (tens of microseconds of
    LAPACK wrapper each) would otherwise sit on the critical path of small filters.  Values are treated as read-only."""
    key = (tag, C.shape, C.tobytes())
    hit = _FACTOR_CACHE.get(key)
    if hit is None:
        if len(_FACTOR_CACHE) > 512:
            _FACTOR_CACHE.clear()
        hit = _FACTOR_CACHE[key] = make(C)
    return hit


def gaussian_noise_factor(C):
    """numpy legacy multivariate_normal factor: draws are z @ A + mean with
    (u, s, v) = svd(C); A = sqrt(s)[:, None] * v  (gaussian_distribution.py:122-123)."""
    C = np.ascontiguousarray(np.atleast_2d(_np(C)), dtype=np.float64)

    def make(C):
        (_, s, v) = np.linalg.svd(C)
        return np.sqrt(s)[:, None] * v

    return _memo("svd", C, make)


def gaussian_whitening(C):
    """(W, log|C|) the way scipy.stats.multivariate_normal factorises (eigh; W = V / sqrt(lambda))."""
    C = np.ascontiguousarray(np.atleast_2d(_np(C)), dtype=np.float64)
    return _memo("eigh", C, _gaussian_whitening)


def _gaussian_whitening(C):
    try:
        import scipy.linalg

        s, u = scipy.linalg.eigh(C, lower=True, check_finite=False)
    except ImportError:  # pragma: no cover
        s, u = np.linalg.eigh(C)
    if np.min(s) <= 0:
        raise ValueError("covariance must be positive definite")
    return np.multiply(u, np.sqrt(1.0 / s)), float(np.sum(np.log(s)))


def log_i0(kappa):
    """log I0(kappa), stable for large kappa."""
    try:
        from scipy.special import ive

        return float(np.log(ive(0, kappa)) + kappa)
    except ImportError:  # pragma: no cover
        import torch

        return float(torch.log(torch.special.i0e(torch.tensor(kappa, dtype=torch.float64))) + kappa)


def _log_iv(nu, kappa):
    from scipy.special import ive

    return float(np.log(ive(nu, kappa)) + kappa)


def vmf_log_norm_const(kappa, D):
    """log C_D(kappa) of VonMisesFisherDistribution.__init__ (von_mises_fisher_distribution.py:41-49)."""
    dim = D - 1
    if dim == 2:
        # kappa / (4 pi sinh kappa), written to survive large kappa
        return math.log(kappa) - math.log(2.0 * math.pi) - kappa - math.log1p(-math.exp(-2.0 * kappa))
    half = (dim + 1) / 2.0
    return (half - 1.0) * math.log(kappa) - half * math.log(TWO_PI) - _log_iv(half - 1.0, kappa)


_IMAGE_TABLE_CACHE: dict = {}


class device_put:
    """callable host array -> contiguous device tensor, hashable per device for the table cache"""

    def __init__(self, device):
        self.device = device
        self.cache_key = str(device)

    def __call__(self, table):
        import torch

        return torch.as_tensor(table, device=self.device).contiguous()


def wn_image_table_cached(mu, P, device_put=None):
    """(table, n_images, device copy) memoised on (mu, P): a filter re-uses the same measurement-noise
    covariance every step and often revisits measurements, and the pruning below costs milliseconds of
    host time that would otherwise sit on the critical path of every update."""
    key = (np.asarray(mu, dtype=np.float64).tobytes(), np.asarray(P, dtype=np.float64).tobytes(),
           None if device_put is None else getattr(device_put, "cache_key", id(device_put)))
    hit = _IMAGE_TABLE_CACHE.get(key)
    if hit is None:
        table, n_img = wn_image_table(mu, P)
        # at most 32 images: 32-bit masks and a 16-per-dimension grid (4096 cells of 5 bytes for T^3: fewer candidates
        # per cell, 3.0 -> 2.4 on average and 5.6 -> 3.6 for the slowest lane of a warp); else 8 per dimension
        narrow = n_img <= 32
        G = WN_CELL_GRID_FINE if narrow else WN_CELL_GRID
        cells = wn_cell_masks(mu, table, G=G)
        if cells is not None:
            cells = wn_cell_table(mu, table, cells, G=G, narrow=narrow)
        flat = table.reshape(-1) if cells is None else np.concatenate([table.reshape(-1), cells.view(np.float64)])
        dev = device_put(flat) if device_put is not None else None
        if len(_IMAGE_TABLE_CACHE) > 256:
            _IMAGE_TABLE_CACHE.clear()
        hit = _IMAGE_TABLE_CACHE[key] = (table, n_img, dev, 0 if cells is None else G, 2 if (cells is not None and narrow) else 0)
    return hit


WN_CELL_GRID = 8       # cells per dimension of the candidate grid
WN_CELL_GRID_FINE = 16 # ... when the masks are 32-bit (at most 32 images)
WN_SCREEN_CUT = 78.0   # kScreenCut of pfb_predict_loglik.cu


def _cell_boxes(lo, n_cells, wid, d):
    idx = np.stack(np.meshgrid(*([np.arange(n_cells)] * d), indexing="ij"), axis=-1).reshape(-1, d)  # (cells, d)
    flat = (idx * (n_cells ** np.arange(d))[None, :]).sum(axis=1)  # c_0 + n c_1 + n^2 c_2
    a = lo + idx * wid - 1e-6
    return idx, flat, a, a + wid + 2e-6


def wn_cell_masks(mu, table, G=WN_CELL_GRID, any_mean=False):
    """Candidate masks of the image sum.  The kernel adds the images k with t_k <= min_j t_j + 38,
    t_k(dev) = g_k.dev + h_k / 2 linear in dev (r_k = 2 t_k below).  Over a cell r_k is bounded by its corner
    values, so image k can only matter somewhere in the cell if min_cell r_k <= min_j max_cell r_j + cut: bit k of the
    cell's uint64 mask.  A superset of what matters at every point of the cell (+1 of slack, cells inflated by 1e-6
    for the rounding of the cell index) that contains the minimiser.
    The cells cut the box dev in [-pi - mu, pi - mu) of ONE measurement into G cells per dimension (index
    c_0 + G c_1 + G^2 c_2), or -- any_mean -- the box [-3 pi, pi)^d that holds dev for every mu in [0, 2 pi)^d into 2 G
    cells per dimension (pfb_lik_params.cell_span = 1: one table for all runs of a batched launch)."""
    d = table.shape[1] // 2
    K = table.shape[0]
    if d > 3 or K > 64:
        return None
    g, h = table[:, :d], table[:, d]
    lo = np.full(d, -3.0 * math.pi) if any_mean else -math.pi - _np(mu).reshape(-1)
    n_cells = 2 * G if any_mean else G
    idx, flat, a, b = _cell_boxes(lo, n_cells, TWO_PI / G, d)
    ga, gb = 2.0 * g[None, :, :] * a[:, None, :], 2.0 * g[None, :, :] * b[:, None, :]
    rmin = h[None, :] + np.minimum(ga, gb).sum(axis=-1)   # (cells, K)
    rmax = h[None, :] + np.maximum(ga, gb).sum(axis=-1)
    keep = rmin <= rmax.min(axis=1, keepdims=True) + WN_SCREEN_CUT + 1.0
    bits = (keep.astype(np.uint64) << np.arange(K, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint64)
    out = np.zeros(n_cells ** d, dtype=np.uint64)
    out[flat] = bits
    return out


def wn_cell_table(mu, table, masks, G=WN_CELL_GRID, limit=60.0, any_mean=False, narrow=False):
    """Device form of the candidate grid: the uint64 masks followed by one reference byte per cell (padded to 8
    bytes) -- the image that is best at the cell's centre, t_k = g_k.dev + h_k / 2 smallest.  The kernel anchors the
    image sum there: exact quadratic form of the reference image, every other candidate through exp(-(t_k - t_ref)).
    Returns None (no grid: the kernel finds the best image itself) if some candidate could beat the reference by more
    than e^limit somewhere in its cell (tiny covariances): the image sum `lin` then stays below K e^limit, so that the
    slice-scaled weight exp(a - c) * lin of the weighting kernel (a - c <= 600 by its rescaling rule) cannot overflow."""
    d = table.shape[1] // 2
    K = table.shape[0]
    g, h = table[:, :d], table[:, d]
    lo = np.full(d, -3.0 * math.pi) if any_mean else -math.pi - _np(mu).reshape(-1)
    n_cells = 2 * G if any_mean else G
    wid = TWO_PI / G
    idx, flat, a, b = _cell_boxes(lo, n_cells, wid, d)
    bits = ((masks[flat][:, None] >> np.arange(K, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)  # (cells, K)
    centre = lo + (idx + 0.5) * wid
    t = centre @ g.T + 0.5 * h[None, :]
    ref = np.where(bits, t, np.inf).argmin(axis=1)                                                         # (cells,)
    dg = g[ref][:, None, :] - g[None, :, :]                                                                # (cells, K, d)
    gain = np.maximum(dg * a[:, None, :], dg * b[:, None, :]).sum(axis=-1) + 0.5 * (h[ref][:, None] - h[None, :])
    if np.any(np.where(bits, gain, -np.inf) > limit):
        return None
    nc = n_cells ** d
    if narrow:  # uint32 masks, then the reference bytes, padded to 8 bytes in all
        assert K <= 32
        out = np.zeros(((nc * 5 + 7) // 8) * 8, dtype=np.uint8)
        out[: nc * 4] = masks.astype(np.uint32).view(np.uint8)
        out[nc * 4 + flat] = ref.astype(np.uint8)
        return out.view(np.uint64)
    refs = np.zeros(((nc + 7) // 8) * 8, dtype=np.uint8)
    refs[flat] = ref.astype(np.uint8)
    return np.concatenate([masks.view(np.uint8), refs]).view(np.uint64)


def wn_image_table(mu, P, m=3, cut=None):
    """Image list of HypertoroidalWrappedNormalDistribution.pdf
    (hypertoroidal_wrapped_normal_distribution.py:80-91): all offsets 2 pi k, k in {-m..m}^d, pruned to
    those that can matter.  maha_k(x) - maha_j(x) = 2 (g_k - g_j).dev + (h_k - h_j) is LINEAR in
    dev = x - mu, so its minimum over the box dev in [-pi - mu, pi - mu) is attained at a corner.
    Image k is dropped iff some image j beats it by more than `cut` everywhere in the box: then
    pdf_k <= e^{-cut/2} pdf_j and the (2m+1)^d dropped terms together stay below 2^-53 of the sum.
    Rows: (g_k = P o_k, h_k = o_k^T P o_k, o_k)."""
    mu = _np(mu).reshape(-1)
    d = mu.shape[0]
    ks = np.array(list(itertools.product(range(-m, m + 1), repeat=d)), dtype=np.float64)
    offs = ks * TWO_PI
    g = offs @ P
    h = np.einsum("kd,kd->k", g, offs)
    if cut is None:
        cut = 2.0 * (math.log(len(offs)) + 53.0 * math.log(2.0))
    lo = -math.pi - mu
    hi = math.pi - mu
    dg = g[:, None, :] - g[None, :, :]              # (k, j, d)
    lin_min = np.sum(np.minimum(2.0 * dg * lo, 2.0 * dg * hi), axis=-1)
    gap_min = lin_min + (h[:, None] - h[None, :])   # min over the box of maha_k - maha_j
    keep = ~np.any(gap_min > cut, axis=1)
    table = np.concatenate([g[keep], h[keep, None], offs[keep]], axis=1)
    table = table[np.argsort(h[keep], kind="stable")]  # nearest images first: the kernel's running minimum settles early
    return np.ascontiguousarray(table), int(keep.sum())


def predict_noise_spec(noise, dim):
    """Reduce a noise distribution to (kind, fields).  kinds: None, 'gauss', 'wn', 'vm', 'vmf'."""
    if noise is None:
        return None, {}
    k = kind_of(noise)
    if k == "gauss":
        mu = _np(noise.mu).reshape(-1)
        return "gauss", {"mu": mu, "A": gaussian_noise_factor(noise.C)}
    if k in ("wn_multi", "wn_1d"):
        mu = _np(noise.mu).reshape(-1)
        C = np.atleast_2d(_np(noise.C))
        return "wn", {"mu": np.mod(mu, TWO_PI), "A": gaussian_noise_factor(C)}
    if k == "vm":
        return "vm", {"mu": np.atleast_1d(_np(noise.mu)).reshape(-1), "kappa": np.atleast_1d(_np(noise.kappa)).reshape(-1)}
    if k == "vmf":
        return "vmf", {"kappa": float(_np(noise.kappa))}
    raise NotImplementedError(
        f"noise distribution {type(noise).__name__} is not in the device registry "
        "(Gaussian, (Hypertoroidal)WrappedNormal, VonMises, VonMisesFisher)"
    )


def lik_params(dist, dim, images_to_device=None):
    """Fill a LikParams from a registry distribution; `dim` = stored components of the filter state.
    Returns (LikParams, keepalive) -- keepalive holds the device image table, if any."""
    k = kind_of(dist)
    p = L.LikParams()
    p.dim = dim
    keep = None
    if k == "gauss":
        mu = _np(dist.mu).reshape(-1)
        if mu.shape[0] != dim:
            raise AssertionError("likelihood dimension does not match the filter state")
        W, logdet = gaussian_whitening(dist.C)
        p.kind = L.LIK_GAUSS
        p.mu[:dim] = mu.tolist()
        p.W[: dim * dim] = W.reshape(-1).tolist()
        p.logc = -0.5 * (dim * math.log(TWO_PI) + logdet)
    elif k == "wn_multi":
        mu = _np(dist.mu).reshape(-1)
        if mu.shape[0] != dim:
            raise AssertionError("likelihood dimension does not match the filter state")
        C = np.atleast_2d(_np(dist.C))
        W, logdet = gaussian_whitening(C)
        table, n_img, dev_table, cell_grid, cell_flags = wn_image_table_cached(mu, W @ W.T, images_to_device)
        p.kind = L.LIK_WN_MULTI
        p.cell_grid = cell_grid if dev_table is not None else 0
        p.cell_span = cell_flags if dev_table is not None else 0
        p.mu[:dim] = mu.tolist()
        p.W[: dim * dim] = W.reshape(-1).tolist()
        p.logc = -0.5 * (dim * math.log(TWO_PI) + logdet)
        p.n_images = n_img
        if dev_table is not None:
            keep = dev_table
            p.images_dev = keep.data_ptr()
        else:
            keep = table
    elif k == "wn_1d":
        if dim != 1:
            raise AssertionError("WrappedNormalDistribution is circular (dim 1)")
        sigma = float(np.sqrt(_np(dist.C).reshape(-1)[0])) if hasattr(dist, "C") else float(_np(dist.sigma))
        if sigma <= 0:
            raise ValueError(f"sigma must be >0, but received {sigma}.")
        p.kind = L.LIK_WN_1D
        p.mu[0] = float(_np(dist.mu).reshape(-1)[0])
        p.sigma = sigma
        p.logc = math.log(1.0 / math.sqrt(TWO_PI) / sigma)
    elif k == "vm":
        if dim != 1:
            raise AssertionError("VonMisesDistribution is circular (dim 1)")
        kappa = float(_np(dist.kappa))
        p.kind = L.LIK_VM
        p.mu[0] = float(_np(dist.mu).reshape(-1)[0])
        p.kappa = kappa
        p.logc = -(math.log(TWO_PI) + log_i0(kappa))
    elif k == "vmf":
        mu = _np(dist.mu).reshape(-1)
        if mu.shape[0] != dim:
            raise AssertionError("likelihood dimension does not match the filter state")
        kappa = float(_np(dist.kappa))
        p.kind = L.LIK_VMF
        p.mu[:dim] = mu.tolist()
        p.kappa = kappa
        p.logc = vmf_log_norm_const(kappa, dim)
    elif k == "uniform":
        # a constant density: the von Mises / von Mises-Fisher kernels with kappa = 0 evaluate exactly log c
        name = type(dist).__name__
        if "spher" in name:
            # surface area of S^(dim-1) embedded in R^dim: 2 pi^(dim/2) / Gamma(dim/2)  (half of it for a hemisphere)
            log_area = math.log(2.0) + 0.5 * dim * math.log(math.pi) - math.lgamma(0.5 * dim)
            if "hemi" in name:
                log_area -= math.log(2.0)
            p.logc = -log_area
        else:
            p.logc = -dim * math.log(TWO_PI)
        p.kind = L.LIK_VM if dim == 1 else L.LIK_VMF
        p.kappa = 0.0
    else:
        raise NotImplementedError(f"{type(dist).__name__} is not a registry likelihood")
    return p, keep
