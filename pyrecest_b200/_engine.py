"""Thin typed wrapper: torch CUDA tensors -> raw pointers -> libpfb.so (include/pfb.h).

torch is plumbing here (device memory, streams); every computation of the hot path happens in
the hand-written kernels behind the C ABI.  One Engine per device; scratch (workspace, stats)
is owned here and grown on demand, so the C library never allocates."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_DT = {torch.float64: L.PFB_F64, torch.float32: L.PFB_F32}
_ENGINES: dict = {}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "pyrecest_b200 needs a CUDA device: the particle-filter path runs only in the sm_100a kernels "
                "of libpfb.so (there is no CPU fallback)."
            )
        self.device = torch.device(device)
        self.lib = L.lib()
        self.launches = 0  # kernels of libpfb launched through this engine
        self._ws = None
        self._ws_n = -1
        self._tile_stats_key = None  # (logw ptr, n) whose per-tile stats are in the workspace
        self.stats = torch.zeros(L.STAT_COUNT, dtype=torch.float64, device=self.device)
        self._ensure_ws(0)

    # -- scratch --------------------------------------------------------------------------------
    def _ensure_ws(self, n):
        if n > self._ws_n:
            nbytes = int(self.lib.pfb_workspace_bytes(n))
            if self._ws is not None and self._ws.numel() >= nbytes:  # e.g. grown by a batched call
                self._ws_n = n
                return _ptr(self._ws), self._ws.numel()
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            assert self._ws.data_ptr() % 256 == 0
            self._ws_n = n
            self._tile_stats_key = None
            L.check(self.lib.pfb_workspace_init(_ptr(self._ws), nbytes, _stream(self.device)))
        return _ptr(self._ws), self._ws.numel()

    def scan_error_flags(self):
        """debug word of the sequential-exact scan (0 = all invariants held)."""
        return int(self._ws[8:12].view(torch.int32).item())

    # -- kernels -----------------------------------------------------------------------------------
    def predict(self, params, x_in, x_out, noise):
        n = x_in.shape[0]
        L.check(self.lib.pfb_predict(C.byref(params), _DT[x_in.dtype], n, _ptr(x_in), _ptr(x_out), _ptr(noise),
                                     _stream(self.device)))
        self.launches += 1

    def loglik(self, lik, x, logw, w_prev=None, stats=None, scaled=False):
        """scaled: write PFB_WEIGHTS_SCALED (slice-scaled linear weights, readable only by the scan that follows)
        instead of log-weights"""
        n = x.shape[0]
        ws, wsb = self._ensure_ws(n)
        st = self.stats if stats is None else stats
        lik.out_format = L.WEIGHTS_SCALED if scaled else L.WEIGHTS_LOG
        self._tile_stats_key = (logw.data_ptr(), n, st.data_ptr(), bool(scaled))
        L.check(self.lib.pfb_loglik(C.byref(lik), _DT[x.dtype], n, _ptr(x), _ptr(w_prev), _ptr(logw), _ptr(st), ws,
                                    wsb, _stream(self.device)))
        self.launches += 1

    def predict_loglik(self, params, lik, x_in, x_out, noise, logw, stats=None, scaled=False):
        n = x_in.shape[0]
        ws, wsb = self._ensure_ws(n)
        st = self.stats if stats is None else stats
        lik.out_format = L.WEIGHTS_SCALED if scaled else L.WEIGHTS_LOG
        self._tile_stats_key = (logw.data_ptr(), n, st.data_ptr(), bool(scaled))
        L.check(self.lib.pfb_predict_loglik(C.byref(params), C.byref(lik), _DT[x_in.dtype], n, _ptr(x_in),
                                            _ptr(x_out), _ptr(noise), _ptr(logw), _ptr(st), ws, wsb,
                                            _stream(self.device)))
        self.launches += 1

    def normalize(self, logw, w_out=None, stats=None):
        n = logw.shape[0]
        ws, wsb = self._ensure_ws(0)
        st = self.stats if stats is None else stats
        L.check(self.lib.pfb_normalize(n, _ptr(logw), _ptr(w_out), _ptr(st), ws, wsb, _stream(self.device)))
        self.launches += 1 if w_out is None else 2

    def scan(self, cdf, w=None, logw=None, exact=True, stats=None):
        n = cdf.shape[0]
        ws, wsb = self._ensure_ws(n)
        st = self.stats if stats is None else stats
        mode = L.SCAN_SEQEXACT if exact else L.SCAN_FAST
        key = self._tile_stats_key
        have = logw is not None and key is not None and key[:3] == (logw.data_ptr(), n, st.data_ptr())
        if have:
            mode |= L.SCAN_TILE_STATS_VALID | (L.SCAN_SLICE_SCALED if key[3] else 0)
        else:  # pfb_scan_cdf recomputes the per-tile stats for this input (which must hold log-weights)
            self._tile_stats_key = None if logw is None else (logw.data_ptr(), n, st.data_ptr(), False)
        L.check(self.lib.pfb_scan_cdf(mode, n, _ptr(w), _ptr(logw), _ptr(cdf), _ptr(st), ws, wsb,
                                      _stream(self.device)))
        # tile_combine | tile_stats; tile_prefix (2 launches for long inputs, +1 for the exponent maximum of scaled
        # weights); exact: maps + chain + final, long inputs also the 3 segment launches around the chain; fast: final
        long_input = (n + 2047) // 2048 > 8192
        scaled = bool(mode & L.SCAN_SLICE_SCALED)
        self.launches += 1 + ((2 + (1 if scaled else 0)) if long_input else 1) + ((3 + (3 if long_input else 0)) if exact else 1)

    def resample(self, cdf, u, x_in, x_out, n_out, dim, systematic=False, idx_out=None, est_kind=L.EST_NONE,
                 est_out=None, x_in_stride=0):
        n = cdf.shape[0]
        ws, wsb = self._ensure_ws(n)
        dt = _DT[x_in.dtype] if x_in is not None else L.PFB_F64
        if x_in_stride and x_in_stride != dim:  # padded gather source (pfb_predict_params.x_pad_out)
            L.check(self.lib.pfb_resample_strided(dt, dim, n, n_out, _ptr(cdf), _ptr(u), _ptr(x_in), x_in_stride, _ptr(x_out),
                                                  _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb, _stream(self.device)))
            self.launches += 3
            return
        L.check(self.lib.pfb_resample(dt, dim, n, n_out, _ptr(cdf), _ptr(u), 1 if systematic else 0, _ptr(x_in),
                                      _ptr(x_out), _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb,
                                      _stream(self.device)))
        self.launches += 1 if systematic else 3  # scatter | guide + bucket records + resample

    def resample_rng(self, cdf, seed, offset, x_in, x_out, n_out, dim, systematic=False, idx_out=None,
                     est_kind=L.EST_NONE, est_out=None, x_in_stride=0):
        n = cdf.shape[0]
        ws, wsb = self._ensure_ws(n)
        dt = _DT[x_in.dtype] if x_in is not None else L.PFB_F64
        if x_in_stride and x_in_stride != dim:
            L.check(self.lib.pfb_resample_rng_strided(dt, dim, n, n_out, _ptr(cdf), seed, offset, _ptr(x_in), x_in_stride,
                                                      _ptr(x_out), _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb,
                                                      _stream(self.device)))
            self.launches += 3
            return
        L.check(self.lib.pfb_resample_rng(dt, dim, n, n_out, _ptr(cdf), seed, offset, 1 if systematic else 0,
                                          _ptr(x_in), _ptr(x_out), _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb,
                                          _stream(self.device)))
        self.launches += 1 if systematic else 3

    def resample_msort(self, cdf, seed, offset, x_in, x_out, n_out, dim, idx_out=None, est_kind=L.EST_NONE, est_out=None,
                       x_in_stride=0):
        """exact multinomial resampling, offspring in CDF order (pfb_resample_msort)"""
        n = cdf.shape[0]
        ws, wsb = self._ensure_ws(n)
        dt = _DT[x_in.dtype] if x_in is not None else L.PFB_F64
        L.check(self.lib.pfb_resample_msort(dt, dim, n, n_out, _ptr(cdf), seed, offset, _ptr(x_in), x_in_stride,
                                            _ptr(x_out), _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb,
                                            _stream(self.device)))
        self.launches += 3  # count tree, source ranges of the coarse bins, resample

    def msort_uniforms(self, out, seed, offset):
        ws, wsb = self._ensure_ws(0)
        L.check(self.lib.pfb_msort_uniforms(out.numel(), seed, offset, _ptr(out), ws, wsb, _stream(self.device)))
        self.launches += 2

    def resample_sharded(self, cdf, cdf_offset, total, u0, j_begin, count, n_total_out, x_in, x_out, dim, idx_out=None,
                         est_kind=L.EST_NONE, est_out=None):
        dt = _DT[x_in.dtype] if x_in is not None else L.PFB_F64
        ws, wsb = self._ensure_ws(0)
        L.check(self.lib.pfb_resample_sharded(dt, dim, cdf.shape[0], _ptr(cdf), float(cdf_offset), float(total), float(u0),
                                              int(j_begin), int(count), int(n_total_out), _ptr(x_in), _ptr(x_out),
                                              _ptr(idx_out), est_kind, _ptr(est_out), ws, wsb, _stream(self.device)))
        self.launches += 1

    def resample_sharded_dev(self, cdf, totals_dev, world, rank, u0, peer_ptrs, x_in, dim, est_kind, est_sums, status):
        """device-planned sharded scatter (pfb_resample_sharded_dev): peer_ptrs = raw device pointers of every rank's
        output buffer"""
        ws, wsb = self._ensure_ws(0)
        arr = (C.c_void_p * world)(*[C.c_void_p(p) for p in peer_ptrs])
        L.check(self.lib.pfb_resample_sharded_dev(_DT[x_in.dtype], dim, cdf.shape[0], _ptr(cdf), _ptr(totals_dev), world,
                                                  rank, float(u0), arr, _ptr(x_in), est_kind, _ptr(est_sums),
                                                  _ptr(status), ws, wsb, _stream(self.device)))
        self.launches += 2

    def resample_sharded_pull(self, cdf_ptrs, x_ptrs, totals_dev, world, rank, u0, x_out, n, dim, est_kind, est_sums, status):
        """pull-form sharded resampling (pfb_resample_sharded_pull): cdf_ptrs / x_ptrs = raw device pointers of every
        rank's CDF shard and particle buffer (peer-mapped for the other ranks); the output is local"""
        ws, wsb = self._ensure_ws(0)
        ca = (C.c_void_p * world)(*[C.c_void_p(p) for p in cdf_ptrs])
        xa = (C.c_void_p * world)(*[C.c_void_p(p) for p in x_ptrs])
        L.check(self.lib.pfb_resample_sharded_pull(_DT[x_out.dtype], dim, n, ca, xa, _ptr(totals_dev), world, rank, float(u0),
                                                   _ptr(x_out), est_kind, _ptr(est_sums), _ptr(status), ws, wsb,
                                                   _stream(self.device)))
        self.launches += 3

    def rng_uniform(self, out, seed, offset):
        L.check(self.lib.pfb_rng_uniform(out.numel(), seed, offset, _ptr(out), _stream(self.device)))
        self.launches += 1

    # -- batched runs (n_runs independent filters of run_len particles in one launch) ------------------------
    def _ensure_ws_batched(self, n_runs, run_len):
        need = int(self.lib.pfb_workspace_bytes_batched(n_runs, run_len))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws_n = -1  # single-filter calls re-check their own size against numel below
            self._tile_stats_key = None
            L.check(self.lib.pfb_workspace_init(_ptr(self._ws), need, _stream(self.device)))
        return _ptr(self._ws), self._ws.numel()

    @staticmethod
    def run_pad(run_len):
        return (run_len + 2047) // 2048 * 2048

    def loglik_batched(self, lik, x, n_runs, run_len, logw_pad, scaled=False):
        ws, wsb = self._ensure_ws_batched(n_runs, run_len)
        lik.out_format = L.WEIGHTS_SCALED if scaled else L.WEIGHTS_LOG
        L.check(self.lib.pfb_loglik_batched(C.byref(lik), _DT[x.dtype], n_runs, run_len, _ptr(x), _ptr(logw_pad),
                                            _ptr(self.stats), ws, wsb, _stream(self.device)))
        self.launches += 1

    def predict_loglik_batched(self, params, lik, x_in, x_out, noise, n_runs, run_len, logw_pad, scaled=False):
        ws, wsb = self._ensure_ws_batched(n_runs, run_len)
        lik.out_format = L.WEIGHTS_SCALED if scaled else L.WEIGHTS_LOG
        L.check(self.lib.pfb_predict_loglik_batched(C.byref(params), C.byref(lik), _DT[x_in.dtype], n_runs, run_len,
                                                    _ptr(x_in), _ptr(x_out), _ptr(noise), _ptr(logw_pad),
                                                    _ptr(self.stats), ws, wsb, _stream(self.device)))
        self.launches += 1

    def scan_batched(self, logw_pad, cdf_pad, n_runs, run_len, exact=True, scaled=False):
        """scaled: logw_pad holds the slice-scaled linear weights a *_batched weighting call with scaled=True left"""
        ws, wsb = self._ensure_ws_batched(n_runs, run_len)
        mode = (L.SCAN_SEQEXACT if exact else L.SCAN_FAST) | L.SCAN_TILE_STATS_VALID | (L.SCAN_SLICE_SCALED if scaled else 0)
        L.check(self.lib.pfb_scan_cdf_batched(mode, n_runs, run_len, _ptr(logw_pad), _ptr(cdf_pad), _ptr(self.stats),
                                              ws, wsb, _stream(self.device)))
        self.launches += 5 if exact else 3

    def resample_batched(self, cdf_pad, x_in, x_out, n_runs, run_len, dim, u=None, seed=0, offset=0, systematic=False,
                         idx_out=None, n_out_run=None):
        ws, wsb = self._ensure_ws_batched(n_runs, run_len)
        n_out_run = run_len if n_out_run is None else n_out_run
        dt = _DT[x_in.dtype] if x_in is not None else L.PFB_F64
        L.check(self.lib.pfb_resample_batched(dt, dim, n_runs, run_len, n_out_run, _ptr(cdf_pad), _ptr(u), seed, offset,
                                              1 if systematic else 0, _ptr(x_in), _ptr(x_out), _ptr(idx_out), ws, wsb,
                                              _stream(self.device)))
        self.launches += 2 if systematic else 3  # guide + resample | record build + fill + resample

    def moments_batched(self, kind, x, n_runs, run_len, dim, order=1, out=None):
        if out is None:
            out = torch.empty((n_runs, dim * dim + 2 * dim + 2), dtype=torch.float64, device=self.device)
        L.check(self.lib.pfb_moments_batched(kind, _DT[x.dtype], dim, n_runs, run_len, _ptr(x), order, _ptr(out),
                                             _stream(self.device)))
        self.launches += 1
        return out

    def moments(self, kind, x, dim, w=None, order=1, center=None, out=None):
        n = x.shape[0]
        ws, wsb = self._ensure_ws(0)
        if out is None:
            out = torch.empty(dim * dim + 2 * dim + 2, dtype=torch.float64, device=self.device)
        L.check(self.lib.pfb_moments(kind, _DT[x.dtype], dim, n, _ptr(x), _ptr(w), order, _ptr(center), _ptr(out),
                                     ws, wsb, _stream(self.device)))
        self.launches += 1
        return out

    def wrap_2pi(self, x_in, x_out):
        L.check(self.lib.pfb_wrap_2pi(_DT[x_in.dtype], x_in.numel(), _ptr(x_in), _ptr(x_out), _stream(self.device)))
        self.launches += 1


def engine(device=None) -> Engine:
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cuda")
    device = torch.device(device)
    if device.type == "cuda" and device.index is None:
        device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    key = str(device)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(device)
    return _ENGINES[key]
