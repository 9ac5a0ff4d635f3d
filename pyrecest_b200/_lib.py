"""ctypes binding of libpfb.so (include/pfb.h).  No torch types cross this boundary: only raw
device pointers, sizes and the CUDA stream handle.

The library is REQUIRED: there is no CPU or pure-PyTorch fallback.  If libpfb.so is missing or
a symbol is absent, importing this module's `lib()` raises immediately."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFB_LIB_PATH") or os.path.join(HERE, "lib", "libpfb.so")  # (override: A/B builds of the library)
CSRC = os.path.join(HERE, "csrc")

PFB_MAXD = 6
ABI_VERSION = 8
VM_STRIPS = 4096
RNG_NONE, RNG_NORMAL, RNG_VONMISES, RNG_VMF = range(4)
PFB_F64, PFB_F32 = 0, 1
# predict modes
PRED_EUCLID, PRED_TORUS_SHIFT, PRED_TORUS_ADD, PRED_SPHERE_VMF, PRED_SPHERE_ADD = range(5)
# likelihood kinds
LIK_GAUSS, LIK_WN_MULTI, LIK_WN_1D, LIK_VM, LIK_VMF = range(5)
# stats words
STAT_MAX, STAT_SUMEXP, STAT_SUMSQ, STAT_FLAGS, STAT_TOTAL, STAT_COUNT = 0, 1, 2, 3, 4, 8
SCAN_FAST, SCAN_SEQEXACT = 0, 1
SCAN_TILE_STATS_VALID = 0x100
SCAN_SLICE_SCALED = 0x200
WEIGHTS_LOG, WEIGHTS_SCALED = 0, 1
EST_NONE, EST_SUM, EST_TRIG = 0, 1, 2
MOM_MEAN, MOM_TRIG, MOM_SCATTER, MOM_CENTRAL = 0, 1, 2, 3


class PredictParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("dim", C.c_int32), ("has_F", C.c_int32), ("has_b", C.c_int32),
        ("noise_cols", C.c_int32), ("has_A", C.c_int32), ("has_mu_noise", C.c_int32),
        ("renormalize", C.c_int32),
        ("F", C.c_double * (PFB_MAXD * PFB_MAXD)), ("b", C.c_double * PFB_MAXD),
        ("A", C.c_double * (PFB_MAXD * PFB_MAXD)), ("mu_noise", C.c_double * PFB_MAXD),
        ("kappa", C.c_double), ("exp_m2k", C.c_double),
        ("rng_kind", C.c_int32), ("rng_vm_table_stride", C.c_int32), ("rng_seed", C.c_uint64), ("rng_offset", C.c_uint64),
        ("rng_kappa", C.c_double * PFB_MAXD), ("rng_vm_area", C.c_double * PFB_MAXD),
        ("rng_vm_table_dev", C.c_void_p), ("rng_vm_table_idx", C.c_int32 * PFB_MAXD),
        ("x_pad_out", C.c_void_p), ("x_pad_stride", C.c_int32), ("reserved2", C.c_int32),
    ]

    def set_von_mises_rng(self, kappas, device="cuda:0"):
        """per-dimension concentration + the strip tables of the in-kernel sampler (pfb_vm_strip_table), uploaded
        once per (set of kappas, device) and shared between dimensions of equal kappa"""
        kappas = [float(k) for k in kappas]
        distinct = sorted({k for k in kappas if k >= 1e-8})
        tables = vm_tables_on_device(tuple(distinct), str(device)) if distinct else None
        for c, k in enumerate(kappas):
            self.rng_kappa[c] = k
            if k >= 1e-8:
                self.rng_vm_table_idx[c] = distinct.index(k)
                self.rng_vm_area[c] = vm_strip_table(k)[1]
        self.rng_vm_table_dev = tables.data_ptr() if tables is not None else None
        self.rng_vm_table_stride = tables.shape[1] if tables is not None else 0
        self._vm_tables = tables  # keeps the device memory alive as long as the params


class LikParams(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dim", C.c_int32), ("n_images", C.c_int32), ("cell_grid", C.c_int32),
        ("mu", C.c_double * PFB_MAXD), ("W", C.c_double * (PFB_MAXD * PFB_MAXD)),
        ("logc", C.c_double), ("kappa", C.c_double), ("sigma", C.c_double),
        ("images_dev", C.c_void_p), ("mu_batch_dev", C.c_void_p),
        ("out_format", C.c_int32), ("cell_span", C.c_int32),
    ]


_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32
# name -> (restype, argtypes); every symbol include/pfb.h declares
SIGNATURES = {
    "pfb_abi_version": (C.c_int, []),
    "pfb_last_error": (C.c_char_p, []),
    "pfb_workspace_bytes": (_i64, [_i64]),
    "pfb_workspace_bytes_batched": (_i64, [_i64, _i64]),
    "pfb_workspace_init": (C.c_int, [_vp, _i64, _vp]),
    "pfb_predict": (C.c_int, [C.POINTER(PredictParams), C.c_int, _i64, _vp, _vp, _vp, _vp]),
    "pfb_loglik": (C.c_int, [C.POINTER(LikParams), C.c_int, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_predict_loglik": (C.c_int, [C.POINTER(PredictParams), C.POINTER(LikParams), C.c_int, _i64, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_loglik_batched": (C.c_int, [C.POINTER(LikParams), C.c_int, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_predict_loglik_batched": (C.c_int, [C.POINTER(PredictParams), C.POINTER(LikParams), C.c_int, _i64, _i64, _vp,
                                             _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_scan_cdf_batched": (C.c_int, [_i32, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_resample_batched": (C.c_int, [C.c_int, _i32, _i64, _i64, _i64, _vp, _vp, C.c_uint64, C.c_uint64, _i32, _vp, _vp,
                                       _vp, _vp, _i64, _vp]),
    "pfb_moments_batched": (C.c_int, [_i32, C.c_int, _i32, _i64, _i64, _vp, _i32, _vp, _vp]),
    "pfb_normalize": (C.c_int, [_i64, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_scan_cdf": (C.c_int, [_i32, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "pfb_resample": (C.c_int, [C.c_int, _i32, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "pfb_resample_rng": (C.c_int, [C.c_int, _i32, _i64, _i64, _vp, C.c_uint64, C.c_uint64, _i32, _vp, _vp, _vp, _i32, _vp,
                                   _vp, _i64, _vp]),
    "pfb_resample_strided": (C.c_int, [C.c_int, _i32, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "pfb_resample_rng_strided": (C.c_int, [C.c_int, _i32, _i64, _i64, _vp, C.c_uint64, C.c_uint64, _vp, _i32, _vp, _vp, _i32,
                                           _vp, _vp, _i64, _vp]),
    "pfb_resample_msort": (C.c_int, [C.c_int, _i32, _i64, _i64, _vp, C.c_uint64, C.c_uint64, _vp, _i32, _vp, _vp, _i32, _vp,
                                     _vp, _i64, _vp]),
    "pfb_msort_uniforms": (C.c_int, [_i64, C.c_uint64, C.c_uint64, _vp, _vp, _i64, _vp]),
    "pfb_resample_sharded": (C.c_int, [C.c_int, _i32, _i64, _vp, C.c_double, C.c_double, C.c_double, _i64, _i64, _i64,
                                       _vp, _vp, _vp, _i32, _vp, _vp, _i64, _vp]),
    "pfb_resample_sharded_dev": (C.c_int, [C.c_int, _i32, _i64, _vp, _vp, _i32, _i32, C.c_double, C.POINTER(C.c_void_p), _vp, _i32,
                                           _vp, _vp, _vp, _i64, _vp]),
    "pfb_resample_sharded_pull": (C.c_int, [C.c_int, _i32, _i64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp, _i32, _i32,
                                            C.c_double, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "pfb_rng_uniform": (C.c_int, [_i64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "pfb_moments": (C.c_int, [_i32, C.c_int, _i32, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "pfb_wrap_2pi": (C.c_int, [C.c_int, _i64, _vp, _vp, _vp]),
    "pfb_vm_strip_table": (C.c_int, [C.c_double, _vp, _vp]),
    "pfb_vm_strip_thresholds": (C.c_int, [C.c_double, _vp, C.c_double, _vp]),
}

_LIB = None


class PfbError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libpfb.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise PfbError("building libpfb.so failed")
    return LIB_PATH


def lib():
    """The loaded library with typed entry points.  Raises if the CUDA library is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PfbError(
                f"{LIB_PATH} not found: the CUDA library is required (no fallback). "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C pyrecest_b200/csrc`."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if handle.pfb_abi_version() != ABI_VERSION:
            raise PfbError("libpfb ABI version mismatch")
        _LIB = handle
    return _LIB


_VM_HOST = {}
_VM_DEV = {}


def vm_strip_table(kappa: float):
    """(table (VM_STRIPS, 2) float64 [start, width], hat area A) for one kappa >= 1e-8; host code of libpfb"""
    import numpy as np

    kappa = float(kappa)
    if kappa not in _VM_HOST:
        table = np.zeros((VM_STRIPS, 2), dtype=np.float64)
        area = C.c_double()
        check(lib().pfb_vm_strip_table(kappa, table.ctypes.data, C.addressof(area)))
        _VM_HOST[kappa] = (table, area.value)
    return _VM_HOST[kappa]


def vm_tables_on_device(kappas: tuple, device: str):
    import numpy as np
    import torch

    key = (kappas, device)
    if key not in _VM_DEV:
        rows = []
        for k in kappas:  # strips, then the quick-accept thresholds (PFB_VM_TABLE_DOUBLES doubles per kappa)
            table, area = vm_strip_table(k)
            thr = np.zeros(VM_STRIPS, dtype=np.uint32)
            check(lib().pfb_vm_strip_thresholds(k, table.ctypes.data, area, thr.ctypes.data))
            rows.append(np.concatenate([table.reshape(-1), thr.view(np.float64)]))
        host = np.stack(rows)
        _VM_DEV[key] = torch.as_tensor(host, device=device).contiguous()
        if len(_VM_DEV) > 64:
            _VM_DEV.pop(next(iter(_VM_DEV)))
    return _VM_DEV[key]


def check(rc: int):
    if rc != 0:
        msg = lib().pfb_last_error().decode("utf-8", "replace")
        if rc == 4:
            raise NotImplementedError(msg)
        if rc == 1:
            raise ValueError(msg)
        raise PfbError(f"libpfb error {rc}: {msg}")
