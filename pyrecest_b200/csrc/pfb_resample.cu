// K(4b) multinomial / systematic resampling (search + gather + fused estimate), K(5) weighted moments,
// and the small host-side plumbing of the C ABI (errors, workspace).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "pfb_common.cuh"

namespace pfb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
  return PFB_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached[dev] = v;
  }
  return cached[dev];
}

int guide_shift() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PFB_GUIDE_SHIFT");
    v = e ? atoi(e) : kGuideShift;
    if (v < 0 || v > 20) v = kGuideShift;
  }
  return v;
}

int carve_workspace(void* ws, int64_t ws_bytes, int64_t n, Workspace* out) {
  if (!ws) { set_error("workspace pointer is NULL"); return PFB_ERR_WORKSPACE; }
  if (((uintptr_t)ws & 255) != 0) { set_error("workspace must be 256-byte aligned"); return PFB_ERR_WORKSPACE; }
  const int64_t need = pfb_workspace_bytes(n);
  if (ws_bytes < need) {
    set_error("workspace too small: %lld bytes given, %lld needed for n=%lld", (long long)ws_bytes,
              (long long)need, (long long)n);
    return PFB_ERR_WORKSPACE;
  }
  unsigned char* b = static_cast<unsigned char*>(ws);
  out->counters = reinterpret_cast<unsigned int*>(b);
  out->partials = reinterpret_cast<double*>(b + kWsCounters);
  out->tiles = b + ws_header();
  out->tile_bytes = ws_bytes - ws_header();
  out->guide = reinterpret_cast<int32_t*>(b + guide_offset(n));
  out->guide_buckets = guide_buckets(n);
  return PFB_OK;
}

// ---------------------------------------------------------------------------------------------
// resampling
// ---------------------------------------------------------------------------------------------
// idx = #{i : fl(c_i / total) <= u}.  fl(c/total) is monotone in c, so a binary search on the raw
// running sums with the division done per probe equals numpy's `cdf /= cdf[-1]; searchsorted(u,'right')`.
//
// A random binary search over an 800 MB CDF costs ~27 dependent probes, the last ~7 of them DRAM
// sectors.  A guide table removes most of it: G[b] = #{i : v_i < b/M} for M = 2^k buckets of [0, 1)
// (M ~ n/8, int32, sized to stay L2-resident).  For u in bucket b = floor(u M) (exact: M is a power
// of two) the answer lies in [G[b], G[b+1]], typically ~8 consecutive CDF entries.
__device__ int g_ldmode = 0;  // experiment switch (PFB_LDMODE): 0 ld.global.nc, 1 ld.global.cg, 2 ld.global.ca, 3 nc + L1::no_allocate
__device__ __forceinline__ double ld_rand_f64(const double* p) {
  double v;
  switch (g_ldmode) {
    case 1: asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p)); break;
    case 2: asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p)); break;
    case 3: asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); break;
    default: v = __ldg(p);
  }
  return v;
}
template <bool RIGHT>
__device__ __forceinline__ int64_t search_range(const double* __restrict__ cdf, int64_t lo, int64_t hi,
                                                double total, double q) {
  // RIGHT: first i in [lo, hi) with v_i > q (count of v_i <= q);  LEFT: first i with v_i >= q
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    const double v = __ddiv_rn(ld_rand_f64(cdf + mid), total);
    const bool go_right = RIGHT ? (v <= q) : (v < q);
    if (go_right) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// bucket-driven: thread b searches b/M over the whole CDF; consecutive threads probe neighbouring
// entries, so the probes are cache-line coalesced and the CDF is streamed from HBM about once.
__global__ void __launch_bounds__(kThreads) guide_kernel(int64_t n, const double* __restrict__ cdf,
                                                         const double* __restrict__ total_ptr,
                                                         int32_t* __restrict__ guide, int64_t M) {
  const double total = *total_ptr;
  const double inv = 1.0 / (double)M;  // exact, M = 2^k
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= M; b += stride) {
    const double q = (double)b * inv;
    guide[b] = (int32_t)search_range<false>(cdf, 0, n, total, q);
  }
}

template <typename T, int D, int EST>
__global__ void __launch_bounds__(kThreads) resample_kernel(int64_t n, int64_t n_out,
                                                            const double* __restrict__ cdf,
                                                            const double* __restrict__ u, int systematic,
                                                            const double* __restrict__ total_ptr,
                                                            const T* __restrict__ x_in, T* __restrict__ x_out,
                                                            int64_t* __restrict__ idx_out, double* est_out,
                                                            Workspace ws, const int32_t* __restrict__ guide,
                                                            int64_t M) {
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  const double total = *total_ptr;
  const double u0 = systematic ? u[0] : 0.0;
  const double dn = (double)n_out;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += stride) {
    const double uj = systematic ? __ddiv_rn(__dadd_rn((double)j, u0), dn) : __ldg(u + j);
    int64_t idx;
    if (uj >= 1.0) {
      idx = n - 1;
    } else {
      const int64_t b = (int64_t)(uj * (double)M);  // exact: M is a power of two
      idx = search_range<true>(cdf, (int64_t)__ldg(guide + b), (int64_t)__ldg(guide + b + 1), total, uj);
      if (idx > n - 1) idx = n - 1;
    }
    if (idx_out) idx_out[j] = idx;
    if (x_in != nullptr) {
      double r[D];
      if (sizeof(T) == 8 && g_ldmode != 0) {
#pragma unroll
        for (int c = 0; c < D; ++c) r[c] = ld_rand_f64(reinterpret_cast<const double*>(x_in) + idx * D + c);
      } else {
        load_rec<T, D>(x_in, idx, r);
      }
      store_rec<T, D>(x_out, j, r);
      if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] += r[c];
      } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double sn, cs;
          sincos(r[c], &sn, &cs);
          acc[2 * c] += cs;
          acc[2 * c + 1] += sn;
        }
      }
    }
  }
  if constexpr (EST != PFB_EST_NONE) {
    block_sum<NE>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < NE; ++k) ws.partials[(int64_t)blockIdx.x * kPartialDoubles + k] = acc[k];
    }
    if (block_is_last(ws.counters)) {
      double t[NE];
#pragma unroll
      for (int k = 0; k < NE; ++k) t[k] = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NE; ++k) t[k] += ws.partials[(int64_t)b * kPartialDoubles + k];
      }
      block_sum<NE>(t, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NE; ++k) est_out[k] = t[k] / dn;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weighted moments
// ---------------------------------------------------------------------------------------------
template <int D, int KIND>
struct MomentCount {
  static constexpr int value = (KIND == PFB_MOM_MEAN) ? D + 1
                               : (KIND == PFB_MOM_TRIG) ? 2 * D + 1
                                                        : D * (D + 1) / 2 + 1;
};

template <typename T, int D, int KIND>
__global__ void __launch_bounds__(kThreads) moments_kernel(int64_t n, const T* __restrict__ x,
                                                           const double* __restrict__ w, int order,
                                                           const double* __restrict__ center, double* out,
                                                           Workspace ws) {
  constexpr int NE = MomentCount<D, KIND>::value;
  static_assert(NE <= kPartialDoubles, "partials too small");
  __shared__ double red[NE * 32];
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  double ctr[D];
#pragma unroll
  for (int c = 0; c < D; ++c) ctr[c] = (KIND == PFB_MOM_CENTRAL && center) ? center[c] : 0.0;
  const double wu = 1.0 / (double)n;
  const double ord = (double)order;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double r[D];
    load_rec<T, D>(x, i, r);
    const double wi = w ? __ldg(w + i) : wu;
    if constexpr (KIND == PFB_MOM_MEAN) {
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] = fma(wi, r[c], acc[c]);
    } else if constexpr (KIND == PFB_MOM_TRIG) {
#pragma unroll
      for (int c = 0; c < D; ++c) {
        double sn, cs;
        sincos(ord * r[c], &sn, &cs);
        acc[2 * c] = fma(wi, cs, acc[2 * c]);
        acc[2 * c + 1] = fma(wi, sn, acc[2 * c + 1]);
      }
    } else {
      int k = 0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const double da = (r[a] - ctr[a]) * wi;
#pragma unroll
        for (int b = a; b < D; ++b) {
          acc[k] = fma(da, r[b] - ctr[b], acc[k]);
          ++k;
        }
      }
    }
    acc[NE - 1] += wi;
  }
  block_sum<NE>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NE; ++k) ws.partials[(int64_t)blockIdx.x * kPartialDoubles + k] = acc[k];
  }
  if (block_is_last(ws.counters)) {
    double t[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) t[k] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
      for (int k = 0; k < NE; ++k) t[k] += ws.partials[(int64_t)b * kPartialDoubles + k];
    }
    block_sum<NE>(t, red);
    if (threadIdx.x == 0) {
      if constexpr (KIND == PFB_MOM_MEAN || KIND == PFB_MOM_TRIG) {
#pragma unroll
        for (int k = 0; k < NE; ++k) out[k] = t[k];
      } else {
        const double div = (KIND == PFB_MOM_CENTRAL) ? t[NE - 1] : 1.0;
        int k = 0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
          for (int b = a; b < D; ++b) {
            out[a * D + b] = t[k] / div;
            out[b * D + a] = t[k] / div;
            ++k;
          }
        }
        out[D * D] = t[NE - 1];
      }
    }
  }
}

template <typename T, int D>
static int launch_resample(int64_t n, int64_t n_out, const double* cdf, const double* u, int systematic,
                           const double* stats, const void* x_in, void* x_out, int64_t* idx_out, int est_kind,
                           double* est_out, Workspace W, cudaStream_t st, const int32_t* guide, int64_t M) {
  const int grid = stream_grid(n_out);
  switch (est_kind) {
    case PFB_EST_NONE:
      resample_kernel<T, D, PFB_EST_NONE><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M);
      break;
    case PFB_EST_SUM:
      resample_kernel<T, D, PFB_EST_SUM><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M);
      break;
    case PFB_EST_TRIG:
      resample_kernel<T, D, PFB_EST_TRIG><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M);
      break;
    default:
      set_error("bad est_kind %d", est_kind);
      return PFB_ERR_INVALID;
  }
  return PFB_OK;
}

template <typename T, int D>
static int launch_moments(int kind, int64_t n, const void* x, const double* w, int order, const double* center,
                          double* out, Workspace W, cudaStream_t st) {
  const int grid = stream_grid(n, 2);
  switch (kind) {
    case PFB_MOM_MEAN: moments_kernel<T, D, PFB_MOM_MEAN><<<grid, kThreads, 0, st>>>(n, (const T*)x, w, order, center, out, W); break;
    case PFB_MOM_TRIG: moments_kernel<T, D, PFB_MOM_TRIG><<<grid, kThreads, 0, st>>>(n, (const T*)x, w, order, center, out, W); break;
    case PFB_MOM_SCATTER: moments_kernel<T, D, PFB_MOM_SCATTER><<<grid, kThreads, 0, st>>>(n, (const T*)x, w, order, center, out, W); break;
    case PFB_MOM_CENTRAL: moments_kernel<T, D, PFB_MOM_CENTRAL><<<grid, kThreads, 0, st>>>(n, (const T*)x, w, order, center, out, W); break;
    default: set_error("bad moment kind %d", kind); return PFB_ERR_INVALID;
  }
  return PFB_OK;
}

}  // namespace pfb

using namespace pfb;

extern "C" int pfb_abi_version(void) { return PFB_ABI_VERSION; }
extern "C" const char* pfb_last_error(void) { return g_err; }

extern "C" int64_t pfb_workspace_bytes(int64_t n) {
  if (n < 0) n = 0;
  int64_t b = guide_offset(n) + (guide_buckets(n) + 2) * (int64_t)sizeof(int32_t);
  return (b + 255) & ~(int64_t)255;
}

extern "C" int pfb_workspace_init(void* ws, int64_t ws_bytes, void* stream) {
  if (!ws || ws_bytes < pfb_workspace_bytes(0)) { set_error("workspace too small"); return PFB_ERR_WORKSPACE; }
  if (const char* e = getenv("PFB_LDMODE")) {
    int v = atoi(e);
    PFB_CUDA_OK(cudaMemcpyToSymbol(g_ldmode, &v, sizeof(int)));
  }
  if (const char* e = getenv("PFB_L2GRAN")) {
    PFB_CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(e)));
  }
  PFB_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)kWsCounters, (cudaStream_t)stream));
  return PFB_OK;
}

extern "C" int pfb_resample(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                            const double* u, int32_t systematic, const void* x_in, void* x_out,
                            int64_t* idx_out, int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes,
                            void* stream) {
  if (n <= 0 || n_out <= 0) return PFB_OK;
  if (!cdf || !u) { set_error("NULL cdf / uniforms"); return PFB_ERR_INVALID; }
  if ((x_in == nullptr) != (x_out == nullptr)) { set_error("x_in and x_out must both be given or both NULL"); return PFB_ERR_INVALID; }
  if (x_in && x_in == x_out) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE && (!est_out || !x_in)) { set_error("estimate requested without output / particles"); return PFB_ERR_INVALID; }
  if (n >= ((int64_t)1 << 31)) { set_error("more than 2^31-1 particles per device are not supported"); return PFB_ERR_UNSUPPORTED; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, n, &W);
  if (rc) return rc;
  // the raw total c_{n-1} is read on the device from the CDF itself
  cudaStream_t st = (cudaStream_t)stream;
  const double* stats_total = cdf + (n - 1);
  const int64_t M = W.guide_buckets;
  guide_kernel<<<stream_grid(M + 1), kThreads, 0, st>>>(n, cdf, stats_total, W.guide, M);
  PFB_LAUNCH_OK("guide_kernel");
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_resample<T, D>(n, n_out, cdf, u, systematic, stats_total, x_in, x_out, idx_out, est_kind, est_out, W, st, W.guide, M))));
  if (rc) return rc;
  PFB_LAUNCH_OK("resample_kernel");
  return PFB_OK;
}

extern "C" int pfb_moments(int32_t kind, pfb_dtype dtype, int32_t dim, int64_t n, const void* x, const double* w,
                           int32_t order, const double* center, double* out, void* ws, int64_t ws_bytes,
                           void* stream) {
  if (n <= 0) return PFB_OK;
  if (!x || !out) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_moments<T, D>(kind, n, x, w, order, center, out, W, st))));
  if (rc) return rc;
  PFB_LAUNCH_OK("moments_kernel");
  return PFB_OK;
}
