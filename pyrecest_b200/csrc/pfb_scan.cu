// K(3) max / log-sum-exp normalisation and K(4a) the CDF scan (fast and sequential-exact).
//
// Sequential-exact scan: reproduces numpy.cumsum(p) (c_i = fl(c_{i-1} + p_i), fp64, ties to
// even, p_i >= 0) bit for bit in ONE pass with a chained (decoupled look-back) scan:
//   * per-tile sums (emitted by the weighting kernel, or by tile_stats_kernel) are prefix-summed
//     by one CTA first; that approximate prefix bounds the running sum at every tile boundary to a
//     relative 4 n 2^-53, which tells most tiles with certainty which binade [2^e, 2^(e+1)) the
//     sequential running sum is in while it crosses them ("safe" tiles);
//   * inside one binade a sequential fp add is an INTEGER add in units of ulp = 2^(e-52):
//       S_i = S_{i-1} + k_i + [rem_i > half] + [rem_i == half and (S_{i-1} + k_i) odd]
//     so each element is a map parity(S_{i-1}) -> increment, stored as (inc_even, inc_odd); maps
//     compose associatively, (g o f)(pi) = f(pi) + g((pi + f(pi)) & 1), which is the operator of
//     the look-back; a safe tile publishes its composed map before it knows its prefix;
//   * tiles that may contain a binade crossing (about one per binade) or start at 0 are "hard":
//     they wait for the exact incoming value and add sequentially with real fp64 adds.
// SURVEY.md appendix B has the derivation; tests/test_scan_model.py holds a numpy model of the
// same tile logic and tests/test_gpu_resample.py checks the kernel against np.cumsum.
#include "pfb_common.cuh"

namespace pfb {

constexpr int kItems = kScanTile / kThreads;  // 8 consecutive elements per thread
static_assert(kItems * kThreads == kScanTile, "tile shape");

struct TileState {
  unsigned long long* flagB;  // low 2 bits: 0 none, 1 aggregate map, 2 inclusive; bits 8..: e + 2048
  long long* aggE;
  long long* aggO;
  double* inclB;
  double* tmax;      // per-tile max of the log-weights (or the shift the sums were taken with)
  double* tsum;      // per-tile sum of exp(logw - tmax)  (or of the weights)
  double* aprefix;   // ntiles + 1 approximate exclusive prefix sums, relative to the global shift
  unsigned int* ticket;
};

inline TileState make_tile_state(unsigned char* base, int64_t ntiles, unsigned int* ticket) {
  TileState s;
  s.flagB = reinterpret_cast<unsigned long long*>(tile_col(base, ntiles, 0));
  s.aggE = reinterpret_cast<long long*>(tile_col(base, ntiles, 1));
  s.aggO = reinterpret_cast<long long*>(tile_col(base, ntiles, 2));
  s.inclB = tile_col(base, ntiles, 3);
  s.tmax = tile_col(base, ntiles, 4);
  s.tsum = tile_col(base, ntiles, 5);
  s.aprefix = tile_col(base, ntiles, 6);  // ntiles + 1 entries: the last one uses the spare eighth column
  s.ticket = ticket;
  return s;
}

// ---- parity maps ---------------------------------------------------------------------------------
struct PMap {
  long long e, o;  // increment when the incoming integer sum is even / odd
};
// Increments saturate at 2^60: anything >= 2^53 only ever means "this crosses into the next binade".
constexpr long long kSat = 0x1000000000000000LL;
__device__ __forceinline__ long long sat(long long v) { return v < kSat ? v : kSat; }
__device__ __forceinline__ PMap pmap_then(PMap f, PMap g) {  // apply f first, then g
  PMap h;
  h.e = sat(f.e + ((f.e & 1LL) ? g.o : g.e));
  h.o = sat(f.o + (((f.o + 1LL) & 1LL) ? g.o : g.e));
  return h;
}
__device__ __forceinline__ long long pmap_apply(PMap f, long long S) { return sat(S + ((S & 1LL) ? f.o : f.e)); }

// map of one addend p >= 0 in the binade whose ulp is 2^ue  (ue = e - 52)
__device__ __forceinline__ PMap pmap_of(double p, int ue) {
  const long long bits = __double_as_longlong(p);
  const int ef = (int)((bits >> 52) & 0x7ff);
  long long M = bits & 0x000fffffffffffffLL;
  int q;
  if (ef == 0) { q = -1074; } else { M |= 0x0010000000000000LL; q = ef - 1075; }
  PMap r;
  if (M == 0) { r.e = 0; r.o = 0; return r; }
  const int sh = ue - q;
  if (sh <= 0) {
    long long k = (sh > -7) ? (M << (-sh)) : kSat;  // M < 2^53: up to 2^59; beyond that certainly a crossing
    r.e = k; r.o = k;
  } else if (sh >= 63) {
    r.e = 0; r.o = 0;
  } else {
    const long long k = M >> sh;
    const long long rem = M & ((1LL << sh) - 1LL);
    const long long half = 1LL << (sh - 1);
    const long long up = rem > half ? 1LL : 0LL;
    const bool tie = rem == half;
    r.e = k + up + ((tie && (k & 1LL)) ? 1LL : 0LL);
    r.o = k + up + ((tie && !(k & 1LL)) ? 1LL : 0LL);
  }
  return r;
}

// The same addend as a plain integer increment; *tie is set when the addend sits exactly half way between two
// representable sums (then the increment depends on the parity of the running sum and the map form is needed).
// For weights that are not dyadic rationals of the running sum this essentially never happens.
__device__ __forceinline__ long long inc_of(double p, int ue, bool* tie) {
  const long long bits = __double_as_longlong(p);
  const int ef = (int)((bits >> 52) & 0x7ff);
  long long M = bits & 0x000fffffffffffffLL;
  int q;
  if (ef == 0) { q = -1074; } else { M |= 0x0010000000000000LL; q = ef - 1075; }
  const int sh = ue - q;
  if (M == 0 || sh >= 63) return 0;
  if (sh <= 0) return (sh > -7) ? (M << (-sh)) : kSat;
  const long long rem = M & ((1LL << sh) - 1LL);
  const long long half = 1LL << (sh - 1);
  if (rem == half) *tie = true;
  return (M >> sh) + (rem > half ? 1LL : 0LL);
}

// inc_of in floating point, for addends of a SAFE tile (every p_i <= the tile sum < 2^(e+1), so p / ulp < 2^54) whose
// binade is not at the bottom of the range (ue >= -1000: 2^-ue is a normal double and p 2^-ue is exact):
// k = p / ulp exactly, inc = floor(k) + [frac > 1/2], tie iff frac == 1/2 -- seven instructions instead of ~40 of
// 64-bit integer work.
__device__ __forceinline__ long long inc_of_fast(double p, double inv_ulp, bool* tie) {
  const double k = p * inv_ulp;
  const double f = floor(k);
  const double d = k - f;  // exact
  if (d == 0.5) *tie = true;
  return __double2ll_rz(f) + (d > 0.5 ? 1LL : 0LL);
}
constexpr int kFastUlpMin = -1000;

__device__ __forceinline__ double ulp_value(int ue) {  // 2^ue, ue >= -1074
  long long b = (ue >= -1022) ? ((long long)(ue + 1023) << 52) : (1LL << (ue + 1074));
  return __longlong_as_double(b);
}
// integer sum S with c = S * 2^(e-52), for c in [2^e, 2^(e+1)) or subnormal with e == -1022
__device__ __forceinline__ long long int_of(double c, int e, bool* ok) {
  const long long bits = __double_as_longlong(c);
  const int ef = (int)((bits >> 52) & 0x7ff);
  const long long mant = bits & 0x000fffffffffffffLL;
  if (ef == 0) { *ok = (e == -1022); return mant; }
  *ok = (ef - 1023 == e);
  return mant | 0x0010000000000000LL;
}

// ---- K(3) normalise ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lse_reduce_kernel(int64_t n, const double* __restrict__ logw,
                                                              double* stats, Workspace ws) {
  __shared__ double red[3 * 32];
  double m = -INFINITY, s = 0.0, q = 0.0, nanf = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double l = logw[i];
    if (l != l) { nanf = 1.0; continue; }
    if (l > m) {
      double sc = exp(m - l);  // m = -inf -> 0
      s = s * sc + 1.0;
      q = q * (sc * sc) + 1.0;
      m = l;
    } else if (l > -INFINITY) {
      double e = exp(l - m);
      s += e;
      q += e * e;
    }
  }
  // block combine: common max, rescale, sum
  double bm = block_max(m, red);
  double sc = (m == -INFINITY) ? 0.0 : exp(m - bm);
  double v[3] = {s * sc, q * sc * sc, nanf};
  block_sum<3>(v, red);
  if (threadIdx.x == 0) {
    ws.partials[4 * blockIdx.x + 0] = bm;
    ws.partials[4 * blockIdx.x + 1] = v[0];
    ws.partials[4 * blockIdx.x + 2] = v[1];
    ws.partials[4 * blockIdx.x + 3] = v[2];
  }
  if (block_is_last(ws.counters)) {
    double M = -INFINITY;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) M = fmax(M, ws.partials[4 * b]);
    M = block_max(M, red);
    double t[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      double pm = ws.partials[4 * b];
      double f = (pm == -INFINITY) ? 0.0 : exp(pm - M);
      t[0] += ws.partials[4 * b + 1] * f;
      t[1] += ws.partials[4 * b + 2] * f * f;
      t[2] += ws.partials[4 * b + 3];
    }
    block_sum<3>(t, red);
    if (threadIdx.x == 0) {
      stats[PFB_STAT_MAX] = M;
      stats[PFB_STAT_SUMEXP] = t[0];
      stats[PFB_STAT_SUMSQ] = t[1];
      double flags = 0.0;
      if (t[2] > 0.0) flags += 1.0;
      if (M == -INFINITY || !(t[0] > 0.0)) flags += 2.0;
      stats[PFB_STAT_FLAGS] = flags;
    }
  }
}

__global__ void __launch_bounds__(kThreads) weights_kernel(int64_t n, const double* __restrict__ logw,
                                                           double* __restrict__ w_out,
                                                           const double* __restrict__ stats) {
  const double M = stats[PFB_STAT_MAX], S = stats[PFB_STAT_SUMEXP];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    w_out[i] = exp(logw[i] - M) / S;
}

// ---- K(4a) scan ------------------------------------------------------------------------------------------
// block-wide exclusive scan of one double per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ double block_excl_scan(double v, double* smem /*33*/, double* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    double w = (lane < (kThreads / 32)) ? smem[lane] : 0.0;
    double wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < (kThreads / 32)) smem[lane] = wi - w;  // exclusive warp offsets
    if (lane == (kThreads / 32) - 1) smem[32] = wi;
  }
  __syncthreads();
  *total = smem[32];
  double exc = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) exc = 0.0;
  return smem[warp] + exc;
}

// block-wide exclusive scan of one int64 per thread
__device__ __forceinline__ long long block_excl_scan_ll(long long v, long long* smem /*34*/, long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    long long w = (lane < (kThreads / 32)) ? smem[lane] : 0;
    long long wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < (kThreads / 32)) smem[lane] = wi - w;
    if (lane == (kThreads / 32) - 1) smem[33] = wi;
  }
  __syncthreads();
  *total = smem[33];
  return smem[warp] + (inc - v);
}

// block-wide exclusive scan of parity maps (ordered composition)
__device__ __forceinline__ PMap block_excl_scan_pmap(PMap v, long long* smem /* 2*33 */, PMap* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PMap inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    PMap t;
    t.e = __shfl_up_sync(0xffffffffu, inc.e, o);
    t.o = __shfl_up_sync(0xffffffffu, inc.o, o);
    if (lane >= o) inc = pmap_then(t, inc);
  }
  // exclusive within warp: inclusive of the previous lane
  PMap exc;
  exc.e = __shfl_up_sync(0xffffffffu, inc.e, 1);
  exc.o = __shfl_up_sync(0xffffffffu, inc.o, 1);
  if (lane == 0) { exc.e = 0; exc.o = 0; }
  __syncthreads();
  if (lane == 31) { smem[2 * warp] = inc.e; smem[2 * warp + 1] = inc.o; }
  __syncthreads();
  if (threadIdx.x == 0) {
    PMap run; run.e = 0; run.o = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      PMap cur; cur.e = smem[2 * w]; cur.o = smem[2 * w + 1];
      smem[2 * w] = run.e; smem[2 * w + 1] = run.o;  // exclusive warp prefix
      run = pmap_then(run, cur);
    }
    smem[64] = run.e; smem[65] = run.o;
  }
  __syncthreads();
  PMap wp; wp.e = smem[2 * warp]; wp.o = smem[2 * warp + 1];
  total->e = smem[64]; total->o = smem[65];
  return pmap_then(wp, exc);
}

__device__ __forceinline__ int block_min_int(int v, int* smem /* 33 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  int x = (lane < (kThreads / 32)) ? smem[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

// ---- tile sums and their prefix ------------------------------------------------------------------
// input of the scan: plain weights, log-weights (addend exp(l - shift)), or the slice-scaled linear weights of
// PFB_WEIGHTS_SCALED (addend p 2^(e_slice - E): an exact scaling, no exp)
enum { kInW = 0, kInLog = 1, kInScaled = 2 };
__device__ __forceinline__ double pow2_of(double d) {  // 2^d for integer-valued d (0 below the subnormals)
  if (!(d > -1100.0)) return 0.0;
  return scalbn(1.0, (int)d);
}

template <bool FROM_LOG>
__global__ void __launch_bounds__(kThreads) tile_stats_kernel(int64_t n, const double* __restrict__ in,
                                                              const double* __restrict__ stats, TileState ts,
                                                              int64_t ntiles) {
  __shared__ double red[32];
  const double shift = FROM_LOG ? stats[PFB_STAT_MAX] : 0.0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t base = tile * kScanTile;
    double v[1] = {0.0};
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
      const int64_t i = base + k * kThreads + threadIdx.x;
      if (i < n) v[0] += FROM_LOG ? exp(__ldg(in + i) - shift) : __ldg(in + i);
    }
    block_sum<1>(v, red);
    if (threadIdx.x == 0) { ts.tmax[tile] = shift; ts.tsum[tile] = v[0]; }
    __syncthreads();
  }
}

// (max, sum exp) of a tile from the 8 per-warp partials the weighting kernels left in the workspace
__global__ void __launch_bounds__(kThreads) tile_combine_kernel(TileState ts, const double* __restrict__ wpart,
                                                                int64_t ntiles) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride) {
    double m = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = fmax(m, wpart[t * 16 + 2 * w]);
    double sacc = 0.0;
    if (m > -INFINITY) {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const double mw = wpart[t * 16 + 2 * w];
        if (mw > -INFINITY) sacc += wpart[t * 16 + 2 * w + 1] * exp(mw - m);
      }
    }
    ts.tmax[t] = m;
    ts.tsum[t] = sacc;
  }
}

// the same for PFB_WEIGHTS_SCALED partials (slice exponent, sum p): tile exponent = largest slice exponent
__global__ void __launch_bounds__(kThreads) tile_combine_scaled_kernel(TileState ts, const double* __restrict__ wpart,
                                                                       int64_t ntiles) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride) {
    double m = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = fmax(m, wpart[t * 16 + 2 * w]);
    double sacc = 0.0;
    if (m > -INFINITY) {
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const double ew = wpart[t * 16 + 2 * w];
        if (ew > -INFINITY) sacc += wpart[t * 16 + 2 * w + 1] * pow2_of(ew - m);
      }
    }
    ts.tmax[t] = m;
    ts.tsum[t] = sacc;
  }
}

// Segmented variant for batched runs: one warp per run.  The shift of a run is the max over its tiles.
__global__ void __launch_bounds__(kThreads) tile_prefix_runs_kernel(TileState ts, int64_t n_runs, int tpr,
                                                                    int from_log) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t run = warp_global; run < n_runs; run += nwarps) {
    const int64_t t0 = run * tpr;
    double M = 0.0;
    if (from_log) {
      M = -INFINITY;
      for (int t = lane; t < tpr; t += 32) M = fmax(M, ts.tmax[t0 + t]);
      M = warp_max(M);
    }
    double carry = 0.0;
    for (int base = 0; base < tpr; base += 32) {
      const int t = base + lane;
      double v = 0.0;
      if (t < tpr) {
        const double tm = ts.tmax[t0 + t];
        const double sc = from_log ? ((tm == -INFINITY) ? 0.0 : (from_log == kInScaled ? pow2_of(tm - M) : exp(tm - M))) : 1.0;
        v = ts.tsum[t0 + t] * sc;
      }
      double inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
      }
      if (t < tpr) {
        ts.aprefix[t0 + t] = carry + (inc - v);
        ts.tsum[t0 + t] = v;
        ts.tmax[t0 + t] = M;
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
}

// one CTA: aprefix[t] = sum_{s<t} tsum[s] exp(tmax[s] - M), t = 0 .. ntiles - 1
__global__ void __launch_bounds__(1024) tile_prefix_kernel(TileState ts, int64_t ntiles,
                                                           const double* __restrict__ max_ptr, int scaled) {
  __shared__ double s_warp[33];
  __shared__ double s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double M = max_ptr ? *max_ptr : 0.0;
  if (scaled) {  // the common exponent E = largest tile exponent
    double m = -INFINITY;
    for (int64_t t = threadIdx.x; t < ntiles; t += 1024) m = fmax(m, ts.tmax[t]);
    m = warp_max(m);
    if (lane == 0) s_warp[warp] = m;
    __syncthreads();
    m = s_warp[lane];
    M = warp_max(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) s_carry = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < ntiles; base += 1024) {
    const int64_t t = base + threadIdx.x;
    double v = 0.0;
    if (t < ntiles) {
      const double tm = ts.tmax[t];
      const double sc = scaled ? ((tm == -INFINITY) ? 0.0 : pow2_of(tm - M)) : (max_ptr ? ((tm == -INFINITY) ? 0.0 : exp(tm - M)) : 1.0);
      v = ts.tsum[t] * sc;
    }
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double x = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += x;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      double w = s_warp[lane];
      double wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double x = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += x;
      }
      s_warp[lane] = wi - w;
      if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    const double carry = s_carry;
    const double excl = carry + s_warp[warp] + (inc - v);
    if (t < ntiles) {
      ts.aprefix[t] = excl;
      ts.tsum[t] = v;    // tile sum relative to the shift the scan will use
      ts.tmax[t] = M;    // ... which is that shift
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_warp[32];
    __syncthreads();
  }
}

// The same prefix for long inputs, three small launches instead of one CTA walking all tiles (88 us of an otherwise
// idle GPU at 49 k tiles): chunk maxima of the tile exponents (scaled input only) -> prefix inside every chunk of 1024
// tiles + chunk totals -> carries.  `part`: 2 * nchunks doubles of the workspace's reduction partials.
__global__ void __launch_bounds__(1024) tile_prefix_max_kernel(TileState ts, int64_t ntiles, double* __restrict__ part) {
  __shared__ double s_warp[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  double m = warp_max(t < ntiles ? ts.tmax[t] : -INFINITY);
  if (lane == 0) s_warp[warp] = m;
  __syncthreads();
  m = warp_max(s_warp[lane]);
  if (threadIdx.x == 0) part[blockIdx.x] = m;
}
__global__ void __launch_bounds__(1024) tile_prefix_local_kernel(TileState ts, int64_t ntiles,
                                                                 const double* __restrict__ max_ptr, int scaled, int nchunks,
                                                                 double* __restrict__ part) {
  __shared__ double s_warp[33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double M = max_ptr ? *max_ptr : 0.0;
  if (scaled) {  // the common exponent E = largest tile exponent = largest chunk maximum
    double m = -INFINITY;
    for (int c = threadIdx.x; c < nchunks; c += 1024) m = fmax(m, part[c]);
    m = warp_max(m);
    if (lane == 0) s_warp[warp] = m;
    __syncthreads();
    M = warp_max(s_warp[lane]);
    __syncthreads();
  }
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  double v = 0.0;
  if (t < ntiles) {
    const double tm = ts.tmax[t];
    const double sc = scaled ? ((tm == -INFINITY) ? 0.0 : pow2_of(tm - M)) : (max_ptr ? ((tm == -INFINITY) ? 0.0 : exp(tm - M)) : 1.0);
    v = ts.tsum[t] * sc;
  }
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double x = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const double w = s_warp[lane];
    double wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double x = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += x;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) s_warp[32] = wi;
  }
  __syncthreads();
  if (t < ntiles) {
    ts.aprefix[t] = s_warp[warp] + (inc - v);  // (inside the chunk; the carry is added by the next launch)
    ts.tsum[t] = v;
    ts.tmax[t] = M;
  }
  if (threadIdx.x == 0) part[nchunks + blockIdx.x] = s_warp[32];
}
__global__ void __launch_bounds__(1024) tile_prefix_carry_kernel(TileState ts, int64_t ntiles, int nchunks,
                                                                 const double* __restrict__ part) {
  __shared__ double s_carry;
  if (threadIdx.x == 0) {  // chunk totals added in chunk order (<= 1024 of them)
    double carry = 0.0;
    for (int c = 0; c < (int)blockIdx.x; ++c) carry += part[nchunks + c];
    s_carry = carry;
  }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (t < ntiles && blockIdx.x > 0) ts.aprefix[t] = s_carry + ts.aprefix[t];
}

// ---- the scan proper: three kernels, no inter-CTA waiting ---------------------------------------------
//   scan_maps_kernel   every tile, in parallel: classify (safe binade e / hard) from the approximate prefix and,
//                      for safe tiles, reduce the tile to its composed parity map
//   scan_chain_kernel  one CTA per run: walk the tiles in order carrying the EXACT running value; stretches of
//                      safe tiles are an ordered scan of their maps (256 threads), hard tiles (one per binade)
//                      are loaded and walked element-wise; leaves cstart[tile] and writes the hard tiles' CDF
//   scan_final_kernel  every safe tile, in parallel: exact start value + in-tile map scan -> CDF
// (A single-pass chained look-back version measured 2.2 ms at 1e8 elements: all tiles of a wave publish their
//  maps together and the look-back has to walk a whole wave back; this pipeline reads the weights twice instead.)
constexpr long long kHardTile = -0x7fffffffLL;  // classification value of a hard tile (otherwise: the binade e)
constexpr long long kTieBit = 1LL << 40;          // OR-ed into the binade of a safe tile that contains a tie
__device__ __forceinline__ bool cls_has_tie(long long c) { return c > (kTieBit >> 1); }  // binades are |e| < 1100
__device__ __forceinline__ long long cls_binade(long long c) { return cls_has_tie(c) ? c - kTieBit : c; }

template <int IN>
__device__ __forceinline__ void load_tile(const double* __restrict__ in, int64_t n, int64_t tile, double shift,
                                          double (&p)[kItems], const double* __restrict__ wpart = nullptr) {
  const int64_t base = tile * kScanTile + (int64_t)threadIdx.x * kItems;
  if (base + kItems <= n) {
    const double2* q = reinterpret_cast<const double2*>(in + base);
#pragma unroll
    for (int k = 0; k < kItems / 2; ++k) {
      double2 v = __ldg(q + k);
      p[2 * k] = v.x; p[2 * k + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k) p[k] = (base + k < n) ? in[base + k] : (IN == kInLog ? -INFINITY : 0.0);
  }
  if constexpr (IN == kInLog) {
#pragma unroll
    for (int k = 0; k < kItems; ++k) p[k] = exp(p[k] - shift);  // exp(-inf) = 0
  } else if constexpr (IN == kInScaled) {
    // the 8 elements of a thread lie in one 256-particle slice (warp w of the weighting kernel): one power of two
    const double ew = wpart[tile * 16 + 2 * (threadIdx.x >> 5)];
    const double f = (ew > -INFINITY) ? pow2_of(ew - shift) : 0.0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) p[k] *= f;
  }
}

__device__ __forceinline__ void store_tile(double* __restrict__ cdf, int64_t n, int64_t tile, const double (&c)[kItems]) {
  const int64_t base = tile * kScanTile + (int64_t)threadIdx.x * kItems;
  if (base + kItems <= n) {
    double2* q = reinterpret_cast<double2*>(cdf + base);
#pragma unroll
    for (int k = 0; k < kItems / 2; ++k) q[k] = make_double2(c[2 * k], c[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (base + k < n) cdf[base + k] = c[k];
  }
}

// safe (returns the binade) or hard: every running sum inside the tile lies in [a0 (1-m), (a0 + s)(1+m)]
__device__ __forceinline__ long long classify_tile(const TileState& ts, int64_t tile, int64_t lt, double margin) {
  if (lt == 0) return kHardTile;  // the first tile of a run starts from 0
  const double a0 = ts.aprefix[tile];
  const double lo = a0 * (1.0 - margin);
  const double hi = (a0 + ts.tsum[tile]) * (1.0 + margin);
  if (!(lo >= 2.2250738585072014e-308) || !(hi < 1.0e300)) return kHardTile;
  const int elo = (int)((__double_as_longlong(lo) >> 52) & 0x7ff) - 1023;
  const int ehi = (int)((__double_as_longlong(hi) >> 52) & 0x7ff) - 1023;
  return elo == ehi ? (long long)elo : kHardTile;
}

template <int IN>
__global__ void __launch_bounds__(kThreads) scan_maps_kernel(int64_t n, const double* __restrict__ in, TileState ts,
                                                             int64_t ntiles, int64_t tpr, double margin,
                                                             const double* __restrict__ wpart) {
  __shared__ long long s_ll[72];
  long long* cls = reinterpret_cast<long long*>(ts.flagB);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long e = classify_tile(ts, tile, tile % tpr, margin);
    if (threadIdx.x == 0) cls[tile] = e;
    if (e == kHardTile) continue;
    double p[kItems];
    load_tile<IN>(in, n, tile, IN != kInW ? ts.tmax[tile] : 0.0, p, wpart);
    const int ue = (int)e - 52;
    bool tie = false;
    long long inc = 0;
    if (ue >= kFastUlpMin) {  // (block-uniform) floating-point form of the increments
      const double inv_ulp = ulp_value(-ue);
#pragma unroll
      for (int k = 0; k < kItems; ++k) inc += inc_of_fast(p[k], inv_ulp, &tie);
    } else {
#pragma unroll
      for (int k = 0; k < kItems; ++k) inc += inc_of(p[k], ue, &tie);
    }
    if (!__syncthreads_or(tie ? 1 : 0)) {  // no tie in the whole tile: the tile is a plain integer sum
      long long tot;
      block_excl_scan_ll(inc, s_ll, &tot);
      if (threadIdx.x == 0) { ts.aggE[tile] = tot; ts.aggO[tile] = tot; }
    } else {
      PMap f = pmap_of(p[0], ue);
#pragma unroll
      for (int k = 1; k < kItems; ++k) f = pmap_then(f, pmap_of(p[k], ue));
      PMap tot;
      block_excl_scan_pmap(f, s_ll, &tot);
      if (threadIdx.x == 0) { ts.aggE[tile] = tot.e; ts.aggO[tile] = tot.o; cls[tile] = e + kTieBit; }
    }
    __syncthreads();
  }
}

// walks the binades inside one (hard) tile starting from the exact value `cur`; fills c[] and returns the
// exact value at the end of the tile (identical in every thread)
__device__ __forceinline__ double hard_tile_walk(const double (&p)[kItems], int cnt, double cur, double (&c)[kItems],
                                                 long long* s_ll, int* s_int, double* s_cur, bool lead) {
  const int tb = threadIdx.x * kItems;
  int pos = 0;
  if (lead) {
    // (block-uniform) first tile of a run: the running sum starts at 0 and doubles every few elements, a binade --
    // one CTA-wide pass of the loop below -- each.  The first 32 * kItems elements sit in warp 0's registers: it walks
    // them with plain sequential adds, lane after lane (256 dependent adds, ~1.5 us), and the binade loop takes over
    // where the crossings have become sparse.
    if (threadIdx.x < 32) {
      double v = cur;
      for (int l = 0; l < 32; ++l) {
        if ((int)threadIdx.x == l) {
#pragma unroll
          for (int k = 0; k < kItems; ++k)
            if (tb + k < cnt) { v = __dadd_rn(v, p[k]); c[k] = v; }
        }
        v = __shfl_sync(0xffffffffu, v, l);
      }
      if (threadIdx.x == 0) *s_cur = v;
    }
    __syncthreads();
    cur = *s_cur;
    pos = 32 * kItems;
    __syncthreads();
    if (pos >= cnt) return cur;
  }
  while (true) {
    if (cur == 0.0) {  // leading zeros: 0 + p = p exactly
      int mine = 0x7fffffff;
#pragma unroll
      for (int k = kItems - 1; k >= 0; --k)
        if (tb + k >= pos && tb + k < cnt && p[k] != 0.0) mine = tb + k;
      const int first = block_min_int(mine, s_int);
#pragma unroll
      for (int k = 0; k < kItems; ++k)
        if (tb + k >= pos && tb + k < first) c[k] = 0.0;
      if (first >= cnt) break;
      if (first >= tb && first < tb + kItems) {
#pragma unroll
        for (int k = 0; k < kItems; ++k)
          if (tb + k == first) { c[k] = p[k]; *s_cur = p[k]; }
      }
      __syncthreads();
      cur = *s_cur;
      pos = first + 1;
      __syncthreads();
      continue;
    }
    int e = (int)((__double_as_longlong(cur) >> 52) & 0x7ff) - 1023;
    if (e < -1022) e = -1022;
    const int ue = e - 52;
    const double u = ulp_value(ue);
    bool ok;
    const long long S0 = int_of(cur, e, &ok);
    int mine = 0x7fffffff;
    long long S_before = 0, S_end;
    // Floating-point form of the increments (as in scan_maps_kernel) when the binade is not at the bottom of the range
    // and no remaining addend is a tie: the pass is then a plain int64 scan -- a third of the cost of the parity-map
    // form.  Increments are clamped at 2^54 (any value >= 2^53 only ever means "this crosses into the next binade",
    // and nothing before the first crossing is clamped), so that the block-wide sums cannot overflow.
    constexpr long long kClamp = 0x0040000000000000LL;
    bool tie = false;
    long long incs[kItems];
    const bool fast = ue >= kFastUlpMin;
    if (fast) {
      const double inv_ulp = ulp_value(-ue);
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        incs[k] = 0;
        if (tb + k >= pos && tb + k < cnt) {
          const double kk = p[k] * inv_ulp;
          incs[k] = (kk < 1.8014398509481984e16) ? inc_of_fast(p[k], inv_ulp, &tie) : kClamp;  // (2^54)
        }
      }
    }
    if (fast && !__syncthreads_or(tie ? 1 : 0)) {
      long long tsum = 0;
#pragma unroll
      for (int k = 0; k < kItems; ++k) { tsum += incs[k]; tsum = tsum < kClamp ? tsum : kClamp; }
      long long tot;
      long long S = S0 + block_excl_scan_ll(tsum, s_ll, &tot);
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        if (tb + k >= pos && tb + k < cnt) {
          const long long Sp = S;
          S += incs[k];
          if (S >= 0x0020000000000000LL) {
            if (mine == 0x7fffffff) { mine = tb + k; S_before = Sp; }
            S = kClamp;  // (stays a crossing for the rest of the thread's elements, without overflow)
          } else {
            c[k] = (double)S * u;
          }
        }
      }
      S_end = S0 + tot;
    } else {
      PMap f; f.e = 0; f.o = 0;
#pragma unroll
      for (int k = 0; k < kItems; ++k)
        if (tb + k >= pos && tb + k < cnt) f = pmap_then(f, pmap_of(p[k], ue));
      PMap tot;
      const PMap g = block_excl_scan_pmap(f, s_ll, &tot);
      long long S = pmap_apply(g, S0);
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        if (tb + k >= pos && tb + k < cnt) {
          const long long Sp = S;
          S = pmap_apply(pmap_of(p[k], ue), S);
          if (S >= 0x0020000000000000LL) {
            if (mine == 0x7fffffff) { mine = tb + k; S_before = Sp; }
          } else {
            c[k] = (double)S * u;
          }
        }
      }
      S_end = pmap_apply(tot, S0);
    }
    const int first = block_min_int(mine, s_int);
    if (first >= cnt) {  // no crossing left: the tile ends inside this binade
      cur = (double)S_end * u;
      break;
    }
    if (first == mine) {
#pragma unroll
      for (int k = 0; k < kItems; ++k)
        if (tb + k == first) {
          const double cx = __dadd_rn((double)S_before * u, p[k]);
          c[k] = cx;
          *s_cur = cx;
        }
    }
    __syncthreads();
    cur = *s_cur;
    pos = first + 1;
    __syncthreads();
  }
  return cur;
}

template <int IN>
__global__ void __launch_bounds__(kThreads) scan_chain_kernel(int64_t n, const double* __restrict__ in,
                                                              double* __restrict__ cdf, double* stats, TileState ts,
                                                              int64_t n_runs, int64_t tpr, unsigned int* err,
                                                              const double* __restrict__ wpart) {
  __shared__ long long s_ll[72];
  __shared__ int s_int[40];
  __shared__ double s_cur;
  const long long* cls = reinterpret_cast<const long long*>(ts.flagB);
  double* cstart = ts.inclB;
  for (int64_t run = blockIdx.x; run < n_runs; run += gridDim.x) {
    const int64_t t_begin = run * tpr, t_end = t_begin + tpr;
    double cur = 0.0;  // exact running value entering tile t
    int64_t t = t_begin;
    while (t < t_end) {
      const long long e = (cls[t] == kHardTile) ? kHardTile : cls_binade(cls[t]);
      if (e == kHardTile) {
        double p[kItems], c[kItems];
        load_tile<IN>(in, n, t, IN != kInW ? ts.tmax[t] : 0.0, p, wpart);
        const int64_t left = n - t * kScanTile;
        const int cnt = (int)(left < kScanTile ? left : kScanTile);
        cur = hard_tile_walk(p, cnt, cur, c, s_ll, s_int, &s_cur, t == t_begin);
        store_tile(cdf, n, t, c);
        if (t * kScanTile + cnt == n && threadIdx.x == 0) stats[PFB_STAT_TOTAL] = cur;
        ++t;
        __syncthreads();
        continue;
      }
      // stretch of safe tiles [t, t2): all in binade e (a tile that could leave it would be hard)
      int64_t t2 = t;
      while (true) {  // first hard tile at or after t, kProbe * kThreads tiles per probe (loads issued together)
        constexpr int kProbe = 8;
        long long cv[kProbe];
#pragma unroll
        for (int q = 0; q < kProbe; ++q) {
          const int64_t cand = t2 + q * kThreads + threadIdx.x;
          cv[q] = (cand >= t_end) ? kHardTile : cls[cand];
        }
        int hit = 0x7fffffff;
#pragma unroll
        for (int q = kProbe - 1; q >= 0; --q)
          if (cv[q] == kHardTile) hit = q * kThreads + (int)threadIdx.x;
        const int first = block_min_int(hit, s_int);
        if (first != 0x7fffffff) { t2 += first; break; }
        t2 += kProbe * kThreads;
      }
      const int64_t L = t2 - t;
      const int64_t chunk = (L + kThreads - 1) / kThreads;
      const int64_t c0 = t + (int64_t)threadIdx.x * chunk;
      const int64_t c1 = (c0 + chunk < t2) ? c0 + chunk : t2;
      PMap f; f.e = 0; f.o = 0;
      constexpr int kAhead = 4;  // the per-thread walk is a chain of dependent compositions: fetch ahead of it
      bool binade_ok = true;
      for (int64_t j = c0; j < c1; j += kAhead) {
        long long ce[kAhead], ae[kAhead], ao[kAhead];
#pragma unroll
        for (int q = 0; q < kAhead; ++q) {
          const bool in = j + q < c1;
          ce[q] = in ? cls[j + q] : 0;
          ae[q] = in ? ts.aggE[j + q] : 0;
          ao[q] = in ? ts.aggO[j + q] : 0;
        }
#pragma unroll
        for (int q = 0; q < kAhead; ++q) {
          if (j + q < c1) {
            binade_ok = binade_ok && cls_binade(ce[q]) == e;
            PMap m; m.e = ae[q]; m.o = ao[q];
            f = pmap_then(f, m);
          }
        }
      }
      if (!binade_ok) atomicOr(err, 4u);
      PMap tot;
      const PMap g = block_excl_scan_pmap(f, s_ll, &tot);
      bool ok;
      const long long S0 = int_of(cur, (int)e, &ok);
      if (!ok) atomicOr(err, 1u);
      const double u = ulp_value((int)e - 52);
      long long S = pmap_apply(g, S0);
      for (int64_t j = c0; j < c1; j += kAhead) {
        long long ae[kAhead], ao[kAhead];
#pragma unroll
        for (int q = 0; q < kAhead; ++q) {
          const bool in = j + q < c1;
          ae[q] = in ? ts.aggE[j + q] : 0;
          ao[q] = in ? ts.aggO[j + q] : 0;
        }
#pragma unroll
        for (int q = 0; q < kAhead; ++q) {
          if (j + q < c1) {
            cstart[j + q] = (double)S * u;
            PMap m; m.e = ae[q]; m.o = ao[q];
            S = pmap_apply(m, S);
          }
        }
      }
      const long long Send = pmap_apply(tot, S0);
      if (Send >= 0x0020000000000000LL) atomicOr(err, 2u);  // a safe stretch left its binade: bound violated
      cur = (double)Send * u;
      t = t2;
      __syncthreads();
    }
  }
}

// ---- the chain of long inputs, with the stretches taken out of the serial part ---------------------------------------
// scan_chain_kernel spends ~6 us per stretch of safe tiles (probe, composition pass, write-back pass) on top of every
// hard tile's own walk.  None of that depends on the running value: the increment of a tie-free safe tile is the plain
// integer aggE (aggE == aggO), so the sum of the increments since the last hard tile is a segmented prefix over all
// tiles, made by two parallel launches ahead of the chain (tiles with a tie count as hard: hard_tile_walk handles any
// tile).  The serial kernel then only walks the hard tiles in order (the stretch before each is ONE addition), and a
// last parallel launch turns (value after the preceding hard tile, increments since) into every safe tile's start.
//   columns reused after scan_maps (exact mode reads neither again):  tsum -> rel[t]  int64 increments since the last
//   break before t;  aggO -> pb[t]  tile index of that break;  aprefix -> the list of breaks in order.
struct SegVal {
  long long s;   // sum of aggE since the last break (saturating)
  long long lb;  // tile index of the last break, -1: none yet
  long long c;   // number of breaks
};
__device__ __forceinline__ SegVal seg_then(SegVal a, SegVal b) {  // a first, then b
  SegVal r;
  r.s = b.lb >= 0 ? b.s : sat(a.s + b.s);
  r.lb = b.lb >= 0 ? b.lb : a.lb;
  r.c = a.c + b.c;
  return r;
}
__device__ __forceinline__ SegVal seg_shfl_up(SegVal v, int o) {
  SegVal t;
  t.s = __shfl_up_sync(0xffffffffu, v.s, o);
  t.lb = __shfl_up_sync(0xffffffffu, v.lb, o);
  t.c = __shfl_up_sync(0xffffffffu, v.c, o);
  return t;
}
// prefix inside chunks of 1024 tiles; part[3 b ..] = the chunk's total
__global__ void __launch_bounds__(1024) scan_seg_local_kernel(TileState ts, int64_t ntiles, long long* __restrict__ part) {
  __shared__ long long s_w[3 * 33];
  long long* cls = reinterpret_cast<long long*>(ts.flagB);
  long long* rel = reinterpret_cast<long long*>(ts.tsum);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  SegVal v{0, -1, 0};
  if (t < ntiles) {
    const long long c = cls[t];
    const bool brk = c == kHardTile || cls_has_tie(c);
    if (brk && c != kHardTile) cls[t] = kHardTile;  // a tile with a tie is walked by the chain like a hard one
    if (brk) { v.s = 0; v.lb = t; v.c = 1; } else { v.s = sat(ts.aggE[t]); }
  }
  SegVal inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const SegVal x = seg_shfl_up(inc, o);
    if (lane >= o) inc = seg_then(x, inc);
  }
  if (lane == 31) { s_w[3 * warp] = inc.s; s_w[3 * warp + 1] = inc.lb; s_w[3 * warp + 2] = inc.c; }
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive prefix over the 32 warp totals, in order
    SegVal run{0, -1, 0};
    for (int w = 0; w < 32; ++w) {
      const SegVal cur{s_w[3 * w], s_w[3 * w + 1], s_w[3 * w + 2]};
      s_w[3 * w] = run.s; s_w[3 * w + 1] = run.lb; s_w[3 * w + 2] = run.c;
      run = seg_then(run, cur);
    }
    part[3 * blockIdx.x] = run.s; part[3 * blockIdx.x + 1] = run.lb; part[3 * blockIdx.x + 2] = run.c;
  }
  __syncthreads();
  SegVal exc = seg_shfl_up(inc, 1);
  if (lane == 0) exc = SegVal{0, -1, 0};
  const SegVal wp{s_w[3 * warp], s_w[3 * warp + 1], s_w[3 * warp + 2]};
  exc = seg_then(wp, exc);
  if (t < ntiles) {
    rel[t] = exc.s;
    ts.aggO[t] = exc.lb;                                  // (-1: no break before t inside this chunk)
    ts.aggE[t] = v.c ? -(exc.c + 1) : ts.aggE[t];          // a break remembers its rank inside the chunk (negative: never a sum)
  }
}
// carries across the chunks; the ordered list of breaks; part[3 nchunks] <- number of breaks
__global__ void __launch_bounds__(1024) scan_seg_carry_kernel(TileState ts, int64_t ntiles, int nchunks,
                                                              long long* __restrict__ part) {
  __shared__ long long s_c[3];
  long long* cls = reinterpret_cast<long long*>(ts.flagB);
  long long* rel = reinterpret_cast<long long*>(ts.tsum);
  long long* blist = reinterpret_cast<long long*>(ts.aprefix);
  if (threadIdx.x == 0) {
    SegVal run{0, -1, 0};
    for (int c = 0; c < (int)blockIdx.x; ++c) run = seg_then(run, SegVal{part[3 * c], part[3 * c + 1], part[3 * c + 2]});
    s_c[0] = run.s; s_c[1] = run.lb; s_c[2] = run.c;
    if (blockIdx.x == (unsigned)nchunks - 1) {
      const SegVal all = seg_then(run, SegVal{part[3 * blockIdx.x], part[3 * blockIdx.x + 1], part[3 * blockIdx.x + 2]});
      part[3 * nchunks] = all.c;
      part[3 * nchunks + 1] = all.s;   // increments after the last break
    }
  }
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (t >= ntiles) return;
  if (ts.aggO[t] < 0) {  // no break before t inside the chunk: the carry's
    rel[t] = sat(s_c[0] + rel[t]);
    ts.aggO[t] = s_c[1];
  }
  if (cls[t] == kHardTile) blist[s_c[2] + (-ts.aggE[t] - 1)] = t;
}
// safe tiles: start value = (value after the preceding break, in ulps of the tile's binade) + increments since
__global__ void __launch_bounds__(1024) scan_seg_apply_kernel(TileState ts, int64_t ntiles, unsigned int* err) {
  const long long* cls = reinterpret_cast<const long long*>(ts.flagB);
  const long long* rel = reinterpret_cast<const long long*>(ts.tsum);
  const int64_t t = (int64_t)blockIdx.x * 1024 + threadIdx.x;
  if (t >= ntiles) return;
  const long long c = cls[t];
  if (c == kHardTile) return;
  const int e = (int)c;
  const long long pb = ts.aggO[t];
  bool ok = pb >= 0;
  const long long S0 = ok ? int_of(ts.inclB[pb], e, &ok) : 0;
  const long long S = S0 + rel[t];
  if (!ok) atomicOr(err, 4u);                              // the tile's binade is not the running value's
  if (S >= 0x0020000000000000LL) atomicOr(err, 2u);        // a safe stretch left its binade: bound violated
  ts.inclB[t] = (double)S * ulp_value(e - 52);
}
// the serial part: the breaks in order; inclB[b] <- the exact value AFTER tile b
template <int IN>
__global__ void __launch_bounds__(kThreads) scan_chain_seg_kernel(int64_t n, const double* __restrict__ in,
                                                                  double* __restrict__ cdf, double* stats, TileState ts,
                                                                  int64_t ntiles, const long long* __restrict__ part,
                                                                  int nchunks, unsigned int* err,
                                                                  const double* __restrict__ wpart) {
  __shared__ long long s_ll[72];
  __shared__ int s_int[40];
  __shared__ double s_cur;
  const long long* cls = reinterpret_cast<const long long*>(ts.flagB);
  const long long* rel = reinterpret_cast<const long long*>(ts.tsum);
  const long long* blist = reinterpret_cast<const long long*>(ts.aprefix);
  const long long nb = part[3 * nchunks];
  double cur = 0.0;
  long long prev = -1;
  for (long long k = 0; k < nb; ++k) {
    const long long b = blist[k];
    if (b > prev + 1) {  // the stretch of safe tiles before this break: one addition
      const int e = (int)cls[prev + 1];
      bool ok;
      const long long S0 = int_of(cur, e, &ok);
      const long long Send = S0 + rel[b];
      if (threadIdx.x == 0) {
        if (!ok) atomicOr(err, 1u);
        if (Send >= 0x0020000000000000LL) atomicOr(err, 2u);
      }
      cur = (double)Send * ulp_value(e - 52);
    }
    double p[kItems], c[kItems];
    load_tile<IN>(in, n, b, IN != kInW ? ts.tmax[b] : 0.0, p, wpart);
    const int64_t left = n - b * kScanTile;
    const int cnt = (int)(left < kScanTile ? left : kScanTile);
    cur = hard_tile_walk(p, cnt, cur, c, s_ll, s_int, &s_cur, b == 0);
    store_tile(cdf, n, b, c);
    if (threadIdx.x == 0) {
      ts.inclB[b] = cur;
      if (b * kScanTile + cnt == n) stats[PFB_STAT_TOTAL] = cur;
    }
    prev = b;
    __syncthreads();
  }
}

template <bool EXACT, int IN>
__global__ void __launch_bounds__(kThreads) scan_final_kernel(int64_t n, const double* __restrict__ in,
                                                              double* __restrict__ cdf, double* stats, TileState ts,
                                                              int64_t ntiles, const double* __restrict__ wpart) {
  __shared__ double s_dbl[40];
  __shared__ long long s_ll[72];
  const long long* cls = reinterpret_cast<const long long*>(ts.flagB);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long e = 0;
    bool has_tie = false;
    if (EXACT) {
      e = cls[tile];
      if (e == kHardTile) continue;  // written by scan_chain_kernel
      has_tie = cls_has_tie(e);
      e = cls_binade(e);
    }
    double p[kItems], c[kItems];
    load_tile<IN>(in, n, tile, IN != kInW ? ts.tmax[tile] : 0.0, p, wpart);
    if constexpr (!EXACT) {
      double tsum = p[0];
#pragma unroll
      for (int k = 1; k < kItems; ++k) tsum += p[k];
      double total;
      const double toff = block_excl_scan(tsum, s_dbl, &total);
      double run = ts.aprefix[tile] + toff;
#pragma unroll
      for (int k = 0; k < kItems; ++k) { run += p[k]; c[k] = run; }
    } else {
      const int ue = (int)e - 52;
      const double u = ulp_value(ue);
      bool ok;
      const long long S0 = int_of(ts.inclB[tile], (int)e, &ok);
      if (!has_tie) {  // plain integer prefix sums (the usual case)
        long long inc[kItems];
        bool tie = false;
        long long tsum = 0;
        if (ue >= kFastUlpMin) {
          const double inv_ulp = ulp_value(-ue);
#pragma unroll
          for (int k = 0; k < kItems; ++k) { inc[k] = inc_of_fast(p[k], inv_ulp, &tie); tsum += inc[k]; }
        } else {
#pragma unroll
          for (int k = 0; k < kItems; ++k) { inc[k] = inc_of(p[k], ue, &tie); tsum += inc[k]; }
        }
        long long tot;
        long long S = S0 + block_excl_scan_ll(tsum, s_ll, &tot);
#pragma unroll
        for (int k = 0; k < kItems; ++k) { S += inc[k]; c[k] = (double)S * u; }
      } else {
        PMap f = pmap_of(p[0], ue);
#pragma unroll
        for (int k = 1; k < kItems; ++k) f = pmap_then(f, pmap_of(p[k], ue));
        PMap tot;
        const PMap g = block_excl_scan_pmap(f, s_ll, &tot);
        long long S = pmap_apply(g, S0);
#pragma unroll
        for (int k = 0; k < kItems; ++k) {
          S = pmap_apply(pmap_of(p[k], ue), S);
          c[k] = (double)S * u;
        }
      }
    }
    store_tile(cdf, n, tile, c);
    if (tile == ntiles - 1) {
      const int64_t last = n - 1 - tile * kScanTile;
      // the elements past the end of the data are zeros (load_tile), so the thread's last entry equals the entry of
      // element n - 1; a dynamic c[last % kItems] would move the whole array to local memory
      if ((int64_t)threadIdx.x == last / kItems) stats[PFB_STAT_TOTAL] = c[kItems - 1];
    }
    __syncthreads();
  }
}

}  // namespace pfb

using namespace pfb;

extern "C" int pfb_normalize(int64_t n, const double* logw, double* w_out, double* stats, void* ws,
                             int64_t ws_bytes, void* stream) {
  if (n <= 0) return PFB_OK;
  if (!logw || !stats) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = stream_grid(n, 4);
  lse_reduce_kernel<<<grid, kThreads, 0, st>>>(n, logw, stats, W);
  PFB_LAUNCH_OK("lse_reduce_kernel");
  if (w_out) {
    weights_kernel<<<stream_grid(n), kThreads, 0, st>>>(n, logw, w_out, stats);
    PFB_LAUNCH_OK("weights_kernel");
  }
  return PFB_OK;
}

namespace pfb {
// n: elements in the (possibly padded) array; n_runs runs of tpr tiles each (n_runs = 1: tpr = all tiles)
static int scan_impl(int32_t mode, int64_t n, const double* in, int in_kind, double* cdf, double* stats,
                     Workspace W, int64_t n_runs, int64_t tpr, bool have_tile_stats, cudaStream_t st) {
  const int64_t ntiles = n_runs * tpr;
  unsigned int* ticket = W.counters + 1;
  unsigned int* err = W.counters + 2;
  TileState ts = make_tile_state(W.tiles, ntiles, ticket);
  const bool from_log = in_kind == kInLog, scaled = in_kind == kInScaled;
  // 1. per-tile sums: combine the partials the weighting kernel left in the workspace, or make them here
  if (scaled) {
    tile_combine_scaled_kernel<<<stream_grid(ntiles), kThreads, 0, st>>>(ts, W.wpart, ntiles);
    PFB_LAUNCH_OK("tile_combine_scaled_kernel");
  } else if (have_tile_stats) {
    tile_combine_kernel<<<stream_grid(ntiles), kThreads, 0, st>>>(ts, W.wpart, ntiles);
    PFB_LAUNCH_OK("tile_combine_kernel");
  } else {
    const int g = (int)(ntiles < (int64_t)sm_count() * 8 ? ntiles : (int64_t)sm_count() * 8);
    if (from_log) tile_stats_kernel<true><<<g, kThreads, 0, st>>>(n, in, stats, ts, ntiles);
    else tile_stats_kernel<false><<<g, kThreads, 0, st>>>(n, in, stats, ts, ntiles);
    PFB_LAUNCH_OK("tile_stats_kernel");
  }
  // 2. approximate prefix of the tile sums (one CTA; one warp per run when batched)
  if (n_runs == 1 && ntiles > 8192) {
    const int nchunks = (int)((ntiles + 1023) / 1024);  // (<= 2^20 tiles for 2^31 particles: 1024 chunks)
    double* part = W.partials;
    if (scaled) tile_prefix_max_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, part);
    tile_prefix_local_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, from_log ? stats + PFB_STAT_MAX : nullptr, scaled ? 1 : 0, nchunks, part);
    tile_prefix_carry_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, nchunks, part);
    PFB_LAUNCH_OK("tile_prefix_local_kernel");
  } else if (n_runs == 1) {
    tile_prefix_kernel<<<1, 1024, 0, st>>>(ts, ntiles, from_log ? stats + PFB_STAT_MAX : nullptr, scaled ? 1 : 0);
    PFB_LAUNCH_OK("tile_prefix_kernel");
  } else {
    const int64_t warps = (int64_t)sm_count() * 8 * (kThreads / 32);
    const int g = (int)((n_runs < warps ? n_runs : warps) * 32 + kThreads - 1) / kThreads;
    tile_prefix_runs_kernel<<<g, kThreads, 0, st>>>(ts, n_runs, (int)tpr, in_kind);
    PFB_LAUNCH_OK("tile_prefix_runs_kernel");
  }
  // 3. the scan proper
  const double margin = 4.0 * (double)(n + kScanTile) * 1.1102230246251565e-16;
  const int tg = (int)(ntiles < (int64_t)sm_count() * 8 ? ntiles : (int64_t)sm_count() * 8);
  const double* wp = W.wpart;
  if (mode == PFB_SCAN_SEQEXACT) {
    const int rg = (int)(n_runs < (int64_t)sm_count() * 8 ? n_runs : (int64_t)sm_count() * 8);
    // long single inputs: the stretches of safe tiles leave the serial chain (scan_seg_* above)
    static const int64_t seg_min = getenv("PFB_SCAN_SEG_MIN") ? atoll(getenv("PFB_SCAN_SEG_MIN")) : 8192;
    const bool seg = n_runs == 1 && ntiles > seg_min;
    const int nchunks = (int)((ntiles + 1023) / 1024);
    long long* spart = reinterpret_cast<long long*>(W.partials);
#define PFB_SCAN_EXACT(IN)                                                                            \
    scan_maps_kernel<IN><<<tg, kThreads, 0, st>>>(n, in, ts, ntiles, tpr, margin, wp);                \
    if (seg) {                                                                                        \
      scan_seg_local_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, spart);                             \
      scan_seg_carry_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, nchunks, spart);                    \
      scan_chain_seg_kernel<IN><<<1, kThreads, 0, st>>>(n, in, cdf, stats, ts, ntiles, spart, nchunks, err, wp); \
      scan_seg_apply_kernel<<<nchunks, 1024, 0, st>>>(ts, ntiles, err);                               \
    } else {                                                                                          \
      scan_chain_kernel<IN><<<rg, kThreads, 0, st>>>(n, in, cdf, stats, ts, n_runs, tpr, err, wp);    \
    }                                                                                                 \
    scan_final_kernel<true, IN><<<tg, kThreads, 0, st>>>(n, in, cdf, stats, ts, ntiles, wp);
    if (scaled) { PFB_SCAN_EXACT(kInScaled) } else if (from_log) { PFB_SCAN_EXACT(kInLog) } else { PFB_SCAN_EXACT(kInW) }
#undef PFB_SCAN_EXACT
  } else {
    if (scaled) scan_final_kernel<false, kInScaled><<<tg, kThreads, 0, st>>>(n, in, cdf, stats, ts, ntiles, wp);
    else if (from_log) scan_final_kernel<false, kInLog><<<tg, kThreads, 0, st>>>(n, in, cdf, stats, ts, ntiles, wp);
    else scan_final_kernel<false, kInW><<<tg, kThreads, 0, st>>>(n, in, cdf, stats, ts, ntiles, wp);
  }
  PFB_LAUNCH_OK("scan kernels");
  return PFB_OK;
}
}  // namespace pfb

extern "C" int pfb_scan_cdf(int32_t mode, int64_t n, const double* w, const double* logw, double* cdf,
                            double* stats, void* ws, int64_t ws_bytes, void* stream) {
  if (n <= 0) return PFB_OK;
  if ((w == nullptr) == (logw == nullptr)) { set_error("exactly one of w / logw must be given"); return PFB_ERR_INVALID; }
  if (!cdf || !stats) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  const bool have_tile_stats = (mode & PFB_SCAN_TILE_STATS_VALID) != 0 && logw != nullptr;
  const bool scaled = (mode & PFB_SCAN_SLICE_SCALED) != 0;
  if (scaled && !have_tile_stats) { set_error("PFB_SCAN_SLICE_SCALED needs the slice exponents a weighting call left in the workspace (PFB_SCAN_TILE_STATS_VALID, logw)"); return PFB_ERR_INVALID; }
  mode &= ~(PFB_SCAN_TILE_STATS_VALID | PFB_SCAN_SLICE_SCALED);
  if (mode != PFB_SCAN_FAST && mode != PFB_SCAN_SEQEXACT) { set_error("bad scan mode %d", mode); return PFB_ERR_INVALID; }
  const double* in = w ? w : logw;
  if (((uintptr_t)in & 15) || ((uintptr_t)cdf & 15)) { set_error("scan buffers must be 16-byte aligned"); return PFB_ERR_INVALID; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, n, &W);
  if (rc) return rc;
  return scan_impl(mode, n, in, scaled ? kInScaled : (logw != nullptr ? kInLog : kInW), cdf, stats, W, 1, scan_tiles(n), have_tile_stats, (cudaStream_t)stream);
}

extern "C" int pfb_scan_cdf_batched(int32_t mode, int64_t n_runs, int64_t run_len, const double* logw_pad,
                                    double* cdf_pad, double* stats, void* ws, int64_t ws_bytes, void* stream) {
  if (n_runs <= 0 || run_len <= 0) return PFB_OK;
  if (!logw_pad || !cdf_pad || !stats) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  if (!(mode & PFB_SCAN_TILE_STATS_VALID)) { set_error("pfb_scan_cdf_batched must follow pfb_loglik_batched / pfb_predict_loglik_batched on the same buffers"); return PFB_ERR_INVALID; }
  const bool scaled = (mode & PFB_SCAN_SLICE_SCALED) != 0;
  mode &= ~(PFB_SCAN_TILE_STATS_VALID | PFB_SCAN_SLICE_SCALED);
  if (mode != PFB_SCAN_FAST && mode != PFB_SCAN_SEQEXACT) { set_error("bad scan mode %d", mode); return PFB_ERR_INVALID; }
  if (((uintptr_t)logw_pad & 15) || ((uintptr_t)cdf_pad & 15)) { set_error("scan buffers must be 16-byte aligned"); return PFB_ERR_INVALID; }
  const int64_t tpr = scan_tiles(run_len);
  const int64_t n_pad = n_runs * tpr * kScanTile;
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, n_pad, &W);
  if (rc) return rc;
  return scan_impl(mode, n_pad, logw_pad, scaled ? kInScaled : kInLog, cdf_pad, stats, W, n_runs, tpr, true, (cudaStream_t)stream);
}
