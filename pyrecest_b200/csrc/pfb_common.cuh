// Shared device/host helpers for libpfb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pfb.h"

namespace pfb {

constexpr int kThreads = 256;            // CTA size of the streaming kernels
constexpr int kMaxPartialBlocks = 4096;  // upper bound on reduction partials per launch
constexpr double kTwoPi = 6.283185307179586476925286766559;  // == 2.0 * math.pi in fp64

// ---- error text (thread-local) ---------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define PFB_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return pfb::cuda_fail(_e, #expr); \
  } while (0)

#define PFB_LAUNCH_OK(name)                                   \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return pfb::cuda_fail(_e, name);   \
  } while (0)

int sm_count();

// ---- workspace layout --------------------------------------------------------------------
// [0, 256)            : unsigned int counters (zeroed once by pfb_workspace_init, self-cleaning)
// [256, +partials)    : kMaxPartialBlocks * 64 doubles of reduction partials
// [.., +tiles)        : scan tile state (flags re-zeroed by every scan call)
constexpr int64_t kWsCounters = 256;
constexpr int64_t kPartialDoubles = 64;  // doubles per block partial (>= D*D + D + 1 ...)
constexpr int64_t kWsPartials = (int64_t)kMaxPartialBlocks * kPartialDoubles * 8;
constexpr int kScanTile = 2048;          // elements per scan tile
// per tile: flagA, sumA_agg, sumA_incl, flagB, agg_even, agg_odd, incl  (7 x 8 bytes) -> pad to 64
constexpr int64_t kTileStateBytes = 192;  // 8 state columns + 8 warps x (max, sum) partials of the weighting kernels
constexpr int kGuideShift = 3;           // guide table: one bucket per 2^3 = 8 cdf entries (rounded to pow2)

inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }
constexpr int kWsJobCap = 16384;                          // (source, first slot, count) triples of heavy particles
constexpr int64_t kWsJobs = (int64_t)kWsJobCap * 3 * 8;
// count tree of the sorted-order multinomial resampling (pfb_msort.cu): node counts and left-child accumulators of
// the top levels (2^18 ints each: up to 2^17 coarse bins) and the exclusive offsets of the coarse bins (2^17 + 64)
constexpr int kMsTopMaxLevels = 17;
constexpr int64_t kMsTreeInts = (int64_t)1 << (kMsTopMaxLevels + 1);
constexpr int64_t kWsMs = (2 * kMsTreeInts + ((int64_t)1 << kMsTopMaxLevels) + 64) * 4;
inline int64_t ws_header() { return kWsCounters + kWsPartials + kWsJobs + kWsMs; }
// guide table of the resampling search: M = 2^k buckets, M >= n / 2^kGuideShift (PFB_GUIDE_SHIFT overrides)
int guide_shift();
inline int64_t guide_buckets(int64_t n) {
  int64_t want = n >> guide_shift();
  int64_t M = 1;
  while (M < want) M <<= 1;
  return M;
}
inline int64_t guide_offset(int64_t n) {
  return (ws_header() + scan_tiles(n) * kTileStateBytes + 255) & ~(int64_t)255;
}

struct Workspace {
  unsigned int* counters;
  double* partials;
  long long* jobs;  // kWsJobCap triples (systematic resampling: offspring runs that are spread over the grid)
  int32_t* ms;      // kWsMs bytes: count tree of pfb_resample_msort
  unsigned char* tiles;
  int64_t tile_bytes;
  int32_t* guide;
  int64_t guide_buckets;
  double* tmax;   // per scan tile: max log-weight   } filled by tile_combine / tile_stats,
  double* tsum;   // per scan tile: sum exp(l - max)  } read by the tile-prefix kernels
  double* wpart;  // per scan tile: 8 warps x (max, sum exp): what the weighting kernels leave behind
};
// tile-state columns (ntiles entries each): 0 flag, 1 aggE, 2 aggO, 3 incl, 4 tmax, 5 tsum, 6 aprefix (+ spare)
inline double* tile_col(unsigned char* base, int64_t ntiles, int col) {
  return reinterpret_cast<double*>(base) + (int64_t)col * ntiles;
}
int carve_workspace(void* ws, int64_t ws_bytes, int64_t n, Workspace* out);

// ---- numpy.mod(x, 2*pi) ----------------------------------------------------------------------
// numpy/_core/src/npymath: mod = fmod(a, b); if (mod) { if ((b < 0) != (mod < 0)) mod += b; }
// else mod = copysign(0, b).   fmod is exact, the add is one IEEE rounding (can return 2*pi).
// general case of np_mod_2pi, out of line: fmod is ~150 instructions and the wrap is inlined at nine places of the
// fused predict + weight kernel, whose instruction-cache misses were its largest stall reason
static __device__ __noinline__ double np_mod_2pi_slow(double a) {
  double m = fmod(a, kTwoPi);
  if (m != 0.0) {
    if (m < 0.0) m += kTwoPi;
  } else {
    m = 0.0;  // +0.0 also for -0.0 / negative multiples
  }
  return m;
}

// (fp64 literals are materialised with two UMOVs next to every instruction that uses them; values read from
//  constant memory come as one LDCU per pair and stay in uniform registers)
static __constant__ double kAngleConst[4] = {6.283185307179586476925286766559, 12.566370614359172953850573533118,
                                             -6.283185307179586476925286766559, 3.141592653589793};
__device__ __forceinline__ double np_mod_2pi(double a) {
  // fast exact paths for -2 pi < a < 4 pi (every in-range particle +- noise lands here), branch-free -- lanes of a
  // warp fall on both sides of 2 pi all the time: r = a + {0, -2 pi, +2 pi}.
  //   [0, 2pi): fmod = a; a + 0.0 also turns -0.0 into +0.0 like copysign(0, b).
  //   [2pi, 4pi): fmod = a - 2pi, exact by Sterbenz.   (-2pi, 0): fmod = a, then += 2pi (one rounding, can return 2 pi).
  double adj = (a >= kAngleConst[0]) ? kAngleConst[2] : 0.0;
  adj = (a < 0.0) ? kAngleConst[0] : adj;
  if (!(a > kAngleConst[2] && a < kAngleConst[1])) return np_mod_2pi_slow(a);  // (also NaN)
  return a + adj;
}

// The same fast paths for the hot kernels, where a particle takes 3 D wraps: no branch per wrap -- the arguments
// outside the fast range only raise `bad`, and the caller redoes the particle's wraps with np_mod_2pi_slow under ONE
// branch (essentially never taken).  Results are those of np_mod_2pi wherever bad stays false.
//   mod2pi_pos: +0 <= a < 4 pi (checked on the high word: negative values, -0.0, NaN and inf fail).
//   mod2pi_any: -2 pi < a < 4 pi; ZERO_SAFE adds the a + 0.0 that turns -0.0 into +0.0 (call sites whose argument is
//               a sum with a non-negative wrapped value cannot see -0.0).
__device__ __forceinline__ double mod2pi_pos(double a, bool& bad) {
  bad |= (unsigned)__double2hiint(a) >= 0x402921FBu;  // 4 pi = 0x402921FB54442D18
  double r = a;
  if (a >= kAngleConst[0]) r = a - kAngleConst[0];
  return r;
}
template <bool ZERO_SAFE = false>
__device__ __forceinline__ double mod2pi_any(double a, bool& bad) {
  bad |= !(fabs(a - kAngleConst[3]) < 9.42477795);  // 3 pi - 1e-8 (NaN fails)
  double r = ZERO_SAFE ? a + 0.0 : a;
  if (a >= kAngleConst[0]) r = a - kAngleConst[0];
  if (a < 0.0) r = a + kAngleConst[0];
  return r;
}

// ---- counter-based device RNG (Philox4x32-10) ------------------------------------------------------
// Draws are a pure function of (seed, offset, particle index, slot): no state to carry, the same
// particle gets the same draw whatever the grid shape.  oracle/pf_oracle.py restates it in numpy.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
struct RngKey {
  uint64_t seed;    // stream of this filter
  uint64_t offset;  // call counter (advanced by the host once per predict / resample)
};
__device__ __forceinline__ uint4 rng_block(const RngKey& k, uint64_t index, uint32_t slot) {
  return philox4x32_10(make_uint4((uint32_t)index, (uint32_t)(index >> 32), slot, (uint32_t)k.offset),
                       make_uint2((uint32_t)k.seed, (uint32_t)(k.seed >> 32) ^ (uint32_t)(k.offset >> 32)));
}
// 53-bit uniform in [0, 1) from two 32-bit words
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  return (double)((((uint64_t)a << 32) | b) >> 11) * 1.1102230246251565e-16;
}
// ---- log and sincospi for the Box-Muller transform ------------------------------------------------------------
// The library versions cost ~60 and ~70 instructions here (every fp64 literal is materialised with two moves next to
// the FMA that uses it); the kernels that draw their own normals (R^d filters, wrapped-normal noise, the batched and
// the sharded filter) are issue-bound on exactly that.  Same accuracy class (< 2 ulp), coefficients in constant memory.
//
// fast_log(x), 2^-1022 <= x <= 1 (normal numbers): x = m 2^e, m in [1, 2); j = round((m - 1) 128), table entry
// (A_j, L_j) with m A_j = 1 + r, |r| <= 2^-8, log x = (e + [j >= 64]) ln 2 + L_j + log1p(r).  Entries 0 and 128 are
// exact (A = 1 resp. 1/2, L = 0), so log(1) = 0 and values just below 1 lose nothing to cancellation.
__device__ const double2 kLogTab[129] = {
#include "pfb_log_table.inc"
};
static __constant__ double kLogCoef[10] = {
    6.93147180369123816490e-01, 1.90821492927058770002e-10,  // ln 2: high part (32 significant bits), low part
    // log1p(r) / r on |r| <= 2^-8 (Chebyshev interpolant, degree 7, error 3e-19 of 2^-8)
    -0.3446376251677561, 0.14308502155395733, -0.1666603624887342, 0.19999999345180994, -0.2500000000492417,
    0.33333333333338416, -0.49999999999999994, 1.0};
__device__ __forceinline__ double fast_log(double x) {
  const long long bits = __double_as_longlong(x);
  const int hi = (int)(bits >> 32);
  const int j = (((hi >> 12) & 0xff) + 1) >> 1;                 // round((m - 1) 128): top 8 mantissa bits, rounded
  const int e = ((hi >> 20) & 0x7ff) - 1023 + (j >> 6 ? 1 : 0);  // (j >= 64 also covers j == 128)
  const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
  const double2 t = __ldg(&kLogTab[j]);
  const double r = fma(m, t.x, -1.0);
  double q = kLogCoef[2];
#pragma unroll
  for (int k = 3; k < 10; ++k) q = fma(q, r, kLogCoef[k]);
  const double ed = (double)e;
  return fma(ed, kLogCoef[0], t.y + fma(r, q, ed * kLogCoef[1]));
}
// sin(pi t), cos(pi t) for 0 <= t < 2: t = k / 2 + r, |r| <= 1/4; sin(pi r) / r and cos(pi r) as polynomials in r^2
// (Chebyshev interpolants, absolute error 3e-17), then the quadrant.
static __constant__ double kTrigCoef[17] = {
    -3.0361360815105725e-05, 0.0004681824775472955, -0.007370594536623111, 0.08214589369866764, -0.5992645294795359,
    2.5501640398790455, -5.1677127800499765, 3.141592653589793,                                          // sin(pi r) / r
    -1.960666668533104e-05, -8.73282354687988e-05, 0.0019264845203986242, -0.025806646050006207, 0.23533062036880426,
    -1.3352627686441232, 4.058712126414695, -4.934802200544673, 1.0};                                     // cos(pi r)
__device__ __forceinline__ void fast_sincospi_02(double t, double* sn, double* cs) {
  const double kd = rint(t + t);        // 0 .. 4
  const int k = (int)kd;
  const double r = fma(kd, -0.5, t);    // exact
  const double w = r * r;
  double ps = kTrigCoef[0], pc = kTrigCoef[8];
#pragma unroll
  for (int i = 1; i < 8; ++i) ps = fma(ps, w, kTrigCoef[i]);
#pragma unroll
  for (int i = 9; i < 17; ++i) pc = fma(pc, w, kTrigCoef[i]);
  const double s = ps * r, c = pc;
  // angle = k pi / 2 + pi r:  k odd swaps sine and cosine; signs by quadrant
  const double s1 = (k & 1) ? c : s, c1 = (k & 1) ? s : c;
  *sn = (k & 2) ? -s1 : s1;
  *cs = ((k + 1) & 2) ? -c1 : c1;
}
// two independent standard normals from one Philox block (Box-Muller; u1 in (0, 1])
__device__ __forceinline__ void normal_pair(const uint4& r, double* z0, double* z1) {
  const double u1 = 1.0 - u53(r.x, r.y);
  const double u2 = u53(r.z, r.w);
  const double rad = sqrt(-2.0 * fast_log(u1));
  double sn, cs;
  fast_sincospi_02(2.0 * u2, &sn, &cs);
  *z0 = rad * cs;
  *z1 = rad * sn;
}
// Zero-mean von Mises draws in (-pi, pi) for D dimensions at once: vertical-strip rejection.  g(t) =
// exp(kappa (cos t - 1)) is decreasing on [0, pi]; the host cuts [0, pi] into PFB_VM_STRIPS strips [a_k, a_k + w_k)
// whose hats g(a_k) w_k all have the same area A (pfb_vm_strip_table).  An attempt: strip k, position u in [0, 1),
// t = a_k + u w_k, sign, acceptance uniform v in (0, 1): accept iff v A < g(t) w_k.
// ~99.9 % of the attempts are accepted, so a warp almost never loops (the wrapped-Cauchy rejection it replaces
// accepted ~2/3 and cost three times the instructions through divergence).  The test is decided in fp32 when it
// is clear of the fp32 error bound and in fp64 otherwise, so the decision is the fp64 one (oracle/philox.py).
// Bit budget: an attempt needs 85 random bits -- acceptance uniform 32, strip 12, sign 1, position 2 x 20 (a 40-bit
// position inside a strip of width ~1e-3 rad resolves 1e-15 rad, the spacing of doubles near 1) -- so the first
// attempts of all D dimensions are cut from ONE bit stream of ceil(85 D / 128) Philox blocks (slots 16, 17, ..:
// 2 blocks instead of 3 for D = 3); the rare retry of dimension c takes a block of its own (slot 48 + 8 attempt + c).
template <int NW>
__device__ __forceinline__ uint32_t stream_bits(const uint32_t (&w)[NW], int start, int len) {
  const int i = start >> 5, o = start & 31;
  const uint64_t lo = w[i], hi = (i + 1 < NW) ? w[i + 1] : 0u;
  return (uint32_t)((lo | (hi << 32)) >> o) & (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u));
}

__device__ __forceinline__ bool vonmises_attempt(const pfb_predict_params& p, int c, uint32_t v32, uint32_t strip,
                                                 uint32_t uhi, uint32_t ulo, double* th_out) {
  const double u = (double)(((uint64_t)uhi << 20) | ulo) * 9.094947017729282e-13;  // 2^-40
  const double kap = p.rng_kappa[c];
  if (kap < 1e-8) {
    *th_out = 3.141592653589793 * u;  // uniform on the circle
    return true;
  }
  const int stride = p.rng_vm_table_stride > 0 ? p.rng_vm_table_stride : 2 * PFB_VM_STRIPS;
  const double* tab = p.rng_vm_table_dev + (int64_t)p.rng_vm_table_idx[c] * stride;
  const double2 t = __ldg(reinterpret_cast<const double2*>(tab) + strip);
  const double th = t.x + u * t.y;
  *th_out = th;
  // quick accept: the acceptance word is below the strip's threshold (the density at the strip's far end bounds the
  // density at th from below) -- the outcome of the full test, without evaluating the density
  if (stride == PFB_VM_TABLE_DOUBLES &&
      v32 < __ldg(reinterpret_cast<const uint32_t*>(tab + 2 * PFB_VM_STRIPS) + strip))
    return true;
  const float kf = (float)kap;
  const float sh = __sinf(0.5f * (float)th);  // |error| < 5e-7 on [0, pi/2]: inside the margin below
  const float rhs = __expf(-2.0f * kf * sh * sh) * (float)t.y;
  const float lhs = ((float)v32 + 0.5f) * 2.3283064365386963e-10f * (float)p.rng_vm_area[c];
  const float margin = (4e-6f * kf + 8e-6f) * rhs;
  if (lhs > rhs + margin) return false;
  if (lhs >= rhs - margin) {  // too close for fp32
    const double sd = sinpi(th * 0.15915494309189535);  // sin(th / 2) without the large-argument path
    const double vd = ((double)v32 + 0.5) * 2.3283064365386963e-10;
    return vd * p.rng_vm_area[c] < exp(-2.0 * kap * sd * sd) * t.y;
  }
  return true;
}

template <int D>
__device__ __forceinline__ void vonmises_draws(const RngKey& k, uint64_t index, const pfb_predict_params& p,
                                               double (&out)[D]) {
  constexpr int NB = (85 * D + 127) / 128;
  uint32_t w[4 * NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const uint4 r = rng_block(k, index, 16u + (uint32_t)b);
    w[4 * b] = r.x; w[4 * b + 1] = r.y; w[4 * b + 2] = r.z; w[4 * b + 3] = r.w;
  }
  uint32_t pending = 0u;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const int s0 = 85 * c;
    double th;
    const bool ok = vonmises_attempt(p, c, stream_bits(w, s0, 32), stream_bits(w, s0 + 32, 12),
                                     stream_bits(w, s0 + 45, 20), stream_bits(w, s0 + 65, 20), &th);
    out[c] = stream_bits(w, s0 + 44, 1) ? th : -th;
    if (!ok) pending |= 1u << c;
  }
  for (uint32_t attempt = 1; pending != 0u; ++attempt) {  // ~1.3e-3 of the draws get here
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if (!((pending >> c) & 1u)) continue;
      const uint4 r = rng_block(k, index, 48u + 8u * attempt + (uint32_t)c);
      const uint32_t ww[4] = {r.x, r.y, r.z, r.w};
      double th;
      const bool ok = vonmises_attempt(p, c, stream_bits(ww, 0, 32), stream_bits(ww, 32, 12), stream_bits(ww, 45, 20),
                                       stream_bits(ww, 65, 20), &th) || attempt >= 64u;
      if (ok) {
        out[c] = stream_bits(ww, 44, 1) ? th : -th;
        pending &= ~(1u << c);
      }
    }
  }
}

// ---- record I/O ---------------------------------------------------------------------------
// Particle records are row-major (n, D).  v1: straight per-thread loads; consecutive lanes read
// consecutive records so a warp touches one contiguous 32*D*sizeof(T) byte span.
template <typename T, int D>
__device__ __forceinline__ void load_rec(const T* __restrict__ base, int64_t i, double (&r)[D]) {
  const T* p = base + i * D;
  if constexpr (sizeof(T) == 8 && D % 2 == 0) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) {
      double2 v = __ldg(q + c);
      r[2 * c] = v.x;
      r[2 * c + 1] = v.y;
    }
  } else if constexpr (sizeof(T) == 4 && D % 4 == 0) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 v = __ldg(q + c);
      r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
  } else if constexpr (sizeof(T) == 4 && D % 2 == 0) {
    const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) {
      float2 v = __ldg(q + c);
      r[2 * c] = v.x; r[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) r[c] = (double)__ldg(p + c);
  }
}

// the same through ordinary (coherent) loads and without a restrict promise: for kernels that may run in place
// (x_in == x_out in the predict kernels), where ld.global.nc through a restrict pointer would be undefined behaviour
template <typename T, int D>
__device__ __forceinline__ void load_rec_inplace(const T* base, int64_t i, double (&r)[D]) {
  const T* p = base + i * D;
  if constexpr (sizeof(T) == 8 && D % 2 == 0) {
    const double2* q = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) {
      double2 v = q[c];
      r[2 * c] = v.x;
      r[2 * c + 1] = v.y;
    }
  } else if constexpr (sizeof(T) == 4 && D % 4 == 0) {
    const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int c = 0; c < D / 4; ++c) {
      float4 v = q[c];
      r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
    }
  } else if constexpr (sizeof(T) == 4 && D % 2 == 0) {
    const float2* q = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) {
      float2 v = q[c];
      r[2 * c] = v.x; r[2 * c + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) r[c] = (double)p[c];
  }
}

template <typename T, int D>
__device__ __forceinline__ void store_rec(T* __restrict__ base, int64_t i, const double (&r)[D]) {
  T* p = base + i * D;
  if constexpr (sizeof(T) == 8 && D % 2 == 0) {
    double2* q = reinterpret_cast<double2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) q[c] = make_double2(r[2 * c], r[2 * c + 1]);
  } else if constexpr (sizeof(T) == 4 && D % 4 == 0) {
    float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
    for (int c = 0; c < D / 4; ++c)
      q[c] = make_float4((float)r[4 * c], (float)r[4 * c + 1], (float)r[4 * c + 2], (float)r[4 * c + 3]);
  } else if constexpr (sizeof(T) == 4 && D % 2 == 0) {
    float2* q = reinterpret_cast<float2*>(p);
#pragma unroll
    for (int c = 0; c < D / 2; ++c) q[c] = make_float2((float)r[2 * c], (float)r[2 * c + 1]);
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) p[c] = (T)r[c];
  }
}

// ---- warp / block reductions ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Sum K values per thread across the CTA; result valid in thread 0 (array `v` is overwritten).
template <int K>
__device__ __forceinline__ void block_sum(double (&v)[K], double* smem /* >= K * 32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) smem[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
      double x = (lane < nwarp) ? smem[k * 32 + lane] : 0.0;
      v[k] = warp_sum(x);
    }
  }
}

__device__ __forceinline__ double block_max(double v, double* smem /* >= 32 doubles */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  double x = (lane < nwarp) ? smem[lane] : -INFINITY;
  return warp_max(x);  // valid in warp 0; broadcast by caller if needed
}

// True (for every thread of the CTA) in the last CTA of the grid to arrive.  Call after the
// CTA's partial results were written to global memory by any of its threads.
__device__ __forceinline__ bool block_is_last(unsigned int* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;  // self-cleaning for the next launch on this stream
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// grid sizing for streaming kernels: enough CTAs for ~8 resident per SM, capped for the partials
inline int stream_grid(int64_t n, int per_thread = 1) {
  int64_t want = (n + (int64_t)kThreads * per_thread - 1) / ((int64_t)kThreads * per_thread);
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// grid for the tile-structured kernels: persistent CTAs striding over the scan tiles
inline int tile_grid(int64_t n) {
  int64_t tiles = scan_tiles(n);
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kMaxPartialBlocks) cap = kMaxPartialBlocks;
  if (tiles < 1) tiles = 1;
  return (int)(tiles < cap ? tiles : cap);
}

// ---- acquire / release helpers for the look-back scans ------------------------------------------
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_s64(long long* p, long long v) {
  asm volatile("st.relaxed.gpu.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_s64(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.gpu.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace pfb

// dispatch helpers -------------------------------------------------------------------------------
#define PFB_DISPATCH_DIM(dim, ...)                                  \
  switch (dim) {                                                    \
    case 1: { constexpr int D = 1; __VA_ARGS__; } break;            \
    case 2: { constexpr int D = 2; __VA_ARGS__; } break;            \
    case 3: { constexpr int D = 3; __VA_ARGS__; } break;            \
    case 4: { constexpr int D = 4; __VA_ARGS__; } break;            \
    case 5: { constexpr int D = 5; __VA_ARGS__; } break;            \
    case 6: { constexpr int D = 6; __VA_ARGS__; } break;            \
    default: pfb::set_error("dim %d outside 1..%d", (int)(dim), PFB_MAXD); return PFB_ERR_INVALID; \
  }

#define PFB_DISPATCH_DTYPE(dtype, ...)                              \
  if ((dtype) == PFB_F64) { using T = double; __VA_ARGS__; }        \
  else if ((dtype) == PFB_F32) { using T = float; __VA_ARGS__; }    \
  else { pfb::set_error("bad dtype %d", (int)(dtype)); return PFB_ERR_INVALID; }
