// K(4b) multinomial / systematic resampling (search + gather + fused estimate), K(5) weighted moments,
// and the small host-side plumbing of the C ABI (errors, workspace).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "pfb_common.cuh"
#include "pfb_resample.cuh"

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
  out->jobs = reinterpret_cast<long long*>(b + kWsCounters + kWsPartials);
  out->ms = reinterpret_cast<int32_t*>(b + kWsCounters + kWsPartials + kWsJobs);
  out->tiles = b + ws_header();
  out->tile_bytes = ws_bytes - ws_header();
  out->guide = reinterpret_cast<int32_t*>(b + guide_offset(n));
  out->guide_buckets = guide_buckets(n);
  out->tmax = tile_col(out->tiles, scan_tiles(n), 4);
  out->tsum = tile_col(out->tiles, scan_tiles(n), 5);
  out->wpart = tile_col(out->tiles, scan_tiles(n), 8);  // 16 doubles per tile
  return PFB_OK;
}

// ---------------------------------------------------------------------------------------------
// resampling
// ---------------------------------------------------------------------------------------------
// idx_j = #{i : fl(c_i / total) <= u_j}: numpy's `cdf /= cdf[-1]; cdf.searchsorted(u, 'right')`, with the
// division applied per comparison instead of materialising the normalised CDF.
//
// A plain binary search over an 800 MB CDF is ~27 dependent probes per particle.  Instead:
//   * guide table G[b] = #{i : v_i < b/M}, v_i = fl(c_i/total), for M = 2^k buckets of [0, 1) (M ~ n/8).
//     For u in bucket b = floor(u M) (exact: M is a power of two) the answer lies in [G[b], G[b+1]].
//   * the ~8 candidate CDF entries of a bucket are fetched with independent loads (one or two cache lines,
//     one memory round trip) and counted; only windows longer than kWin fall back to a binary search.
//   * v_i <= u is decided without a division for all but ties: with t = fl(u total),
//       c_i <= t (1 - 2^-52)  =>  c_i/total <= u            =>  v_i <= u
//       c_i >= t (1 + 2^-51)  =>  c_i/total >= u (1+2^-52)   =>  v_i >  u
//     and the exact fl(c_i/total) <= u only in between.
//   * each thread carries kBatch particles through the phases so the dependent chain
//     u -> guide -> CDF window -> particle record is paid once per batch.
constexpr int kWin = 8;
constexpr int kBatch = 2;

// guide entries are the only data of the kernel that is re-used (every bucket is hit ~2^shift times): ask L2
// to keep them (evict_last) while the streams and the random CDF / particle lines pass through
__device__ __forceinline__ uint64_t make_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ int32_t ld_guide(const int32_t* p, uint64_t pol) {
  int32_t v;
  asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}


// Element-driven guide build: b_i = floor(v_i M) is non-decreasing, and G[b] = first i with b_i >= b, so
// element i owns the buckets (b_{i-1}, b_i].  The CDF is streamed once; long runs (a heavy particle
// spans many buckets) are filled by the whole warp.
__global__ void __launch_bounds__(kThreads) guide_kernel(int64_t n, const double* __restrict__ cdf,
                                                         int32_t* __restrict__ guide, int64_t M,
                                                         int64_t run_pad) {
  // n: elements over all runs (padded); run r occupies cdf[r run_pad, (r+1) run_pad) and guide row r (M + 2 ints).
  // Each thread owns kG consecutive elements (loads issued together); run_pad is a multiple of 32 kG when there
  // is more than one run, so neither a thread nor a warp straddles two runs.
  constexpr int kG = 4;
  const double dM = (double)M;
  const int lane = threadIdx.x & 31;
  const int64_t nquads = (n + kG - 1) / kG;
  const int64_t nround = (nquads + 31) & ~(int64_t)31;  // whole warps stay converged for the shuffles
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t qd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < nround; qd += stride) {
    const int64_t i0 = qd * kG;
    const int64_t run = (i0 < n ? i0 : n - 1) / run_pad;
    const int64_t k0 = i0 - run * run_pad;  // position of the first element inside the run
    const int64_t last = (run + 1) * run_pad - 1;
    const double total = __ldg(cdf + (last < n ? last : n - 1));
    int32_t* g = guide + run * (M + 2);
    double c[kG];
#pragma unroll
    for (int k = 0; k < kG; ++k) c[k] = (i0 + k < n) ? __ldg(cdf + i0 + k) : 0.0;
    int64_t b[kG];
#pragma unroll
    for (int k = 0; k < kG; ++k) b[k] = (i0 + k < n) ? (int64_t)(__ddiv_rn(c[k], total) * dM) : M;  // past the end: owns nothing
    int64_t bprev = __shfl_up_sync(0xffffffffu, b[kG - 1], 1);
    if (lane == 0 || k0 == 0) {
      if (k0 == 0) bprev = -1;
      else bprev = (i0 <= n) ? (int64_t)(__ddiv_rn(__ldg(cdf + i0 - 1), total) * dM) : M;
    }
#pragma unroll
    for (int k = 0; k < kG; ++k) {
      const int64_t bp = (k == 0) ? bprev : b[k - 1];
      int64_t len = (i0 + k < n) ? b[k] - bp : 0;
      // short runs of buckets: the owner writes them; long ones: warp-cooperative
      if (len > 0 && len <= 4) {
        for (int64_t bb = bp + 1; bb <= b[k]; ++bb) g[bb] = (int32_t)(k0 + k);
      }
      unsigned longmask = __ballot_sync(0xffffffffu, len > 4);
      while (longmask) {
        const int src = __ffs(longmask) - 1;
        longmask &= longmask - 1;
        const int64_t b0 = __shfl_sync(0xffffffffu, bp, src) + 1;
        const int64_t b1 = __shfl_sync(0xffffffffu, b[k], src);
        const int32_t val = (int32_t)__shfl_sync(0xffffffffu, k0 + k, src);
        const int64_t grow = __shfl_sync(0xffffffffu, run, src);
        int32_t* gs = guide + grow * (M + 2);
        for (int64_t bb = b0 + lane; bb <= b1; bb += 32) gs[bb] = val;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bucket records (single-filter multinomial path).  The guide + CDF-window lookup above costs two dependent
// random DRAM fetches (one 64-byte chunk of the guide, two of the CDF).  A bucket record folds both into ONE
// aligned 64-byte fetch: for bucket b of Mr = 2^k >= n/16 buckets of [0, 1)
//     word 0        : G[b]                       (int32, first CDF index of the bucket)
//     halfword 2    : min(count, 65535)           count = G[b+1] - G[b]
//     halfwords 3.. : q_k = floor((v_{G[b]+k} Mr - b) 2^16), k < min(count, 29)   (0xFFFF = unused)
// v Mr, the subtraction of b and the scaling by 2^16 are all exact, so with p = floor((u Mr - b) 2^16)
//     q_k < p  =>  v_k < u   (counted)        q_k > p  =>  v_k > u   (not counted)
// and only q_k == p (probability ~ count / 65536) or count > 29 needs the exact search in the CDF itself.
// ---------------------------------------------------------------------------------------------
static bool two_pass_record_build() {
  const char* e = getenv("PFB_RECORD_BUILD");  // "2pass": guide_kernel + bucket_record_kernel (kept as a cross-check)
  return e && e[0] == '2';
}
static bool bucket_records_enabled() {
  const char* e = getenv("PFB_BUCKET_RECORDS");  // read per call: the tests run both paths
  return e ? atoi(e) != 0 : true;
}

__global__ void __launch_bounds__(kThreads) bucket_record_kernel(int64_t n, const double* __restrict__ cdf,
                                                                 const int32_t* __restrict__ guide, int64_t M,
                                                                 uint4* __restrict__ recs) {
  const double total = __ldg(cdf + n - 1);
  const double dM = (double)M;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= M; b += stride) {
    const int32_t lo = __ldg(guide + b);
    const int32_t hi = (b < M) ? __ldg(guide + b + 1) : lo;
    const int32_t cnt = hi - lo;
    uint32_t w[16];
    w[0] = (uint32_t)lo;
#pragma unroll
    for (int k = 1; k < 16; ++k) w[k] = 0xFFFFFFFFu;
    w[1] = 0xFFFF0000u | (uint32_t)(cnt < 65535 ? cnt : 65535);
    const int m = cnt < kRecEntries ? cnt : kRecEntries;
    const double db = (double)b;
#pragma unroll
    for (int k = 0; k < kRecEntries; ++k) {
      if (k < m) {
        const double v = __ddiv_rn(__ldg(cdf + lo + k), total);
        const uint32_t q = (uint32_t)((v * dM - db) * 65536.0) & 0xFFFFu;
        const int h = k + 3;  // halfword slot
        if (h & 1) w[h >> 1] = (w[h >> 1] & 0x0000FFFFu) | (q << 16);
        else w[h >> 1] = (w[h >> 1] & 0xFFFF0000u) | q;
      }
    }
    uint4* r = recs + b * 4;
    r[0] = make_uint4(w[0], w[1], w[2], w[3]);
    r[1] = make_uint4(w[4], w[5], w[6], w[7]);
    r[2] = make_uint4(w[8], w[9], w[10], w[11]);
    r[3] = make_uint4(w[12], w[13], w[14], w[15]);
  }
}

// ---- one-pass record build ------------------------------------------------------------------------------
// guide_kernel + bucket_record_kernel stream the CDF twice and divide twice per element.  This kernel builds the
// records straight from the CDF, one 2048-element tile per CTA step:
//   A  every element: bucket b_i = floor(v_i M) and 16-bit in-bucket position q_i into shared memory, plus 32
//      elements of look-ahead (a record holds at most 29 entries, so the look-ahead always decides between an
//      exact count <= 29 and "long") and the bucket of the element before the tile;
//   B  every bucket head (b_i != b_{i-1}) writes its own record -- base i, count, the q of its entries -- and
//      owns the empty buckets (b_{i-1}, b_i): up to 4 it writes itself, longer gaps go to a queue in shared
//      memory that the warps drain with coalesced 16-byte stores; gaps above kBigGap buckets (one particle
//      holding > 4096/M of the weight) are handed to record_fill_kernel, which spreads them over the grid.
// Long buckets store count 65535: the query then takes its upper search bound from the next record's base.
constexpr int kRecTile = 2048;
constexpr int kRecLook = 32;
constexpr int kBigGap = 4096;

__device__ __forceinline__ void store_empty_record(uint4* __restrict__ recs, int64_t bucket, int32_t base) {
  uint4* r = recs + bucket * 4;
  r[0] = make_uint4((uint32_t)base, 0xFFFF0000u, 0xFFFFFFFFu, 0xFFFFFFFFu);
  r[1] = r[2] = r[3] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
}

// Runs: tile t belongs to run t / tpr, whose CDF is cdf[run run_pad, +run_len) and whose records start at
// recs + run (M + 2); indices stored in the records are local to the run (a plain call is one run).
__global__ void __launch_bounds__(kThreads) record_build_kernel(int64_t n_runs, int64_t run_len, int64_t run_pad,
                                                                const double* __restrict__ cdf_all, int64_t M,
                                                                uint4* __restrict__ recs_all, int32_t* __restrict__ jobs,
                                                                unsigned int* __restrict__ job_count) {
  constexpr int kE = kRecTile / kThreads;  // elements per thread
  __shared__ int32_t sb_raw[1 + kRecTile + kRecLook];
  __shared__ __align__(4) uint16_t sq_raw[4 + kRecTile + kRecLook + 4];
  __shared__ int32_t s_gap[kRecTile][3];  // (first bucket, length, base) of the medium gaps of this tile
  __shared__ uint16_t s_head[kRecTile + kRecLook + 2];
  __shared__ int s_wsum[kThreads / 32];
  __shared__ int s_ngap;
  int32_t* sb = sb_raw + 1;  // sb[-1]: bucket of the element before the tile
  uint16_t* sq = sq_raw + 4;  // (front padding: a record is copied from 3 halfwords before its first entry)
  const double dM = (double)M;
  const int64_t tpr = (run_len + kRecTile - 1) / kRecTile;
  const int64_t ntiles = n_runs * tpr;
  const int64_t n = run_len;  // (the body below works inside one run)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double total = 1.0, scale = 1.0;
  // b = floor(v M), q = floor((v M - b) 2^16) with v = fl(c/total), i.e. the high and low bits of
  // floor(A), A = v M 2^16 (an exact scaling of v).  T = c * fl(M 2^16 / total) differs from A by less than
  // 2^-50 A (three roundings), so floor(T) = floor(A) unless T lies that close to an integer: only then divide.
  auto bucket_of = [&](double c, uint32_t* q) -> int32_t {
    double t = c * scale;
    long long f = (long long)t;
    const double fr = t - (double)f;
    const double tol = t * 8.881784197001252e-16;  // 2^-50
    if (fr <= tol || 1.0 - fr <= tol) f = (long long)(__ddiv_rn(c, total) * (dM * 65536.0));
    *q = (uint32_t)f & 0xFFFFu;
    const long long b = f >> 16;
    return (int32_t)(b < 0 ? 0 : (b > M ? M : b));  // (only a corrupt CDF gets clamped; keeps the stores in bounds)
  };
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t run = tile / tpr;
    const double* __restrict__ cdf = cdf_all + run * run_pad;
    uint4* __restrict__ recs = recs_all + run * (M + 2) * 4;
    total = __ldg(cdf + n - 1);
    scale = __ddiv_rn(dM * 65536.0, total);
    const int64_t i0 = (tile - run * tpr) * kRecTile;
    const int T = (int)((n - i0 < kRecTile) ? n - i0 : kRecTile);  // elements of this tile
    if (threadIdx.x == 0) s_ngap = 0;
    // phase A
    const int e0 = threadIdx.x * kE;
    int32_t bmine[kE];
    {
      double c[kE];
      if (i0 + e0 + kE <= n) {
        const double2* src = reinterpret_cast<const double2*>(cdf + i0 + e0);
#pragma unroll
        for (int k = 0; k < kE / 2; ++k) { const double2 v = __ldg(src + k); c[2 * k] = v.x; c[2 * k + 1] = v.y; }
      } else {
#pragma unroll
        for (int k = 0; k < kE; ++k) c[k] = (i0 + e0 + k < n) ? __ldg(cdf + i0 + e0 + k) : total;
      }
#pragma unroll
      for (int k = 0; k < kE; ++k) {
        uint32_t q;
        const int32_t b = bucket_of(c[k], &q);
        bmine[k] = (e0 + k < T) ? b : 0x7fffffff;
        sb[e0 + k] = bmine[k];
        sq[e0 + k] = (uint16_t)q;
      }
      if (threadIdx.x < kRecLook) {  // look-ahead into the next tile
        const int64_t i = i0 + kRecTile + threadIdx.x;
        uint32_t q = 0;
        const int32_t b = (T == kRecTile && i < n) ? bucket_of(__ldg(cdf + i), &q) : 0x7fffffff;
        sb[kRecTile + threadIdx.x] = b;
        sq[kRecTile + threadIdx.x] = (uint16_t)q;
      } else if (threadIdx.x == kRecLook) {
        uint32_t q;
        sb[-1] = (i0 == 0) ? -1 : bucket_of(__ldg(cdf + i0 - 1), &q);
      }
    }
    __syncthreads();
    // phase B: bucket heads (b_i != b_{i-1}), ~1 in 12 elements.  They are compacted IN ORDER into a list (flags,
    // block scan), the heads of the look-ahead and a sentinel appended, so that a head's entry count is the
    // distance to the next head and every lane of a warp builds a record.
    int nh = 0;
    {
      uint32_t flags = 0;
      int32_t prev = sb[e0 - 1];
#pragma unroll
      for (int k = 0; k < kE; ++k) {
        if (e0 + k < T && bmine[k] != prev) flags |= 1u << k;
        prev = bmine[k];
      }
      int incl = __popc(flags);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int off = incl - __popc(flags);
      int tot = 0;
#pragma unroll
      for (int wv = 0; wv < kThreads / 32; ++wv) {
        const int v = s_wsum[wv];
        if (wv < warp) off += v;
        tot += v;
      }
      nh = tot;
      while (flags) {
        const int k = __ffs((int)flags) - 1;
        flags &= flags - 1u;
        s_head[off++] = (uint16_t)(e0 + k);
      }
      if (warp == 0) {  // heads inside the look-ahead, then the sentinel
        const bool hd = sb[kRecTile + lane] != sb[kRecTile + lane - 1] && T == kRecTile;
        const uint32_t bal = __ballot_sync(0xffffffffu, hd);
        if (hd) s_head[tot + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)(kRecTile + lane);
        // sentinel: past the look-ahead (the bucket is long), or the end of the data in the last, partial tile
        if (lane == 0) s_head[tot + __popc(bal)] = (uint16_t)(T == kRecTile ? T + kRecLook + 1 : T);
      }
    }
    __syncthreads();
    {
      for (int hdx = threadIdx.x; hdx < nh; hdx += kThreads) {
        const int e = s_head[hdx];
        const int32_t b = sb[e], bp = sb[e - 1];
        const int cnt = (int)s_head[hdx + 1] - e;  // > kRecEntries: long bucket
        // halfword h of the record (h >= 3) is entry h - 3 = sq[e - 3 + h]: the record is a copy of the 32 halfwords
        // from sq[e - 3], taken as aligned words (funnel-shifted when e - 3 is odd), masked to 0xFFFF past the entries
        uint32_t w[16];
        {
          const int s0 = e - 3 + 4;  // halfword index into sq_raw
          const uint32_t* sw = reinterpret_cast<const uint32_t*>(sq_raw) + (s0 >> 1);
          const int valid = (cnt < kRecEntries ? cnt : kRecEntries) + 3;  // halfwords [3, valid) hold entries
          uint32_t prev = sw[0];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t next = sw[j + 1];
            uint32_t v = (s0 & 1) ? __funnelshift_r(prev, next, 16) : prev;
            prev = next;
            if (2 * j >= valid) v |= 0x0000FFFFu;
            if (2 * j + 1 >= valid) v |= 0xFFFF0000u;
            w[j] = v;
          }
        }
        w[0] = (uint32_t)(i0 + e);
        w[1] = (w[1] & 0xFFFF0000u) | (uint32_t)(cnt <= kRecEntries ? cnt : 65535);
        uint4* r = recs + (int64_t)b * 4;
        r[0] = make_uint4(w[0], w[1], w[2], w[3]);
        r[1] = make_uint4(w[4], w[5], w[6], w[7]);
        r[2] = make_uint4(w[8], w[9], w[10], w[11]);
        r[3] = make_uint4(w[12], w[13], w[14], w[15]);
        // the empty buckets in front of it
        const int32_t gap = b - bp - 1;
        if (gap > 0 && gap <= 4) {
          for (int32_t g = bp + 1; g < b; ++g) store_empty_record(recs, g, (int32_t)(i0 + e));
        } else if (gap > 4) {
          const int slot = atomicAdd(&s_ngap, 1);
          s_gap[slot][0] = bp + 1;
          s_gap[slot][1] = gap;
          s_gap[slot][2] = (int32_t)(i0 + e);
        }
      }
    }
    __syncthreads();
    // phase C: medium gaps, one warp per gap, coalesced 16-byte stores; big gaps to the global list
    {
      const int ng = s_ngap;
      for (int j = warp; j < ng; j += kThreads / 32) {
        const int32_t g0 = s_gap[j][0], len = s_gap[j][1], base = s_gap[j][2];
        if (len > kBigGap) {
          if (lane == 0) {
            const unsigned int slot = atomicAdd(job_count, 1u);
            jobs[3 * slot] = (int32_t)(run * (M + 2)) + g0; jobs[3 * slot + 1] = len; jobs[3 * slot + 2] = base;
          }
          continue;
        }
        uint4* r = recs + (int64_t)g0 * 4;
        for (int32_t x = lane; x < len * 4; x += 32)
          r[x] = (x & 3) ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)
                         : make_uint4((uint32_t)base, 0xFFFF0000u, 0xFFFFFFFFu, 0xFFFFFFFFu);
      }
    }
    __syncthreads();
  }
}

// the big gaps (a few heavy particles): every job is spread over the whole grid
__global__ void __launch_bounds__(kThreads) record_fill_kernel(uint4* __restrict__ recs, const int32_t* __restrict__ jobs,
                                                               unsigned int* __restrict__ job_count,
                                                               unsigned int* __restrict__ done_count) {
  const unsigned int nj = *reinterpret_cast<volatile unsigned int*>(job_count);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (unsigned int j = 0; j < nj; ++j) {
    const int32_t g0 = jobs[3 * j], len = jobs[3 * j + 1], base = jobs[3 * j + 2];
    uint4* r = recs + (int64_t)g0 * 4;
    for (int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; x < (int64_t)len * 4; x += stride)
      r[x] = (x & 3) ? make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu)
                     : make_uint4((uint32_t)base, 0xFFFF0000u, 0xFFFFFFFFu, 0xFFFFFFFFu);
  }
  if (block_is_last(done_count) && threadIdx.x == 0) *job_count = 0u;  // self-cleaning
}


// phase 4 of the resampling kernels: gather the kBatch source records, store them, accumulate the fused estimate
template <typename T, int D, int EST, int NE>
__device__ __forceinline__ void gather_store(const T* __restrict__ x_in, T* __restrict__ x_out,
                                             int64_t* __restrict__ idx_out, const int64_t (&src)[kBatch],
                                             const int64_t (&dst)[kBatch], const int64_t (&idx)[kBatch],
                                             const bool (&live)[kBatch], double (&acc)[NE], int64_t x_in_stride = D) {
  if (x_in != nullptr) {
    double r[kBatch][D];
#pragma unroll
    for (int q = 0; q < kBatch; ++q)
      if (live[q]) gather_rec<T, D>(x_in, src[q], r[q], x_in_stride);
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      if (!live[q]) continue;
      store_rec<T, D>(x_out, dst[q], r[q]);
      if (idx_out) idx_out[dst[q]] = idx[q];
      if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] += r[q][c];
      } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double sn, cs;
          sincos(r[q][c], &sn, &cs);
          acc[2 * c] += cs;
          acc[2 * c + 1] += sn;
        }
      }
    }
  } else if (idx_out) {
#pragma unroll
    for (int q = 0; q < kBatch; ++q)
      if (live[q]) idx_out[dst[q]] = idx[q];
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
                                                            int64_t M, RngKey key, int64_t run_pad, int64_t run_len,
                                                            int64_t n_out_run) {
  // runs: output j belongs to run j / n_out_run and is drawn from cdf[run run_pad, +run_len) / x_in[run run_len, ..)
  // (a plain call is one run: run_pad = run_len = n, n_out_run = n_out)
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  (void)total_ptr; (void)n;
  const uint64_t pol = make_evict_last_policy();
  const bool use_rng = (u == nullptr);
  const double dn = (double)n_out_run;
  const double dM = (double)M;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < n_out; j0 += stride * kBatch) {
    double uj[kBatch], total[kBatch], frac[kBatch];
    int32_t lo[kBatch], hi[kBatch];
    int64_t run[kBatch];
    bool live[kBatch];
    // phase 1: uniforms, then the guide entries
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int64_t j = j0 + q * stride;
      live[q] = j < n_out;
      run[q] = live[q] ? j / n_out_run : 0;
      total[q] = __ldg(cdf + (run[q] + 1) * run_pad - 1);
      if (!live[q]) {
        uj[q] = 0.0;
      } else if (systematic) {
        double u0;
        if (use_rng) { const uint4 r = rng_block(key, (uint64_t)run[q], 7u); u0 = u53(r.x, r.y); }
        else u0 = __ldg(u + run[q]);
        uj[q] = __ddiv_rn(__dadd_rn((double)(j - run[q] * n_out_run), u0), dn);
      } else if (use_rng) {
        const uint4 r = rng_block(key, (uint64_t)j, 7u);
        uj[q] = u53(r.x, r.y);
      } else {
        uj[q] = __ldg(u + j);
      }
    }
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      if (uj[q] >= 1.0) {  // only a rounded-up systematic point can get here
        lo[q] = hi[q] = (int32_t)(run_len - 1);
        frac[q] = 0.0;
      } else {
        const double ub = uj[q] * dM;             // exact
        int64_t b = (int64_t)ub;
        b = b < 0 ? 0 : (b > M ? M : b);
        frac[q] = ub - (double)b;                 // position of u inside its bucket, [0, 1)
        const int32_t* g = guide + run[q] * (M + 2);
        lo[q] = ld_guide(g + b, pol);
        hi[q] = ld_guide(g + b + 1, pol);
        // (a run whose total is 0 / NaN leaves its guide unwritten: keep every later load inside the run)
        lo[q] = lo[q] < 0 ? 0 : (lo[q] > (int32_t)run_len ? (int32_t)run_len : lo[q]);
        hi[q] = hi[q] < lo[q] ? lo[q] : (hi[q] > (int32_t)run_len ? (int32_t)run_len : hi[q]);
      }
    }
    // phase 2: one 64-byte window of the CDF per particle, all loads independent.  Buckets longer than
    // the window are entered where linear interpolation of u inside the bucket points (the CDF is close to
    // linear over a few dozen randomly ordered particles), aligned to the 64-byte DRAM burst.
    double win[kBatch][kWin];
    int32_t wst[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const int32_t len = hi[q] - lo[q];
      int32_t s0 = lo[q];
      if (len > kWin) {
        s0 = lo[q] + (int32_t)(frac[q] * (double)len) - kWin / 2;
        const int64_t abs0 = (run[q] * run_pad + s0) & ~(int64_t)(kWin - 1);
        s0 = (int32_t)(abs0 - run[q] * run_pad);
        if (s0 < lo[q]) s0 = lo[q];
      }
      wst[q] = s0;
#pragma unroll
      for (int k = 0; k < kWin; ++k) {
        const int32_t i = s0 + k;
        win[q][k] = (i < hi[q]) ? __ldg(cdf + run[q] * run_pad + i) : INFINITY;
      }
    }
    // phase 3: count, with the binary search only for long windows
    int64_t idx[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const double t = uj[q] * total[q];
      const double tl = t * (1.0 - 2.220446049250313e-16);
      const double th = t * (1.0 + 4.440892098500626e-16);
      int cnt = 0;
#pragma unroll
      for (int k = 0; k < kWin; ++k) {
        // windows are sorted: once an entry is > u the rest are too; counting all is branch-free
        if (win[q][k] != INFINITY && cdf_le(win[q][k], total[q], uj[q], tl, th)) cnt = k + 1;
      }
      int64_t r = (int64_t)wst[q] + cnt;
      if (cnt == 0 && wst[q] > lo[q])            // everything in the window is > u: the answer lies to its left
        r = search_range<true>(cdf + run[q] * run_pad, (int64_t)lo[q], (int64_t)wst[q], total[q], uj[q]);
      else if (cnt == kWin && wst[q] + kWin < hi[q])  // everything is <= u: to its right
        r = search_range<true>(cdf + run[q] * run_pad, (int64_t)wst[q] + kWin, (int64_t)hi[q], total[q], uj[q]);
      if (r > run_len - 1) r = run_len - 1;
      if (r < 0) r = 0;
      idx[q] = r;
    }
    // phase 4: gather, store, estimate
    int64_t src[kBatch], dst[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      src[q] = run[q] * run_len + idx[q];
      dst[q] = j0 + q * stride;
    }
    gather_store<T, D, EST, NE>(x_in, x_out, idx_out, src, dst, idx, live, acc);
  }
  if constexpr (EST != PFB_EST_NONE) finish_estimate<NE>(acc, red, ws, est_out, n_out);
}

// multinomial resampling through the bucket records (one run, uniforms from memory or from Philox)
template <typename T, int D, int EST>
__global__ void __launch_bounds__(kThreads) resample_rec_kernel(int64_t n, int64_t n_out,
                                                                const double* __restrict__ cdf,
                                                                const double* __restrict__ u,
                                                                const T* __restrict__ x_in, T* __restrict__ x_out,
                                                                int64_t* __restrict__ idx_out, double* est_out,
                                                                Workspace ws, const uint4* __restrict__ recs,
                                                                int64_t M, RngKey key, int64_t n_runs, int64_t run_pad,
                                                                int64_t n_out_run, int64_t x_in_stride) {
  // runs: output j belongs to run j / n_out_run, drawn from cdf[run run_pad, +n) / x_in[run n, ..) through the
  // records at recs + run (M + 2); a plain call is one run (run_pad = n, n_out_run = n_out)
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  const bool use_rng = (u == nullptr);
  const double dM = (double)M;
  const bool multi = n_runs > 1;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j0 < n_out; j0 += stride * kBatch) {
    double uj[kBatch];
    bool live[kBatch];
    int64_t dst[kBatch], bkt[kBatch], run[kBatch];
    uint32_t p2[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      dst[q] = j0 + q * stride;
      live[q] = dst[q] < n_out;
      run[q] = (multi && live[q]) ? dst[q] / n_out_run : 0;
      if (!live[q]) uj[q] = 0.0;
      else if (use_rng) { const uint4 r = rng_block(key, (uint64_t)dst[q], 7u); uj[q] = u53(r.x, r.y); }
      else uj[q] = __ldg(u + dst[q]);
      const double ub = uj[q] * dM;  // exact (M is a power of two), u in [0, 1)
      bkt[q] = (int64_t)ub;
      bkt[q] = bkt[q] < 0 ? 0 : (bkt[q] > M ? M : bkt[q]);  // (an injected u outside [0, 1) must not leave the records)
      const uint32_t p = (uint32_t)((ub - (double)bkt[q]) * 65536.0);
      p2[q] = p | (p << 16);
    }
    uint64_t rec[kBatch][2][4];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const uint4* r = recs + (run[q] * (M + 2) + bkt[q]) * 4;
      ld_sector(r, rec[q][0]);
      ld_sector(r + 2, rec[q][1]);
    }
    int64_t idx[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      const double* crun = cdf + run[q] * run_pad;
      const int64_t r = record_decode(rec[q][0], rec[q][1], p2[q], recs + run[q] * (M + 2) * 4, bkt[q], crun, n, uj[q]);
      idx[q] = r;
    }
    int64_t src[kBatch];
#pragma unroll
    for (int q = 0; q < kBatch; ++q) src[q] = run[q] * n + idx[q];
    gather_store<T, D, EST, NE>(x_in, x_out, idx_out, src, dst, idx, live, acc, x_in_stride);
  }
  if constexpr (EST != PFB_EST_NONE) finish_estimate<NE>(acc, red, ws, est_out, n_out);
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

// per-run moments for batched runs: one CTA per run (deterministic), uniform weights 1/run_len
template <typename T, int D, int KIND>
__global__ void __launch_bounds__(kThreads) moments_batched_kernel(int64_t n_runs, int64_t run_len,
                                                                   const T* __restrict__ x, int order,
                                                                   double* __restrict__ out) {
  constexpr int NE = MomentCount<D, KIND>::value;
  __shared__ double red[NE * 32];
  const double wu = 1.0 / (double)run_len;
  const double ord = (double)order;
  for (int64_t run = blockIdx.x; run < n_runs; run += gridDim.x) {
    double acc[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) acc[k] = 0.0;
    for (int64_t i = threadIdx.x; i < run_len; i += blockDim.x) {
      double r[D];
      load_rec<T, D>(x, run * run_len + i, r);
      if constexpr (KIND == PFB_MOM_MEAN) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma(wu, r[c], acc[c]);
      } else if constexpr (KIND == PFB_MOM_TRIG) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double sn, cs;
          sincos(ord * r[c], &sn, &cs);
          acc[2 * c] = fma(wu, cs, acc[2 * c]);
          acc[2 * c + 1] = fma(wu, sn, acc[2 * c + 1]);
        }
      } else {
        int k = 0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
          for (int b = a; b < D; ++b) { acc[k] = fma(r[a] * wu, r[b], acc[k]); ++k; }
        }
      }
      acc[NE - 1] += wu;
    }
    block_sum<NE>(acc, red);
    if (threadIdx.x == 0) {
      double* o = out + run * (D * D + 2 * D + 2);
      if constexpr (KIND == PFB_MOM_MEAN || KIND == PFB_MOM_TRIG) {
#pragma unroll
        for (int k = 0; k < NE; ++k) o[k] = acc[k];
      } else {
        int k = 0;
#pragma unroll
        for (int a = 0; a < D; ++a) {
#pragma unroll
          for (int b = a; b < D; ++b) { o[a * D + b] = acc[k]; o[b * D + a] = acc[k]; ++k; }
        }
        o[D * D] = acc[NE - 1];
      }
    }
    __syncthreads();
  }
}

template <typename T, int D>
static int launch_moments_batched(int kind, int64_t n_runs, int64_t run_len, const void* x, int order, double* out,
                                  cudaStream_t st) {
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(n_runs < cap ? n_runs : cap);
  switch (kind) {
    case PFB_MOM_MEAN: moments_batched_kernel<T, D, PFB_MOM_MEAN><<<grid, kThreads, 0, st>>>(n_runs, run_len, (const T*)x, order, out); break;
    case PFB_MOM_TRIG: moments_batched_kernel<T, D, PFB_MOM_TRIG><<<grid, kThreads, 0, st>>>(n_runs, run_len, (const T*)x, order, out); break;
    case PFB_MOM_SCATTER: moments_batched_kernel<T, D, PFB_MOM_SCATTER><<<grid, kThreads, 0, st>>>(n_runs, run_len, (const T*)x, order, out); break;
    default: set_error("bad moment kind %d for a batched call", kind); return PFB_ERR_INVALID;
  }
  return PFB_OK;
}

// ---------------------------------------------------------------------------------------------
// systematic resampling, source-centric: the comb u_j = fl(fl(j + u0)/n_comb) is monotone in j, so the offspring
// of source i are the contiguous slots [J(v_{i-1}), J(v_i)), J(v) = #{j : u_j < v}, v_i = fl(C_i/total) -- the
// same idx_j = #{i : v_i <= u_j} as the search formulation, without any search: J has a closed-form guess that is
// corrected with exact comparisons.  One thread per source: the CDF and the particles are streamed once, every
// offspring is one coalesced store; heavy particles (many offspring) are written by the whole warp.
// Global CDF value of entry i: C_i = fl(cdf_offset + c_i), normalised by `total` (cdf_offset = 0, total = c_{n-1}
// for a single filter; shard offset / global total for a sharded one, offspring restricted to
// [j_begin, j_begin + count)).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double comb_point(int64_t j, double u0, double dn) {
  return __ddiv_rn(__dadd_rn((double)j, u0), dn);
}
// u_j < v, i.e. fl(fl(j + u0) / dn) < v, decided without the division unless t = fl(j + u0) lies within 2^-51
// (relative) of p = fl(v dn):  t <= p (1 - 2^-51)  =>  t / dn < v (1 - 2^-52)  =>  fl(t / dn) < v, and symmetrically.
__device__ __forceinline__ bool comb_lt(int64_t j, double v, double p, double u0, double dn) {
  const double t = __dadd_rn((double)j, u0);
  if (t <= p * (1.0 - 4.440892098500626e-16)) return true;
  if (t >= p * (1.0 + 4.440892098500626e-16)) return false;
  return __ddiv_rn(t, dn) < v;
}
// #{j in [0, n_comb) : u_j < v}
__device__ __forceinline__ int64_t comb_count_below(double v, double u0, double dn, int64_t n_comb) {
  const double p = v * dn;
  double g = ceil(p - u0);
  int64_t j = g <= 0.0 ? 0 : (g >= dn ? n_comb : (int64_t)g);
  while (j > 0 && !comb_lt(j - 1, v, p, u0, dn)) --j;
  while (j < n_comb && comb_lt(j, v, p, u0, dn)) ++j;
  return j;
}

template <typename T, int D, int EST>
__global__ void __launch_bounds__(kThreads) systematic_scatter_kernel(
    int64_t n, const double* __restrict__ cdf, double cdf_offset, double total_imm, double u0_imm,
    const double* __restrict__ u0_ptr, RngKey key, int u0_mode, int64_t j_begin, int64_t count, int64_t n_comb,
    const T* __restrict__ x_in, T* __restrict__ x_out, int64_t* __restrict__ idx_out, double* est_out, Workspace ws,
    int64_t big_cnt) {
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  const double total = (total_imm > 0.0) ? total_imm : __dadd_rn(cdf_offset, cdf[n - 1]);
  double u0 = u0_imm;
  if (u0_mode == 1) u0 = u0_ptr[0];
  else if (u0_mode == 2) { const uint4 r = rng_block(key, 0ULL, 7u); u0 = u53(r.x, r.y); }
  const double dn = (double)n_comb;
  const int64_t j_end = j_begin + count;
  const int lane = threadIdx.x & 31;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  // A piece [j_begin, j_end) of the comb (sharded filter: the offspring that go to one rank) comes from a contiguous
  // range of sources: J(v_i) is monotone, so the first source with J(v_i) > j_begin and the first with
  // J(v_i) >= j_end are found by bisection and only the sources in between are visited -- a small piece costs
  // ~50 dependent probes instead of a pass over the whole shard.
  __shared__ int64_t s_src[2];
  int64_t i_first = 0, i_last = n - 1;
  if (j_begin > 0 || j_end < n_comb) {
    if (threadIdx.x == 0) {
      auto J = [&](int64_t i) { return comb_count_below(__ddiv_rn(__dadd_rn(cdf_offset, __ldg(cdf + i)), total), u0, dn, n_comb); };
      int64_t lo = 0, hi = n - 1;  // first i in [0, n-1) with J(v_i) > j_begin, else n-1
      while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (J(mid) > j_begin) hi = mid; else lo = mid + 1; }
      s_src[0] = lo;
      hi = n - 1;                  // first i >= that with J(v_i) >= j_end, else n-1
      while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (J(mid) >= j_end) hi = mid; else lo = mid + 1; }
      s_src[1] = lo;
    }
    __syncthreads();
    i_first = s_src[0] & ~(int64_t)31;  // warps stay aligned to 32 consecutive sources
    i_last = s_src[1];
  }
  const int64_t nround = (i_last + 1 + 31) & ~(int64_t)31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = i_first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    // e = first slot that no longer belongs to source i
    int64_t e = j_end;
    if (i < n - 1) {
      const double v = __ddiv_rn(__dadd_rn(cdf_offset, __ldg(cdf + i)), total);
      e = comb_count_below(v, u0, dn, n_comb);
      e = e < j_begin ? j_begin : (e > j_end ? j_end : e);
    }
    int64_t sbeg = __shfl_up_sync(0xffffffffu, e, 1);
    if (lane == 0) {
      if (i == 0) sbeg = j_begin;
      else if (i < n) {
        const double vp = __ddiv_rn(__dadd_rn(cdf_offset, __ldg(cdf + i - 1)), total);
        sbeg = comb_count_below(vp, u0, dn, n_comb);
        sbeg = sbeg < j_begin ? j_begin : (sbeg > j_end ? j_end : sbeg);
      }
    }
    int64_t cnt = (i < n) ? e - sbeg : 0;
    double r[D];
#pragma unroll
    for (int c = 0; c < D; ++c) r[c] = 0.0;
    if (cnt > 0 && x_in != nullptr) load_rec<T, D>(x_in, i, r);
    if (cnt > 0) {
      if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma((double)cnt, r[c], acc[c]);
      } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double sn, cs;
          sincospi(r[c] * 0.3183098861837907, &sn, &cs);  // angles in [0, 2 pi): no large-argument reduction needed (1e-16 abs error)
          acc[2 * c] = fma((double)cnt, cs, acc[2 * c]);
          acc[2 * c + 1] = fma((double)cnt, sn, acc[2 * c + 1]);
        }
      }
    }
    // few offspring: the owner stores them; many: the whole warp does
    if (cnt > 0 && cnt <= 2) {
      for (int64_t t = 0; t < cnt; ++t) {
        const int64_t k = sbeg - j_begin + t;
        if (x_in != nullptr) store_rec<T, D>(x_out, k, r);
        if (idx_out) idx_out[k] = i;
      }
    }
    // a degenerate weight distribution (one particle holding > 1e-4 of the weight): queued, copied by the whole grid
    if (cnt > big_cnt) {
      const unsigned int slot = atomicAdd(ws.counters + 6, 1u);
      if (slot < (unsigned int)kWsJobCap) { ws.jobs[3 * slot] = i; ws.jobs[3 * slot + 1] = sbeg - j_begin; ws.jobs[3 * slot + 2] = cnt; }
    }
    unsigned longmask = __ballot_sync(0xffffffffu, cnt > 2 && cnt <= big_cnt);
    while (longmask) {
      const int src = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int64_t s0 = __shfl_sync(0xffffffffu, sbeg, src) - j_begin;
      const int64_t c0 = __shfl_sync(0xffffffffu, cnt, src);
      const int64_t isrc = __shfl_sync(0xffffffffu, i, src);
      double rb[D];
#pragma unroll
      for (int c = 0; c < D; ++c) rb[c] = __shfl_sync(0xffffffffu, r[c], src);
      for (int64_t t = lane; t < c0; t += 32) {
        if (x_in != nullptr) store_rec<T, D>(x_out, s0 + t, rb);
        if (idx_out) idx_out[s0 + t] = isrc;
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
        for (int k = 0; k < NE; ++k) est_out[k] = t[k] / (double)count;
      }
    }
  }
}

// heavy particles: every queued run of offspring is copied by the whole grid
template <typename T, int D>
__global__ void __launch_bounds__(kThreads) systematic_fill_kernel(const T* __restrict__ x_in, T* __restrict__ x_out,
                                                                   int64_t* __restrict__ idx_out, Workspace ws) {
  const unsigned int nj_raw = *reinterpret_cast<volatile unsigned int*>(ws.counters + 6);
  const unsigned int nj = nj_raw < (unsigned int)kWsJobCap ? nj_raw : (unsigned int)kWsJobCap;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (unsigned int j = 0; j < nj; ++j) {
    const int64_t src = ws.jobs[3 * j], s0 = ws.jobs[3 * j + 1], cnt = ws.jobs[3 * j + 2];
    double r[D];
    if (x_in != nullptr) load_rec<T, D>(x_in, src, r);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += stride) {
      if (x_in != nullptr) store_rec<T, D>(x_out, s0 + t, r);
      if (idx_out) idx_out[s0 + t] = src;
    }
  }
  if (block_is_last(ws.counters + 7) && threadIdx.x == 0) ws.counters[6] = 0u;  // self-cleaning
}

// ---------------------------------------------------------------------------------------------
// Sharded filter, device-side plan: ONE launch per rank and step produces every offspring of the rank's own sources
// and stores each one straight into the buffer of the rank that owns its output slot (peer pointers over NVLink, or
// the rank's own buffer).  Nothing is read back to the host: the shard offsets O_g = fl(O_{g-1} + T_{g-1}) and the
// global total come from the all-gathered totals on the device, the comb offset u0 from a generator all ranks share.
// Global slot j lives on rank j / n_loc at row j % n_loc.  est_sums receives the SUMS over the offspring produced
// here (an all-reduce over the ranks makes the global estimate); status[0] <- 1 if the global total is not a positive
// finite number (nothing is written then).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxShards = 16;
struct PeerTable {
  void* out[kMaxShards];
};
struct ShardGeom {
  double offset, total;
};
__device__ __forceinline__ ShardGeom shard_geometry(const double* __restrict__ totals, int world, int rank) {
  ShardGeom g{0.0, 0.0};
  double o = 0.0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) g.offset = o;
    o = __dadd_rn(o, totals[r]);
  }
  g.total = o;
  return g;
}
template <typename T, int D>
__device__ __forceinline__ void store_slot(const PeerTable& peers, int64_t n_loc, int64_t slot, const double (&r)[D]) {
  int64_t rk = (int64_t)((double)slot / (double)n_loc);
  int64_t row = slot - rk * n_loc;
  if (row < 0) { --rk; row += n_loc; }
  else if (row >= n_loc) { ++rk; row -= n_loc; }
  store_rec<T, D>(static_cast<T*>(peers.out[rk]), row, r);
}

template <typename T, int D, int EST>
__global__ void __launch_bounds__(kThreads) sharded_scatter_kernel(int64_t n, const double* __restrict__ cdf,
                                                                   const double* __restrict__ totals, int world, int rank,
                                                                   double u0, PeerTable peers, const T* __restrict__ x_in,
                                                                   double* est_sums, double* status, Workspace ws,
                                                                   int64_t big_cnt) {
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  const ShardGeom geo = shard_geometry(totals, world, rank);
  const bool bad = !(geo.total > 0.0) || !(geo.total < INFINITY);
  if (bad) {
    if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = 1.0;
    return;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) status[0] = 0.0;
  const int64_t n_comb = n * world;
  const double dn = (double)n_comb;
  const double total = geo.total, cdf_offset = geo.offset;
  // this shard's offspring are the slots [J(O_g / total), J(O_{g+1} / total)) -- implied by the per-source counts below,
  // the first source starts where the previous shard's last source ended
  const int lane = threadIdx.x & 31;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  const int64_t nround = (n + 31) & ~(int64_t)31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool last_shard = rank == world - 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
    int64_t e = 0;  // first slot that no longer belongs to source i
    if (i < n) {
      if (i == n - 1 && last_shard) e = n_comb;
      else e = comb_count_below(__ddiv_rn(__dadd_rn(cdf_offset, __ldg(cdf + i)), total), u0, dn, n_comb);
    }
    int64_t sbeg = __shfl_up_sync(0xffffffffu, e, 1);
    if (lane == 0 && i < n) {
      const double cprev = (i == 0) ? cdf_offset : __dadd_rn(cdf_offset, __ldg(cdf + i - 1));
      sbeg = (i == 0 && rank == 0) ? 0 : comb_count_below(__ddiv_rn(cprev, total), u0, dn, n_comb);
    }
    const int64_t cnt = (i < n) ? e - sbeg : 0;
    double r[D];
#pragma unroll
    for (int c = 0; c < D; ++c) r[c] = 0.0;
    if (cnt > 0) {
      load_rec<T, D>(x_in, i, r);
      if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
        for (int c = 0; c < D; ++c) acc[c] = fma((double)cnt, r[c], acc[c]);
      } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
        for (int c = 0; c < D; ++c) {
          double sn, cs;
          sincospi(r[c] * 0.3183098861837907, &sn, &cs);
          acc[2 * c] = fma((double)cnt, cs, acc[2 * c]);
          acc[2 * c + 1] = fma((double)cnt, sn, acc[2 * c + 1]);
        }
      }
    }
    if (cnt > 0 && cnt <= 2) {
      for (int64_t t = 0; t < cnt; ++t) store_slot<T, D>(peers, n, sbeg + t, r);
    }
    if (cnt > big_cnt) {
      const unsigned int slot = atomicAdd(ws.counters + 6, 1u);
      if (slot < (unsigned int)kWsJobCap) { ws.jobs[3 * slot] = i; ws.jobs[3 * slot + 1] = sbeg; ws.jobs[3 * slot + 2] = cnt; }
    }
    unsigned longmask = __ballot_sync(0xffffffffu, cnt > 2 && cnt <= big_cnt);
    while (longmask) {
      const int src = __ffs(longmask) - 1;
      longmask &= longmask - 1;
      const int64_t s0 = __shfl_sync(0xffffffffu, sbeg, src);
      const int64_t c0 = __shfl_sync(0xffffffffu, cnt, src);
      double rb[D];
#pragma unroll
      for (int c = 0; c < D; ++c) rb[c] = __shfl_sync(0xffffffffu, r[c], src);
      for (int64_t t = lane; t < c0; t += 32) store_slot<T, D>(peers, n, s0 + t, rb);
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
        for (int k = 0; k < NE; ++k) est_sums[k] = t[k];
      }
    }
  }
}

// heavy particles of the sharded scatter: every queued run of offspring is written by the whole grid
template <typename T, int D>
__global__ void __launch_bounds__(kThreads) sharded_fill_kernel(int64_t n, const T* __restrict__ x_in, PeerTable peers,
                                                                Workspace ws) {
  const unsigned int nj_raw = *reinterpret_cast<volatile unsigned int*>(ws.counters + 6);
  const unsigned int nj = nj_raw < (unsigned int)kWsJobCap ? nj_raw : (unsigned int)kWsJobCap;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (unsigned int j = 0; j < nj; ++j) {
    const int64_t src = ws.jobs[3 * j], s0 = ws.jobs[3 * j + 1], cnt = ws.jobs[3 * j + 2];
    double r[D];
    load_rec<T, D>(x_in, src, r);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += stride) store_slot<T, D>(peers, n, s0 + t, r);
  }
  if (block_is_last(ws.counters + 7) && threadIdx.x == 0) ws.counters[6] = 0u;  // self-cleaning
}

template <typename T, int D>
static int launch_sharded_dev(int64_t n, const double* cdf, const double* totals, int world, int rank, double u0,
                              const PeerTable& peers, const void* x_in, int est_kind, double* est_sums, double* status,
                              Workspace W, cudaStream_t st) {
  const int grid = stream_grid(n);
  int64_t big_cnt = (n * world) >> 13;
  if (big_cnt < 8192) big_cnt = 8192;
#define PFB_SH_LAUNCH(E) sharded_scatter_kernel<T, D, E><<<grid, kThreads, 0, st>>>(n, cdf, totals, world, rank, u0, peers, (const T*)x_in, est_sums, status, W, big_cnt)
  switch (est_kind) {
    case PFB_EST_NONE: PFB_SH_LAUNCH(PFB_EST_NONE); break;
    case PFB_EST_SUM: PFB_SH_LAUNCH(PFB_EST_SUM); break;
    case PFB_EST_TRIG: PFB_SH_LAUNCH(PFB_EST_TRIG); break;
    default: set_error("bad est_kind %d", est_kind); return PFB_ERR_INVALID;
  }
#undef PFB_SH_LAUNCH
  sharded_fill_kernel<T, D><<<sm_count() * 4, kThreads, 0, st>>>(n, (const T*)x_in, peers, W);
  return PFB_OK;
}

// ---------------------------------------------------------------------------------------------
// Sharded filter, PULL form: every rank produces exactly its own n_loc output slots, reading the sources they come
// from wherever they live -- the CDF shard and the particle buffer of every rank are peer-mapped (torch symmetric
// memory), remote ones are read over NVLink with ordinary loads, all stores are local.  The push form above makes
// the rank that OWNS the heavy sources produce all their offspring: with the weight concentrated in one shard that
// rank does world x the work while the others idle.  Here the work per rank is its n_loc slots whatever the
// weights; what the skew costs is the NVLink read of the remote sources, inside the same kernel as the local work.
//   sharded_pull_plan_kernel   one warp per source shard: the range [a_s, b_s) of shard s's sources whose offspring
//                              intersect this rank's slots (two bisections on the peer's CDF)
//   sharded_pull_kernel        the sources of every non-empty range, one thread each (as in the scatter kernel),
//                              offspring clipped to the rank's slot range
// ---------------------------------------------------------------------------------------------
struct PeerInputs {
  const double* cdf[kMaxShards];
  const void* x[kMaxShards];
};
struct PullGeom {
  double offset[kMaxShards + 1];  // O_0 = 0, O_{s+1} = fl(O_s + T_s); offset[world] = the global total
};
__device__ __forceinline__ PullGeom pull_geometry(const double* __restrict__ totals, int world) {
  PullGeom g;
  double o = 0.0;
  for (int r = 0; r < world; ++r) { g.offset[r] = o; o = __dadd_rn(o, totals[r]); }
  g.offset[world] = o;
  return g;
}
// first slot that no longer belongs to source i of shard s
__device__ __forceinline__ int64_t pull_end_slot(const double* cdf_s, int64_t i, int64_t n, bool last_shard, double off,
                                                 double total, double u0, double dn, int64_t n_comb) {
  if (i == n - 1 && last_shard) return n_comb;
  return comb_count_below(__ddiv_rn(__dadd_rn(off, cdf_s[i]), total), u0, dn, n_comb);
}

__global__ void __launch_bounds__(32 * kMaxShards) sharded_pull_plan_kernel(int64_t n, PeerInputs in, const double* __restrict__ totals,
                                                                     int world, int rank, double u0, long long* plan,
                                                                     double* status) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const PullGeom geo = pull_geometry(totals, world);
  const double total = geo.offset[world];
  const bool bad = !(total > 0.0) || !(total < INFINITY);
  if (threadIdx.x == 0) status[0] = bad ? 1.0 : 0.0;
  if (warp >= world || lane != 0) return;
  const int s = warp;
  long long a = 0, b = 0;
  if (!bad) {
    const int64_t n_comb = n * world;
    const double dn = (double)n_comb;
    const int64_t S0 = (int64_t)rank * n, S1 = S0 + n;
    const double* cdf_s = in.cdf[s];
    const bool last = s == world - 1;
    auto E = [&](int64_t i) { return pull_end_slot(cdf_s, i, n, last, geo.offset[s], total, u0, dn, n_comb); };
    // first slot of the shard, and one past its last: empty intersections need no remote read at all
    const int64_t first_slot = s == 0 ? 0 : comb_count_below(__ddiv_rn(geo.offset[s], total), u0, dn, n_comb);
    const int64_t end_slot = last ? n_comb : comb_count_below(__ddiv_rn(geo.offset[s + 1], total), u0, dn, n_comb);
    if (first_slot < S1 && end_slot > S0) {
      int64_t lo = 0, hi = n;  // a = first i with E(i) > S0 (its offspring reach into the rank's slots)
      while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (E(mid) > S0) hi = mid; else lo = mid + 1; }
      a = lo;
      hi = n;                  // first i >= a with E(i) >= S1: the last source that can own a slot below S1
      while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (E(mid) >= S1) hi = mid; else lo = mid + 1; }
      b = lo < n ? lo + 1 : n;
    }
  }
  plan[2 * s] = a;
  plan[2 * s + 1] = b;
}

template <typename T, int D, int EST, int kU>
__global__ void __launch_bounds__(kThreads) sharded_pull_kernel(int64_t n, PeerInputs in, const double* __restrict__ totals,
                                                                int world, int rank, double u0,
                                                                const long long* __restrict__ plan, T* __restrict__ x_out,
                                                                double* est_sums, const double* __restrict__ status,
                                                                Workspace ws, int64_t big_cnt) {
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  __shared__ double red[NE * 32];
  if (status[0] != 0.0) return;  // (the plan kernel found the global total invalid: nothing is written)
  const PullGeom geo = pull_geometry(totals, world);
  const double total = geo.offset[world];
  const int64_t n_comb = n * world;
  const double dn = (double)n_comb;
  const int64_t S0 = (int64_t)rank * n, S1 = S0 + n;
  const int lane = threadIdx.x & 31;
  double acc[NE];
#pragma unroll
  for (int k = 0; k < NE; ++k) acc[k] = 0.0;
  // A warp takes kU x 32 consecutive sources per round and issues the loads of all kU groups together: a remote
  // source costs an NVLink round trip, and one 8-byte load per thread in flight keeps only ~1/3 of the link's
  // bandwidth-delay product busy.
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int s = 0; s < world; ++s) {
    const int64_t a = plan[2 * s], b = plan[2 * s + 1];
    if (b <= a) continue;
    const double* cdf_s = in.cdf[s];
    const T* x_s = static_cast<const T*>(in.x[s]);
    const double off = geo.offset[s];
    const bool last = s == world - 1;
    const int64_t i0 = a & ~(int64_t)31;  // rounds stay aligned to 32 sources
    const int64_t nchunks = (b - i0 + 32 * kU - 1) / (32 * kU);
    for (int64_t ch = warp_global; ch < nchunks; ch += nwarps) {
      const int64_t base = i0 + ch * (32 * kU);
      double cv[kU];
      bool mine[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t i = base + u * 32 + lane;
        mine[u] = i >= a && i < b;
        cv[u] = mine[u] ? cdf_s[i] : 0.0;
      }
      // the slot where the chunk's first source starts (lane 0 only needs it)
      int64_t prev_end = 0;
      {
        const int64_t ifirst = base > a ? base : a;  // first source of this chunk that is mine
        if (lane == 0) {
          prev_end = (ifirst == 0) ? (s == 0 ? 0 : comb_count_below(__ddiv_rn(off, total), u0, dn, n_comb))
                                   : pull_end_slot(cdf_s, ifirst - 1, n, false, off, total, u0, dn, n_comb);
        }
      }
      int64_t e[kU], row0[kU], cnt[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t i = base + u * 32 + lane;
        e[u] = 0;
        if (mine[u]) e[u] = (i == n - 1 && last) ? n_comb : comb_count_below(__ddiv_rn(__dadd_rn(off, cv[u]), total), u0, dn, n_comb);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t i = base + u * 32 + lane;
        int64_t sbeg = __shfl_up_sync(0xffffffffu, e[u], 1);
        const int64_t carry = __shfl_sync(0xffffffffu, u > 0 ? e[u > 0 ? u - 1 : 0] : prev_end, u > 0 ? 31 : 0);
        const int64_t first_start = __shfl_sync(0xffffffffu, prev_end, 0);  // (every lane takes part in the shuffle)
        if (lane == 0) sbeg = carry;
        if (i == a) sbeg = first_start;  // (the range starts inside a round: the lanes below are not mine)
        const int64_t c0 = sbeg < S0 ? S0 : sbeg, c1 = e[u] > S1 ? S1 : e[u];
        cnt[u] = mine[u] && c1 > c0 ? c1 - c0 : 0;
        row0[u] = c0 - S0;
      }
      double r[kU][D];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
#pragma unroll
        for (int c = 0; c < D; ++c) r[u][c] = 0.0;
        if (cnt[u] > 0) load_rec_inplace<T, D>(x_s, base + u * 32 + lane, r[u]);  // (ordinary loads: a peer's memory)
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (cnt[u] > 0) {
          if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] = fma((double)cnt[u], r[u][c], acc[c]);
          } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double sn, cs;
              sincospi(r[u][c] * 0.3183098861837907, &sn, &cs);
              acc[2 * c] = fma((double)cnt[u], cs, acc[2 * c]);
              acc[2 * c + 1] = fma((double)cnt[u], sn, acc[2 * c + 1]);
            }
          }
        }
        if (cnt[u] > 0 && cnt[u] <= 2) {
          for (int64_t t = 0; t < cnt[u]; ++t) store_rec<T, D>(x_out, row0[u] + t, r[u]);
        }
        if (cnt[u] > big_cnt) {
          const unsigned int slot = atomicAdd(ws.counters + 6, 1u);
          if (slot < (unsigned int)(kWsJobCap - 16)) {
            ws.jobs[3 * slot] = ((long long)s << 48) | (base + u * 32 + lane);
            ws.jobs[3 * slot + 1] = row0[u];
            ws.jobs[3 * slot + 2] = cnt[u];
          }
        }
        unsigned longmask = __ballot_sync(0xffffffffu, cnt[u] > 2 && cnt[u] <= big_cnt);
        while (longmask) {
          const int src = __ffs(longmask) - 1;
          longmask &= longmask - 1;
          const int64_t s0 = __shfl_sync(0xffffffffu, row0[u], src);
          const int64_t k0 = __shfl_sync(0xffffffffu, cnt[u], src);
          double rb[D];
#pragma unroll
          for (int c = 0; c < D; ++c) rb[c] = __shfl_sync(0xffffffffu, r[u][c], src);
          for (int64_t t = lane; t < k0; t += 32) store_rec<T, D>(x_out, s0 + t, rb);
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
      for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) {
#pragma unroll
        for (int k = 0; k < NE; ++k) t[k] += ws.partials[(int64_t)bb * kPartialDoubles + k];
      }
      block_sum<NE>(t, red);
      if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NE; ++k) est_sums[k] = t[k];
      }
    }
  }
}

// heavy particles of the pull kernel: every queued run of offspring is written by the whole grid
template <typename T, int D>
__global__ void __launch_bounds__(kThreads) sharded_pull_fill_kernel(PeerInputs in, T* __restrict__ x_out, Workspace ws) {
  const unsigned int nj_raw = *reinterpret_cast<volatile unsigned int*>(ws.counters + 6);
  const unsigned int nj = nj_raw < (unsigned int)(kWsJobCap - 16) ? nj_raw : (unsigned int)(kWsJobCap - 16);  // (the tail holds the plan)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (unsigned int j = 0; j < nj; ++j) {
    const long long tag = ws.jobs[3 * j];
    const int64_t row0 = ws.jobs[3 * j + 1], cnt = ws.jobs[3 * j + 2];
    double r[D];
    load_rec_inplace<T, D>(static_cast<const T*>(in.x[tag >> 48]), tag & 0xFFFFFFFFFFFFLL, r);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < cnt; t += stride) store_rec<T, D>(x_out, row0 + t, r);
  }
  if (block_is_last(ws.counters + 7) && threadIdx.x == 0) ws.counters[6] = 0u;  // self-cleaning
}

template <typename T, int D>
static int launch_sharded_pull(int64_t n, const PeerInputs& in, const double* totals, int world, int rank, double u0,
                               void* x_out, int est_kind, double* est_sums, double* status, Workspace W, cudaStream_t st) {
  long long* plan = reinterpret_cast<long long*>(W.jobs) + 3 * (int64_t)kWsJobCap - 2 * kMaxShards;  // tail of the job list
  sharded_pull_plan_kernel<<<1, 32 * kMaxShards, 0, st>>>(n, in, totals, world, rank, u0, plan, status);
  const int grid = stream_grid(n);
  int64_t big_cnt = (n * world) >> 13;
  if (big_cnt < 8192) big_cnt = 8192;
  static const int unroll = getenv("PFB_PULL_UNROLL") ? atoi(getenv("PFB_PULL_UNROLL")) : 4;
#define PFB_PULL_LAUNCH(E)                                                                                                          \
  if (unroll == 1) sharded_pull_kernel<T, D, E, 1><<<grid, kThreads, 0, st>>>(n, in, totals, world, rank, u0, plan, (T*)x_out, est_sums, status, W, big_cnt); \
  else sharded_pull_kernel<T, D, E, 4><<<grid, kThreads, 0, st>>>(n, in, totals, world, rank, u0, plan, (T*)x_out, est_sums, status, W, big_cnt)
  switch (est_kind) {
    case PFB_EST_NONE: PFB_PULL_LAUNCH(PFB_EST_NONE); break;
    case PFB_EST_SUM: PFB_PULL_LAUNCH(PFB_EST_SUM); break;
    case PFB_EST_TRIG: PFB_PULL_LAUNCH(PFB_EST_TRIG); break;
    default: set_error("bad est_kind %d", est_kind); return PFB_ERR_INVALID;
  }
#undef PFB_PULL_LAUNCH
  sharded_pull_fill_kernel<T, D><<<sm_count() * 4, kThreads, 0, st>>>(in, (T*)x_out, W);
  return PFB_OK;
}

template <typename T, int D>
static int launch_systematic(int64_t n, const double* cdf, double cdf_offset, double total_imm, double u0_imm,
                             const double* u0_ptr, RngKey key, int u0_mode, int64_t j_begin, int64_t count,
                             int64_t n_comb, const void* x_in, void* x_out, int64_t* idx_out, int est_kind,
                             double* est_out, Workspace W, cudaStream_t st) {
  const int grid = stream_grid(n);
  // heavy sources go through the job list of the workspace (at most n_comb / big_cnt <= 8192 of them); without a
  // workspace their warp writes them
  const bool have_jobs = W.jobs != nullptr && W.counters != nullptr;
  int64_t big_cnt = n_comb >> 13;
  if (big_cnt < 8192) big_cnt = 8192;
  if (!have_jobs) big_cnt = (int64_t)1 << 30;
#define PFB_SYS_LAUNCH(E) systematic_scatter_kernel<T, D, E><<<grid, kThreads, 0, st>>>(n, cdf, cdf_offset, total_imm, u0_imm, u0_ptr, key, u0_mode, j_begin, count, n_comb, (const T*)x_in, (T*)x_out, idx_out, est_out, W, big_cnt)
  switch (est_kind) {
    case PFB_EST_NONE: PFB_SYS_LAUNCH(PFB_EST_NONE); break;
    case PFB_EST_SUM: PFB_SYS_LAUNCH(PFB_EST_SUM); break;
    case PFB_EST_TRIG: PFB_SYS_LAUNCH(PFB_EST_TRIG); break;
    default: set_error("bad est_kind %d", est_kind); return PFB_ERR_INVALID;
  }
#undef PFB_SYS_LAUNCH
  if (have_jobs) systematic_fill_kernel<T, D><<<sm_count() * 4, kThreads, 0, st>>>((const T*)x_in, (T*)x_out, idx_out, W);
  return PFB_OK;
}

template <typename T, int D>
static int launch_resample_rec(int64_t n, int64_t n_out, const double* cdf, const double* u, const void* x_in,
                               void* x_out, int64_t* idx_out, int est_kind, double* est_out, Workspace W,
                               cudaStream_t st, const uint4* recs, int64_t M, RngKey key, int64_t n_runs = 1,
                               int64_t run_pad = 0, int64_t n_out_run = 0, int64_t x_in_stride = 0) {
  if (n_runs <= 1) { n_runs = 1; run_pad = n; n_out_run = n_out; }
  if (x_in_stride <= 0) x_in_stride = D;
  const int grid = stream_grid(n_out);
  switch (est_kind) {
    case PFB_EST_NONE:
      resample_rec_kernel<T, D, PFB_EST_NONE><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, (const T*)x_in, (T*)x_out, idx_out, est_out, W, recs, M, key, n_runs, run_pad, n_out_run, x_in_stride);
      break;
    case PFB_EST_SUM:
      resample_rec_kernel<T, D, PFB_EST_SUM><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, (const T*)x_in, (T*)x_out, idx_out, est_out, W, recs, M, key, n_runs, run_pad, n_out_run, x_in_stride);
      break;
    case PFB_EST_TRIG:
      resample_rec_kernel<T, D, PFB_EST_TRIG><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, (const T*)x_in, (T*)x_out, idx_out, est_out, W, recs, M, key, n_runs, run_pad, n_out_run, x_in_stride);
      break;
    default:
      set_error("bad est_kind %d", est_kind);
      return PFB_ERR_INVALID;
  }
  return PFB_OK;
}

template <typename T, int D>
static int launch_resample(int64_t n, int64_t n_out, const double* cdf, const double* u, int systematic,
                           const double* stats, const void* x_in, void* x_out, int64_t* idx_out, int est_kind,
                           double* est_out, Workspace W, cudaStream_t st, const int32_t* guide, int64_t M, RngKey key,
                           int64_t run_pad, int64_t run_len, int64_t n_out_run) {
  const int grid = stream_grid(n_out);
  switch (est_kind) {
    case PFB_EST_NONE:
      resample_kernel<T, D, PFB_EST_NONE><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M, key, run_pad, run_len, n_out_run);
      break;
    case PFB_EST_SUM:
      resample_kernel<T, D, PFB_EST_SUM><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M, key, run_pad, run_len, n_out_run);
      break;
    case PFB_EST_TRIG:
      resample_kernel<T, D, PFB_EST_TRIG><<<grid, kThreads, 0, st>>>(n, n_out, cdf, u, systematic, stats, (const T*)x_in, (T*)x_out, idx_out, est_out, W, guide, M, key, run_pad, run_len, n_out_run);
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
  int64_t b = pfb::rec_offset(n) + pfb::rec_bytes(n);
  return (b + 255) & ~(int64_t)255;
}

extern "C" int pfb_workspace_init(void* ws, int64_t ws_bytes, void* stream) {
  if (!ws || ws_bytes < pfb_workspace_bytes(0)) { set_error("workspace too small"); return PFB_ERR_WORKSPACE; }
  PFB_CUDA_OK(cudaMemsetAsync(ws, 0, (size_t)kWsCounters, (cudaStream_t)stream));
  return PFB_OK;
}

namespace pfb {
// bucket records of one CDF in the workspace (record_build_kernel + record_fill_kernel, or the two-pass cross-check)
int launch_record_build(int64_t n, const double* cdf, void* ws, const Workspace& W, cudaStream_t st, const uint4** recs_out,
                        int64_t* Mr_out) {
  const int64_t Mr = rec_buckets(n);  // the int32 guide of the Mr buckets is scratch for the record build
  uint4* recs = reinterpret_cast<uint4*>(static_cast<unsigned char*>(ws) + rec_offset(n));
  if (two_pass_record_build()) {
    guide_kernel<<<stream_grid((n + 3) / 4, 2), kThreads, 0, st>>>(n, cdf, W.guide, Mr, n);
    PFB_LAUNCH_OK("guide_kernel");
    bucket_record_kernel<<<stream_grid(Mr + 1, 2), kThreads, 0, st>>>(n, cdf, W.guide, Mr, recs);
    PFB_LAUNCH_OK("bucket_record_kernel");
  } else {  // (the big-gap job list lives in the guide area: at most Mr / kBigGap + 1 jobs of 3 ints, and none if Mr <= kBigGap)
    const int64_t tiles = (n + kRecTile - 1) / kRecTile;
    const int64_t cap = (int64_t)sm_count() * 4;
    record_build_kernel<<<(int)(tiles < cap ? tiles : cap), kThreads, 0, st>>>(1, n, n, cdf, Mr, recs, W.guide, W.counters + 4);
    PFB_LAUNCH_OK("record_build_kernel");
    record_fill_kernel<<<sm_count() * 4, kThreads, 0, st>>>(recs, W.guide, W.counters + 4, W.counters + 5);
    PFB_LAUNCH_OK("record_fill_kernel");
  }
  *recs_out = recs;
  *Mr_out = Mr;
  return PFB_OK;
}

static int resample_impl(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf, const double* u,
                         RngKey key, int32_t systematic, const void* x_in, void* x_out, int64_t* idx_out,
                         int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes, void* stream,
                         int32_t x_in_stride = 0) {
  if (x_in_stride != 0 && x_in_stride != dim) {
    if (x_in_stride < dim || systematic) { set_error("x_in_stride %d: must be >= dim, multinomial resampling only", x_in_stride); return PFB_ERR_INVALID; }
    if (!bucket_records_enabled()) { set_error("strided gather sources need the bucket-record path"); return PFB_ERR_UNSUPPORTED; }
  }
  if (n <= 0 || n_out <= 0) return PFB_OK;
  if (!cdf) { set_error("NULL cdf"); return PFB_ERR_INVALID; }
  if ((x_in == nullptr) != (x_out == nullptr)) { set_error("x_in and x_out must both be given or both NULL"); return PFB_ERR_INVALID; }
  if (x_in && x_in == x_out) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE && (!est_out || !x_in)) { set_error("estimate requested without output / particles"); return PFB_ERR_INVALID; }
  if (n >= ((int64_t)1 << 31)) { set_error("more than 2^31-1 particles per device are not supported"); return PFB_ERR_UNSUPPORTED; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, n, &W);
  if (rc) return rc;
  // the raw total c_{n-1} is read on the device from the CDF itself
  cudaStream_t st = (cudaStream_t)stream;
  if (systematic) {  // monotone comb: tiled merge, no guide table
    PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_systematic<T, D>(n, cdf, 0.0, 0.0, 0.0, u, key, u ? 1 : 2, 0, n_out, n_out, x_in, x_out, idx_out, est_kind, est_out, W, st))));
    if (rc) return rc;
    PFB_LAUNCH_OK("systematic_scatter_kernel");
    return PFB_OK;
  }
  if (bucket_records_enabled() && rec_buckets(n) <= W.guide_buckets) {
    const uint4* recs = nullptr;
    int64_t Mr = 0;
    rc = launch_record_build(n, cdf, ws, W, st, &recs, &Mr);
    if (rc) return rc;
    PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_resample_rec<T, D>(n, n_out, cdf, u, x_in, x_out, idx_out, est_kind, est_out, W, st, recs, Mr, key, 1, 0, 0, x_in_stride))));
    if (rc) return rc;
    PFB_LAUNCH_OK("resample_rec_kernel");
    return PFB_OK;
  }
  const double* stats_total = cdf + (n - 1);
  const int64_t M = W.guide_buckets;
  guide_kernel<<<stream_grid((n + 3) / 4, 2), kThreads, 0, st>>>(n, cdf, W.guide, M, n);
  PFB_LAUNCH_OK("guide_kernel");
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_resample<T, D>(n, n_out, cdf, u, systematic, stats_total, x_in, x_out, idx_out, est_kind, est_out, W, st, W.guide, M, key, n, n, n_out))));
  if (rc) return rc;
  PFB_LAUNCH_OK("resample_kernel");
  return PFB_OK;
}

__global__ void __launch_bounds__(kThreads) rng_uniform_kernel(int64_t n, RngKey key, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint4 r = rng_block(key, (uint64_t)j, 7u);
    out[j] = u53(r.x, r.y);
  }
}
}  // namespace pfb

extern "C" int pfb_resample(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                            const double* u, int32_t systematic, const void* x_in, void* x_out,
                            int64_t* idx_out, int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes,
                            void* stream) {
  if (n > 0 && n_out > 0 && !u) { set_error("NULL uniforms (use pfb_resample_rng for in-kernel draws)"); return PFB_ERR_INVALID; }
  return resample_impl(dtype, dim, n, n_out, cdf, u, RngKey{0, 0}, systematic, x_in, x_out, idx_out, est_kind, est_out,
                       ws, ws_bytes, stream);
}

extern "C" int pfb_resample_rng(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                                uint64_t rng_seed, uint64_t rng_offset, int32_t systematic, const void* x_in,
                                void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out, void* ws,
                                int64_t ws_bytes, void* stream) {
  return resample_impl(dtype, dim, n, n_out, cdf, nullptr, RngKey{rng_seed, rng_offset}, systematic, x_in, x_out,
                       idx_out, est_kind, est_out, ws, ws_bytes, stream);
}

extern "C" int pfb_resample_strided(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                                    const double* u, const void* x_in, int32_t x_in_stride, void* x_out,
                                    int64_t* idx_out, int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes,
                                    void* stream) {
  if (n > 0 && n_out > 0 && !u) { set_error("NULL uniforms (use pfb_resample_rng_strided for in-kernel draws)"); return PFB_ERR_INVALID; }
  return resample_impl(dtype, dim, n, n_out, cdf, u, RngKey{0, 0}, 0, x_in, x_out, idx_out, est_kind, est_out, ws,
                       ws_bytes, stream, x_in_stride);
}

extern "C" int pfb_resample_rng_strided(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                                        uint64_t rng_seed, uint64_t rng_offset, const void* x_in, int32_t x_in_stride,
                                        void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out, void* ws,
                                        int64_t ws_bytes, void* stream) {
  return resample_impl(dtype, dim, n, n_out, cdf, nullptr, RngKey{rng_seed, rng_offset}, 0, x_in, x_out, idx_out,
                       est_kind, est_out, ws, ws_bytes, stream, x_in_stride);
}

extern "C" int64_t pfb_workspace_bytes_batched(int64_t n_runs, int64_t run_len) {
  if (n_runs < 1) n_runs = 1;
  if (run_len < 0) run_len = 0;
  const int64_t n_pad = n_runs * scan_tiles(run_len) * kScanTile;
  int64_t b = guide_offset(n_pad) + n_runs * (guide_buckets(run_len) + 2) * (int64_t)sizeof(int32_t);
  b = (b + 255) & ~(int64_t)255;
  b += n_runs * (pfb::rec_buckets(run_len) + 2) * 64;  // per-run bucket records
  b = (b + 255) & ~(int64_t)255;
  const int64_t single = pfb_workspace_bytes(n_pad);  // the weighting / scan calls size themselves on n_pad
  return b > single ? b : single;
}

extern "C" int pfb_resample_batched(pfb_dtype dtype, int32_t dim, int64_t n_runs, int64_t run_len, int64_t n_out_run,
                                    const double* cdf_pad, const double* u, uint64_t rng_seed, uint64_t rng_offset,
                                    int32_t systematic, const void* x_in, void* x_out, int64_t* idx_out, void* ws,
                                    int64_t ws_bytes, void* stream) {
  if (n_runs <= 0 || run_len <= 0 || n_out_run <= 0) return PFB_OK;
  if (!cdf_pad) { set_error("NULL cdf"); return PFB_ERR_INVALID; }
  if ((x_in == nullptr) != (x_out == nullptr)) { set_error("x_in and x_out must both be given or both NULL"); return PFB_ERR_INVALID; }
  if (x_in && x_in == x_out) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
  const int64_t run_pad = scan_tiles(run_len) * kScanTile;
  const int64_t n_pad = n_runs * run_pad;
  if (ws_bytes < pfb_workspace_bytes_batched(n_runs, run_len)) { set_error("workspace too small for %lld runs of %lld", (long long)n_runs, (long long)run_len); return PFB_ERR_WORKSPACE; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  int32_t* guide = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(ws) + guide_offset(n_pad));
  const int64_t M = guide_buckets(run_len);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_out = n_runs * n_out_run;
  if (!systematic && bucket_records_enabled()) {  // multinomial: per-run bucket records (the job list lives in the guide area)
    const int64_t Mr = rec_buckets(run_len);
    const int64_t goff = guide_offset(n_pad) + n_runs * (M + 2) * (int64_t)sizeof(int32_t);
    uint4* recs = reinterpret_cast<uint4*>(static_cast<unsigned char*>(ws) + ((goff + 255) & ~(int64_t)255));
    const int64_t tiles = n_runs * scan_tiles(run_len);
    const int64_t cap = (int64_t)sm_count() * 4;
    record_build_kernel<<<(int)(tiles < cap ? tiles : cap), kThreads, 0, st>>>(n_runs, run_len, run_pad, cdf_pad, Mr, recs, guide, W.counters + 4);
    PFB_LAUNCH_OK("record_build_kernel");
    record_fill_kernel<<<sm_count() * 4, kThreads, 0, st>>>(recs, guide, W.counters + 4, W.counters + 5);
    PFB_LAUNCH_OK("record_fill_kernel");
    PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_resample_rec<T, D>(run_len, n_out, cdf_pad, u, x_in, x_out, idx_out, PFB_EST_NONE, nullptr, W, st, recs, Mr, RngKey{rng_seed, rng_offset}, n_runs, run_pad, n_out_run))));
    if (rc) return rc;
    PFB_LAUNCH_OK("resample_rec_kernel");
    return PFB_OK;
  }
  guide_kernel<<<stream_grid((n_pad + 3) / 4, 2), kThreads, 0, st>>>(n_pad, cdf_pad, guide, M, run_pad);
  PFB_LAUNCH_OK("guide_kernel");
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_resample<T, D>(n_pad, n_out, cdf_pad, u, systematic, cdf_pad, x_in, x_out, idx_out, PFB_EST_NONE, nullptr, W, st, guide, M, RngKey{rng_seed, rng_offset}, run_pad, run_len, n_out_run))));
  if (rc) return rc;
  PFB_LAUNCH_OK("resample_kernel");
  return PFB_OK;
}

extern "C" int pfb_moments_batched(int32_t kind, pfb_dtype dtype, int32_t dim, int64_t n_runs, int64_t run_len,
                                   const void* x, int32_t order, double* out, void* stream) {
  if (n_runs <= 0 || run_len <= 0) return PFB_OK;
  if (!x || !out) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  int rc = PFB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_moments_batched<T, D>(kind, n_runs, run_len, x, order, out, st))));
  if (rc) return rc;
  PFB_LAUNCH_OK("moments_batched_kernel");
  return PFB_OK;
}

extern "C" int pfb_resample_sharded(pfb_dtype dtype, int32_t dim, int64_t n, const double* cdf, double cdf_offset,
                                    double total, double u0, int64_t j_begin, int64_t count, int64_t n_total_out,
                                    const void* x_in, void* x_out, int64_t* idx_out, int32_t est_kind,
                                    double* est_out, void* ws, int64_t ws_bytes, void* stream) {
  if (n <= 0 || count <= 0) return PFB_OK;
  if (!cdf) { set_error("NULL cdf"); return PFB_ERR_INVALID; }
  if ((x_in == nullptr) != (x_out == nullptr)) { set_error("x_in and x_out must both be given or both NULL"); return PFB_ERR_INVALID; }
  if (!(total > 0.0) || n_total_out <= 0) { set_error("total weight and comb size must be positive"); return PFB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  Workspace W{};
  int rc = PFB_OK;
  if (est_kind != PFB_EST_NONE && (!est_out || !x_in)) { set_error("estimate requested without output / particles"); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE || ws != nullptr) {  // (the workspace also carries the job list of heavy particles)
    rc = carve_workspace(ws, ws_bytes, 0, &W);
    if (rc) return rc;
  }
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_systematic<T, D>(n, cdf, cdf_offset, total, u0, nullptr, RngKey{0, 0}, 0, j_begin, count, n_total_out, x_in, x_out, idx_out, est_kind, est_out, W, st))));
  if (rc) return rc;
  PFB_LAUNCH_OK("systematic_scatter_kernel");
  return PFB_OK;
}

extern "C" int pfb_resample_sharded_dev(pfb_dtype dtype, int32_t dim, int64_t n, const double* cdf,
                                        const double* totals_dev, int32_t world, int32_t rank, double u0,
                                        void* const* peer_out, const void* x_in, int32_t est_kind, double* est_sums,
                                        double* status, void* ws, int64_t ws_bytes, void* stream) {
  if (n <= 0) return PFB_OK;
  if (!cdf || !totals_dev || !peer_out || !x_in || !status) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  if (world < 1 || world > kMaxShards || rank < 0 || rank >= world) { set_error("world %d / rank %d: 1..%d shards", world, rank, kMaxShards); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE && !est_sums) { set_error("estimate requested without output"); return PFB_ERR_INVALID; }
  if (n * (int64_t)world >= ((int64_t)1 << 40)) { set_error("comb too long"); return PFB_ERR_UNSUPPORTED; }
  PeerTable peers{};
  for (int r = 0; r < world; ++r) {
    if (!peer_out[r]) { set_error("peer_out[%d] is NULL", r); return PFB_ERR_INVALID; }
    if (peer_out[r] == x_in) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
    peers.out[r] = peer_out[r];
  }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_sharded_dev<T, D>(n, cdf, totals_dev, world, rank, u0, peers, x_in, est_kind, est_sums, status, W, st))));
  if (rc) return rc;
  PFB_LAUNCH_OK("sharded_scatter_kernel");
  return PFB_OK;
}

extern "C" int pfb_resample_sharded_pull(pfb_dtype dtype, int32_t dim, int64_t n, const double* const* peer_cdf,
                                         const void* const* peer_x, const double* totals_dev, int32_t world,
                                         int32_t rank, double u0, void* x_out, int32_t est_kind, double* est_sums,
                                         double* status, void* ws, int64_t ws_bytes, void* stream) {
  if (n <= 0) return PFB_OK;
  if (!peer_cdf || !peer_x || !totals_dev || !x_out || !status || !ws) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  if (world < 1 || world > kMaxShards || rank < 0 || rank >= world) { set_error("world %d / rank %d: 1..%d shards", world, rank, kMaxShards); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE && !est_sums) { set_error("estimate requested without output"); return PFB_ERR_INVALID; }
  if (n * (int64_t)world >= ((int64_t)1 << 40)) { set_error("comb too long"); return PFB_ERR_UNSUPPORTED; }
  PeerInputs in{};
  for (int r = 0; r < world; ++r) {
    if (!peer_cdf[r] || !peer_x[r]) { set_error("peer_cdf[%d] / peer_x[%d] is NULL", r, r); return PFB_ERR_INVALID; }
    if (peer_x[r] == x_out) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
    in.cdf[r] = peer_cdf[r];
    in.x[r] = peer_x[r];
  }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_sharded_pull<T, D>(n, in, totals_dev, world, rank, u0, x_out, est_kind, est_sums, status, W, st))));
  if (rc) return rc;
  PFB_LAUNCH_OK("sharded_pull_kernel");
  return PFB_OK;
}

extern "C" int pfb_rng_uniform(int64_t n, uint64_t rng_seed, uint64_t rng_offset, double* out, void* stream) {
  if (n <= 0) return PFB_OK;
  if (!out) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  rng_uniform_kernel<<<stream_grid(n), kThreads, 0, (cudaStream_t)stream>>>(n, RngKey{rng_seed, rng_offset}, out);
  PFB_LAUNCH_OK("rng_uniform_kernel");
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
