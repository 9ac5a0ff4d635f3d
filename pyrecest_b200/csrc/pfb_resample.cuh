// Device helpers shared by the resampling kernels (pfb_resample.cu, pfb_msort.cu).
#pragma once
#include "pfb_common.cuh"

namespace pfb {

// (the exact test is out of line on purpose: inlined, the compiler folds the band test into a predicate and runs the
//  division for every lane whose c is above the lower cut)
static __device__ __noinline__ bool cdf_le_exact(double c, double total, double u) { return __ddiv_rn(c, total) <= u; }
__device__ __forceinline__ bool cdf_le(double c, double total, double u, double tl, double th) {
  if (c <= tl) return true;
  if (c >= th) return false;
  return cdf_le_exact(c, total, u);
}

template <bool RIGHT>
__device__ __forceinline__ int64_t search_range(const double* __restrict__ cdf, int64_t lo, int64_t hi,
                                                double total, double q) {
  // RIGHT: first i in [lo, hi) with v_i > q (count of v_i <= q);  LEFT: first i with v_i >= q
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    const double v = __ddiv_rn(__ldg(cdf + mid), total);
    const bool go_right = RIGHT ? (v <= q) : (v < q);
    if (go_right) lo = mid + 1; else hi = mid;
  }
  return lo;
}

constexpr int kRecShift = 4;
constexpr int kRecEntries = 29;

inline int64_t rec_buckets(int64_t n) {
  int64_t want = n >> kRecShift;
  int64_t M = 1;
  while (M < want) M <<= 1;
  return M;
}
inline int64_t rec_offset(int64_t n) {
  return (guide_offset(n) + (guide_buckets(n) + 2) * (int64_t)sizeof(int32_t) + 255) & ~(int64_t)255;
}
inline int64_t rec_bytes(int64_t n) { return (rec_buckets(n) + 2) * 64; }

// One 32-byte sector in ONE load instruction (LDG.256, 32-byte aligned address).
__device__ __forceinline__ void ld_sector(const void* p, uint64_t (&w)[4]) {
  asm volatile("ld.global.nc.L2::64B.v4.u64 {%0, %1, %2, %3}, [%4];"
               : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3])
               : "l"(p));
}

// Random gather of one particle record.  Component-wise loads of a record send one request per component
// to L2 for what is one (or two) 32-byte sectors, and L2 does not merge misses in flight: the 24-byte T^3
// record cost ~150 bytes of DRAM reads.  Here every covering sector is loaded exactly once (sector-aligned,
// so the reads stay inside the 256-byte allocation granule) and the components are selected from registers.
template <typename T, int D>
__device__ __forceinline__ void gather_rec(const T* __restrict__ x, int64_t i, double (&r)[D], int64_t stride = D) {
  constexpr int R = D * (int)sizeof(T);
  constexpr int WPS = 32 / (int)sizeof(T);                  // components per sector
  constexpr int NS = (R - (int)sizeof(T)) / 32 + 2;         // sectors a record can touch
  const uintptr_t a = reinterpret_cast<uintptr_t>(x + i * stride);
  const int off = (int)(a & 31) / (int)sizeof(T);
  const unsigned char* base = reinterpret_cast<const unsigned char*>(a & ~(uintptr_t)31);
  uint64_t raw[NS][4];
#pragma unroll
  for (int sct = 0; sct < NS; ++sct) {
    if (sct == 0 || off * (int)sizeof(T) + R > 32 * sct) ld_sector(base + 32 * sct, raw[sct]);
    else raw[sct][0] = raw[sct][1] = raw[sct][2] = raw[sct][3] = 0ull;
  }
  auto word = [&](int k) -> double {  // component k of the sector-aligned span (k is a compile-time constant)
    if constexpr (sizeof(T) == 8) {
      return __longlong_as_double((long long)raw[k / 4][k % 4]);
    } else {
      const uint64_t v = raw[k / 8][(k % 8) / 2];
      return (double)__uint_as_float((k & 1) ? (uint32_t)(v >> 32) : (uint32_t)v);
    }
  };
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double v = word(c);
#pragma unroll
    for (int o = 1; o < WPS; ++o)
      if (c + o < NS * WPS && off == o) v = word(c + o);
    r[c] = v;
  }
}

// deterministic grid reduction of the fused estimate: per-CTA partials, the last CTA sums them in a fixed order
template <int NE>
__device__ __forceinline__ void finish_estimate(double (&acc)[NE], double* red, Workspace ws, double* est_out,
                                                int64_t n_out) {
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
      for (int k = 0; k < NE; ++k) est_out[k] = t[k] / (double)n_out;
    }
  }
}

// Decode one 64-byte bucket record (two sectors already in registers) for a uniform u whose in-bucket position is
// p (p2 = p | p << 16): index = #{i : fl(c_i / total) <= u}, clamped to [0, n - 1].  The exact search in the CDF runs
// only for a tie on the 16-bit grid or a crowded bucket; its bounds are clamped so that a corrupt record (a run
// whose total is 0 or NaN leaves records unwritten) can never send a load out of the CDF.
__device__ __forceinline__ int64_t record_decode(const uint64_t (&s0)[4], const uint64_t (&s1)[4], uint32_t p2,
                                                 const uint4* __restrict__ recs_run, int64_t bkt,
                                                 const double* __restrict__ cdf_run, int64_t n, double u) {
  uint32_t w[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    w[2 * k] = (uint32_t)s0[k];
    w[2 * k + 1] = (uint32_t)(s0[k] >> 32);
    w[8 + 2 * k] = (uint32_t)s1[k];
    w[8 + 2 * k + 1] = (uint32_t)(s1[k] >> 32);
  }
  const int32_t base = (int32_t)w[0];
  const uint32_t cnt = w[1] & 0xFFFFu;
  uint32_t below = __vcmpltu2(w[1], p2) & 0xFFFF0000u;
  uint32_t eq = __vcmpeq2(w[1], p2) & 0xFFFF0000u;
  int nb = __popc(below);
#pragma unroll
  for (int k = 2; k < 16; ++k) {
    nb += __popc(__vcmpltu2(w[k], p2));
    eq |= __vcmpeq2(w[k], p2);
  }
  int64_t r = (int64_t)base + (nb >> 4);
  if (eq != 0u || cnt > (uint32_t)kRecEntries) {  // a tie in the 16-bit grid, or a crowded bucket: exact search
    int64_t lo = (int64_t)base, hi = (int64_t)base + cnt;
    if (cnt == 65535u) hi = (int64_t)(int32_t)__ldg(&recs_run[(bkt + 1) * 4].x);
    lo = lo < 0 ? 0 : (lo > n ? n : lo);
    hi = hi < lo ? lo : (hi > n ? n : hi);
    r = search_range<true>(cdf_run, lo, hi, __ldg(cdf_run + n - 1), u);
  }
  return r < 0 ? 0 : (r > n - 1 ? n - 1 : r);
}

// host side (pfb_resample.cu): build the bucket records of one CDF in the workspace
int launch_record_build(int64_t n, const double* cdf, void* ws, const Workspace& W, cudaStream_t st, const uint4** recs_out,
                        int64_t* Mr_out);

}  // namespace pfb
