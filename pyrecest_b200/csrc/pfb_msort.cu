// K(4b'), "multinomial_sorted": exact multinomial resampling whose offspring come out in CDF order.
//
// The reference draws n_out i.i.d. uniforms and emits x[searchsorted(cdf, u_j)] at slot j
// (pyrecest/distributions/abstract_dirac_distribution.py:82-84, pyrecest/_backend/_shared_numpy/random.py:28-29):
// a random gather whose two dependent DRAM misses per particle bound pfb_resample at ~0.16 of the HBM roofline.
// With in-kernel Philox draws there is no reference stream to match, only the distribution: n_out i.i.d. uniforms.
// The same distribution is produced here in a different ORDER:
//   * [0, 1) is cut into K = 2^k equal leaves (about 16..32 offspring each).  How many of the n_out uniforms fall into
//     each leaf is multinomial(n_out; 1/K, ..) -- drawn exactly by binomial splitting with p = 1/2: a node that holds m
//     uniforms gives popcount(first m bits of its own Philox bit stream) to its left child (Binomial(m, 1/2) is the
//     number of ones among m fair bits: exact, no floating point);
//   * the uniforms of leaf b are b/K + r/K with fresh (53 - k)-bit draws r: conditioned on the leaf, a 53-bit uniform
//     is exactly that;
//   * offspring are emitted leaf by leaf: slot = (exclusive prefix of the leaf counts) + t.
// The result is a permutation of what n_out unsorted uniforms would give (every later step treats the particles as an
// exchangeable set), the record / CDF / particle accesses of neighbouring threads fall into the same cache lines
// and the stores stream.  oracle/philox.py (msort_*) restates the tree and the draws bit for bit.
//
//   msort_top_kernel       levels [0, k_top) of the tree: every node's bit stream is popcounted by the whole grid
//                          (cooperative launch, grid barrier between levels; the work per level is n_out bits);
//                          leaves counts + exclusive offsets of the 2^k_top coarse bins.  Independent of the data.
//   msort_resample_kernel  one CTA per coarse bin: the remaining k - k_top <= 9 levels in shared memory, then one
//                          thread per offspring: leaf, uniform, index by bisection of the CDF inside the leaf's own
//                          source range (a handful of entries, cache resident), gather, store, fused estimate.
#include "pfb_resample.cuh"

namespace pfb {

constexpr int kMsLeafTarget = 32;  // expected offspring per leaf in (16, 32]
constexpr int kMsSubMax = 9;       // levels expanded inside a CTA (512 leaves per coarse bin)
constexpr int kMsSubMin = 4;
constexpr uint32_t kMsSlotTree = 64u;  // Philox slot of tree level l: 64 + l
constexpr uint32_t kMsSlotLeaf = 33u;  // Philox slot of the in-leaf draws

struct MsPlan {
  int k;      // leaf level: K = 2^k leaves
  int k_top;  // levels done by msort_top_kernel: 2^k_top coarse bins
};
static MsPlan ms_plan(int64_t n_out) {
  MsPlan p{0, 0};
  while (((int64_t)kMsLeafTarget << p.k) < n_out) ++p.k;
  if (p.k > kMsSubMin) {
    const int a = p.k - kMsSubMax, b = (p.k - kMsSubMin < kMsSubMax) ? p.k - kMsSubMin : kMsSubMax;
    p.k_top = a > b ? a : b;
  }
  return p;
}

// number of ones among the first `bits` (1..128) bits of a Philox block, words in the order x, y, z, w, low bit first
__device__ __forceinline__ int popc_prefix(const uint4& r, int bits) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  int pc = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int b = bits - 32 * j;
    if (b >= 32) pc += __popc(w[j]);
    else if (b > 0) pc += __popc(w[j] & ((1u << b) - 1u));
  }
  return pc;
}

// all CTAs of a cooperative launch; `bar` counts arrivals monotonically (zeroed by the host before the launch)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int* phase) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    *phase += 1u;
    const unsigned int target = *phase * gridDim.x;
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (seen < target) __nanosleep(32);
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

// cnt / left: level l occupies [2^l - 1, 2^(l+1) - 1)
__global__ void __launch_bounds__(kThreads) msort_top_kernel(int64_t n_out, int k_top, RngKey key, int32_t* cnt,
                                                             int32_t* left, int32_t* offs, unsigned int* bar) {
  __shared__ unsigned int s_phase;
  __shared__ int32_t s_scan[kThreads / 32 + 1];
  if (threadIdx.x == 0) s_phase = 0u;
  const int lane = threadIdx.x & 31;
  const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t gwarp = gtid >> 5, gwarps = gthreads >> 5;
  const int64_t internal = ((int64_t)1 << k_top) - 1;
  for (int64_t i = gtid; i < internal; i += gthreads) left[i] = 0;
  if (gtid == 0) cnt[0] = (int32_t)n_out;
  if (k_top > 0) grid_barrier(bar, &s_phase);
  for (int l = 0; l < k_top; ++l) {
    const int64_t N = (int64_t)1 << l, base = N - 1;
    // chunks of 32 Philox blocks (4096 bits, one block per lane); a node holds n_out / 2^l +- a few sigma items:
    // the last chunk of a node keeps going until the node's bits are used up, whatever its count
    const double m_est = (double)n_out / (double)N;
    const int64_t cpn = (int64_t)((m_est + 12.0 * sqrt(m_est) + 64.0) / 4096.0) + 1;
    // every warp takes a contiguous run of items, so that it adds into a node once, not once per chunk (the chunks
    // of the few nodes near the root would otherwise queue tens of thousands of atomics on one address)
    const int64_t items = N * cpn, ipw = (items + gwarps - 1) / gwarps;
    const int64_t w_end = (gwarp + 1) * ipw < items ? (gwarp + 1) * ipw : items;
    int64_t node_acc = -1;
    int acc = 0;
    for (int64_t w = gwarp * ipw; w < w_end; ++w) {
      const int64_t i = w / cpn, q = w - i * cpn;
      // the node's own count from its parent (final since the last barrier); the q == 0 warp records it
      int64_t m = n_out;
      if (l > 0) {
        const int64_t pi = (N >> 1) - 1 + (i >> 1);
        const int32_t pm = __ldcg(cnt + pi), pl = __ldcg(left + pi);
        m = (i & 1) ? pm - pl : pl;
      }
      if (q == 0 && lane == 0) cnt[base + i] = (int32_t)m;
      if (i != node_acc) {
        if (node_acc >= 0) {
          acc = __reduce_add_sync(0xffffffffu, acc);
          if (lane == 0 && acc != 0) atomicAdd(left + base + node_acc, acc);
        }
        node_acc = i;
        acc = 0;
      }
      for (int64_t start = q * 4096; start < m; start += 4096) {
        const int64_t mine = m - start - (int64_t)lane * 128;
        if (mine > 0) {
          const uint4 r = rng_block(key, ((uint64_t)i << 32) | (uint64_t)(start / 128 + lane), kMsSlotTree + (uint32_t)l);
          acc += popc_prefix(r, mine >= 128 ? 128 : (int)mine);
        }
        if (q != cpn - 1) break;
      }
    }
    if (node_acc >= 0) {
      acc = __reduce_add_sync(0xffffffffu, acc);
      if (lane == 0 && acc != 0) atomicAdd(left + base + node_acc, acc);
    }
    grid_barrier(bar, &s_phase);
  }
  // counts of the coarse bins (level k_top) from their parents
  {
    const int64_t N = (int64_t)1 << k_top;
    if (k_top > 0) {
      for (int64_t i = gtid; i < N; i += gthreads) {
        const int64_t pi = (N >> 1) - 1 + (i >> 1);
        const int32_t pm = __ldcg(cnt + pi), pl = __ldcg(left + pi);
        cnt[N - 1 + i] = (i & 1) ? pm - pl : pl;
      }
      grid_barrier(bar, &s_phase);
    }
  }
  // exclusive offsets of the coarse bins
  if (blockIdx.x == 0) {
    const int64_t KC = (int64_t)1 << k_top;
    const int32_t* c = cnt + (KC - 1);
    int32_t carry = 0;
    for (int64_t b0 = 0; b0 < KC; b0 += kThreads) {
      const int64_t i = b0 + threadIdx.x;
      const int32_t v = i < KC ? __ldcg(c + i) : 0;
      int32_t inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncthreads();
      if (lane == 31) s_scan[threadIdx.x >> 5] = inc;
      __syncthreads();
      int32_t woff = 0, tot = 0;
#pragma unroll
      for (int wv = 0; wv < kThreads / 32; ++wv) {
        const int32_t sv = s_scan[wv];
        if (wv < (int)(threadIdx.x >> 5)) woff += sv;
        tot += sv;
      }
      if (i < KC) offs[i] = carry + woff + inc - v;
      carry += tot;
    }
    if (threadIdx.x == 0) offs[KC] = carry;
  }
}

// sin and cos of an angle in [0, 2 pi] for the fused circular estimate.  x = k (2 pi / 256) + r with |r| <= pi / 256:
// (cos, sin) of the grid angle come from a 256-entry table in shared memory, sin r and cos r from three-term
// series (truncation < 1e-18), combined with the angle-sum formulas: ~20 fp64 instructions against ~75 for sincos().
// Errors stay below 2 ulp, far inside the 1e-12 the estimate has to meet.
constexpr int kTrigTab = 256;
__device__ __forceinline__ void trig_table_fill(double2* tab) {
  for (int i = threadIdx.x; i < kTrigTab; i += blockDim.x) {
    double sn, cs;
    sincospi((double)i * (2.0 / kTrigTab), &sn, &cs);
    tab[i] = make_double2(cs, sn);
  }
}
// (fp64 literals cost two UMOVs each next to the DFMA that uses them; from constant memory a pair comes with one LDCU)
__constant__ double kMsConst[12] = {
    kTrigTab / 6.283185307179586, 0.02454369260617026, 9.567553118338697e-19,          // 0: 256 / 2 pi, 2 pi / 256 hi + tail
    8.333333333333333e-3,         -1.6666666666666666e-1,                                 // 3: sin series
    -1.3888888888888889e-3,       4.1666666666666664e-2, -0.5,                            // 5: cos series
    1.0 - 8.881784197001252e-16,  1.0 + 8.881784197001252e-16, 1e-290, 1.1102230246251565e-16};  // 8: band, 2^-53
__device__ __forceinline__ void sincos_tab(const double2* __restrict__ tab, double x, double* sn, double* cs) {
  const double kq = rint(x * kMsConst[0]);
  double r = fma(-kq, kMsConst[1], x);                        // (exact product: kq < 2^9)
  r = fma(-kq, kMsConst[2], r);
  const double2 t = tab[(int)kq & (kTrigTab - 1)];
  const double z = r * r;
  const double s = fma(r * z, fma(z, kMsConst[3], kMsConst[4]), r);
  const double c = fma(z, fma(z, fma(z, kMsConst[5], kMsConst[6]), kMsConst[7]), 1.0);
  *cs = fma(t.x, c, -(t.y * s));
  *sn = fma(t.y, c, t.x * s);
}

// fl(c / total) <= q (LE) or < q (!LE), decided without the division unless c lies within 2^-50 (relative) of
// q * total:  c <= t (1 - 2^-50)  =>  c / total < q (1 - 2^-51)  =>  fl(c / total) < q,  and symmetrically above.
struct CdfCut {
  double lo, hi, q, total;
};
__device__ __forceinline__ CdfCut cdf_cut(double q, double total) {
  const double t = q * total;
  CdfCut k{t * kMsConst[8], t * kMsConst[9], q, total};
  if (!(t > kMsConst[10])) { k.lo = -INFINITY; k.hi = INFINITY; }  // tiny / zero products: always the exact test
  return k;
}
// (out of line on purpose: inlined, the compiler turns the band test into a predicate and runs the division for
//  every lane whose c is above the lower cut -- half of all comparisons)
template <bool LE>
__device__ __noinline__ bool cdf_below_exact(double c, double total, double q) {
  const double v = __ddiv_rn(c, total);
  return LE ? (v <= q) : (v < q);
}
template <bool LE>
__device__ __forceinline__ bool cdf_below(const CdfCut& k, double c) {
  if (c <= k.lo) return true;
  if (c >= k.hi) return false;
  return cdf_below_exact<LE>(c, k.total, k.q);
}
// first i in [lo, hi) whose normalised CDF value is not below the cut (LE: count of v_i <= q; !LE: of v_i < q)
template <bool LE>
__device__ __forceinline__ int64_t cdf_search(const double* __restrict__ cdf, int64_t lo, int64_t hi, const CdfCut& k) {
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (cdf_below<LE>(k, __ldg(cdf + mid))) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// L2 prefetch of the CDF entries and particle records of sources [ia, ib): one 128-byte line of each per thread
// (a hint for the next round of a CTA, ~600 sources: 40 lines of CDF, 120 of particles; longer ranges are cut)
template <typename T>
__device__ __forceinline__ void prefetch_sources(const double* __restrict__ cdf, const T* __restrict__ x_in,
                                                 int x_in_stride, int ia, int ib) {
  const uintptr_t pc = (reinterpret_cast<uintptr_t>(cdf + ia) & ~(uintptr_t)127) + (uintptr_t)threadIdx.x * 128;
  if (pc < reinterpret_cast<uintptr_t>(cdf + ib)) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc));
  if (x_in != nullptr) {
    const uintptr_t px = (reinterpret_cast<uintptr_t>(x_in + (int64_t)ia * x_in_stride) & ~(uintptr_t)127) + (uintptr_t)threadIdx.x * 128;
    if (px < reinterpret_cast<uintptr_t>(x_in + (int64_t)ib * x_in_stride)) asm volatile("prefetch.global.L2 [%0];" ::"l"(px));
  }
}

// source index range of every coarse bin: G[B] = #{i : fl(c_i / total) < B / KC}; the offspring of bin B have
// indices in [G[B], G[B + 1]]
__global__ void __launch_bounds__(kThreads) msort_guide_kernel(int64_t n, const double* __restrict__ cdf, int k_top,
                                                               int32_t* __restrict__ guide) {
  const int64_t KC = (int64_t)1 << k_top;
  const double total = __ldg(cdf + n - 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t B = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; B <= KC; B += stride) {
    int64_t g = n;
    if (B < KC) g = B == 0 ? 0 : cdf_search<false>(cdf, 0, n, cdf_cut((double)B / (double)KC, total));
    guide[B] = (int32_t)g;
  }
}

template <typename T, int D, int EST>
__global__ void __launch_bounds__(kThreads, 4) msort_resample_kernel(
    int64_t n, int64_t n_out, const double* __restrict__ cdf, const T* __restrict__ x_in, int64_t x_in_stride,
    T* __restrict__ x_out, int64_t* __restrict__ idx_out, double* est_out, Workspace ws, RngKey key, int k, int k_top,
    const int32_t* __restrict__ cnt_top, const int32_t* __restrict__ offs, const int32_t* __restrict__ guide) {
  constexpr int NE = (EST == PFB_EST_SUM) ? D : (EST == PFB_EST_TRIG ? 2 * D : 1);
  constexpr int kLeavesMax = 1 << kMsSubMax;
  __shared__ double red[NE * 32];
  __shared__ int32_t s_cnt[2][kLeavesMax];
  __shared__ int32_t s_left[kLeavesMax / 2];
  __shared__ int32_t s_off[kLeavesMax + 1];
  __shared__ int32_t s_src[kLeavesMax + 1];  // first source index of every leaf (+ the end of the bin)
  __shared__ int32_t s_red[kThreads / 32];
  __shared__ double2 s_trig[EST == PFB_EST_TRIG ? kTrigTab : 1];
  if constexpr (EST == PFB_EST_TRIG) trig_table_fill(s_trig);  // (visible after the first barrier of the bin loop)
  const int sub = k - k_top;
  const int NL = 1 << sub;
  const int64_t KC = (int64_t)1 << k_top;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int low_bits = 53 - k;
  const double total = __ldg(cdf + n - 1);
  const double dK = (double)((int64_t)1 << k);
  double acc[NE];
#pragma unroll
  for (int q = 0; q < NE; ++q) acc[q] = 0.0;
  for (int64_t B = blockIdx.x; B < KC; B += gridDim.x) {
    const int32_t m0 = __ldg(cnt_top + B);
    const int64_t o0 = (int64_t)__ldg(offs + B);
    const int64_t g0 = (int64_t)__ldg(guide + B), g1 = (int64_t)__ldg(guide + B + 1);
    __syncthreads();  // (the previous bin's tables are no longer read)
    if (threadIdx.x == 0) s_cnt[0][0] = m0;
    if (m0 > 0) {  // the sources of the first round, by a linear estimate inside the bin's range
      const int64_t est = g0 + (g1 - g0) * (2 * kThreads + 64) / m0 + 64;
      prefetch_sources<T>(cdf, x_in, (int)x_in_stride, (int)g0, (int)(est < g1 ? est : g1));
    }
    // first source of every leaf: v_i < b / K is monotone in i, bisection inside the bin's range (independent of the
    // count tree: the loads are in flight while the tree is built)
    for (int lb = threadIdx.x; lb <= NL; lb += kThreads) {
      int64_t g = lb == 0 ? g0 : g1;
      if (lb > 0 && lb < NL) g = cdf_search<false>(cdf, g0, g1, cdf_cut((double)(B * NL + lb) / dK, total));
      s_src[lb] = (int32_t)g;
    }
    int cur = 0;
    // the remaining levels of the count tree, in shared memory
    for (int j = 0; j < sub; ++j) {
      const int Nj = 1 << j;
      __syncthreads();
      int mmax = 0;
      for (int i = threadIdx.x; i < Nj; i += kThreads) { mmax = max(mmax, s_cnt[cur][i]); s_left[i] = 0; }
      mmax = __reduce_max_sync(0xffffffffu, mmax);
      if (lane == 0) s_red[warp] = mmax;
      __syncthreads();
      mmax = 0;
#pragma unroll
      for (int wv = 0; wv < kThreads / 32; ++wv) mmax = max(mmax, s_red[wv]);
      const int bpn = (mmax + 127) >> 7;  // Philox blocks per node
      for (int f = threadIdx.x; f < Nj * bpn; f += kThreads) {
        const int i = f / bpn, q = f - i * bpn;
        const int mine = s_cnt[cur][i] - q * 128;
        if (mine > 0) {
          const uint4 r = rng_block(key, ((uint64_t)(B * Nj + i) << 32) | (uint64_t)q, kMsSlotTree + (uint32_t)(k_top + j));
          atomicAdd(&s_left[i], popc_prefix(r, mine >= 128 ? 128 : mine));
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < Nj; i += kThreads) {
        const int32_t m = s_cnt[cur][i], L = s_left[i];
        s_cnt[cur ^ 1][2 * i] = L;
        s_cnt[cur ^ 1][2 * i + 1] = m - L;
      }
      cur ^= 1;
    }
    __syncthreads();
    // exclusive offsets of the NL leaves (two per thread)
    {
      const int a = (2 * (int)threadIdx.x < NL) ? s_cnt[cur][2 * threadIdx.x] : 0;
      const int b = (2 * (int)threadIdx.x + 1 < NL) ? s_cnt[cur][2 * threadIdx.x + 1] : 0;
      int inc = a + b;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) s_red[warp] = inc;
      __syncthreads();
      int woff = 0;
#pragma unroll
      for (int wv = 0; wv < kThreads / 32; ++wv)
        if (wv < warp) woff += s_red[wv];
      const int excl = woff + inc - (a + b);
      if (2 * (int)threadIdx.x < NL) s_off[2 * threadIdx.x] = excl;
      if (2 * (int)threadIdx.x + 1 < NL) s_off[2 * threadIdx.x + 1] = excl + a;
      if (threadIdx.x == kThreads - 1) s_off[NL] = woff + inc;
    }
    __syncthreads();
    // two adjacent offspring per thread: slots (2 j, 2 j + 1) of the bin share one Philox block
    const float leaf_per_slot = (float)NL / (float)(m0 > 0 ? m0 : 1);
    const uint64_t leaf0 = (uint64_t)B * (uint64_t)NL;
    const int nm1 = (int)n - 1;
    const int xs = (int)x_in_stride;
    for (int s0 = 2 * (int)threadIdx.x; s0 < m0; s0 += 2 * kThreads) {
      {  // software prefetch of the sources of the NEXT round of the CTA: demand misses of a dependent search
         // chain cannot keep enough lines in flight to stream from DRAM
        const int sa = s0 - 2 * (int)threadIdx.x + 2 * kThreads;
        if (sa < m0) {
          int la = (int)((float)sa * leaf_per_slot) - 2, lb = (int)((float)(sa + 2 * kThreads) * leaf_per_slot) + 3;
          la = la < 0 ? 0 : la;
          lb = lb > NL ? NL : lb;
          prefetch_sources<T>(cdf, x_in, xs, s_src[la], s_src[lb]);
        }
      }
      int idx[2];
      bool live[2];
      int sbase[2], slen[2];
      CdfCut cut[2];
      const uint4 blk = rng_block(key, ((uint64_t)B << 32) | (uint64_t)(s0 >> 1), kMsSlotLeaf);
      int lf = (int)((float)s0 * leaf_per_slot);  // leaf offsets are close to linear in the slot: guess, then walk
      lf = lf > NL - 1 ? NL - 1 : lf;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int s = s0 + q;
        live[q] = s < m0;
        sbase[q] = 0;
        slen[q] = 0;
        cut[q] = CdfCut{0.0, 0.0, 0.0, 1.0};
        if (!live[q]) continue;
        while (s_off[lf] > s) --lf;
        while (s_off[lf + 1] <= s) ++lf;
        const uint64_t r64 = q ? (((uint64_t)blk.z << 32) | blk.w) : (((uint64_t)blk.x << 32) | blk.y);
        const uint64_t R = ((leaf0 + (uint64_t)lf) << low_bits) | (r64 >> (11 + k));
        cut[q] = cdf_cut((double)R * kMsConst[11], total);
        sbase[q] = s_src[lf];
        slen[q] = s_src[lf + 1] - sbase[q];
      }
      // both bisections in one loop: their loads are issued together
      while ((slen[0] | slen[1]) > 0) {
        double c[2];
        int half[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          half[q] = slen[q] >> 1;
          const int at = sbase[q] + half[q];
          c[q] = __ldg(cdf + (at < nm1 ? at : nm1));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (slen[q] > 0) {
            if (cdf_below<true>(cut[q], c[q])) { sbase[q] += half[q] + 1; slen[q] -= half[q] + 1; }
            else slen[q] = half[q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) idx[q] = sbase[q] > nm1 ? nm1 : sbase[q];
      if (x_in != nullptr) {
        double r[2][D];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (!live[q]) continue;
          const T* src = x_in + (int64_t)idx[q] * xs;
#pragma unroll
          for (int c = 0; c < D; ++c) r[q][c] = (double)__ldg(src + c);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          if (!live[q]) continue;
          store_rec<T, D>(x_out, o0 + s0 + q, r[q]);
          if constexpr (EST == PFB_EST_SUM) {
#pragma unroll
            for (int c = 0; c < D; ++c) acc[c] += r[q][c];
          } else if constexpr (EST == PFB_EST_TRIG) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
              double sn, cs;
              if (r[q][c] >= 0.0 && r[q][c] <= 6.3) sincos_tab(s_trig, r[q][c], &sn, &cs);
              else sincos(r[q][c], &sn, &cs);  // (particles of a torus filter are wrapped: not taken there)
              acc[2 * c] += cs;
              acc[2 * c + 1] += sn;
            }
          }
        }
      }
      if (idx_out) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
          if (live[q]) idx_out[o0 + s0 + q] = idx[q];
      }
    }
  }
  if constexpr (EST != PFB_EST_NONE) finish_estimate<NE>(acc, red, ws, est_out, n_out);
}

// writes the uniforms of the sorted-order draw in slot order (tests / debugging): the same tree and leaf draws
__global__ void __launch_bounds__(kThreads) msort_uniform_kernel(int64_t n_out, RngKey key, int k, int k_top,
                                                                 const int32_t* __restrict__ cnt_top,
                                                                 const int32_t* __restrict__ offs,
                                                                 double* __restrict__ u_out) {
  // (a plain restatement: every thread walks its own path down the sub-tree; only used on small inputs)
  const int sub = k - k_top;
  const int64_t KC = (int64_t)1 << k_top;
  for (int64_t B = blockIdx.x; B < KC; B += gridDim.x) {
    const int32_t m0 = cnt_top[B];
    const int64_t o0 = offs[B];
    for (int s = threadIdx.x; s < m0; s += blockDim.x) {
      // descend: at every level the slot either stays in the left child (s < L) or moves to the right one
      int64_t node = B;
      int32_t m = m0, pos = s;
      for (int j = 0; j < sub; ++j) {
        int32_t L = 0;
        for (int32_t start = 0; start < m; start += 128) {
          const uint4 r = rng_block(key, ((uint64_t)node << 32) | (uint64_t)(start / 128), kMsSlotTree + (uint32_t)(k_top + j));
          L += popc_prefix(r, m - start >= 128 ? 128 : m - start);
        }
        if (pos < L) { node = 2 * node; m = L; }
        else { node = 2 * node + 1; pos -= L; m -= L; }
      }
      const uint64_t b = (uint64_t)node;
      const uint4 r = rng_block(key, ((uint64_t)B << 32) | (uint64_t)(s >> 1), kMsSlotLeaf);  // slots (2 j, 2 j + 1) of a bin
      const uint64_t r64 = (s & 1) ? (((uint64_t)r.z << 32) | r.w) : (((uint64_t)r.x << 32) | r.y);
      const uint64_t R = (b << (53 - k)) | (r64 >> (11 + k));
      u_out[o0 + s] = (double)R * 1.1102230246251565e-16;
    }
  }
}

static int launch_top(int64_t n_out, RngKey key, const MsPlan& pl, const Workspace& W, cudaStream_t st) {
  int32_t* cnt = W.ms;
  int32_t* left = W.ms + kMsTreeInts;
  int32_t* offs = W.ms + 2 * kMsTreeInts;
  unsigned int* bar = W.counters + 8;
  static int resident = 0;
  if (resident == 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, msort_top_kernel, kThreads, 0) != cudaSuccess || occ < 1) occ = 1;
    resident = (occ > 2 ? 2 : occ) * sm_count();
  }
  // enough warps for the items of one level (n_out / 4096 chunks + one per node)
  int64_t want = ((n_out >> 12) + ((int64_t)1 << pl.k_top)) / (kThreads / 32) + 1;
  int grid = (int)(want < resident ? want : resident);
  if (pl.k_top == 0) grid = 1;
  PFB_CUDA_OK(cudaMemsetAsync(bar, 0, 4, st));
  int k_top = pl.k_top;
  void* args[] = {&n_out, &k_top, &key, &cnt, &left, &offs, &bar};
  if (grid > 1) {
    PFB_CUDA_OK(cudaLaunchCooperativeKernel((const void*)msort_top_kernel, dim3(grid), dim3(kThreads), args, 0, st));
  } else {
    msort_top_kernel<<<1, kThreads, 0, st>>>(n_out, k_top, key, cnt, left, offs, bar);
    PFB_LAUNCH_OK("msort_top_kernel");
  }
  return PFB_OK;
}

template <typename T, int D>
static int launch_msort(int64_t n, int64_t n_out, const double* cdf, const void* x_in, int64_t x_in_stride, void* x_out,
                        int64_t* idx_out, int est_kind, double* est_out, const Workspace& W, cudaStream_t st,
                        RngKey key, const MsPlan& pl) {
  const int64_t KC = (int64_t)1 << pl.k_top;
  const int32_t* cnt_top = W.ms + (KC - 1);
  const int32_t* offs = W.ms + 2 * kMsTreeInts;
  int32_t* guide = W.ms + kMsTreeInts;  // (the left-child accumulators of the top tree are dead by now)
  msort_guide_kernel<<<(int)((KC + 1 + kThreads - 1) / kThreads), kThreads, 0, st>>>(n, cdf, pl.k_top, guide);
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(KC < cap ? KC : cap);
  if (x_in_stride <= 0) x_in_stride = D;
#define PFB_MS_LAUNCH(E) msort_resample_kernel<T, D, E><<<grid, kThreads, 0, st>>>(n, n_out, cdf, (const T*)x_in, x_in_stride, (T*)x_out, idx_out, est_out, W, key, pl.k, pl.k_top, cnt_top, offs, guide)
  switch (est_kind) {
    case PFB_EST_NONE: PFB_MS_LAUNCH(PFB_EST_NONE); break;
    case PFB_EST_SUM: PFB_MS_LAUNCH(PFB_EST_SUM); break;
    case PFB_EST_TRIG: PFB_MS_LAUNCH(PFB_EST_TRIG); break;
    default: set_error("bad est_kind %d", est_kind); return PFB_ERR_INVALID;
  }
#undef PFB_MS_LAUNCH
  return PFB_OK;
}

}  // namespace pfb

using namespace pfb;

extern "C" int pfb_resample_msort(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                                  uint64_t rng_seed, uint64_t rng_offset, const void* x_in, int32_t x_in_stride,
                                  void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out, void* ws,
                                  int64_t ws_bytes, void* stream) {
  if (n <= 0 || n_out <= 0) return PFB_OK;
  if (!cdf) { set_error("NULL cdf"); return PFB_ERR_INVALID; }
  if ((x_in == nullptr) != (x_out == nullptr)) { set_error("x_in and x_out must both be given or both NULL"); return PFB_ERR_INVALID; }
  if (x_in && x_in == x_out) { set_error("resampling cannot run in place"); return PFB_ERR_INVALID; }
  if (!x_in && !idx_out) { set_error("nothing to produce: give particles and / or idx_out"); return PFB_ERR_INVALID; }
  if (est_kind != PFB_EST_NONE && (!est_out || !x_in)) { set_error("estimate requested without output / particles"); return PFB_ERR_INVALID; }
  if (x_in_stride != 0 && x_in_stride < dim) { set_error("x_in_stride %d: must be 0 or >= dim", x_in_stride); return PFB_ERR_INVALID; }
  if (n >= ((int64_t)1 << 31) || n_out >= ((int64_t)1 << 31)) { set_error("more than 2^31-1 particles per device are not supported"); return PFB_ERR_UNSUPPORTED; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, n, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const RngKey key{rng_seed, rng_offset};
  const MsPlan pl = ms_plan(n_out);
  rc = launch_top(n_out, key, pl, W, st);
  if (rc) return rc;
  PFB_DISPATCH_DTYPE(dtype, PFB_DISPATCH_DIM(dim, (rc = launch_msort<T, D>(n, n_out, cdf, x_in, x_in_stride, x_out, idx_out, est_kind, est_out, W, st, key, pl))));
  if (rc) return rc;
  PFB_LAUNCH_OK("msort_resample_kernel");
  return PFB_OK;
}

extern "C" int pfb_msort_uniforms(int64_t n_out, uint64_t rng_seed, uint64_t rng_offset, double* out, void* ws,
                                  int64_t ws_bytes, void* stream) {
  if (n_out <= 0) return PFB_OK;
  if (!out) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  if (n_out >= ((int64_t)1 << 31)) { set_error("more than 2^31-1 draws are not supported"); return PFB_ERR_UNSUPPORTED; }
  Workspace W;
  int rc = carve_workspace(ws, ws_bytes, 0, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const RngKey key{rng_seed, rng_offset};
  const MsPlan pl = ms_plan(n_out);
  rc = launch_top(n_out, key, pl, W, st);
  if (rc) return rc;
  const int64_t KC = (int64_t)1 << pl.k_top;
  const int64_t cap = (int64_t)sm_count() * 8;
  msort_uniform_kernel<<<(int)(KC < cap ? KC : cap), kThreads, 0, st>>>(n_out, key, pl.k, pl.k_top, W.ms + (KC - 1), W.ms + 2 * kMsTreeInts, out);
  PFB_LAUNCH_OK("msort_uniform_kernel");
  return PFB_OK;
}
