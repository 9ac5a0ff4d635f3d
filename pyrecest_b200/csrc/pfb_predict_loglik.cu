// K(1) predict, K(2) log-likelihood weighting, and their fusion.
//
// One thread per particle record (row-major (n, D)), grid-stride.  Arithmetic is float64
// whatever the storage type.  Operation order follows oracle/pf_oracle.py so that the pure
// add / numpy.mod paths are bit-exact (the library is compiled with -fmad=false; the fused
// multiply-adds that are wanted are written explicitly).
#include <mutex>
#include <unordered_map>

#include "pfb_common.cuh"

namespace pfb {

// ---------------------------------------------------------------------------------------------
// predict
// ---------------------------------------------------------------------------------------------
// PLAIN (compile time): identity transition without offset and noise rows used as drawn (no factor, no mean) -- the
// torus / sphere filters with von Mises(-Fisher) noise.  The small matrix products behind the run-time flags are
// if-converted by ptxas: predicated off they still issue (~70 instructions per T^3 particle).
template <int D, bool PLAIN>
__device__ __forceinline__ void apply_transition(const pfb_predict_params& p, const double (&x)[D],
                                                 double (&y)[D]) {
  if constexpr (PLAIN) {
#pragma unroll
    for (int r = 0; r < D; ++r) y[r] = x[r];
    return;
  }
  if (p.has_F) {
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double acc = p.F[r * D] * x[0];
#pragma unroll
      for (int c = 1; c < D; ++c) acc = fma(p.F[r * D + c], x[c], acc);
      y[r] = acc;
    }
  } else {
#pragma unroll
    for (int r = 0; r < D; ++r) y[r] = x[r];
  }
  if (p.has_b) {
#pragma unroll
    for (int r = 0; r < D; ++r) y[r] += p.b[r];
  }
}

// noise row in state space: rows as given, or z A (+ mu) for standard-normal input.  RNG: the draws come from the
// in-kernel generator (p.rng_kind); a compile-time switch, so that the array-fed kernels do not carry the
// samplers' registers.
template <typename T, int D, bool RNG, bool PLAIN>
__device__ __forceinline__ void load_noise(const pfb_predict_params& p, const T* __restrict__ noise,
                                           int64_t i, double (&n)[D]) {
  double z[D];
  if constexpr (!RNG) {
    load_rec<T, D>(noise, i, z);
  } else if (p.rng_kind == PFB_RNG_NORMAL) {
    const RngKey key{p.rng_seed, p.rng_offset};
#pragma unroll
    for (int c = 0; c < D; c += 2) {
      double z0, z1;
      normal_pair(rng_block(key, (uint64_t)i, (uint32_t)(c >> 1)), &z0, &z1);
      z[c] = z0;
      if (c + 1 < D) z[c + 1] = z1;
    }
  } else if (p.rng_kind == PFB_RNG_VONMISES) {
    const RngKey key{p.rng_seed, p.rng_offset};
    vonmises_draws<D>(key, (uint64_t)i, p, z);
  }
  if constexpr (sizeof(T) == 4 && RNG) {
    {
#pragma unroll
      for (int c = 0; c < D; ++c) z[c] = (double)(float)z[c];  // what an fp32 noise array would have held
    }
  }
  if constexpr (PLAIN) {
#pragma unroll
    for (int c = 0; c < D; ++c) n[c] = z[c];
    return;
  }
  if (p.has_A) {
#pragma unroll
    for (int c = 0; c < D; ++c) {
      double acc = z[0] * p.A[c];
#pragma unroll
      for (int k = 1; k < D; ++k) acc = fma(z[k], p.A[k * D + c], acc);
      n[c] = acc;
    }
  } else {
#pragma unroll
    for (int c = 0; c < D; ++c) n[c] = z[c];
  }
  if (p.has_mu_noise) {
#pragma unroll
    for (int c = 0; c < D; ++c) n[c] += p.mu_noise[c];
  }
}

template <typename T, int D, int MODE, bool RNG, bool PLAIN>
__device__ __forceinline__ void predict_one(const pfb_predict_params& p, const T* x_in,
                                            const T* __restrict__ noise, int64_t i, double (&out)[D]) {
  double x[D], y[D];
  load_rec_inplace<T, D>(x_in, i, x);  // (x_in may be x_out: every thread reads its own record before it writes it)
  apply_transition<D, PLAIN>(p, x, y);
  const bool has_noise = (noise != nullptr || RNG) && p.noise_cols > 0;
  if constexpr (MODE == PFB_PRED_EUCLID) {
    if (has_noise) {
      double n[D];
      load_noise<T, D, RNG, PLAIN>(p, noise, i, n);
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = y[c] + n[c];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = y[c];
    }
  } else if constexpr (MODE == PFB_PRED_TORUS_SHIFT) {
    bool bad = false;
    if (has_noise) {
      double n[D];
      load_noise<T, D, RNG, PLAIN>(p, noise, i, n);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        const double mean = mod2pi_pos(y[c], bad);        // set_mean wraps the particle
        const double s = mod2pi_any(n[c] + mean, bad);    // sample(): draw + mean, wrapped: s in [0, 2 pi]
        out[c] = (s == kAngleConst[0]) ? 0.0 : s;         // filter wraps again: mod(s) = s, except mod(2 pi) = 0
      }
      if (bad) {  // some argument outside the fast range of the wraps: the general numpy.mod for this particle
#pragma unroll
        for (int c = 0; c < D; ++c) {
          const double s = np_mod_2pi_slow(n[c] + np_mod_2pi_slow(y[c]));
          out[c] = (s == kAngleConst[0]) ? 0.0 : s;
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = mod2pi_pos(y[c], bad);
      if (bad) {
#pragma unroll
        for (int c = 0; c < D; ++c) out[c] = np_mod_2pi_slow(y[c]);
      }
    }
  } else if constexpr (MODE == PFB_PRED_TORUS_ADD) {
    bool bad = false;
    if (has_noise) {
      double n[D];
      load_noise<T, D, RNG, PLAIN>(p, noise, i, n);          // includes the (host-wrapped) noise mean
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = mod2pi_any(y[c] + mod2pi_any<true>(n[c], bad), bad);
      if (bad) {
#pragma unroll
        for (int c = 0; c < D; ++c) out[c] = np_mod_2pi_slow(y[c] + np_mod_2pi_slow(n[c]));
      }
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = mod2pi_pos(y[c], bad);
      if (bad) {
#pragma unroll
        for (int c = 0; c < D; ++c) out[c] = np_mod_2pi_slow(y[c]);
      }
    }
  } else if constexpr (MODE == PFB_PRED_SPHERE_VMF) {
    static_assert(D == 3 || MODE != PFB_PRED_SPHERE_VMF, "VMF noise is implemented on S^2");
    if (has_noise) {
      double t[D];
      if constexpr (RNG) {
        const RngKey key{p.rng_seed, p.rng_offset};
        const uint4 r0 = rng_block(key, (uint64_t)i, 0u);
        t[0] = u53(r0.x, r0.y);
        double z0, z1;
        normal_pair(rng_block(key, (uint64_t)i, 1u), &z0, &z1);
        t[1 % D] = z0;
        t[2 % D] = z1;
      } else {
        load_rec<T, D>(noise, i, t);             // (u, z1, z2)
      }
      // scipy _rvs_3d: canonical draw around e1
      double xc = 1.0 + log(t[0] + (1.0 - t[0]) * p.exp_m2k) / p.kappa;
      double temp = sqrt(1.0 - xc * xc);
      double nrm = sqrt(t[1] * t[1] + t[2] * t[2]);
      double s[3] = {xc, temp * (t[1] / nrm), temp * (t[2] / nrm)};
      // scipy _rotate_samples: +-Householder reflector taking e1 to the mode y
      double alpha = y[0];
      double xn2 = y[1] * y[1] + y[2] * y[2];
      double r[3];
      bool positive;
      if (xn2 == 0.0) {
        r[0] = s[0]; r[1] = s[1]; r[2] = s[2];
        positive = alpha > 0.0;
      } else {
        double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
        double tau = (beta - alpha) / beta;
        double den = alpha - beta;
        double v1 = y[1] / den, v2 = y[2] / den;
        double vs = (s[0] + v1 * s[1]) + v2 * s[2];
        double tv = tau * vs;
        r[0] = s[0] - tv;
        r[1] = s[1] - tv * v1;
        r[2] = s[2] - tv * v2;
        positive = beta > 0.0;
      }
      if (!positive) { r[0] = -r[0]; r[1] = -r[1]; r[2] = -r[2]; }
      if (p.renormalize & 1) {
        double nn = sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
        r[0] /= nn; r[1] /= nn; r[2] /= nn;
      }
      if ((p.renormalize & 2) && r[2] < 0.0) { r[0] = -r[0]; r[1] = -r[1]; r[2] = -r[2]; }  // upper hemisphere
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = r[c < 3 ? c : 0];
    } else {
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = y[c];
    }
  } else {  // PFB_PRED_SPHERE_ADD
    double n[D];
    if (has_noise) {
      load_noise<T, D, RNG, PLAIN>(p, noise, i, n);
#pragma unroll
      for (int c = 0; c < D; ++c) y[c] += n[c];
    }
    double ss = y[0] * y[0];
#pragma unroll
    for (int c = 1; c < D; ++c) ss += y[c] * y[c];
    double nn = sqrt(ss);
    if ((p.renormalize & 2) && y[D - 1] < 0.0) nn = -nn;  // antipodal points are identified: keep the last coordinate >= 0
#pragma unroll
    for (int c = 0; c < D; ++c) out[c] = y[c] / nn;
  }
}

// ---------------------------------------------------------------------------------------------
// log-likelihoods
// ---------------------------------------------------------------------------------------------
template <int D>
__device__ __forceinline__ double whitened_sq(const pfb_lik_params& L, const double (&dev)[D]) {
  double maha = 0.0;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double acc = dev[0] * L.W[c];
#pragma unroll
    for (int r = 1; r < D; ++r) acc = fma(dev[r], L.W[r * D + c], acc);
    maha = fma(acc, acc, maha);
  }
  return maha;
}

constexpr double kImageCut = 90.0;  // screen: drop images whose maha exceeds the best one by this much (e^-45)
constexpr double kTermCut = 76.0;   // a term e^-38 below the best one cannot change the fp64 sum

// log pdf in two parts: the return value a and a linear factor `lin` >= 1 with pdf = exp(a) * lin (lin is 1 except
// for the image sum of the hypertoroidal wrapped normal, whose callers can then skip the log of the sum).
template <int D, int KIND>
__device__ __forceinline__ double loglik_parts(const pfb_lik_params& L, const double* __restrict__ mu,
                                               const double (&x)[D],
                                               const double* __restrict__ images /* smem or null */, double& lin) {
  lin = 1.0;
  if constexpr (KIND == PFB_LIK_GAUSS) {
    double dev[D];
#pragma unroll
    for (int c = 0; c < D; ++c) dev[c] = x[c] - mu[c];
    return L.logc - 0.5 * whitened_sq<D>(L, dev);
  } else if constexpr (KIND == PFB_LIK_WN_MULTI) {
    // xs = (xs + pi) % (2 pi) - pi; sum over the image table of mvn.pdf(xs + 2 pi k).
    // maha_k - maha_0 = 2 g_k.dev + h_k is linear in dev = xs - mu, so t_k = g_k.dev + h_k / 2 (three fp64 FMAs)
    // orders the images: the best one (smallest t) gets the exact quadratic form of the reference, every other
    // image enters the sum through exp(-(t_k - t_best)) -- in fp32 when the term is below 2e-9 of the sum, through
    // its own exact quadratic form when it is not (rare), not at all beyond e^-38.
    double dev[D], xs[D];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < D; ++c) xs[c] = mod2pi_pos(x[c] + kAngleConst[3], bad);
    if (bad) {
#pragma unroll
      for (int c = 0; c < D; ++c) xs[c] = np_mod_2pi_slow(x[c] + kAngleConst[3]);
    }
#pragma unroll
    for (int c = 0; c < D; ++c) dev[c] = (xs[c] - kAngleConst[3]) - mu[c];
    const int K = L.n_images;
    constexpr int ROW = 2 * D + 1;
    if constexpr (D <= 3) {
      if (K <= 64) {
        const double2* lt = reinterpret_cast<const double2*>(images + K * ROW + ((K * ROW) & 1));  // (g0, g1), (g2, h/2)
        const double d0 = dev[0], d1 = D > 1 ? dev[D > 1 ? 1 : 0] : 0.0, d2 = D > 2 ? dev[D > 2 ? 2 : 0] : 0.0;
        uint32_t mlo, mhi;  // candidate bits of images 0..31 / 32..63
        int kref;           // the image whose exact quadratic form anchors the sum
        auto lin_form = [&](int k) -> double {
          const double2 a = lt[2 * k], b = lt[2 * k + 1];
          return fma(a.x, d0, fma(a.y, d1, fma(b.x, d2, b.y)));
        };
        if (L.cell_grid > 0 && (L.mu_batch_dev == nullptr || (L.cell_span & 1))) {
          // the host has listed, per cell, the images that can come within e^-39 of the best one anywhere in the cell
          // (registry.wn_cell_masks: a superset that contains the minimiser) and the image that is best at the cell's
          // centre (no image beats it by more than e^60 anywhere in the cell).  The cells cut the box of xs for one
          // measurement (cell_span 0), or the box dev in [-3 pi, pi)^D that holds every measurement (cell_span 1:
          // twice the cells per dimension, one table for all runs of a batched launch)
          const int G = (L.cell_span & 1) ? 2 * L.cell_grid : L.cell_grid;
          const double inv = (double)L.cell_grid * 0.15915494309189535;  // cells per unit: cell_grid / (2 pi)
          int ci = 0, mul = 1;
#pragma unroll
          for (int c = 0; c < D; ++c) {
            int q = (int)(((L.cell_span & 1) ? dev[c] + 9.42477796076938 : xs[c]) * inv);
            q = q < 0 ? 0 : (q >= G ? G - 1 : q);
            ci += q * mul;
            mul *= G;
          }
          if (L.cell_span & 2) {  // at most 32 images: 32-bit masks (a 16^3 grid fits the shared memory of a CTA)
            const uint32_t* cells = reinterpret_cast<const uint32_t*>(lt + 2 * K);
            mlo = cells[ci];
            mhi = 0u;
            kref = reinterpret_cast<const unsigned char*>(cells + mul)[ci];
          } else {
            const uint2* cells = reinterpret_cast<const uint2*>(lt + 2 * K);
            const uint2 cand = cells[ci];
            mlo = cand.x;
            mhi = cand.y;
            kref = reinterpret_cast<const unsigned char*>(cells + mul)[ci];
          }
        } else {
          mlo = K >= 32 ? 0xFFFFFFFFu : ((1u << K) - 1u);
          mhi = K > 32 ? (K >= 64 ? 0xFFFFFFFFu : ((1u << (K - 32)) - 1u)) : 0u;
          double tmin = INFINITY;
          kref = 0;
          for (int k = 0; k < K; ++k) {
            const double t = lin_form(k);
            if (t < tmin) { tmin = t; kref = k; }
          }
        }
        const double tref = lin_form(kref);
        double sh[D];
        {
          const double* row = images + kref * ROW;
#pragma unroll
          for (int c = 0; c < D; ++c) sh[c] = dev[c] + row[D + 1 + c];
        }
        const double mbest = whitened_sq<D>(L, sh);
        double extra = 0.0;
        if (kref < 32) mlo &= ~(1u << kref); else mhi &= ~(1u << (kref - 32));
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t m = half ? mhi : mlo;
          while (m) {
            const int k = __ffs((int)m) - 1 + 32 * half;
            m &= m - 1u;
            const double d = lin_form(k) - tref;
            if (d <= 0.5 * kTermCut) {
              if (d > 20.0) {
                extra += (double)__expf(-(float)d);
              } else {  // a term above 2e-9 of the anchor's: its own quadratic form, like the reference
                const double* row = images + k * ROW;
                double s2[D];
#pragma unroll
                for (int c = 0; c < D; ++c) s2[c] = dev[c] + row[D + 1 + c];
                extra += exp(-0.5 * (whitened_sq<D>(L, s2) - mbest));
              }
            }
          }
        }
        lin = 1.0 + extra;
        return L.logc - 0.5 * mbest;
      }
    }
    double best = INFINITY, sum = 0.0;
    // exp(-d/2) of a term relative to the best one: fp64 where it matters, fp32 once the term is below
    // 2e-9 of the sum (relative error 1e-6 of 2e-9), nothing beyond e^-38
    auto rel_term = [](double d) -> double {
      if (d <= 40.0) return exp(-0.5 * d);
      return d <= kTermCut ? (double)__expf(-0.5f * (float)d) : 0.0;
    };
    auto accumulate = [&](const double* row) {
      double sh[D];
#pragma unroll
      for (int c = 0; c < D; ++c) sh[c] = dev[c] + row[D + 1 + c];
      const double ma = whitened_sq<D>(L, sh);
      if (ma < best) {
        sum = (best == INFINITY ? 0.0 : sum * rel_term(best - ma)) + 1.0;  // rescale to the new best
        best = ma;
      } else {
        sum += rel_term(ma - best);
      }
    };
    // generic path (D > 3 or more than 64 images): the same screen in fp64, two passes
    const double q0 = whitened_sq<D>(L, dev);
    double mmin = INFINITY;
    for (int k = 0; k < K; ++k) {
      const double* row = images + k * ROW;
      double lk = row[0] * dev[0];
#pragma unroll
      for (int c = 1; c < D; ++c) lk = fma(row[c], dev[c], lk);
      mmin = fmin(mmin, fma(2.0, lk, q0) + row[D]);
    }
    for (int k = 0; k < K; ++k) {
      const double* row = images + k * ROW;
      double lk = row[0] * dev[0];
#pragma unroll
      for (int c = 1; c < D; ++c) lk = fma(row[c], dev[c], lk);
      if (fma(2.0, lk, q0) + row[D] - mmin <= kImageCut) accumulate(row);
    }
    lin = sum;
    return L.logc - 0.5 * best;
  } else if constexpr (KIND == PFB_LIK_WN_1D) {
    if (L.sigma > 10.0) return -log(kTwoPi);
    double xx = np_mod_2pi(x[0]);
    xx -= mu[0];
    const double tmp = -1.0 / (2.0 * (L.sigma * L.sigma));
    double result = exp(xx * xx * tmp);
    for (int k = 1; k <= 1000; ++k) {
      double xp = xx + kTwoPi * k;
      double xm = xx - kTwoPi * k;
      double old = result;
      result += exp(xp * xp * tmp) + exp(xm * xm * tmp);
      if (result == old) break;
    }
    return log(result) + L.logc;  // logc = log(1/sqrt(2 pi)/sigma)
  } else if constexpr (KIND == PFB_LIK_VM) {
    return fma(L.kappa, cos(x[0] - mu[0]), L.logc);  // logc = -log(2 pi I0(kappa))
  } else {  // PFB_LIK_VMF
    double dot = mu[0] * x[0];
#pragma unroll
    for (int c = 1; c < D; ++c) dot = fma(mu[c], x[c], dot);
    return fma(L.kappa, dot, L.logc);  // logc = log C(kappa)
  }
}

// log(lin), lin = 1 + s with s >= 0: the series is exact to fp64 for small s
__device__ __forceinline__ double log_lin(double lin) {
  const double sx = lin - 1.0;
  if (sx == 0.0) return 0.0;
  if (sx < 1e-5) return sx * (1.0 - sx * (0.5 - sx * (1.0 / 3.0)));
  return log(lin);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// second, padded copy of a predicted particle (pfb_predict_params.x_pad_out): whole sectors, vector stores
template <typename T, int D>
__device__ __forceinline__ void store_padded(const pfb_predict_params& p, int64_t i, const double (&out)[D]) {
  constexpr int DP = (D <= 4) ? 4 : 8;
  if (p.x_pad_out == nullptr || p.x_pad_stride != DP) return;
  double q[DP];
#pragma unroll
  for (int c = 0; c < DP; ++c) q[c] = (c < D) ? out[c] : 0.0;
  store_rec<T, DP>(static_cast<T*>(p.x_pad_out), i, q);
}

template <typename T, int D, int MODE, bool RNG, bool PLAIN>
__global__ void __launch_bounds__(kThreads) predict_kernel(pfb_predict_params p, int64_t n,
                                                           const T* x_in, T* x_out,
                                                           const T* __restrict__ noise) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double out[D];
    predict_one<T, D, MODE, RNG, PLAIN>(p, x_in, noise, i, out);
    store_rec<T, D>(x_out, i, out);
    store_padded<T, D>(p, i, out);
  }
}

// finish the max reduction: partial (max, nan) per CTA -> stats by the last CTA
__device__ __forceinline__ void finish_max(double tmax, double tnan, double* stats, Workspace ws,
                                           double* smem) {
  double bm = block_max(tmax, smem);
  double bn = block_max(tnan, smem);
  if (threadIdx.x == 0) {
    ws.partials[2 * blockIdx.x] = bm;
    ws.partials[2 * blockIdx.x + 1] = bn;
  }
  if (block_is_last(ws.counters)) {
    double m = -INFINITY, nn = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      m = fmax(m, ws.partials[2 * b]);
      nn = fmax(nn, ws.partials[2 * b + 1]);
    }
    m = block_max(m, smem);
    nn = block_max(nn, smem);
    if (threadIdx.x == 0) {
      stats[PFB_STAT_MAX] = m;
      double flags = 0.0;
      if (nn > 0.0) flags += 1.0;
      if (m == -INFINITY) flags += 2.0;
      stats[PFB_STAT_FLAGS] = flags;
    }
  }
}

template <int D>
__device__ __forceinline__ const double* stage_images(const pfb_lik_params& L, double* smem_img) {
  if (L.kind != PFB_LIK_WN_MULTI) return nullptr;
  constexpr int ROW = 2 * D + 1;
  const int K = L.n_images;
  const int count = K * ROW;
  for (int k = threadIdx.x; k < count; k += blockDim.x) smem_img[k] = L.images_dev[k];
  if (D <= 3 && K <= 64) {  // linear-form rows (g0, g1, g2, h / 2) after the table (16-byte aligned: count*8 + pad)
    double* lt = smem_img + count + (count & 1);
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      const double* row = L.images_dev + k * ROW;
      lt[4 * k] = row[0];
      lt[4 * k + 1] = D > 1 ? row[D > 1 ? 1 : 0] : 0.0;
      lt[4 * k + 2] = D > 2 ? row[D > 2 ? 2 : 0] : 0.0;
      lt[4 * k + 3] = 0.5 * row[D];
    }
    if (L.cell_grid > 0) {  // candidate masks of the cell grid, after the linear rows
      uint64_t* cells = reinterpret_cast<uint64_t*>(lt + 4 * K);
      const uint64_t* src = reinterpret_cast<const uint64_t*>(L.images_dev + count);
      const int G = (L.cell_span & 1) ? 2 * L.cell_grid : L.cell_grid;
      int nc = G;
      if (D > 1) nc *= G;
      if (D > 2) nc *= G;
      const int words = (L.cell_span & 2) ? (nc * 5 + 7) / 8 : nc + (nc + 7) / 8;  // masks (4 or 8 bytes), then the reference bytes
      for (int k = threadIdx.x; k < words; k += blockDim.x) cells[k] = src[k];
    }
  }
  __syncthreads();
  return smem_img;
}

// Both weighting kernels walk the particle array in the scan's tiles (kScanTile consecutive particles
// per CTA step, thread t takes particles t, t+256, ... of the tile: coalesced) and leave, per tile,
// (max log-weight, sum exp(logw - max)) in the workspace: pfb_scan_cdf turns those into the approximate
// prefix its sequential-exact scan needs, without another pass over the log-weights.

// runs of a batched launch: n_runs independent filters of run_len particles each; their log-weights
// live at logw[run * run_pad + k] (run_pad = tpr * kScanTile, padded with -inf, so that no scan tile
// straddles two runs).  A plain launch is one run with run_pad = run_len.
struct Seg {
  int64_t n_runs, run_len, run_pad, tpr;
};
inline Seg seg_single(int64_t n) { return Seg{1, n, n, (n + kScanTile - 1) / kScanTile}; }
inline Seg seg_batched(int64_t n_runs, int64_t run_len) {
  const int64_t tpr = (run_len + kScanTile - 1) / kScanTile;
  return Seg{n_runs, run_len, tpr * kScanTile, tpr};
}

// No CTA-wide barrier inside the tile loop: the likelihoods have data-dependent cost (image candidates,
// series lengths), so every warp owns a contiguous 256-particle slice of the tile, reduces its own
// (max, sum exp) with shuffles and stores it; pfb_scan_cdf combines the 8 partials of a tile.
//
// Two output formats (pfb_lik_params.out_format):
//   PFB_WEIGHTS_LOG     logw_i = log pdf_i; per slice (max, sum exp(l - max)).  One exp per particle here and one
//                       more in every pass of the scan that turns log-weights into addends.
//   PFB_WEIGHTS_SCALED  p_i = pdf_i 2^(-e_s), e_s an integer exponent of the slice: ceil(max(l) log2 e) over the first
//                       32 particles of the slice that have a non-zero weight (so p_i is of order 1; a later particle more
//                       than e^600 above it makes the warp rescale what it has stored -- exact power-of-two scalings);
//                       per slice (e_s, sum p).  The scan forms its addends as p_i 2^(e_s - E) -- an exact scaling, no
//                       exp -- so the whole step spends ONE exp per particle, in a single pass, and keeps the range of
//                       the log domain (a slice of far-away particles scales to 0 only relative to the best slice, like
//                       exp(l - max) would).
constexpr double kLog2e = 1.4426950408889634, kLn2 = 0.6931471805599453;
__device__ __forceinline__ double slice_exponent(double m) {  // integer-valued; -inf for an empty slice
  if (!(m > -INFINITY)) return -INFINITY;
  const double e = ceil(m * kLog2e);
  return e < -4000.0 ? -4000.0 : (e > 4000.0 ? 4000.0 : e);
}

// a later batch of the slice lies more than e^600 above the slice exponent: move the exponent up and scale what the
// warp has stored so far (lane-private columns of the output array).  Returns the new exponent; *factor is the
// power of two the caller's running sum and maximum have to be scaled by.
static __device__ __noinline__ double rescale_slice(double* mine, int k, int64_t lbase, int64_t run_pad, double eg,
                                                    double a, double* factor) {
  const double en = slice_exponent(warp_max(a));
  const double d = eg - en;
  const double f = (d > -1100.0) ? scalbn(1.0, (int)d) : 0.0;
  for (int j = 0; j < k; ++j)
    if (lbase + (int64_t)j * 32 < run_pad) mine[j * 32] *= f;
  *factor = f;
  return en;
}

// exp(x) for x <= 700 (0 below -700: such a weight is 1e-304 of its slice's largest).  The usual reduction
// x = k ln 2 + r, |r| <= ln 2 / 2, and a degree-11 polynomial (< 1 ulp in all); the coefficients sit in constant memory
// (literals would be materialised with two moves in front of every DFMA: the library exp issues ~50 instructions).
static __constant__ double kExpCoef[16] = {
    1.4426950408889634,        // log2 e
    6755399441055744.0,        // 1.5 * 2^52: adding it rounds to an integer in the low word
    6.93147180559945286227e-01,  // ln 2, high part
    2.31904681384629955842e-17,  // ln 2, low part
    // Chebyshev interpolant of exp on [-ln 2 / 2, ln 2 / 2], degree 11, monomial form (approximation error 7e-18)
    2.510158556723003e-08, 2.7632488519513105e-07, 2.7557267919196622e-06, 2.480148595312923e-05,
    0.00019841269859107206, 0.0013888888951800668, 0.008333333333334735, 0.04166666666649038,
    0.16666666666666652, 0.5000000000000018, 1.0, 1.0};
__device__ __forceinline__ double exp_le700(double x) {
  const double t = fma(x, kExpCoef[0], kExpCoef[1]);
  const int k = __double2loint(t);
  const double kd = t - kExpCoef[1];
  double r = fma(kd, -kExpCoef[2], x);
  r = fma(kd, -kExpCoef[3], r);
  double q = kExpCoef[4];
#pragma unroll
  for (int j = 5; j < 16; ++j) q = fma(q, r, kExpCoef[j]);
  const double y = __hiloint2double(__double2hiint(q) + (k << 20), __double2loint(q));
  return (x < -700.0) ? 0.0 : y;
}

// eval(i, run, lin) returns a with pdf_i = exp(a) * lin
template <bool SCALED, typename EVAL>
__device__ __forceinline__ void weight_tiles(Seg sg, double* logw, double* stats, Workspace ws,
                                             double* red, EVAL eval) {
  const int64_t ntiles = sg.n_runs * sg.tpr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarpSpan = kScanTile / (kThreads / 32);  // 256 particles per warp and tile
  double gmax = -INFINITY, gnan = 0.0;
  double bp = 0.0, bc = 0.0;  // SCALED: the largest weight so far is bp e^bc
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t run = tile / sg.tpr;
    const int64_t lbase = (tile - run * sg.tpr) * kScanTile + warp * kWarpSpan + lane;
    if constexpr (SCALED) {
      constexpr int kPer = kWarpSpan / 32;
      double eg = -INFINITY, cg = 0.0, sacc = 0.0, pmax = 0.0;
      double* mine = logw + run * sg.run_pad + lbase;
#pragma unroll 1
      for (int k = 0; k < kPer; ++k) {  // (not unrolled: the evaluation is long)
        const int64_t local = lbase + k * 32;
        double a = -INFINITY, lin = 1.0;
        if (local < sg.run_len) {
          a = eval(run * sg.run_len + local, run, lin);
          if (a != a || lin != lin) { gnan = 1.0; a = -INFINITY; }
        }
        if (!(eg > -INFINITY)) {  // (warp-uniform) no exponent yet: first batch, or only zero weights so far
          eg = slice_exponent(warp_max(a));
          cg = eg * kLn2;
        } else if (__any_sync(0xffffffffu, a - cg > 600.0)) {
          double f;
          eg = rescale_slice(mine, k, lbase, sg.run_pad, eg, a, &f);
          cg = eg * kLn2;
          sacc *= f;
          pmax *= f;
        }
        const double p = (a > -INFINITY) ? exp_le700(a - cg) * lin : 0.0;  // (a = cg = -inf, a slice of zero weights so far, would be NaN)
        if (local < sg.run_pad) mine[k * 32] = p;  // (padding of a batched run: weight 0)
        sacc += p;
        pmax = (p > pmax) ? p : pmax;
      }
      const double wsum = warp_sum(sacc);
      const double wp = warp_max(pmax);
      if (lane == 0) {
        ws.wpart[tile * 16 + 2 * warp] = eg;
        ws.wpart[tile * 16 + 2 * warp + 1] = wsum;
      }
      if (wp > 0.0) {  // running maximum of the weights, kept as (factor, log offset): logs only when the offsets differ
        if (bp == 0.0 || cg == bc) {
          bp = fmax(bp, wp);
          bc = cg;
        } else if (log(wp) + cg > log(bp) + bc) {
          bp = wp;
          bc = cg;
        }
      }
    } else {
      double m = -INFINITY, sacc = 0.0;  // running max and sum exp(l - m) of this thread's particles
#pragma unroll 1
      for (int k = 0; k < kWarpSpan / 32; ++k) {
        const int64_t local = lbase + k * 32;
        if (local < sg.run_len) {
          double lin;
          double v = eval(run * sg.run_len + local, run, lin);
          v += log_lin(lin);
          logw[run * sg.run_pad + local] = v;
          if (v != v) {
            gnan = 1.0;
          } else if (v > -INFINITY) {
            const double e = exp(-fabs(v - m));  // one exp for both cases (m = -inf: e = 0, sacc = 0 * 0 + 1)
            sacc = (v > m) ? sacc * e + 1.0 : sacc + e;
            m = fmax(m, v);
          }
        } else if (local < sg.run_pad) {
          logw[run * sg.run_pad + local] = -INFINITY;  // padding of a batched run: weight 0
        }
      }
      const double wm = warp_max(m);
      const double wsum = warp_sum((m > -INFINITY) ? sacc * exp(m - wm) : 0.0);
      if (lane == 0) {
        ws.wpart[tile * 16 + 2 * warp] = wm;
        ws.wpart[tile * 16 + 2 * warp + 1] = wsum;
      }
      gmax = fmax(gmax, wm);
    }
  }
  if constexpr (SCALED) gmax = (bp > 0.0) ? log(bp) + bc : -INFINITY;
  finish_max(gmax, gnan, stats, ws, red);
}

template <typename T, int D, int KIND, bool SCALED>
__global__ void __launch_bounds__(kThreads, 4) loglik_kernel(pfb_lik_params L, Seg sg,
                                                          const T* __restrict__ x,
                                                          const double* __restrict__ w_prev,
                                                          double* logw, double* stats,
                                                          Workspace ws) {
  extern __shared__ __align__(16) double smem_dyn[];
  __shared__ double red[32];
  const double* images = stage_images<D>(L, smem_dyn);
  weight_tiles<SCALED>(sg, logw, stats, ws, red, [&](int64_t i, int64_t run, double& lin) {
    double r[D];
    load_rec<T, D>(x, i, r);
    const double* mu = L.mu_batch_dev ? L.mu_batch_dev + run * D : L.mu;
    double l = loglik_parts<D, KIND>(L, mu, r, images, lin);
    if (w_prev != nullptr) l += log(w_prev[i]);
    return l;
  });
}

template <typename T, int D, int MODE, int KIND, bool RNG, bool SCALED, bool PLAIN>
__global__ void __launch_bounds__(kThreads, 4) predict_loglik_kernel(pfb_predict_params p,
                                                                  pfb_lik_params L, Seg sg,
                                                                  const T* x_in, T* x_out,
                                                                  const T* __restrict__ noise,
                                                                  double* logw,
                                                                  double* stats, Workspace ws) {
  extern __shared__ __align__(16) double smem_dyn[];
  __shared__ double red[32];
  const double* images = stage_images<D>(L, smem_dyn);
  weight_tiles<SCALED>(sg, logw, stats, ws, red, [&](int64_t i, int64_t run, double& lin) {
    double out[D];
    predict_one<T, D, MODE, RNG, PLAIN>(p, x_in, noise, i, out);
    if (x_out != nullptr) store_rec<T, D>(x_out, i, out);  // (NULL: only the padded copy is wanted, see pfb.h)
    store_padded<T, D>(p, i, out);
    if constexpr (sizeof(T) == 4) {
      // weight what was stored (the fp32 particle), as a separate pfb_loglik call would
#pragma unroll
      for (int c = 0; c < D; ++c) out[c] = (double)(float)out[c];
    }
    const double* mu = L.mu_batch_dev ? L.mu_batch_dev + run * D : L.mu;
    return loglik_parts<D, KIND>(L, mu, out, images, lin);
  });
}

template <typename T>
__global__ void __launch_bounds__(kThreads) wrap_kernel(int64_t count, const T* __restrict__ x_in,
                                                        T* x_out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    x_out[i] = (T)np_mod_2pi((double)x_in[i]);
}

// ---------------------------------------------------------------------------------------------
// host-side validation + dispatch
// ---------------------------------------------------------------------------------------------
static int check_predict(const pfb_predict_params* p) {
  if (!p) { set_error("predict params NULL"); return PFB_ERR_INVALID; }
  if (p->dim < 1 || p->dim > PFB_MAXD) { set_error("dim %d outside 1..%d", p->dim, PFB_MAXD); return PFB_ERR_INVALID; }
  if (p->mode < PFB_PRED_EUCLID || p->mode > PFB_PRED_SPHERE_ADD) { set_error("bad predict mode %d", p->mode); return PFB_ERR_INVALID; }
  if (p->mode == PFB_PRED_SPHERE_VMF && p->dim != 3) { set_error("VMF noise needs dim 3 (S^2), got %d", p->dim); return PFB_ERR_UNSUPPORTED; }
  if (p->noise_cols != 0 && p->noise_cols != p->dim) { set_error("noise_cols %d must equal dim %d", p->noise_cols, p->dim); return PFB_ERR_INVALID; }
  if (p->mode == PFB_PRED_SPHERE_VMF && p->noise_cols && !(p->kappa > 0.0)) { set_error("kappa must be > 0"); return PFB_ERR_INVALID; }
  if (p->x_pad_out != nullptr && p->x_pad_stride != (p->dim <= 4 ? 4 : 8)) {
    set_error("x_pad_stride %d: the padded copy has 4 components per record for dim <= 4, else 8", p->x_pad_stride);
    return PFB_ERR_INVALID;
  }
  if (p->rng_kind < PFB_RNG_NONE || p->rng_kind > PFB_RNG_VMF) { set_error("bad rng_kind %d", p->rng_kind); return PFB_ERR_INVALID; }
  if (p->rng_kind != PFB_RNG_NONE && (p->rng_kind == PFB_RNG_VMF) != (p->mode == PFB_PRED_SPHERE_VMF)) {
    set_error("rng_kind %d does not go with predict mode %d (VMF triples <-> PFB_PRED_SPHERE_VMF)", p->rng_kind, p->mode);
    return PFB_ERR_INVALID;
  }
  if (p->rng_kind == PFB_RNG_VONMISES) {
    for (int c = 0; c < p->dim; ++c) {
      if (!(p->rng_kappa[c] >= 0.0)) { set_error("von Mises kappa[%d] must be >= 0", c); return PFB_ERR_INVALID; }
      if (p->rng_kappa[c] < 1e-8) continue;
      if ((p->rng_vm_table_stride != 0 && p->rng_vm_table_stride != 2 * PFB_VM_STRIPS && p->rng_vm_table_stride != PFB_VM_TABLE_DOUBLES) ||
          !p->rng_vm_table_dev || !(p->rng_vm_area[c] > 0.0) || p->rng_vm_table_idx[c] < 0 || p->rng_vm_table_idx[c] >= PFB_MAXD) {
        set_error("von Mises draws need the strip table of kappa[%d] (pfb_vm_strip_table) on the device", c);
        return PFB_ERR_INVALID;
      }
    }
  }
  return PFB_OK;
}

// Strip table of the von Mises sampler (pfb_common.cuh: vonmises_draws).  With g(t) = exp(-2 kappa sin^2(t/2)):
// a_0 = 0, a_{k+1} = a_k + A / g(a_k); A is the smallest area for which the K-th strip reaches pi (bisection:
// the end point is increasing in A), and the last strip is clipped to pi, which only raises its hat.
static double vm_strip_end(double kappa, double A, int K, double* table) {
  const double kPiH = 3.141592653589793;
  double a = 0.0;
  for (int k = 0; k < K; ++k) {
    const double sh = sin(0.5 * a);
    double w = A / exp(-2.0 * kappa * sh * sh);
    if (table) {  // fill mode: clip at pi (a strip past pi keeps zero width: it is chosen and always rejected)
      if (!(a < kPiH)) { a = kPiH; w = 0.0; }
      else if (!(a + w < kPiH) || k == K - 1) w = kPiH - a;
      table[2 * k] = a;
      table[2 * k + 1] = w;
    } else if (!(a + w < kPiH)) {
      return 4.0;  // reaches pi (at the last strip: just enough; earlier: A is larger than needed)
    }
    a += w;
  }
  return a;
}

#ifndef PFB_PL_F32_TU
extern "C" int pfb_vm_strip_table(double kappa, double* table, double* area) {
  if (!table || !area) { set_error("NULL table / area"); return PFB_ERR_INVALID; }
  if (!(kappa >= 1e-8) || !(kappa < 1e300)) { set_error("strip tables exist for 1e-8 <= kappa < inf (smaller kappas draw uniformly)"); return PFB_ERR_INVALID; }
  const double kPiH = 3.141592653589793;
  const int K = PFB_VM_STRIPS;
  double lo = 0.0, hi = kPiH;  // end(lo) = 0 < pi <= end(hi)
  for (int it = 0; it < 200 && lo < hi; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    if (vm_strip_end(kappa, mid, K, nullptr) >= kPiH) hi = mid; else lo = mid;
  }
  vm_strip_end(kappa, hi, K, table);
  *area = hi;
  return PFB_OK;
}


extern "C" int pfb_vm_strip_thresholds(double kappa, const double* table, double area, uint32_t* thr) {
  if (!table || !thr || !(area > 0.0) || !(kappa >= 1e-8)) { set_error("NULL table / thresholds, or no strip table for this kappa"); return PFB_ERR_INVALID; }
  for (int k = 0; k < PFB_VM_STRIPS; ++k) {
    const double end = table[2 * k] + table[2 * k + 1];
    const double sh = sin(0.5 * end);
    // v32 < thr  =>  (v32 + 1/2) 2^-32 A < g(end) w (1 - 1e-9) <= g(t) w for every t of the strip, rounding included
    const double r = exp(-2.0 * kappa * sh * sh) * table[2 * k + 1] / area * 4294967296.0 * (1.0 - 1e-9) - 2.0;
    thr[k] = r <= 0.0 ? 0u : (r >= 4294967295.0 ? 4294967295u : (uint32_t)r);
  }
  return PFB_OK;
}
#endif

static size_t lik_smem(const pfb_lik_params* L) {
  if (L->kind != PFB_LIK_WN_MULTI) return 0;
  size_t cells = 0;
  if (L->cell_grid > 0) {
    cells = 1;
    for (int c = 0; c < L->dim; ++c) cells *= (size_t)((L->cell_span & 1) ? 2 * L->cell_grid : L->cell_grid);
    cells = (L->cell_span & 2) ? ((cells * 5 + 7) / 8) * 8 : cells * 8 + ((cells + 7) / 8) * 8;  // masks + one reference byte per cell
  }
  return (size_t)L->n_images * (2 * L->dim + 1) * 8 + 8 + (size_t)L->n_images * 32 + cells;
}

static int check_lik(const pfb_lik_params* L, int dim) {
  if (!L) { set_error("likelihood params NULL"); return PFB_ERR_INVALID; }
  if (L->dim != dim) { set_error("likelihood dim %d != particle dim %d", L->dim, dim); return PFB_ERR_INVALID; }
  if (L->out_format != PFB_WEIGHTS_LOG && L->out_format != PFB_WEIGHTS_SCALED) { set_error("bad out_format %d", L->out_format); return PFB_ERR_INVALID; }
  if (L->kind < PFB_LIK_GAUSS || L->kind > PFB_LIK_VMF) { set_error("unknown likelihood kind %d (arbitrary callables are not supported)", L->kind); return PFB_ERR_UNSUPPORTED; }
  if ((L->kind == PFB_LIK_WN_1D || L->kind == PFB_LIK_VM) && dim != 1) { set_error("likelihood kind %d is circular (dim 1), got %d", L->kind, dim); return PFB_ERR_INVALID; }
  if (L->kind == PFB_LIK_WN_MULTI) {
    if (L->n_images < 1 || !L->images_dev) { set_error("wn_multi needs an image table"); return PFB_ERR_INVALID; }
    if (L->cell_span < 0 || L->cell_span > 3 || ((L->cell_span & 2) && L->n_images > 32)) { set_error("bad cell_span %d (bit 1, 32-bit masks, needs at most 32 images)", L->cell_span); return PFB_ERR_INVALID; }
    if ((int64_t)lik_smem(L) > 40 * 1024) { set_error("image table too large (%d rows, cell grid %d)", L->n_images, L->cell_grid); return PFB_ERR_INVALID; }
    if (L->cell_grid != 0 && (L->cell_grid < 1 || L->cell_grid > 16 || dim > 3 || L->n_images > 64)) {
      set_error("cell_grid %d: candidate masks need 1 <= G <= 16, dim <= 3 and at most 64 images", L->cell_grid);
      return PFB_ERR_INVALID;
    }
  }
  return PFB_OK;
}


#define PFB_DISPATCH_MODE(mode, ...)                                                       \
  switch (mode) {                                                                          \
    case PFB_PRED_EUCLID: { constexpr int MODE = PFB_PRED_EUCLID; __VA_ARGS__; } break;     \
    case PFB_PRED_TORUS_SHIFT: { constexpr int MODE = PFB_PRED_TORUS_SHIFT; __VA_ARGS__; } break; \
    case PFB_PRED_TORUS_ADD: { constexpr int MODE = PFB_PRED_TORUS_ADD; __VA_ARGS__; } break; \
    case PFB_PRED_SPHERE_ADD: { constexpr int MODE = PFB_PRED_SPHERE_ADD; __VA_ARGS__; } break; \
    default: set_error("bad mode"); return PFB_ERR_INVALID;                                 \
  }

#define PFB_DISPATCH_KIND(kind, ...)                                                       \
  switch (kind) {                                                                          \
    case PFB_LIK_GAUSS: { constexpr int KIND = PFB_LIK_GAUSS; __VA_ARGS__; } break;         \
    case PFB_LIK_WN_MULTI: { constexpr int KIND = PFB_LIK_WN_MULTI; __VA_ARGS__; } break;   \
    case PFB_LIK_VMF: { constexpr int KIND = PFB_LIK_VMF; __VA_ARGS__; } break;             \
    default: set_error("bad kind"); return PFB_ERR_INVALID;                                 \
  }

}  // namespace pfb

namespace pfb {

// One wave of CTAs for the tile loops: as many as are resident at once (the kernel's own occupancy), so that no
// CTA starts when the others are half done.  Cached per kernel.
template <typename K>
static int one_wave_grid(K kernel, size_t smem, const Seg& sg) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int resident;
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(reinterpret_cast<const void*>(kernel));
    if (it == cache.end()) {
      int occ = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kThreads, smem) != cudaSuccess || occ < 1) occ = 4;
      it = cache.emplace(reinterpret_cast<const void*>(kernel), occ * sm_count()).first;
    }
    resident = it->second;
  }
  const int64_t tiles = sg.n_runs * sg.tpr;
  int g = resident < kMaxPartialBlocks ? resident : kMaxPartialBlocks;
  if (tiles < g) g = (int)(tiles < 1 ? 1 : tiles);
  return g;
}

static bool plain_predict(const pfb_predict_params* p) {
  return !p->has_F && !p->has_b && !p->has_A && !p->has_mu_noise;
}

// ---- launches of one storage type T.  The fp64 instantiations live in this translation unit, the fp32 ones in
// pfb_predict_loglik_f32.cu (which includes this file with PFB_PL_F32_TU defined): the two halves compile in parallel.
template <typename T>
int predict_launch(const pfb_predict_params* p, int64_t n, const void* x_in, void* x_out, const void* noise,
                   cudaStream_t st) {
  const int grid = stream_grid(n);
  const bool rng = p->rng_kind != PFB_RNG_NONE;
  const bool plain = plain_predict(p);
#define PFB_PREDICT_LAUNCH(DD, MM, RR, PP) \
  predict_kernel<T, DD, MM, RR, PP><<<grid, kThreads, 0, st>>>(*p, n, (const T*)x_in, (T*)x_out, (const T*)noise)
  if (p->mode == PFB_PRED_SPHERE_VMF) {
    if (rng) { if (plain) PFB_PREDICT_LAUNCH(3, PFB_PRED_SPHERE_VMF, true, true); else PFB_PREDICT_LAUNCH(3, PFB_PRED_SPHERE_VMF, true, false); }
    else { if (plain) PFB_PREDICT_LAUNCH(3, PFB_PRED_SPHERE_VMF, false, true); else PFB_PREDICT_LAUNCH(3, PFB_PRED_SPHERE_VMF, false, false); }
  } else if (plain && (p->mode == PFB_PRED_TORUS_SHIFT || p->mode == PFB_PRED_TORUS_ADD)) {
    if (p->mode == PFB_PRED_TORUS_SHIFT) {
      if (rng) { PFB_DISPATCH_DIM(p->dim, PFB_PREDICT_LAUNCH(D, PFB_PRED_TORUS_SHIFT, true, true)); }
      else { PFB_DISPATCH_DIM(p->dim, PFB_PREDICT_LAUNCH(D, PFB_PRED_TORUS_SHIFT, false, true)); }
    } else {
      if (rng) { PFB_DISPATCH_DIM(p->dim, PFB_PREDICT_LAUNCH(D, PFB_PRED_TORUS_ADD, true, true)); }
      else { PFB_DISPATCH_DIM(p->dim, PFB_PREDICT_LAUNCH(D, PFB_PRED_TORUS_ADD, false, true)); }
    }
  } else if (rng) {
    PFB_DISPATCH_DIM(p->dim, PFB_DISPATCH_MODE(p->mode, PFB_PREDICT_LAUNCH(D, MODE, true, false)));
  } else {
    PFB_DISPATCH_DIM(p->dim, PFB_DISPATCH_MODE(p->mode, PFB_PREDICT_LAUNCH(D, MODE, false, false)));
  }
#undef PFB_PREDICT_LAUNCH
  return PFB_OK;
}

template <typename T>
int loglik_launch(const pfb_lik_params* lik, Seg sg, const void* x, const double* w_prev, double* logw, double* stats,
                  Workspace W, cudaStream_t st) {
  const size_t sm = lik_smem(lik);
#define PFB_LOGLIK_ONE(DD, KK, S)                                                                         \
  {                                                                                                       \
    auto k = loglik_kernel<T, DD, KK, S>;                                                                 \
    k<<<one_wave_grid(k, sm, sg), kThreads, sm, st>>>(*lik, sg, (const T*)x, w_prev, logw, stats, W);     \
  }
#define PFB_LOGLIK_LAUNCH(S)                                                                  \
  if (lik->kind == PFB_LIK_WN_1D) {                                                            \
    PFB_LOGLIK_ONE(1, PFB_LIK_WN_1D, S)                                                        \
  } else if (lik->kind == PFB_LIK_VM) {                                                        \
    PFB_LOGLIK_ONE(1, PFB_LIK_VM, S)                                                           \
  } else {                                                                                     \
    PFB_DISPATCH_DIM(lik->dim, PFB_DISPATCH_KIND(lik->kind, PFB_LOGLIK_ONE(D, KIND, S)));      \
  }
  if (lik->out_format == PFB_WEIGHTS_SCALED) { PFB_LOGLIK_LAUNCH(true) } else { PFB_LOGLIK_LAUNCH(false) }
#undef PFB_LOGLIK_LAUNCH
#undef PFB_LOGLIK_ONE
  return PFB_OK;
}

// Only manifold-consistent (mode, likelihood) pairs are instantiated:
//   Euclidean <-> gauss;  torus <-> wn_multi (any D), wn_1d / vm (D = 1);  sphere <-> vmf.
template <typename T, int D, int MODE, int KIND>
static int launch_fused_one(const pfb_predict_params* p, const pfb_lik_params* lik, Seg sg, const void* x_in,
                            void* x_out, const void* noise, double* logw, double* stats, Workspace W,
                            cudaStream_t st) {
  const size_t sm = lik_smem(lik);
  const bool scaled = lik->out_format == PFB_WEIGHTS_SCALED;
  // the PLAIN kernels exist for the torus modes and S^2 von Mises-Fisher noise (identity transition, rows as drawn)
  constexpr bool kHasPlain = MODE == PFB_PRED_TORUS_SHIFT || MODE == PFB_PRED_TORUS_ADD || MODE == PFB_PRED_SPHERE_VMF;
  const bool plain = kHasPlain && plain_predict(p);
#define PFB_FUSED_LAUNCH(R, S, P)                                                                                  \
  {                                                                                                                \
    auto k = predict_loglik_kernel<T, D, MODE, KIND, R, S, P>;                                                     \
    k<<<one_wave_grid(k, sm, sg), kThreads, sm, st>>>(*p, *lik, sg, (const T*)x_in, (T*)x_out, (const T*)noise, logw, stats, W); \
  }
#define PFB_FUSED_RS(P)                                                                  \
  if (p->rng_kind != PFB_RNG_NONE) {                                                     \
    if (scaled) PFB_FUSED_LAUNCH(true, true, P) else PFB_FUSED_LAUNCH(true, false, P)    \
  } else {                                                                               \
    if (scaled) PFB_FUSED_LAUNCH(false, true, P) else PFB_FUSED_LAUNCH(false, false, P)  \
  }
  if constexpr (kHasPlain) {
    if (plain) { PFB_FUSED_RS(true) return PFB_OK; }
  }
  PFB_FUSED_RS(false)
#undef PFB_FUSED_RS
#undef PFB_FUSED_LAUNCH
  return PFB_OK;
}

template <typename T, int D, int MODE>
static int launch_fused(const pfb_predict_params* p, const pfb_lik_params* lik, Seg n, const void* x_in,
                        void* x_out, const void* noise, double* logw, double* stats, Workspace W,
                        cudaStream_t st) {
  constexpr bool torus = (MODE == PFB_PRED_TORUS_SHIFT || MODE == PFB_PRED_TORUS_ADD);
  constexpr bool sphere = (MODE == PFB_PRED_SPHERE_VMF || MODE == PFB_PRED_SPHERE_ADD);
  if constexpr (MODE == PFB_PRED_EUCLID) {
    if (lik->kind == PFB_LIK_GAUSS) return launch_fused_one<T, D, MODE, PFB_LIK_GAUSS>(p, lik, n, x_in, x_out, noise, logw, stats, W, st);
  }
  if constexpr (torus) {
    if (lik->kind == PFB_LIK_WN_MULTI) return launch_fused_one<T, D, MODE, PFB_LIK_WN_MULTI>(p, lik, n, x_in, x_out, noise, logw, stats, W, st);
    if constexpr (D == 1) {
      if (lik->kind == PFB_LIK_WN_1D) return launch_fused_one<T, 1, MODE, PFB_LIK_WN_1D>(p, lik, n, x_in, x_out, noise, logw, stats, W, st);
      if (lik->kind == PFB_LIK_VM) return launch_fused_one<T, 1, MODE, PFB_LIK_VM>(p, lik, n, x_in, x_out, noise, logw, stats, W, st);
    }
  }
  if constexpr (sphere && D >= 2) {
    if (lik->kind == PFB_LIK_VMF) return launch_fused_one<T, D, MODE, PFB_LIK_VMF>(p, lik, n, x_in, x_out, noise, logw, stats, W, st);
  }
  set_error("likelihood kind %d is not defined on the manifold of predict mode %d (dim %d)", lik->kind, MODE, D);
  return PFB_ERR_UNSUPPORTED;
}

template <typename T>
int fused_launch(const pfb_predict_params* p, const pfb_lik_params* lik, Seg sg, const void* x_in, void* x_out,
                 const void* noise, double* logw, double* stats, Workspace W, cudaStream_t st) {
  int rc = PFB_OK;
  if (p->mode == PFB_PRED_SPHERE_VMF) {
    rc = launch_fused<T, 3, PFB_PRED_SPHERE_VMF>(p, lik, sg, x_in, x_out, noise, logw, stats, W, st);
  } else {
    PFB_DISPATCH_DIM(p->dim, PFB_DISPATCH_MODE(p->mode, (rc = launch_fused<T, D, MODE>(p, lik, sg, x_in, x_out, noise, logw, stats, W, st))));
  }
  return rc;
}

template <typename T>
int wrap_launch(int64_t count, const void* x_in, void* x_out, cudaStream_t st) {
  wrap_kernel<T><<<stream_grid(count), kThreads, 0, st>>>(count, (const T*)x_in, (T*)x_out);
  return PFB_OK;
}

#ifdef PFB_PL_F32_TU
template int predict_launch<float>(const pfb_predict_params*, int64_t, const void*, void*, const void*, cudaStream_t);
template int loglik_launch<float>(const pfb_lik_params*, Seg, const void*, const double*, double*, double*, Workspace, cudaStream_t);
template int fused_launch<float>(const pfb_predict_params*, const pfb_lik_params*, Seg, const void*, void*, const void*, double*, double*, Workspace, cudaStream_t);
template int wrap_launch<float>(int64_t, const void*, void*, cudaStream_t);
#else
extern template int predict_launch<float>(const pfb_predict_params*, int64_t, const void*, void*, const void*, cudaStream_t);
extern template int loglik_launch<float>(const pfb_lik_params*, Seg, const void*, const double*, double*, double*, Workspace, cudaStream_t);
extern template int fused_launch<float>(const pfb_predict_params*, const pfb_lik_params*, Seg, const void*, void*, const void*, double*, double*, Workspace, cudaStream_t);
extern template int wrap_launch<float>(int64_t, const void*, void*, cudaStream_t);
#endif
}  // namespace pfb

#ifndef PFB_PL_F32_TU
using namespace pfb;

extern "C" int pfb_predict(const pfb_predict_params* p, pfb_dtype dtype, int64_t n, const void* x_in,
                           void* x_out, const void* noise, void* stream) {
  int rc = check_predict(p);
  if (rc) return rc;
  if (n <= 0) return PFB_OK;
  if (!x_in || !x_out) { set_error("NULL particle pointer"); return PFB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, (rc = predict_launch<T>(p, n, x_in, x_out, noise, st)));
  if (rc) return rc;
  PFB_LAUNCH_OK("predict_kernel");
  return PFB_OK;
}

namespace pfb {
static int loglik_impl(const pfb_lik_params* lik, pfb_dtype dtype, Seg sg, const void* x, const double* w_prev,
                       double* logw, double* stats, void* ws, int64_t ws_bytes, void* stream) {
  if (!lik) { set_error("likelihood params NULL"); return PFB_ERR_INVALID; }
  int rc = check_lik(lik, lik->dim);
  if (rc) return rc;
  if (sg.n_runs <= 0 || sg.run_len <= 0) return PFB_OK;
  if (!x || !logw || !stats) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  Workspace W;
  rc = carve_workspace(ws, ws_bytes, sg.n_runs == 1 ? sg.run_len : sg.n_runs * sg.tpr * (int64_t)kScanTile, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, (rc = loglik_launch<T>(lik, sg, x, w_prev, logw, stats, W, st)));
  if (rc) return rc;
  PFB_LAUNCH_OK("loglik_kernel");
  return PFB_OK;
}

static int predict_loglik_impl(const pfb_predict_params* p, const pfb_lik_params* lik, pfb_dtype dtype, Seg sg,
                               const void* x_in, void* x_out, const void* noise, double* logw, double* stats,
                               void* ws, int64_t ws_bytes, void* stream) {
  int rc = check_predict(p);
  if (rc) return rc;
  rc = check_lik(lik, p->dim);
  if (rc) return rc;
  if (sg.n_runs <= 0 || sg.run_len <= 0) return PFB_OK;
  if (!x_in || !logw || !stats || (!x_out && !p->x_pad_out)) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  Workspace W;
  rc = carve_workspace(ws, ws_bytes, sg.n_runs == 1 ? sg.run_len : sg.n_runs * sg.tpr * (int64_t)kScanTile, &W);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFB_DISPATCH_DTYPE(dtype, (rc = fused_launch<T>(p, lik, sg, x_in, x_out, noise, logw, stats, W, st)));
  if (rc) return rc;
  PFB_LAUNCH_OK("predict_loglik_kernel");
  return PFB_OK;
}
}  // namespace pfb

extern "C" int pfb_loglik(const pfb_lik_params* lik, pfb_dtype dtype, int64_t n, const void* x,
                          const double* w_prev, double* logw, double* stats, void* ws, int64_t ws_bytes,
                          void* stream) {
  if (lik && lik->mu_batch_dev) { set_error("mu_batch_dev is only valid with the *_batched entry points"); return PFB_ERR_INVALID; }
  return loglik_impl(lik, dtype, seg_single(n), x, w_prev, logw, stats, ws, ws_bytes, stream);
}

extern "C" int pfb_predict_loglik(const pfb_predict_params* p, const pfb_lik_params* lik, pfb_dtype dtype,
                                  int64_t n, const void* x_in, void* x_out, const void* noise, double* logw,
                                  double* stats, void* ws, int64_t ws_bytes, void* stream) {
  if (lik && lik->mu_batch_dev) { set_error("mu_batch_dev is only valid with the *_batched entry points"); return PFB_ERR_INVALID; }
  return predict_loglik_impl(p, lik, dtype, seg_single(n), x_in, x_out, noise, logw, stats, ws, ws_bytes, stream);
}

extern "C" int pfb_loglik_batched(const pfb_lik_params* lik, pfb_dtype dtype, int64_t n_runs, int64_t run_len,
                                  const void* x, double* logw_pad, double* stats, void* ws, int64_t ws_bytes,
                                  void* stream) {
  return loglik_impl(lik, dtype, seg_batched(n_runs, run_len), x, nullptr, logw_pad, stats, ws, ws_bytes, stream);
}

extern "C" int pfb_predict_loglik_batched(const pfb_predict_params* p, const pfb_lik_params* lik, pfb_dtype dtype,
                                          int64_t n_runs, int64_t run_len, const void* x_in, void* x_out,
                                          const void* noise, double* logw_pad, double* stats, void* ws,
                                          int64_t ws_bytes, void* stream) {
  return predict_loglik_impl(p, lik, dtype, seg_batched(n_runs, run_len), x_in, x_out, noise, logw_pad, stats, ws,
                             ws_bytes, stream);
}

extern "C" int pfb_wrap_2pi(pfb_dtype dtype, int64_t count, const void* x_in, void* x_out, void* stream) {
  if (count <= 0) return PFB_OK;
  if (!x_in || !x_out) { set_error("NULL pointer"); return PFB_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  int rc = PFB_OK;
  PFB_DISPATCH_DTYPE(dtype, (rc = wrap_launch<T>(count, x_in, x_out, st)));
  if (rc) return rc;
  PFB_LAUNCH_OK("wrap_kernel");
  return PFB_OK;
}
#endif  // PFB_PL_F32_TU
