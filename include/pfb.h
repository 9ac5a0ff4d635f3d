/*
 * pfb.h -- C ABI of libpfb.so, the B200 (sm_100a) particle-filter hot path.
 *
 * The reference (pyRecEst v0.9.1, /root/reference) is 100 % Python and has no FFI; the
 * seam this library plugs into is the set of Python methods listed beside each entry point
 * (file:line relative to /root/reference).  pyrecest_b200/ binds these symbols with ctypes
 * (see INTEGRATION.md for the stub a pyRecEst maintainer would add).
 *
 * Conventions
 *   - every pointer named *_dev / x / w / cdf / u / noise ... is DEVICE memory owned by the
 *     caller (torch); small parameter blocks (pfb_*_params) are HOST structs copied by value.
 *   - particle records are row-major (n, dim) -- the reference's own `.d` layout --
 *     element type `dtype` (PFB_F64 default, PFB_F32 storage option); log-weights, weights,
 *     CDFs, uniforms and all reduction outputs are always float64.
 *   - the library never allocates, frees or synchronises: scratch comes from `ws`
 *     (pfb_workspace_bytes), every call is asynchronous on `stream` (a cudaStream_t).
 *   - return value: 0 = PFB_OK, otherwise a pfb_status; pfb_last_error() gives text
 *     (thread-local).  Device-detected conditions (all-zero weights, NaN, negative
 *     likelihood) are reported through the `stats` words the caller reads back.
 */
#ifndef PFB_H
#define PFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFB_MAXD 6           /* max stored components per particle (R^d, T^d: d; S^d: d+1) */
#define PFB_ABI_VERSION 8

typedef enum { PFB_F64 = 0, PFB_F32 = 1 } pfb_dtype;

typedef enum {
  PFB_OK = 0,
  PFB_ERR_INVALID = 1,      /* bad argument (dim, kind, null pointer, alignment) */
  PFB_ERR_WORKSPACE = 2,    /* workspace too small */
  PFB_ERR_CUDA = 3,         /* a CUDA runtime call failed; see pfb_last_error() */
  PFB_ERR_UNSUPPORTED = 4
} pfb_status;

/* manifolds / predict modes -------------------------------------------------------------- */
typedef enum {
  PFB_PRED_EUCLID = 0,       /* x' = f(x) + n                 euclidean_particle_filter.py:59-69 */
  PFB_PRED_TORUS_SHIFT = 1,  /* x' = mod(mod(n + mod(f(x))))  hypertoroidal_particle_filter.py:43-56,
                                abstract_particle_filter.py:49-50, hypertoroidal_wrapped_normal_distribution.py:51-63,109-115 */
  PFB_PRED_TORUS_ADD = 2,    /* x' = mod(f(x) + mod(n + mu))  abstract_particle_filter.py:44-47 */
  PFB_PRED_SPHERE_VMF = 3,   /* S^2: VMF(mode = f(x), kappa) draw from (u, z1, z2); pattern of
                                hyperhemisphere_cart_prod_particle_filter.py:90-95 + scipy _rvs_3d/_rotate_samples */
  PFB_PRED_SPHERE_ADD = 4    /* x' = (f(x) + n)/|f(x) + n| */
} pfb_predict_mode;

typedef struct {
  int32_t mode;              /* pfb_predict_mode */
  int32_t dim;               /* stored components D */
  int32_t has_F;             /* 0: identity transition; 1: x -> F x + b (registry LinearTransition) */
  int32_t has_b;
  int32_t noise_cols;        /* columns of the noise array E (0 = no noise; 3 for SPHERE_VMF) */
  int32_t has_A;             /* 1: noise rows are standard normals, n = z A + mu_noise (A is E x D, the
                                numpy-legacy factor sqrt(s)[:,None] * Vt of svd(C)); 0: rows used as given */
  int32_t has_mu_noise;
  int32_t renormalize;       /* sphere modes: bit 0 renormalise the rotated VMF draw; bit 1 mirror onto the upper
                                hemisphere x_last >= 0 (hyperhemispherical filters identify antipodes) */
  double F[PFB_MAXD * PFB_MAXD];   /* row-major */
  double b[PFB_MAXD];
  double A[PFB_MAXD * PFB_MAXD];   /* row-major E x D */
  double mu_noise[PFB_MAXD];
  double kappa;              /* SPHERE_VMF */
  double exp_m2k;            /* SPHERE_VMF: exp(-2 kappa), computed by the host like scipy does */
  /* device RNG (production path; the parity path injects draws through `noise`): when rng_kind != 0 the
   * `noise` pointer is ignored and the draws are generated in the kernel with Philox4x32-10 keyed by
   * (rng_seed, rng_offset, particle index) -- nothing is read from or written to HBM for them */
  int32_t rng_kind;          /* pfb_rng_kind */
  int32_t rng_vm_table_stride; /* doubles per strip table in rng_vm_table_dev: 0 or 2 * PFB_VM_STRIPS = strips only;
                                * PFB_VM_TABLE_DOUBLES = strips followed by the quick-accept thresholds */
  uint64_t rng_seed;
  uint64_t rng_offset;
  double rng_kappa[PFB_MAXD]; /* PFB_RNG_VONMISES: concentration per dimension */
  double rng_vm_area[PFB_MAXD];  /* PFB_RNG_VONMISES: hat area A of the dimension's strip table (pfb_vm_strip_table) */
  const double* rng_vm_table_dev; /* device pointer: strip tables, PFB_VM_STRIPS x (start, width) doubles each */
  int32_t rng_vm_table_idx[PFB_MAXD]; /* which table of rng_vm_table_dev each dimension uses (equal kappas share) */
  /* optional second output of the predict kernels: the predicted particles once more as records of x_pad_stride
   * components (zero padded) so that a record never straddles a 64-byte fetch -- the source of the random gather
   * of pfb_resample*_strided (a 24-byte T^3 record straddles 25 % of the time, a 32-byte one never) */
  void* x_pad_out;
  int32_t x_pad_stride;      /* components per padded record: 4 for D = 3, 8 for D = 5, 6; 0 = no padded copy */
  int32_t reserved2;
} pfb_predict_params;

/* von Mises draws use vertical-strip rejection: [0, pi] is cut into PFB_VM_STRIPS strips under a piecewise-
 * constant hat of equal areas A, so that one Philox block gives strip, position, acceptance uniform and sign, and
 * a draw is rejected with probability ~1/2000 only.  pfb_vm_strip_table (host code, no GPU) fills
 * table[2k] = start a_k, table[2k+1] = width w_k and *area = A for one concentration kappa >= 1e-8. */
#define PFB_VM_STRIPS 4096
int pfb_vm_strip_table(double kappa, double* table, double* area);
/* Quick-accept thresholds of the same table (host code): thr[k] is a 32-bit integer such that an acceptance word
 * v32 < thr[k] is accepted anywhere in strip k (the density is decreasing, so its value at the strip's far end bounds it
 * from below) -- ~99 % of the attempts are decided by one integer compare, with exactly the outcome of the full test.
 * A device table with thresholds is the 2 * PFB_VM_STRIPS doubles of pfb_vm_strip_table followed by the PFB_VM_STRIPS
 * uint32 thresholds (PFB_VM_TABLE_DOUBLES doubles in all; pfb_predict_params.rng_vm_table_stride says which). */
#define PFB_VM_TABLE_DOUBLES (2 * PFB_VM_STRIPS + PFB_VM_STRIPS / 2)
int pfb_vm_strip_thresholds(double kappa, const double* table, double area, uint32_t* thr);

typedef enum {
  PFB_RNG_NONE = 0,
  PFB_RNG_NORMAL = 1,     /* standard normals in place of the noise rows (combine with has_A / mu_noise) */
  PFB_RNG_VONMISES = 2,   /* zero-mean von Mises draws per dimension */
  PFB_RNG_VMF = 3         /* (u, z1, z2) triples for PFB_PRED_SPHERE_VMF */
} pfb_rng_kind;

/* likelihood registry ---------------------------------------------------------------------- */
typedef enum {
  PFB_LIK_NONE = -1,
  PFB_LIK_GAUSS = 0,     /* gaussian_distribution.py:43-50 (scipy mvn.pdf)                      */
  PFB_LIK_WN_MULTI = 1,  /* hypertoroidal_wrapped_normal_distribution.py:65-93, image table      */
  PFB_LIK_WN_1D = 2,     /* circle/wrapped_normal_distribution.py:59-127 (series until converged) */
  PFB_LIK_VM = 3,        /* circle/von_mises_distribution.py:42-50                                */
  PFB_LIK_VMF = 4        /* hypersphere_subset/von_mises_fisher_distribution.py:51-54             */
} pfb_lik_kind;

typedef struct {
  int32_t kind;
  int32_t dim;               /* stored components D */
  int32_t n_images;          /* WN_MULTI: rows of images_dev */
  int32_t cell_grid;         /* WN_MULTI, D <= 3, n_images <= 64, G <= 16: G > 0 says that images_dev is
                                followed by G^D uint64 candidate masks, one per cell of the box dev in
                                [-pi - mu, pi - mu)^D cut into G cells per dimension (index c_0 + G c_1 + G^2 c_2):
                                bit k set iff image k can pass the screen somewhere in the cell.  0: screen all. */
  double mu[PFB_MAXD];
  double W[PFB_MAXD * PFB_MAXD]; /* GAUSS / WN_MULTI: whitening matrix (row-major D x D), y = (x - mu) W */
  double logc;               /* log normalising constant added to every log-likelihood */
  double kappa;              /* VM / VMF */
  double sigma;              /* WN_1D */
  const double* images_dev;  /* WN_MULTI: DEVICE table (n_images, 2D + 1): g_k = P o_k (D), h_k = o_k^T P o_k, o_k (D) */
  const double* mu_batch_dev; /* *_batched only: DEVICE (n_runs, D) likelihood centres, one measurement per run
                                 (NULL: `mu` for every run) */
  int32_t out_format;        /* pfb_weight_format of the `logw` output of the weighting calls */
  int32_t cell_span;         /* bit 0: the cell grid covers dev = wrap(x) - mu in [-3 pi, pi)^D, the union over every mu in
                                [0, 2 pi)^D, with 2 cell_grid cells per dimension -- valid for any measurement, so the
                                *_batched calls (one mu per run, mu_batch_dev) can use it; clear: the box of ONE
                                measurement (above).  bit 1: the masks are uint32 (n_images <= 32), followed by the
                                reference bytes -- a 16^3 grid then fits the shared memory of a CTA */
} pfb_lik_params;

/* What the weighting calls (pfb_loglik, pfb_predict_loglik and their *_batched forms) write into `logw`:
 *   PFB_WEIGHTS_LOG     log-weights  l_i = log pdf(x_i) (+ log w_prev_i)
 *   PFB_WEIGHTS_SCALED  linear weights scaled per slice of 256 consecutive particles by a power of two,
 *                       p_i = pdf(x_i) w_prev_i 2^(-e_s), e_s = ceil(max_slice(l) log2 e); the exponents stay in the
 *                       workspace.  Only pfb_scan_cdf / pfb_scan_cdf_batched called next on the same buffers with
 *                       PFB_SCAN_TILE_STATS_VALID | PFB_SCAN_SLICE_SCALED can read this format: its addends are
 *                       p_i 2^(e_s - E), E the largest exponent -- exact scalings, so the step spends one exp per
 *                       particle in all and keeps the range of the log domain. */
typedef enum { PFB_WEIGHTS_LOG = 0, PFB_WEIGHTS_SCALED = 1 } pfb_weight_format;

/* stats words written by the weighting / normalise kernels (device doubles) */
enum {
  PFB_STAT_MAX = 0,      /* max_i logw_i                                   */
  PFB_STAT_SUMEXP = 1,   /* sum_i exp(logw_i - max)                        */
  PFB_STAT_SUMSQ = 2,    /* sum_i exp(logw_i - max)^2  (ESS = S1^2/S2)     */
  PFB_STAT_FLAGS = 3,    /* bit0: NaN seen, bit1: all weights zero/-inf    */
  PFB_STAT_TOTAL = 4,    /* last raw CDF entry c_{N-1}                      */
  PFB_STAT_COUNT = 8
};

typedef enum { PFB_SCAN_FAST = 0, PFB_SCAN_SEQEXACT = 1 } pfb_scan_mode;
/* OR into `mode`: the workspace still holds the per-tile (max, sum exp) that pfb_loglik /
 * pfb_predict_loglik left for exactly this logw array, so pfb_scan_cdf can skip its own pass over it */
#define PFB_SCAN_TILE_STATS_VALID 0x100
/* OR into `mode` (together with PFB_SCAN_TILE_STATS_VALID): `logw` holds PFB_WEIGHTS_SCALED output */
#define PFB_SCAN_SLICE_SCALED 0x200
typedef enum { PFB_EST_NONE = 0, PFB_EST_SUM = 1, PFB_EST_TRIG = 2 } pfb_est_kind;
typedef enum { PFB_MOM_MEAN = 0, PFB_MOM_TRIG = 1, PFB_MOM_SCATTER = 2, PFB_MOM_CENTRAL = 3 } pfb_moment_kind;

int pfb_abi_version(void);
const char* pfb_last_error(void);

/* scratch needed by any entry point for up to n particles: counters, reduction partials, the job list of heavy
 * particles (systematic resampling), scan tile states, and the lookup structure of multinomial resampling (one
 * 64-byte bucket record per ~16 particles: ~4 bytes per particle).  The library never allocates.
 * The first pfb_workspace_bytes(0) bytes must be zeroed once (pfb_workspace_init). */
int64_t pfb_workspace_bytes(int64_t n);
int64_t pfb_workspace_bytes_batched(int64_t n_runs, int64_t run_len);  /* for the *_batched entry points */
int pfb_workspace_init(void* ws, int64_t ws_bytes, void* stream);

/* K(1) predict.  Replaces AbstractParticleFilter.predict_nonlinear
 * (pyrecest/filters/abstract_particle_filter.py:21-54) and its Euclidean / hypertoroidal
 * overrides.  x_in may equal x_out (in place). */
int pfb_predict(const pfb_predict_params* p, pfb_dtype dtype, int64_t n,
                const void* x_in, void* x_out, const void* noise, void* stream);

/* K(2) log-likelihood weighting: logw_i = loglik(x_i) + log(w_prev_i) (w_prev NULL = uniform 1/n
 * is NOT added: a constant shift cancels in the normalisation).  Also reduces max_i into stats[0].
 * Replaces the likelihood evaluation inside AbstractDiracDistribution.reweigh
 * (pyrecest/distributions/abstract_dirac_distribution.py:69-80). */
int pfb_loglik(const pfb_lik_params* lik, pfb_dtype dtype, int64_t n, const void* x,
               const double* w_prev, double* logw, double* stats, void* ws, int64_t ws_bytes,
               void* stream);

/* K(1)+K(2) fused: predict, write x_out, weight the predicted particle.  x_out may be NULL when p->x_pad_out is
 * given: a caller that resamples next reads only the padded copy, so the compact predicted particles are dead. */
int pfb_predict_loglik(const pfb_predict_params* p, const pfb_lik_params* lik, pfb_dtype dtype,
                       int64_t n, const void* x_in, void* x_out, const void* noise,
                       double* logw, double* stats, void* ws, int64_t ws_bytes, void* stream);

/* Batched Monte-Carlo runs (BASELINE.json configs[3]; pyrecest/evaluation/iterate_configs_and_runs.py:33-49 loops
 * over runs sequentially): n_runs independent filters of run_len particles each in ONE launch.  Particles are
 * (n_runs * run_len, D) run-major; per-particle scalars use a padded stride run_pad = ceil(run_len/2048)*2048
 * (logw_pad / cdf_pad hold n_runs * run_pad doubles; padding carries weight 0) so that scan tiles never straddle
 * runs.  The measurement of run r is lik->mu_batch_dev[r]; all other parameters are shared. */
int pfb_loglik_batched(const pfb_lik_params* lik, pfb_dtype dtype, int64_t n_runs, int64_t run_len,
                       const void* x, double* logw_pad, double* stats, void* ws, int64_t ws_bytes, void* stream);
int pfb_predict_loglik_batched(const pfb_predict_params* p, const pfb_lik_params* lik, pfb_dtype dtype,
                               int64_t n_runs, int64_t run_len, const void* x_in, void* x_out,
                               const void* noise, double* logw_pad, double* stats, void* ws, int64_t ws_bytes,
                               void* stream);
/* per-run sequential-exact CDFs (each run's CDF equals numpy.cumsum of that run's weights, shifted by the
 * run's own max); must follow one of the two calls above (mode | PFB_SCAN_TILE_STATS_VALID) */
int pfb_scan_cdf_batched(int32_t mode, int64_t n_runs, int64_t run_len, const double* logw_pad,
                         double* cdf_pad, double* stats, void* ws, int64_t ws_bytes, void* stream);
/* per-run resampling: run r draws n_out_run offspring from its own CDF with uniforms u[r * n_out_run + k]
 * (or in-kernel Philox draws when u == NULL; systematic: one u0 per run, u[r]); x_out is (n_runs*n_out_run, D) */
int pfb_resample_batched(pfb_dtype dtype, int32_t dim, int64_t n_runs, int64_t run_len, int64_t n_out_run,
                         const double* cdf_pad, const double* u, uint64_t rng_seed, uint64_t rng_offset,
                         int32_t systematic, const void* x_in, void* x_out, int64_t* idx_out, void* ws,
                         int64_t ws_bytes, void* stream);
/* per-run moments: out is (n_runs, count) with the layout of pfb_moments per run; uniform weights 1/run_len */
int pfb_moments_batched(int32_t kind, pfb_dtype dtype, int32_t dim, int64_t n_runs, int64_t run_len,
                        const void* x, int32_t order, double* out, void* stream);

/* K(3) max / log-sum-exp normalisation: stats <- (max, sum exp, sum exp^2, flags); if w_out
 * != NULL, w_out_i = exp(logw_i - max) / sumexp  (the posterior weights reweigh() returns). */
int pfb_normalize(int64_t n, const double* logw, double* w_out, double* stats, void* ws,
                  int64_t ws_bytes, void* stream);

/* K(4a) inclusive prefix sum of non-negative weights -> raw CDF c (not divided by c_{n-1}).
 * Input is either w (logw == NULL) or exp(logw_i - stats[PFB_STAT_MAX]) (w == NULL).
 * PFB_SCAN_SEQEXACT reproduces numpy.cumsum's sequential fp64 rounding bit for bit
 * (numpy Generator.choice: cdf = p.cumsum(), reached from
 * pyrecest/_backend/_shared_numpy/random.py:28-29); PFB_SCAN_FAST is a plain blocked scan.
 * stats[PFB_STAT_TOTAL] <- c_{n-1}. */
int pfb_scan_cdf(int32_t mode, int64_t n, const double* w, const double* logw, double* cdf,
                 double* stats, void* ws, int64_t ws_bytes, void* stream);

/* K(4b) resampling: idx_j = #{i : fl(c_i / c_{n-1}) <= u_j} (searchsorted side='right', the
 * division done per comparison so it is the reference's `cdf /= cdf[-1]` bit for bit), then
 * x_out[j] = x_in[idx_j].  systematic != 0: u_j = (j + u[0]) / n_out computed in-kernel, idx
 * clamped to n-1.  idx_out (int64) optional.  est_kind: also reduce the resampled particles
 * (uniform weights 1/n_out) into est_out: PFB_EST_SUM -> mean (D doubles), PFB_EST_TRIG ->
 * (sum cos, sum sin)/n_out per dim (2D doubles).
 * Replaces AbstractDiracDistribution.sample (abstract_dirac_distribution.py:82-84) and the
 * resampling step of update_nonlinear_using_likelihood (abstract_particle_filter.py:122-125). */
int pfb_resample(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                 const double* u, int32_t systematic, const void* x_in, void* x_out,
                 int64_t* idx_out, int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes,
                 void* stream);
/* The same with the uniforms generated in the kernel (u_j = Philox(seed, offset, j), or u0 =
 * Philox(seed, offset, 0) for systematic): the production path, no uniform array in HBM. */
int pfb_resample_rng(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                     uint64_t rng_seed, uint64_t rng_offset, int32_t systematic, const void* x_in,
                     void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out, void* ws,
                     int64_t ws_bytes, void* stream);
/* pfb_resample / pfb_resample_rng with the gather source given as records of x_in_stride components (the padded
 * copy a predict kernel left behind through pfb_predict_params.x_pad_out; x_in_stride = dim is the plain call).
 * Multinomial only. */
int pfb_resample_strided(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                         const double* u, const void* x_in, int32_t x_in_stride, void* x_out, int64_t* idx_out,
                         int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes, void* stream);
int pfb_resample_rng_strided(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                             uint64_t rng_seed, uint64_t rng_offset, const void* x_in, int32_t x_in_stride,
                             void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out, void* ws,
                             int64_t ws_bytes, void* stream);
/* "multinomial_sorted": exact multinomial resampling with in-kernel draws whose offspring are emitted in CDF
 * order instead of the order of the reference's uniform stream (abstract_dirac_distribution.py:82-84 draws n_out
 * i.i.d. uniforms; here the SAME joint distribution is generated leaf by leaf: [0, 1) is cut into K = 2^k equal
 * leaves, the leaf occupancies -- multinomial(n_out; 1/K, ..) -- come from exact binomial splitting with p = 1/2
 * (popcounts of Philox bit streams) and each leaf's uniforms from fresh (53 - k)-bit draws).  The output is a
 * permutation of what pfb_resample_rng would give for the same multiset of uniforms; idx_out (optional) receives the
 * source index of every output slot, non-decreasing up to the order inside a leaf.  Lookups hit neighbouring cache
 * lines and the stores stream, which removes the two DRAM misses per particle of the reference-order gather.
 * x_in_stride: components per source record (0 = dim).  est_kind / est_out as in pfb_resample. */
int pfb_resample_msort(pfb_dtype dtype, int32_t dim, int64_t n, int64_t n_out, const double* cdf,
                       uint64_t rng_seed, uint64_t rng_offset, const void* x_in, int32_t x_in_stride, void* x_out,
                       int64_t* idx_out, int32_t est_kind, double* est_out, void* ws, int64_t ws_bytes, void* stream);
/* fill `out` (n_out doubles, slot order) with the uniforms pfb_resample_msort uses (tests, debugging) */
int pfb_msort_uniforms(int64_t n_out, uint64_t rng_seed, uint64_t rng_offset, double* out, void* ws, int64_t ws_bytes,
                       void* stream);

/* Sharded filter (one filter split over several GPUs): this rank owns particles
 * [start, start + n) of a global CDF C_i = fl(cdf_offset + c_i), normalised by `total` (the sum over all
 * shards).  Produces the `count` systematic offspring j = j_begin .. j_begin + count - 1 of the global
 * comb u_j = (j + u0) / n_total_out:  x_out[k] = x_in[#{i : fl(C_i / total) <= u_j}]  (index local to the
 * shard, clamped to n - 1); est_kind / est_out as in pfb_resample (mean over these `count` offspring).
 * The caller picks [j_begin, j_begin + count) as the offspring that fall into this shard
 * (pyrecest_b200/sharded.py).  x_out may be PEER memory (a pointer into another GPU's buffer, mapped through CUDA
 * IPC / torch symmetric memory): the kernel then stores the offspring that belong to that rank straight into its
 * output slots over NVLink -- compute and exchange in one kernel; otherwise the pieces are exchanged with an
 * all-to-all.  Only the sources of the piece are visited (bisection on the monotone offspring count).
 * ws may be NULL when est_kind == PFB_EST_NONE (runs of > 8192 offspring of one particle are then written by one
 * warp instead of the whole grid). */
int pfb_resample_sharded(pfb_dtype dtype, int32_t dim, int64_t n, const double* cdf, double cdf_offset,
                         double total, double u0, int64_t j_begin, int64_t count, int64_t n_total_out,
                         const void* x_in, void* x_out, int64_t* idx_out, int32_t est_kind, double* est_out,
                         void* ws, int64_t ws_bytes, void* stream);
/* The same with the plan made on the device: ONE call per rank and step.  totals_dev: DEVICE array of the `world`
 * shard totals in rank order (an all-gather of stats[PFB_STAT_TOTAL]); the shard offsets O_g = fl(O_{g-1} + T_{g-1}) and
 * the global total are formed from it in the kernel, so the host never reads a weight.  peer_out: HOST array of `world`
 * pointers, peer_out[r] = the (n, dim) output buffer of rank r (its own for r == rank, peer-mapped otherwise): the
 * offspring of global slot j is stored at row j % n of peer_out[j / n].  Every source of this shard is visited once;
 * est_sums receives the SUMS over the offspring produced here (x, or cos / sin per dimension): all-reduce them and
 * divide by n * world.  status[0] (DEVICE) <- 1 when the global total is not a positive finite number (nothing is
 * written then), else 0. */
int pfb_resample_sharded_dev(pfb_dtype dtype, int32_t dim, int64_t n, const double* cdf, const double* totals_dev,
                             int32_t world, int32_t rank, double u0, void* const* peer_out, const void* x_in,
                             int32_t est_kind, double* est_sums, double* status, void* ws, int64_t ws_bytes,
                             void* stream);
/* The same resampling in PULL form: this rank produces exactly its own n output slots (x_out, local) and reads the
 * sources they come from wherever they live.  peer_cdf[r] / peer_x[r] (HOST arrays of `world` pointers): the CDF
 * shard (n doubles, from pfb_scan_cdf against the global shift) and the (n, dim) particle buffer of rank r -- its
 * own for r == rank, peer-mapped otherwise (remote sources are read over NVLink with ordinary loads; every store is
 * local).  The work per rank is its n slots whatever the weights, where the push form makes the owner of the heavy
 * sources produce all their offspring.  The caller must order the peers' scans before this call (the all-gather of
 * the totals does) and a collective after it before any rank overwrites a buffer a peer may still be reading (the
 * all-reduce of est_sums does).  est_sums: SUMS over this rank's n offspring.  Identical offspring to
 * pfb_resample_sharded_dev, slot for slot. */
int pfb_resample_sharded_pull(pfb_dtype dtype, int32_t dim, int64_t n, const double* const* peer_cdf,
                              const void* const* peer_x, const double* totals_dev, int32_t world, int32_t rank,
                              double u0, void* x_out, int32_t est_kind, double* est_sums, double* status, void* ws,
                              int64_t ws_bytes, void* stream);
/* fill `out` (n doubles) with the uniforms pfb_resample_rng would use (tests, debugging) */
int pfb_rng_uniform(int64_t n, uint64_t rng_seed, uint64_t rng_offset, double* out, void* stream);

/* K(5) weighted moments (w == NULL: uniform 1/n).  out (doubles):
 *   PFB_MOM_MEAN    sum w x              (D)      linear_dirac_distribution.py:19-21, hyperspherical_dirac_distribution.py:43-44
 *   PFB_MOM_TRIG    sum w cos(order x), sum w sin(order x)   (2D)  hypertoroidal_dirac_distribution.py:72-80
 *   PFB_MOM_SCATTER sum w x x^T          (D*D)    abstract_hypersphere_subset_dirac_distribution.py:18-26
 *   PFB_MOM_CENTRAL sum w (x-c)(x-c)^T / sum w, c = center (D DEVICE doubles, e.g. a PFB_MOM_MEAN output)  linear_dirac_distribution.py:73-82
 * plus out[last] = sum w. */
int pfb_moments(int32_t kind, pfb_dtype dtype, int32_t dim, int64_t n, const void* x,
                const double* w, int32_t order, const double* center, double* out, void* ws,
                int64_t ws_bytes, void* stream);

/* element-wise numpy.mod(x, 2 pi) on n*dim values (HypertoroidalDiracDistribution.__init__,
 * hypertoroidal_dirac_distribution.py:41) */
int pfb_wrap_2pi(pfb_dtype dtype, int64_t count, const void* x_in, void* x_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PFB_H */
