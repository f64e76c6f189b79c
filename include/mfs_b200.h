/*
 * mfs_b200.h -- C ABI of the B200-native batched moment filter (drop-in for the hot path of zgbkdlm/mfs).
 *
 * The reference has no native boundary: its "operator API" is the Python signature of
 *   moment_filter_rms(state_cond_raw_moments, measurement_cond_pdf, rms0, ys, stable)      mfs/one_dim/filtering.py:32-36
 *   moment_filter_cms(state_cond_central_moments, state_cond_mean, pdf, cms0, mean0, ys)   mfs/one_dim/filtering.py:92-98
 *   moment_filter_scms(... scms0, mean0, scale0, ys)                                       mfs/one_dim/filtering.py:164-172
 *   moment_quadrature(ms, mean, scale, sort_nodes, ldl)                                    mfs/one_dim/quadtures.py:83-85
 * taking Python callables.  Here the callables become *named device functors* (ids below) with packed double
 * parameters, and one call runs B independent filters (the batch axis the reference fans out over OS processes,
 * dardel/run_benes_bernoulli_mf.sh:26-45).  INTEGRATION.md shows the ctypes / XLA-FFI binding a maintainer adds.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures: device pointers are `void*`/`double*`, the stream is `void*`
 *     (a cudaStream_t; NULL = legacy default stream).
 *   - the caller owns every buffer; `mfs_filter_1d` never allocates.  `mfs_filter_1d_host` (host buffers) owns a
 *     device staging workspace, cached between calls and returned by mfs_release_cached_memory().
 *   - return value 0 = launched/completed; negative = argument / CUDA error, text via mfs_last_error()
 *     (thread-local).  NUMERICAL failure is not an error: like the JAX scan it yields NaN from the failing step
 *     onwards, and status_out[b] = index of the first failed step (-1: none).
 *   - all strides are in ELEMENTS of the pointed-to type.  A stride of 0 on a per-filter input means "shared by all
 *     filters".
 */
#ifndef MFS_B200_H_
#define MFS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFS_ABI_VERSION 3
#define MFS_MAX_N 15          /* quadrature nodes; 2N moments.  Reference sweeps N=2..15 (dardel/run_time_profile.sh:25) */
#define MFS_MAX_PARAMS 4

/* moment representation: mfs/definitions.py (rms / cms / scms) */
enum { MFS_MODE_RAW = 0, MFS_MODE_CENTRAL = 1, MFS_MODE_SCALED = 2 };

/* transition-moment family: the three factories of mfs/one_dim/moments.py */
enum {
  MFS_TRANS_TME = 0,        /* sde_cond_moments_tme          moments.py:141-179  (order 1..3)              */
  MFS_TRANS_TME_NORMAL = 1, /* sde_cond_moments_tme_normal   moments.py:182-219  N(mean,var) from TME      */
  MFS_TRANS_EULER = 2,      /* sde_cond_moments_euler        moments.py:222-255  N(x+a dt, b^2 dt)         */
  MFS_TRANS_NORMAL_AFFINE = 3 /* N(F x, Sigma): exact OU closure, dardel/convergence/convergence_mf.py:86-107;
                                 trans_params = {F, Sigma} */
};

/* drift a(x) of dX = a(X) dt + b dW, constant dispersion b */
enum {
  MFS_DRIFT_BENES = 0,  /* tanh(x)                 mfs/one_dim/ss_models.py:37   params: -            */
  MFS_DRIFT_WELL = 1,   /* x (1 - theta1 x^2)      mfs/one_dim/ss_models.py:71   params: {theta1}     */
  MFS_DRIFT_LINEAR = 2  /* a x  (OU: a = -1/ell)   tests/test_filtering.py:45    params: {a}          */
};

/* measurement model p(y | x) */
enum {
  MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC = 0, /* Bernoulli(1/(1+exp(-(x^3/c0 - c1)))); Benes: {5,0} ss_models.py:43-47 */
  MFS_MEAS_POISSON_SOFTPLUS = 1,         /* Poisson(log(1+exp(theta2 x)))   params {theta2}   ss_models.py:80-84  */
  MFS_MEAS_GAUSSIAN = 2                  /* N(y; h x, r^2)   params {h, r}    tests/test_filtering.py:41-42       */
};

enum { MFS_YS_U8 = 0, MFS_YS_I32 = 1, MFS_YS_F64 = 2 };

/* what the kernel writes back */
enum {
  MFS_OUT_FULL = 0, /* every step: ms_out[b][t][0..2N) (+ mean_out[b][t], scale_out[b][t]) like the reference */
  MFS_OUT_LAST = 1, /* only the final step's moments (ms_out[b][0..2N)), mean, scale                          */
  MFS_OUT_NONE = 2, /* only nell (parameter-estimation objective, dardel/parameter_estimation/mf.py:52)       */
  MFS_OUT_MEANVAR = 3 /* every step: (mean, variance) of the filtering distribution, ms_out[b][t][0..2) -- what the
                         reference's consumers read from the history (dardel/prey_predator/mf.py:84-88,
                         dardel/benes_bernoulli/post_processing_mf.py:41-66): 16 B per step instead of 16 N.
                         raw: (m1, m2 - m1^2); central: (mean, cm2); scaled: (mean, scale^2 scm2).
                         mean_out / scale_out are not written.
                         2-D filter: ms_out[b][t][0..5) = (E x1, E x2, Var x1, Cov(x1, x2), Var x2). */
};

/* flags.
 * By default the prediction half-step does not re-derive its quadrature from the posterior moments: the posterior of
 * the previous update IS an N-atom measure {x_i, w_i l_i / c}, whose N-point Gauss rule is itself, so the atoms are
 * carried over (the Hankel pivots are still checked on the moments, so "matrix not positive definite -> NaN" fires on
 * the same quantity as the reference's Cholesky).  MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE forces the literal recursion
 * of mfs/one_dim/filtering.py:78 (a second moment_quadrature per step), which agrees to rounding. */
#define MFS_FLAG_RECOMPUTE_PREDICT_QUADRATURE 1

typedef struct mfs_filter1d_args {
  int32_t abi_version;   /* MFS_ABI_VERSION */
  int32_t mode;          /* MFS_MODE_*  */
  int32_t N;             /* 2..MFS_MAX_N */
  int32_t stable;        /* 0/1: `stable=True` LDL completion (mfs/utils.py:526-538) */
  int64_t B;             /* independent filters */
  int64_t T;             /* time steps */

  int32_t trans_id;      /* MFS_TRANS_* */
  int32_t drift_id;      /* MFS_DRIFT_* (ignored by NORMAL_AFFINE) */
  int32_t tme_order;     /* 1..3 for TME / TME_NORMAL */
  int32_t meas_id;       /* MFS_MEAS_* */
  double dt;
  double dispersion;     /* constant b */
  const double* trans_params; /* device, [B or 1][MFS_MAX_PARAMS] */
  int64_t trans_param_stride; /* elements between filters; 0 = shared */
  const double* meas_params;  /* device, [B or 1][MFS_MAX_PARAMS] */
  int64_t meas_param_stride;

  const double* ms0;     /* device, initial moments [B or 1][2N] */
  int64_t ms0_stride;
  const double* mean0;   /* CENTRAL/SCALED: [B or 1] */
  int64_t mean0_stride;
  const double* scale0;  /* SCALED: [B or 1] */
  int64_t scale0_stride;

  const void* ys;        /* device, measurement y[b][t] at ys + b*ys_stride_b + t*ys_stride_t */
  int32_t ys_dtype;      /* MFS_YS_* */
  int32_t out_mode;      /* MFS_OUT_* */
  int64_t ys_stride_b;
  int64_t ys_stride_t;

  double* ms_out;        /* FULL: [B][T][2N] via strides below; LAST: [B][2N] (ms_stride_t ignored);
                            MEANVAR: [B][T][2] via the same strides */
  int64_t ms_stride_b;
  int64_t ms_stride_t;
  double* mean_out;      /* FULL: [B][T] (stride_b = aux_stride_b, stride_t = 1); LAST: [B]; may be NULL in RAW mode */
  double* scale_out;     /* same shape as mean_out; SCALED mode only */
  int64_t aux_stride_b;
  double* nell_out;      /* [B] negative log-likelihood */
  int32_t* status_out;   /* [B] first failed step or -1; may be NULL */
  int32_t flags;         /* MFS_FLAG_* */
  int32_t segment_steps; /* segment length of the segmented execution below; 0 = default (64) */
  /* Optional caller-owned device scratch (size from mfs_filter_1d_workspace_bytes).  With a workspace and T >= 2
   * segments the scan runs segment by segment and every segment only launches the filters that are still alive,
   * re-packed densely (a diverged filter otherwise idles its lane -- and its warp -- until the end of the scan).
   * Results are identical with and without; NULL = one launch for the whole scan. */
  void* workspace;
  int64_t workspace_bytes;
  /* Time-chunked execution (checkpoint / resume): the scan of one record set may be cut into calls over consecutive
   * blocks of measurements.  carry_out (device, [B][4N + 4] doubles, may be NULL) receives the filter state after the
   * last step of this call; handing it to the next call as carry_in (may be NULL = start from ms0 / mean0 / scale0,
   * which are then ignored) continues the scan bit-identically to one long call.  nell_out is the running total.
   * Layout of one state: ms[2N], atom weights[N], atom nodes[N], mean, scale, nell, flag (flag < 0: the filter failed
   * at absolute step -flag - 1 and stays NaN).  t_offset = absolute index of this call's first step (only used for
   * the step index reported in status_out).  mfs_filter_1d (device pointers) only. */
  const double* carry_in;
  double* carry_out;
  int64_t t_offset;
  /* Parameter grids over shared records (the theta grid of dardel/parameter_estimation/mf.py's objective): with
   * grid_records = P > 0 the batch is a (B / P) x P grid, filter b = g P + j reads record j (ys + j ys_stride_b) and
   * row g of every per-filter input table (ms0, mean0, scale0, trans_params, meas_params: stride x g).  Outputs stay
   * indexed by b.  0 = one record and one table row per filter.  mfs_filter_1d (device pointers) only. */
  int64_t grid_records;
} mfs_filter1d_args;

/* Scratch size in bytes for mfs_filter_1d's segmented execution (-1 for invalid arguments). */
int64_t mfs_filter_1d_workspace_bytes(int32_t N, int64_t B, int64_t T);

/* ---- d = 2 moment filter (mfs/multi_dims/filtering.py:210-344, quadratures.py:120-178) ------------------------------
 * Moments are ordered graded-lexicographically (mfs/multi_dims/multi_indices.py): position of the multi-index (a, b)
 * is (a+b)(a+b+1)/2 + a, z = N(2N+1) moments with |n| <= 2N-1; `inds` is the (3, s, s) int32 gather table of
 * gram_and_hankel_indices_graded_lexico(N, 2), s = N(N+1)/2.  Transition: the Normal families of
 * mfs/multi_dims/moments.py:257-411 for the Lotka--Volterra SDE of mfs/multi_dims/ss_models.py:40-61
 * (trans_params = {alpha, beta, delta, gamma, sigma}); measurement: MFS_MEAS_BERNOULLI_LOGISTIC_CUBIC on x[obs_dim]
 * (prey--predator: {c0, c1} = {1, 1}, ss_models.py:63-67).  Device pointers. */
typedef struct mfs_filternd_args {
  int32_t abi_version;
  int32_t mode;          /* MFS_MODE_RAW | MFS_MODE_CENTRAL */
  int32_t N;             /* 2..7 (one basis row per lane: N(N+1)/2 <= 32; the reference runs N = 5 and 7) */
  int32_t d;             /* must be 2 */
  int64_t B, T;
  int32_t trans_id;      /* MFS_TRANS_EULER | MFS_TRANS_TME_NORMAL | MFS_TRANS_TME (moments.py:414-479) */
  int32_t tme_order;     /* 1..2 */
  int32_t meas_id;
  int32_t obs_dim;
  double dt;
  const double* trans_params; int64_t trans_param_stride;   /* [B|1][8] */
  const double* meas_params;  int64_t meas_param_stride;    /* [B|1][4] */
  const double* ms0;   int64_t ms0_stride;                  /* [B|1][z] */
  const double* mean0; int64_t mean0_stride;                /* [B|1][2], CENTRAL */
  const uint8_t* ys;                                        /* [B][T] */
  const int32_t* inds;                                      /* [3][s][s] */
  int32_t out_mode;      /* MFS_OUT_* */
  int32_t stable;        /* 0/1: `stable=True`, LDL completion of the Gram factor (mfs/utils.py:526-538) */
  double* ms_out;        /* FULL [B][T][z] | LAST [B][z] | MEANVAR [B][T][5] */
  double* mean_out;      /* FULL [B][T][2] | LAST [B][2] (CENTRAL; not written in MEANVAR / NONE) */
  double* nell_out;      /* [B] */
  int32_t* status_out;   /* [B] or NULL */
} mfs_filternd_args;

/* Enqueue B two-dimensional filters x T steps on `stream` (one warp per filter). */
int mfs_filter_nd(const mfs_filternd_args* a, void* stream);

/* ---- brute-force grid filter (mfs/classical_filters_smoothers/brute_force.py:26-136) ---------------------------------
 * B measurement records filtered on ONE shared spatial grid xs[n].  pred_method: brute_force.py:50-58.  The 'chapman'
 * methods run every integration sub-step as an FP64 tensor-core GEMM  S' = S Pw^T  (S: [B][n] densities, Pw[i][j] =
 * N(x_i; m_j, s_j) * trapezoid weight_j); 'kolmogorov' is the explicit finite-difference Euler scheme.
 * Drift functor + constant dispersion as in mfs_filter1d_args; trans_params are shared by all filters (one operator).
 * Device pointers; the caller owns the workspace (size from mfs_brute_force_workspace_bytes). */
enum { MFS_BF_CHAPMAN_EULER = 0, MFS_BF_CHAPMAN_TME = 1, MFS_BF_KOLMOGOROV = 2 };
/* mfs_brute_force_args.flags.  POWER_OPERATOR ('chapman' methods, integration_steps > 1): the sub-steps of a time step
 * apply the same linear operator (brute_force.py:115-122), so the k-th power of the operator is formed once per call by
 * binary powering on the FP64 tensor-core GEMM and every time step is ONE contraction instead of k.  Same mathematics,
 * different rounding (associativity; all terms are non-negative: relative differences ~1e-13) -- hence opt-in. */
enum { MFS_BF_FLAG_POWER_OPERATOR = 1 };

typedef struct mfs_brute_force_args {
  int32_t abi_version;
  int32_t pred_method;        /* MFS_BF_* */
  int32_t tme_order;          /* 1..3, 'chapman-tme-<order>' */
  int32_t integration_steps;  /* sub-steps between two measurements (>= 1) */
  int32_t n_grid;             /* n >= 3 */
  int32_t drift_id;           /* MFS_DRIFT_* */
  int32_t meas_id;            /* MFS_MEAS_* */
  int32_t ys_dtype;           /* MFS_YS_* */
  int64_t B, T;
  double dt;                  /* time between two measurements */
  double dispersion;
  const double* trans_params; /* [MFS_MAX_PARAMS] */
  const double* meas_params;  /* [B|1][MFS_MAX_PARAMS] */
  int64_t meas_param_stride;
  const double* xs;           /* [n] grid (evenly spaced for 'kolmogorov') */
  const double* init_ps;      /* [B|1][n] initial density values */
  int64_t init_ps_stride;
  const void* ys;             /* y[b][t] at ys + b*ys_stride_b + t*ys_stride_t */
  int64_t ys_stride_b, ys_stride_t;
  int32_t out_mode;           /* MFS_OUT_FULL: pdfs_out[B][T][n]; MFS_OUT_LAST: pdfs_out[B][n] */
  int32_t flags;              /* MFS_BF_FLAG_* (0: the reference's literal recursion) */
  double* pdfs_out;
  double* nell_out;           /* optional [B]: -sum_t log trapz(p(y_t|x) p(x|y_{1:t-1})) (by-product; may be NULL) */
  void* workspace;
  int64_t workspace_bytes;
} mfs_brute_force_args;

/* Workspace size in bytes for mfs_brute_force (-1 for invalid arguments); _ex: for a call with the given flags. */
int64_t mfs_brute_force_workspace_bytes(int32_t n_grid, int64_t B, int32_t pred_method);
int64_t mfs_brute_force_workspace_bytes_ex(int32_t n_grid, int64_t B, int32_t pred_method, int32_t flags);

/* Enqueue the grid filter on `stream`.  Replaces jax.jit(brute_force_filter)(grid, ys) per trajectory
 * (dardel/benes_bernoulli/brute_force.py:51-55, 74). */
int mfs_brute_force(const mfs_brute_force_args* a, void* stream);

/* FP64 tensor-pipe (DMMA m8n8k4) throughput micro-benchmark: the roofline denominator of the grid-filter GEMM. */
int mfs_dmma_peak(int device, int32_t iters, double* flops, double* ms);

/* ABI version of the loaded library. */
int mfs_abi_version(void);

/* Last error text of the calling thread ("" if none). */
const char* mfs_last_error(void);

/* Look up a functor id by name, e.g. kind="drift", name="benes" -> MFS_DRIFT_BENES.  Kinds: "mode", "trans",
 * "drift", "meas".  Returns 0 and writes *id, or -1 for an unknown name. */
int mfs_functor_lookup(const char* kind, const char* name, int32_t* id);

/* Enqueue B filters x T steps on `stream`; every pointer in `a` is a DEVICE pointer.  Asynchronous w.r.t. the host.
 * Replaces: jax.jit(moment_filter_{rms,cms,scms})(ys) per trajectory (dardel/benes_bernoulli/mf.py:51-67). */
int mfs_filter_1d(const mfs_filter1d_args* a, void* stream);

/* nell AND its gradient w.r.t. functor parameters in one pass (forward-mode derivative carried through the scan).
 * Replaces jax.value_and_grad of the objective the reference hands to L-BFGS-B
 * (dardel/parameter_estimation/mf.py:37-73; "differentiable in the parameter", README.md:45).
 * `a` as for mfs_filter_1d (RAW or CENTRAL mode; ms_out / mean_out are not written: out_mode must be MFS_OUT_NONE;
 * `stable` must be 0).  tangent_ids[k] in [0, MFS_MAX_PARAMS) selects trans_params[id], in [MFS_MAX_PARAMS,
 * 2 MFS_MAX_PARAMS) selects meas_params[id - MFS_MAX_PARAMS].  Writes nell_out[b], status_out[b] and
 * grad_out[b][MFS_GRAD_MAX_TANGENTS] (row stride MFS_GRAD_MAX_TANGENTS; slots >= n_tangents are 0); a failed filter
 * gives NaN value and gradient.  Device pointers (tangent_ids is a HOST array). */
#define MFS_GRAD_MAX_TANGENTS 2
int mfs_filter_1d_grad(const mfs_filter1d_args* a, int32_t n_tangents, const int32_t* tangent_ids, double* grad_out,
                       void* stream);

/* Same contract with HOST pointers everywhere (ys, ms0, params, outputs).  The batch is cut into chunks that are
 * pipelined H2D -> kernel -> D2H on `device` (double-buffered, two streams); returns after the last D2H completed.
 * `chunk_filters` = 0 picks a default. */
int mfs_filter_1d_host(const mfs_filter1d_args* a, int device, int64_t chunk_filters);

/* mfs_filter_1d_host keeps its device staging buffers cached between calls (same shapes -> no cudaMalloc/cudaFree in
 * steady state).  This returns every idle cached buffer to the driver. */
int mfs_release_cached_memory(void);

/* Batched moment quadrature (mfs/one_dim/quadtures.py:83-133): ms[B][2N] -> weights[B][N], nodes[B][N] (ascending
 * nodes when sort_nodes != 0).  mean/scale may be NULL (0 / 1).  Device pointers. */
int mfs_moment_quadrature_1d(int32_t N, int64_t B, const double* ms, const double* mean, const double* scale,
                             int32_t sort_nodes, int32_t ldl, double* weights, double* nodes, void* stream);

/* Batched d-dimensional moment quadrature (mfs/multi_dims/quadratures.py:120-178), d = 2, N = 2..7:
 * ms[B][z] (graded-lex order, z = N(2N+1)) -> weights[B][S^2], nodes[B][S^2][2], S = N(N+1)/2, in the reference's
 * Cartesian order (node i*S + j = (lambda1_i, lambda2_j)).  mean / scale: [B][2] or NULL; inds: the (3, S, S) int32
 * table of gram_and_hankel_indices_graded_lexico(N, 2); ldl != 0: `ldl=True`.  Device pointers.  Failure -> NaN. */
int mfs_moment_quadrature_nd(int32_t N, int32_t d, int64_t B, const double* ms, const double* mean, const double* scale,
                             const int32_t* inds, int32_t ldl, double* weights, double* nodes, void* stream);

/* Characteristic function by moments (mfs/one_dim/moments.py:309-337), the post-processing step after the filter
 * (dardel/benes_bernoulli/post_processing_mf.py:37-60): out[b][j] = sum_n w_n exp(i zs[j] x_n) with (w, x) the
 * quadrature of ms[b][0..2N) (mean/scale may be NULL = 0 / 1).  out is complex128 [B][m] as (re, im) pairs.
 * Device pointers.  A failed quadrature gives NaN, like the reference. */
int mfs_characteristic_fn_1d(int32_t N, int64_t B, int64_t m, const double* ms, const double* mean, const double* scale,
                             const double* zs, double* out, void* stream);

/* ---- data simulators: the step before the filter ---------------------------------------------------------------------
 * One trajectory per thread; random numbers from a counter-based Philox4x32-10 stream keyed by `seed` and indexed by
 * (traj_offset + b, time step, draw), so a batch sharded over ranks (each passing its own traj_offset) equals the
 * single-GPU batch.  The reference draws from jax.random keys (rng_keys.npy), which cannot be reproduced without JAX:
 * the LAW of (x0, xs, ys) is the reference's, the stream is this library's. */
enum {
  MFS_SIM_TME = 0,         /* simulate_sde with tme.mean_and_cov(order) Gaussian sub-steps: mfs/utils.py:190-249,
                              mfs/one_dim/ss_models.py:49-54, 86-91 (order 1 = Euler--Maruyama)                     */
  MFS_SIM_BENES_EXACT = 1  /* exact Benes transition: N(x +- dt, dt) w.p. (1 +- tanh x)/2 (bench data; drift ignored) */
};
#define MFS_SIM_MAX_COMPONENTS 8

typedef struct mfs_simulate1d_args {
  int32_t abi_version;
  int32_t scheme;             /* MFS_SIM_* */
  int32_t tme_order;          /* 1..3 */
  int32_t integration_steps;  /* sub-steps per dt (reference: 100) */
  int32_t drift_id;           /* MFS_DRIFT_* */
  int32_t meas_id;            /* MFS_MEAS_*: Bernoulli(logistic) / Poisson(softplus) / h x + r N(0,1) */
  int32_t ys_dtype;           /* MFS_YS_* */
  int32_t n_components;       /* initial Gaussian mixture, GaussianSum1D (mfs/utils.py:31-58) */
  int64_t B, T;
  double dt, dispersion;
  const double* trans_params; int64_t trans_param_stride;   /* device [B|1][MFS_MAX_PARAMS] */
  const double* meas_params;  int64_t meas_param_stride;    /* device [B|1][MFS_MAX_PARAMS] */
  double init_means[MFS_SIM_MAX_COMPONENTS], init_variances[MFS_SIM_MAX_COMPONENTS], init_weights[MFS_SIM_MAX_COMPONENTS];
  uint64_t seed;
  uint64_t traj_offset;       /* global id of trajectory b = traj_offset + b */
  void* ys_out;  int64_t ys_stride_b, ys_stride_t;          /* device, y[b][t]; may be NULL */
  double* xs_out; int64_t xs_stride_b, xs_stride_t;         /* device, x[b][t] = state at t_{k+1}; may be NULL */
  double* x0_out;                                           /* device [B]; may be NULL */
} mfs_simulate1d_args;

/* Enqueue B trajectories x T steps.  Replaces init_cond.sampler + simulate_trajectory + jax.random.bernoulli/poisson
 * per Monte-Carlo run (dardel/benes_bernoulli/mf.py:73-80, dardel/parameter_estimation/mf.py:58-65). */
int mfs_simulate_1d(const mfs_simulate1d_args* a, void* stream);

typedef struct mfs_simulate_lv_args {
  int32_t abi_version;
  int32_t integration_steps;  /* Milstein sub-steps per dt (reference: 100) */
  int32_t n_components;
  int32_t obs_dim;            /* measured coordinate (prey--predator: 0) */
  int64_t B, T;
  double dt;
  const double* trans_params; int64_t trans_param_stride;   /* device [B|1][8]: alpha, beta, delta, gamma, sigma */
  const double* meas_params;  int64_t meas_param_stride;    /* device [B|1][4]: c0, c1 of the logistic-cubic emission */
  double init_means[MFS_SIM_MAX_COMPONENTS][2], init_covs[MFS_SIM_MAX_COMPONENTS][4], init_weights[MFS_SIM_MAX_COMPONENTS];
  uint64_t seed, traj_offset;
  uint8_t* ys_out;            /* device [B][T]; may be NULL */
  double* xs_out;             /* device [B][T][2]; may be NULL */
  double* x0_out;             /* device [B][2]; may be NULL */
} mfs_simulate_lv_args;

/* Lotka--Volterra Milstein simulator + Bernoulli measurements: prey_predator(...).simulate,
 * mfs/multi_dims/ss_models.py:76-93. */
int mfs_simulate_lv(const mfs_simulate_lv_args* a, void* stream);

/* Diagnostic: the kernels' own elementary functions (branch-free exp / log / tanh with constant-bank polynomial
 * coefficients, mfs_b200/csrc/models.cuh) evaluated on x[0..n); any output pointer may be NULL.  Device pointers.
 * Used by the parity tests to bound their error against libm. */
int mfs_math_selftest(int64_t n, const double* x, double* out_exp, double* out_log, double* out_tanh, void* stream);

/* Number of kernel launches issued by this library in the calling process since load (all threads). */
int64_t mfs_launch_count(void);

/* FP64 FMA throughput micro-benchmark used for the roofline denominator: runs `iters` dependent-chain-free DFMA
 * blocks on the whole device and returns achieved FLOP/s (2 flop per FMA) in *flops.  Device pointers none. */
int mfs_fp64_peak(int device, int32_t iters, double* flops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* MFS_B200_H_ */
