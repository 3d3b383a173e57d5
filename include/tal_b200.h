/*
 * tal_b200.h — C ABI of the B200-native discrete-domain GP hot path.
 *
 * The reference (jonhue/transductive-active-learning) has no FFI layer: its L0
 * is jax.numpy/XLA.  This header declares the entry points that take L0's
 * place.  Each export cites the reference call site (path:line under the
 * reference tree) whose arithmetic it replaces.  The Python binding a
 * maintainer would add is a ctypes stub; see INTEGRATION.md.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.
 *  - Every pointer is a DEVICE pointer owned by the caller unless the
 *    parameter name ends in `_host`.
 *  - Every call only enqueues work on `stream` (a cudaStream_t passed as
 *    void*) and returns without synchronising.  Return value: TAL_OK or a
 *    TAL_ERR_* code; `tal_last_error_string()` describes the last failure
 *    on the calling thread.
 *  - Matrices are row-major.  A covariance shard is `cov[q][n_loc][ld]`
 *    (GP-major, then local row, then column; `ld >= N` elements between
 *    rows, `cov_stride_q` elements between GPs).  Local row r is global row
 *    `row0 + r`.  On one GPU row0 = 0 and n_loc = N.
 *  - `dtype` selects the storage/arithmetic type of covariance-sized data
 *    (cov, kbuf, wbuf): TAL_F64 (the reference's only mode) or TAL_F32.
 *    ld must be N rounded up to a multiple of 32 elements; the pad columns
 *    hold zeros.  kbuf and wbuf are [q][ld] of the same dtype.
 *    The O(qN) state vectors (mean, var, l, u, rho, F) are always double.
 *  - Indices are int64 (`long long`), masks are one byte per point.
 *  - q <= TAL_MAX_Q.
 */
#ifndef TAL_B200_H
#define TAL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TAL_ABI_VERSION 3
#define TAL_MAX_Q 16

/* status codes */
#define TAL_OK 0
#define TAL_ERR_CUDA 1        /* CUDA runtime / launch error */
#define TAL_ERR_ARG 5         /* invalid argument */
/* result.status of tal_masked_argmax — lib/model/marginal.py:336-342 */
#define TAL_ARGMAX_OK 0
#define TAL_ARGMAX_EMPTY_SAFE_SET 2 /* "Operating safe set is empty. Cannot make observations." */
#define TAL_ARGMAX_ALL_NEG_INF 3    /* "Objective is negative infinity everywhere." */
#define TAL_ARGMAX_ALL_NAN 4        /* "Objective is NaN everywhere." */
#define TAL_ARGMAX_PEER_ERROR 7     /* row-sharded model: an NVLink peer exchange timed out (flags bit 3) */

/* dtype */
#define TAL_F64 0
#define TAL_F32 1
/* arithmetic of the covariance update c <- c - k*w, OR-ed into the `dtype` argument of
 * tal_step_vectors, tal_rank1_update[_lower], tal_rankb_update, tal_pending_correct_row and
 * the step context: without it the product is rounded before the subtraction (two roundings,
 * what NumPy and an unfused XLA dot + subtract compute: lib/gp/gaussian_distribution.py:36);
 * with it the update is one fused multiply-subtract (one rounding, what a fused XLA:CPU loop
 * compiled with FMA contraction computes).  The matrix, the pending rows and the variance
 * vector always use the same mode, so var == diag(cov) holds bit for bit in either. */
#define TAL_ARITH_FMA 16

/* stationary kernel kinds — lib/gp/kernels/stationary.py:23-55 */
#define TAL_KERNEL_GAUSSIAN 0
#define TAL_KERNEL_LAPLACE 1
#define TAL_KERNEL_MATERN32 2
#define TAL_KERNEL_MATERN52 3
#define TAL_KERNEL_LINEAR 4 /* lib/gp/kernels/non_stationary.py:11-13 */

/* row-reduce scoring rules — lib/algorithms/{vtl,ctl,mm_itl}.py, tru_var/__init__.py */
#define TAL_RULE_VTL 0
#define TAL_RULE_CTL 1
#define TAL_RULE_MMITL 2
#define TAL_RULE_TRUVAR 3

/* result record written by tal_masked_argmax (32 bytes, device memory) */
typedef struct tal_argmax_result {
  long long idx;    /* lowest index among the exact ties at the maximum; -1 if none or status != OK */
  double val;       /* nanmax of the safety-masked objective */
  long long n_ties; /* number of points attaining val exactly */
  int status;       /* TAL_ARGMAX_* (derived from flags) */
  int flags;        /* bit0: safe set non-empty; bit1: some F != -inf; bit2: some safe F not NaN;
                       bit3: a peer exchange of this rank timed out (TAL_ARGMAX_PEER_ERROR) */
} tal_argmax_result;

/* ---- library --------------------------------------------------------- */
int tal_abi_version(void);
const char* tal_last_error_string(void);
/* number of kernels launched by this library in this process (bench.py's
 * `gpu_launches`); reset with tal_reset_launch_count. */
long long tal_launch_count(void);
void tal_reset_launch_count(void);
/* 1 if the library was built for sm_100a and the current device is CC 10.x */
int tal_device_ok(void);

/* ---- Gram construction ------------------------------------------------
 * Replaces Kernel.cross_covariance / covariance (lib/gp/kernels/base.py:17-25)
 * with the scalar kernels of stationary.py:23-55 / non_stationary.py:11-13.
 * out[j][i] = k(X[i], Y[j]) for j in [j0, j0+n_rows): the reference returns
 * [len(Y), len(X)] (vmap over X inside, over Y outside).  X is [n,d], Y is
 * [m,d], both double, row-major.  `out` has `n_rows` rows of `ld` elements.
 * If `mean_out`/`var_out` are non-null they are ignored (reserved).        */
int tal_gram_stationary(int kind, const double* X, long long n, const double* Y, long long m,
                        int d, double variance, double lengthscale, void* out, long long ld,
                        long long j0, long long n_rows, int dtype, void* stream);
/* The covariance K(X, X) of a model in lower-block storage (see tal_rank1_update_lower): rows
 * [j0, j0 + n_rows) of the Gram matrix over X itself, of which only the column tiles that reach into the
 * kept part of a row (the columns below the end of its diagonal block) are computed and written; the rest of
 * `out` is left untouched (nothing reads it before tal_sym_mirror rebuilds it).  j0 % 32 == 0. */
int tal_gram_stationary_lower(int kind, const double* X, long long n, int d, double variance, double lengthscale,
                              void* out, long long ld, long long j0, long long n_rows, int dtype, void* stream);

/* ---- index lookup -------------------------------------------------------
 * Replaces get_indices (lib/utils.py:53-56): out[k] = argmin_j ||X[j]-A[k]||_2
 * with the first minimum winning.  X [n,d], A [m,d] double.                */
int tal_nearest_index(const double* X, long long n, const double* A, long long m, int d,
                      long long* out, void* stream);

/* ---- posterior update, one observation (the BO loop's case) -----------
 * Replaces MarginalModel.step (lib/model/marginal.py:350-368) →
 * GaussianDistribution.posterior (lib/gp/gaussian_distribution.py:100-121) →
 * _compute_posterior_covariance (:16-37) with m = 1, batched over q GPs.
 *
 * Phase 1, tal_extract_row: kbuf[i][:] = cov_i[idx, :] if this shard owns
 *   global row idx, else zeros (so that a sum over shards is a broadcast);
 *   also stashes mean_i[idx] into scal[i] (before it is overwritten).
 *   `idx_dev` points to the observed index (device); if null, `idx_host` is used.
 * Phase 2, tal_step_vectors (needs the complete kbuf): for each GP
 *   s = k[idx] + rho[idx]^2, L = sqrt(s), w = (k / L) / L  (utils.py:22-23, 1x1),
 *   mean' = mean + w (y - mean[idx]),  var' = var - k*w  (same arithmetic as
 *   the matrix diagonal), l' = max(l, mean' - beta sqrt(var')),
 *   u' = min(u, mean' + beta sqrt(var')) (marginal.py:160-164,366-367; NaN
 *   propagating like jnp.maximum).  Writes wbuf[q][N].
 *   y: if y_gather != 0, y_i = y_dev[i*y_stride + idx] else y_dev[i*y_stride].
 * Phase 3, tal_rank1_update: cov_i[a][b] -= k_i[row0+a] * w_i[b] in place
 *   over the shard (gaussian_distribution.py:36).  Pure HBM streaming:
 *   2*q*n_loc*N*sizeof(T) algorithmic bytes.
 *   variant: 0 = default, 1 = vectorised LDG/STG, 2 = bulk-async (TMA) pipeline. */
int tal_extract_row(const void* cov, long long cov_stride_q, long long ld, long long row0,
                    long long n_loc, long long N, int q, const long long* idx_dev,
                    long long idx_host, const double* mean, void* kbuf, double* scal, int dtype,
                    void* stream);
int tal_step_vectors(const void* kbuf, void* wbuf, long long N, int q, const long long* idx_dev,
                     long long idx_host, const double* y_dev, long long y_stride, int y_gather,
                     const double* rho, const double* beta_host, const double* scal,
                     double* mean, double* var, double* l, double* u, int dtype, void* stream);
int tal_rank1_update(void* cov, long long cov_stride_q, long long ld, long long row0,
                     long long n_loc, long long N, int q, const void* kbuf, const void* wbuf,
                     int dtype, int variant, void* stream);

/* ---- lower-block ("symmetric") storage -------------------------------------
 * The posterior covariance is symmetric (up to the rounding of k_a w_b versus
 * k_b w_a), so an opt-in storage mode maintains only the blocks on or below the
 * diagonal: of global row g the columns [0, cend(g)), cend(g) = min(ld,
 * (g / B + 1) * B), B = tal_sym_block(dtype) = 32 elements (the rows of one update CTA).  Entry
 * (g, c) with c >= cend(g) is read as (c, g).  The allocation and addressing are
 * unchanged ([q][n_loc][ld]); the blocks above the diagonal are simply not
 * touched, which halves the HBM traffic of the rank-1 update.
 * tal_rank1_update_lower: as tal_rank1_update on the kept blocks only
 *   (algorithmic bytes 2*q*s*sum_rows cend(row)); row0 must be a multiple of 32.
 * tal_extract_row_lower / tal_gather_rows_lower: as tal_extract_row /
 *   tal_gather_rows, assembling the row from the part this shard keeps
 *   (zeros elsewhere, so that a sum over shards is the row).
 * tal_rankm_update_lower: as tal_rankm_update on the kept blocks.
 * tal_sym_mirror: one GPU holding all rows: fills the blocks above the diagonal
 *   from their mirror images (before handing the matrix to a full-storage reader). */
long long tal_sym_block(int dtype);
int tal_rank1_update_lower(void* cov, long long cov_stride_q, long long ld, long long row0,
                           long long n_loc, long long N, int q, const void* kbuf, const void* wbuf,
                           int dtype, void* stream);
int tal_extract_row_lower(const void* cov, long long cov_stride_q, long long ld, long long row0,
                          long long n_loc, long long N, int q, const long long* idx_dev,
                          long long idx_host, const double* mean, void* kbuf, double* scal, int dtype,
                          void* stream);
int tal_gather_rows_lower(const void* cov, long long ld, long long row0, long long n_loc, long long N,
                          const long long* obs_idx, long long m, double* B, long long ldb, int dtype,
                          void* stream);
int tal_rankm_update_lower(void* cov, long long ld, long long row0, long long n_loc, long long N,
                           const double* B, const double* W, long long ldb, int m, int dtype,
                           void* stream);
int tal_sym_mirror(void* cov, long long cov_stride_q, long long ld, long long N, int q, int dtype,
                   void* stream);

/* ---- temporally blocked posterior update ------------------------------------
 * The (k_i, w_i) of the last p <= TAL_MAX_PENDING observations may be kept pending
 * (K, W: [p][q][ld] of the cov dtype, `pend_stride` elements between slots) instead of streaming the matrix after every one of them:
 * tal_pending_correct_row applies them to the row that the next observation reads
 *   (kbuf [q][ld] as written by tal_extract_row[_lower] / the peer exchange):
 *   k[c] <- ((k[c] - K_0[a] W_0[b]) - K_1[a] W_1[b]) - ... with (a, b) = (idx, c), or
 *   (c, idx) for the mirrored entries of lower-block storage;
 * tal_rankb_update applies all p of them to every maintained entry in ONE pass:
 *   cov[a][b] <- ((cov[a][b] - K_0[a] W_0[b]) - K_1[a] W_1[b]) - ...
 * Each entry sees the same sequence of rounded operations as p calls of
 * tal_rank1_update[_lower]: the result is bit-identical, the HBM traffic per
 * observation is divided by p (the pass stays HBM-bound up to p = 16).            */
#define TAL_MAX_PENDING 32
int tal_rankb_update(void* cov, long long cov_stride_q, long long ld, long long row0, long long n_loc,
                     long long N, int q, const void* K, const void* W, long long pend_stride, int p,
                     int lower, int dtype, void* stream);
int tal_pending_correct_row(void* kbuf, long long ld, long long N, int q, const void* K, const void* W,
                            long long pend_stride, int p, const long long* idx_dev, long long idx_host,
                            int lower, int dtype, void* stream);

/* ---- posterior update, m observations ---------------------------------
 * Replaces GaussianDistribution.posterior for general m (prior_distr,
 * lib/gp/prior.py:52-66; SafeOpt/GoOSE replay).  One GP per call.
 * tal_gather_rows: B[j][:] = cov[obs_idx[j]-row0][:] (rows this shard owns;
 *   others zero), B is [m][ldb] of the cov dtype converted to double.
 * tal_small_posterior_weights: S = B[:, obs_idx] + diag(noise_std^2) (or
 *   jitter^2 when noise_std is null: utils.py:44-50), L = chol(S),
 *   W = S^-1 B (utils.py:22-23), mean' = mean + W^T (y - mean[obs_idx]),
 *   var' = var - sum_j B[j]*W[j].  m <= 64.  NaN on Cholesky failure.
 * tal_rankm_update: cov[a][b] -= sum_j B[j][row0+a] * W[j][b] in place.    */
int tal_gather_rows(const void* cov, long long ld, long long row0, long long n_loc, long long N,
                    const long long* obs_idx, long long m, double* B, long long ldb, int dtype,
                    void* stream);
int tal_small_posterior_weights(const double* B, long long ldb, long long N,
                                const long long* obs_idx, int m, const double* noise_std,
                                double jitter, const double* y, double* W, double* mean,
                                double* var, double* scratch, void* stream);
long long tal_small_scratch_bytes(void); /* size of `scratch` (device) */
int tal_rankm_update(void* cov, long long ld, long long row0, long long n_loc, long long N,
                     const double* B, const double* W, long long ldb, int m, int dtype,
                     void* stream);

/* ---- confidence-bound masks -------------------------------------------
 * Replaces optimistic/pessimistic_safe_set, potential_expanders, max_l,
 * potential_maximizers (lib/model/marginal.py:171-224).  l,u are [q][N].
 * op_safe_in: the operating safe set when a subclass overrides it
 * (unsafe.py:9-13, safe_opt/model.py:34-37); null = pessimistic.
 * Any output pointer may be null.  max_l_out: one double.                 */
int tal_bounds_masks(const double* l, const double* u, int q, int first_constraint, long long N,
                     const unsigned char* op_safe_in, unsigned char* pess, unsigned char* opt,
                     unsigned char* pm, unsigned char* pe, double* max_l_out, void* stream);
/* ordered stream compaction: idx_out = indices where mask != 0 (ascending),
 * *count_out = how many.  Replaces domain[mask] / set_to_idx (utils.py:74-75). */
int tal_mask_to_indices(const unsigned char* mask, long long N, long long* idx_out,
                        long long* count_out, void* stream);
/* scalar-wise subset test of lib/utils.py:65-71 (jnp.isin on flattened
 * coordinates) for X = domain[x_mask], A = domain[roi_idx]: sid[N][d] maps
 * every coordinate to the id of its distinct scalar value; present is an
 * [n_sid] byte scratch.  *out = 1 iff every scalar of X occurs in A;
 * empty-set rules of :66-69 included.                                      */
int tal_scalar_subset(const int* sid, int d, long long N, long long n_sid,
                      const unsigned char* x_mask, const long long* roi_idx,
                      const long long* roi_count_dev, long long roi_count_host,
                      unsigned char* present, int* out, void* stream);

/* ---- scoring ------------------------------------------------------------
 * tal_score_itl_undirected: F[x] = sum_i log(1 + var_i[x]/rho_i[x]^2)
 *   (lib/algorithms/itl.py:28-30).
 *
 * Row-reduce family (one GP per call).  The reference gathers cov[a, x] for
 * a in the ROI through nested vmaps (vtl.py:21-41, marginal.py:129-158,
 * tru_var/__init__.py:31-48); here the m ROI rows are streamed once and
 * reduced per column:  c[x] = sum_{a in roi} elem(cov[a][x]) with
 *   VTL, CTL:  elem = w_a v^2              (w_a = 1, resp. 1/var[a])
 *   MM-ITL:    elem = -0.5 log(1 - w_a v^2 dinv[x]),  w_a = 1/var[a]
 *   TruVar:    elem = max(p0 (w_a - v^2 dinv[x]), p1), w_a = var[a], p0 = beta, p1 = eta^2
 * with dinv[x] = 1/(var[x] + rho[x]^2).
 * tal_colsum_rows reads ROI rows a (coalesced; rows outside this shard are
 *   skipped); `part` is an [n_split][N] scratch, c_out is [N].
 * tal_colsum_cols computes the same sums for the shard's own candidates
 *   x = row0 + r through the symmetric entries cov[r][a] (no communication;
 *   used when the covariance is row-sharded); c_out and dinv are [n_loc].
 * tal_roi_trace: tr = sum_{a in roi} var[a].  tal_roi_weights: w_a.
 * tal_pred_var_inv: dinv.
 * tal_score_rows_finish: F[x] (+)= epilogue over n candidates
 *   VTL: -(tr - c/(var+rho^2)); CTL: c/(var+rho^2); MM-ITL: c; TruVar: -c.   */
int tal_score_itl_undirected(const double* var, const double* rho, int q, long long N, double* F,
                             void* stream);
int tal_colsum_rows(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                    long long N, const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* part, int n_split,
                    double* c_out, int dtype, void* stream);
int tal_colsum_cols(int rule, const void* cov, long long ld, long long n_loc,
                    const long long* roi_idx, long long m, const double* row_weight,
                    const double* dinv, double p0, double p1, double* c_out, int dtype,
                    void* stream);
/* lower-block storage (tal_rank1_update_lower): the same column sums over ALL candidates x from what
 * this shard keeps — the kept part of its ROI rows plus, for its own rows x, the ROI entries mirrored
 * into row x.  Every pair (a, x) is counted exactly once across the shards; a sum over the shards
 * completes c_out [N] (one GPU: it is complete).  dinv is the full [N] vector. */
int tal_colsum_lower(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                     long long N, const long long* roi_idx, long long m, const double* row_weight,
                     const double* dinv, double p0, double p1, double* part, int n_split,
                     double* c_out, int dtype, void* stream);
/* the same for an ROI that covers a large part of the domain (the potential-maximiser ROIs of the Safe-BO
 * configurations, lib/algorithms/__init__.py:140-147: m ~ N): the mirrored part of every row is streamed and
 * masked by roi_mask (one byte per point) instead of gathered entry by entry; w_pt = row_weight by point
 * ([N], NULL with row_weight NULL). */
int tal_colsum_lower_dense(int rule, const void* cov, long long ld, long long row0, long long n_loc,
                           long long N, const long long* roi_idx, long long m, const unsigned char* roi_mask,
                           const double* row_weight, const double* w_pt, const double* dinv, double p0,
                           double p1, double* part, int n_split, double* c_out, int dtype, void* stream);
/* ... and fused with the posterior update that precedes it (a VTL-type step flushes its ONE pending rank-1 update
 * before it scores: lib/model/marginal.py:350-368 followed by lib/algorithms/vtl.py:21-41): one pass applies
 * cov[a][b] <- cov[a][b] - k[a] w[b] (the rounded operation of tal_rank1_update_lower, `dtype` may carry
 * TAL_ARITH_FMA) to every kept entry of ONE GP and accumulates both halves of tal_colsum_lower_dense from the updated
 * values: 2 transfers of the matrix per step instead of 4.  kvec, wvec: [ld] of the cov dtype (the pending slot of this
 * GP); rowpart: tal_update_colsum_scratch_bytes(ld, n_loc, dtype) bytes; part: [n_split][N]. */
long long tal_update_colsum_scratch_bytes(long long ld, long long n_loc, int dtype);
int tal_update_colsum_lower_dense(int rule, void* cov, long long ld, long long row0, long long n_loc, long long N,
                                  const void* kvec, const void* wvec, const unsigned char* roi_mask,
                                  const double* w_pt, const double* dinv, double p0, double p1, double* part,
                                  int n_split, double* rowpart, double* c_out, int dtype, void* stream);
int tal_roi_trace(const double* var, const long long* roi_idx, long long m, double* out,
                  void* stream);
int tal_roi_weights(const double* var, const long long* roi_idx, long long m, int reciprocal,
                    double* out, void* stream);
int tal_pred_var_inv(const double* var, const double* rho, long long N, double* out,
                     void* stream);
int tal_score_rows_finish(int rule, const double* c, const double* var, const double* rho,
                          const double* tr, long long n, int accumulate, double* F,
                          void* stream);

/* ---- directed ITL conditioning (fp64, DMMA tensor path) ----------------
 * Replaces compute_posterior_covariance(cov_i, roi_idx, None) + diag
 * (lib/algorithms/itl.py:41-48; gaussian_distribution.py:16-37;
 * utils.py:10-24,44-50) by: S = cov[A,A] + jitter^2 I, L = chol(S) (NaN on
 * failure, like jnp.linalg.cholesky), V = L^-1 cov[A,:] (forward solve only),
 * post_var[x] = var[x] - sum_j V[j][x]^2.
 * m_pad = m rounded up to 128; column counts are multiples of 64.
 * tal_gather_roi: B[j][x] = cov[roi[j]][x] (double, [m_pad][ldb], pad rows and
 *   columns zero) and S = blockdiag(cov[A,A] + jitter^2 I, I) ([m_pad][lds]).
 * tal_potrf_lower: in-place blocked Cholesky of S (lower).  work: scratch of
 *   tal_potrf_work_bytes(m_pad); it receives the inverses of the diagonal
 *   blocks, which tal_trsm_lower consumes.
 * tal_trsm_lower: B <- L^-1 B, B is [m_pad][n_cols] with row stride ldb.
 * tal_colsumsq_dense: out[x] = sum_j V[j][x]^2 (part: [n_split][n_cols]).
 * tal_score_itl_directed: F[x] (+)= 0.5*max(log((var+rho^2)/(var-cs+rho^2)),0)
 *   (itl.py:50-58) with cs the column sums of squares above.               */
int tal_gather_roi(const void* cov, long long ld, long long N, const long long* roi_idx,
                   long long m, long long m_pad, double jitter, double* S, long long lds,
                   double* B, long long ldb, int dtype, void* stream);
/* row-sharded gather: B[j][x] = cov[roi[j]][row0 + x] read through the symmetric
 * entry cov_loc[x][roi[j]] ([m_pad][ldb], ncols >= n_loc columns written, pads
 * zero); S receives the rows whose roi[j] this shard owns (+ the identity pad
 * on the shard with row0 = 0) and zeros elsewhere: a sum over shards gives S. */
/* lower-block storage, one GPU holding all rows: the same gather through the mirrored entries. */
int tal_gather_roi_lower(const void* cov, long long ld, long long N, const long long* roi_idx,
                         long long m, long long m_pad, double jitter, double* S, long long lds,
                         double* B, long long ldb, int dtype, void* stream);
int tal_gather_roi_sharded(const void* cov, long long ld, long long row0, long long n_loc,
                           const long long* roi_idx, long long m, long long m_pad, double jitter,
                           double* S, long long lds, double* B, long long ldb, long long ncols,
                           int dtype, void* stream);
long long tal_potrf_work_bytes(long long m_pad);
int tal_potrf_lower(double* S, long long lds, long long m_pad, void* work, void* stream);
int tal_trsm_lower(const double* L, long long lds, long long m_pad, const void* work, double* B,
                   long long ldb, long long n_cols, void* stream);
int tal_colsumsq_dense(const double* V, long long ldv, long long m, long long n_cols,
                       double* part, int n_split, double* out, void* stream);
int tal_score_itl_directed(const double* var, const double* cs, const double* rho,
                           long long n, int accumulate, double* F, void* stream);

/* ---- dense blocks of a covariance: sampling, entropy, information gain -----
 * SURVEY.md 8f-4.  All of them are: gather a block, factor it with
 * tal_potrf_lower (blocked DMMA Cholesky; the reference calls SVD / slogdet /
 * cholesky of XLA:CPU), then one small pass over the factor.
 * tal_gather_sym_block: S[j][k] = cov[idx[j]][idx[k]] + (j == k) (add + add_std[j]^2),
 *   identity pad up to m_pad (idx == NULL: the leading m x m block; lower != 0:
 *   lower-block storage, mirrored entries) — GaussianDistribution.sub_distr
 *   (lib/gp/gaussian_distribution.py:182-187), MarginalModel.masked_distr, the
 *   `covariance[ix_(obs, obs)] + noise_covariance_matrix` of .information_gain (:163-173)
 *   and the `+ jitter * eye` of .entropy (:145-148).
 * tal_chol_logdet: out[0] = 2 sum_{i<m} log L[i][i] = slogdet(S)[1] for the factored S
 *   (:141-149, :174-177).
 * tal_tri_sample: out[s][r] = mean[r] + sum_{c<=r} L[r][c] Z[s][c] for k draws Z[k][ldz]
 *   (GaussianDistribution.sample :90-98, MarginalModel.sample lib/model/marginal.py:226-237:
 *   the reference factors the covariance again for every draw).
 * tal_gather_cols_t: Vt[i][j] = V[j][idx[i]] (i < nb; zero rows up to nb_pad) — the columns B of
 *   V = L^-1 cov[A, :] transposed for the SYRK below.
 * tal_syrk_lower_sub: C <- C - A A^T on the 128 x 128 tiles on or below the diagonal
 *   (A is [n][K], n % 128 == 0, K % 16 == 0): the posterior block
 *   Sigma_BB - Sigma_BA (Sigma_AA + jitter^2 I)^-1 Sigma_AB of .information_gain (:164-169)
 *   as Sigma_BB - V_B^T V_B on the fp64 tensor path.                          */
int tal_gather_sym_block(const void* cov, long long ld, long long N, const long long* idx, long long m,
                         long long m_pad, const double* add_std, double add, double* S, long long lds,
                         int lower, int dtype, void* stream);
int tal_chol_logdet(const double* L, long long lds, long long m, double* out, void* stream);
int tal_tri_sample(const double* L, long long lds, long long n, const double* Z, long long ldz, int k,
                   const double* mean, double* out, long long ldo, void* stream);
int tal_gather_cols_t(const double* V, long long ldv, long long m_pad, const long long* idx, long long nb,
                      long long nb_pad, double* Vt, long long ldvt, void* stream);
int tal_syrk_lower_sub(double* C, long long ldc, long long n, const double* A, long long lda, long long K,
                       void* stream);

/* ---- incremental conditioning for a fixed ROI ---------------------------
 * Beyond the reference (which re-runs the m x m Cholesky and the m x N solves
 * of lib/algorithms/itl.py:41-48 at every step): when the ROI is unchanged,
 * V = L^-1 cov[A,:] and cs follow from the previous step's by an exact rank-1
 * downdate (csrc/condinc.cu).  Per observation at index idx:
 * tal_cond_extract_col: pcol[j] = V[j][idx - col0] if this shard holds that
 *   column (else 0, so that a sum over shards broadcasts it).
 * tal_cond_update: V <- T^-1 (V - pcol w^T), cs[x] = sum_j V[j][x]^2, with
 *   T = chol(I - p p^T), p = pcol / sqrt(k[idx] + rho[idx]^2); kbuf_i, wbuf_i,
 *   rho_i are GP i's rows of kbuf / wbuf / rho as left by tal_step_vectors.
 *   V is [m_pad][ncols] (row stride ldv) and holds global columns
 *   [col0, col0 + n_valid); coef_scratch: tal_cond_coef_bytes(m_pad).       */
int tal_cond_extract_col(const double* V, long long ldv, long long m_pad, long long col0,
                         long long ncols, const long long* idx_dev, long long idx_host,
                         double* pcol, void* stream);
int tal_cond_update(double* V, long long ldv, long long m_pad, long long ncols, long long n_valid,
                    long long col0, const double* pcol, const void* kbuf_i, const void* wbuf_i,
                    const double* rho_i, const long long* idx_dev, long long idx_host,
                    void* coef_scratch, double* cs, int chunks, double* chunk_scratch, int dtype,
                    void* stream);
long long tal_cond_coef_bytes(long long m_pad);
/* chunks: 1 = one pass, every thread walks all m rows of its two columns
 * (2 m N 8 bytes of traffic; right when there are >= 64k columns); > 1 = the rows
 * are split into `chunks` chunks handled by different CTAs through the prefix-sum
 * form of the recurrence (one extra read pass over V, but m/chunks-long serial
 * chains: right for a shard with few columns); 0 = choose.  chunk_scratch:
 * tal_cond_chunk_scratch_bytes(ncols) bytes (may be null with chunks <= 1).  */
long long tal_cond_chunk_scratch_bytes(long long ncols);

/* Append form of the same incremental conditioning (csrc/condapp.cu): conditioning on
 * the target set and on the observations commutes, so the conditioned GP is carried along
 * by a growing factor  Wb = [W_0 (m_pad rows = L^-1 Sigma_A. at set-up); g_0; h_0; g_1; h_1; ...]
 * ([cap_rows][ldw] fp64, `rows` rows in use): per observation ONE read-only pass
 * kA = k + sum_j sign_j col_j Wb[j][:] (col = Wb[:, x*], signs -,...,-,(-,+)*), two appended
 * rows g = kA / sqrt(kA[x*] + rho^2), h = k / sqrt(k[x*] + rho^2), and
 * cs += g^2 - k w.  Traffic (m_pad + 2 t) ncols 8 bytes, nothing rewritten, no recurrence.
 * scratch: tal_cond_append_scratch_bytes(ncols, n_split); the caller adds 2 to `rows`.  */
int tal_cond_append_update(double* Wb, long long ldw, long long cap_rows, long long m_pad, long long rows,
                           long long ncols, long long n_valid, long long cand0, const double* col,
                           const void* kbuf_i, const void* wbuf_i, const double* rho_i,
                           const long long* idx_dev, long long idx_host, double* cs, double* scratch,
                           int n_split, int dtype, void* stream);
long long tal_cond_append_scratch_bytes(long long ncols, int n_split);

/* ---- masked arg-max -----------------------------------------------------
 * Replaces DiscreteGreedyAlgorithm._step's sample-region masking
 * (lib/algorithms/__init__.py:80) + MarginalModel.objective_argmax
 * (lib/model/marginal.py:330-348) up to the PRNG draw.  F_io [N] is
 * overwritten with the sample-region-masked objective (the "F" of the result
 * dict).  safe: operating safe set; if null it is computed on the fly as the
 * pessimistic safe set all_{i>=fc} l_i[x] >= 0 (fused safe-set mask).
 * sample: sample-region mask or null (all true).  idx_offset is added to the
 * reported index (shard's first candidate).  scratch: tal_argmax_scratch_bytes().
 * Tie policy: lowest index among exact ties (the reference draws uniformly
 * with jax.random.choice); n_ties reports the size of the tie set.         */
long long tal_argmax_scratch_bytes(void);
int tal_masked_argmax(double* F_io, const unsigned char* safe, const double* l, int q,
                      int first_constraint, long long l_stride, const unsigned char* sample,
                      long long N, long long idx_offset, void* scratch, tal_argmax_result* out,
                      void* stream);
/* lexicographic reduce of G per-shard results (max val, then min idx; statuses
 * combined) — the all-reduce of the per-shard arg-max after an all-gather. */
int tal_argmax_combine(const tal_argmax_result* parts, int G, tal_argmax_result* out,
                       void* stream);

/* ---- exchange steps over NVLink peer memory (row-sharded model) ----
 * One process per GPU.  Every rank owns a MAILBOX: one device allocation
 * (tal_peer_alloc) laid out as [header: tal_peer_header_bytes()][kbuf: q*ld
 * elements at kbuf_off][pcol: q*m_pad doubles at pcol_off] (offsets chosen by
 * the caller, multiples of 256), exported with tal_peer_export (a CUDA IPC
 * handle, exchanged by the host over any channel) and mapped by the peers with
 * tal_peer_import.  `mailboxes` is a DEVICE array of the G mailbox pointers as
 * seen from this rank (entry `rank` = its own).  The kernels replace the two
 * NCCL calls of a step (all-gather of the arg-max records; broadcast of the
 * observed row, lib/gp/gaussian_distribution.py:33 `Kxt`) by P2P stores plus
 * release/acquire sequence numbers kept in the mailboxes; every rank must
 * issue the same call sequence.  Counters advance on the device, so the launch
 * sequence of a step is capturable in a CUDA graph.  Spins are bounded: on
 * expiry the mailbox error word becomes TAL_PEER_ERR_* (tal_peer_error).      */
#define TAL_PEER_MAX 16
#define TAL_PEER_HANDLE_BYTES 64
#define TAL_PEER_ERR_RECORD_TIMEOUT 1
#define TAL_PEER_ERR_READY_TIMEOUT 2
#define TAL_PEER_ERR_ROW_TIMEOUT 3
long long tal_peer_header_bytes(void);
int tal_peer_alloc(long long bytes, void** ptr_host);           /* zero-filled */
int tal_peer_free(void* ptr);
int tal_peer_export(void* ptr, unsigned char* handle_host);     /* TAL_PEER_HANDLE_BYTES */
int tal_peer_import(const unsigned char* handle_host, void** ptr_host);
int tal_peer_unimport(void* ptr);
int tal_peer_enable_access(int peer_device);                    /* same-process peers */
/* all-gather of the per-shard arg-max records: publish = store `rec` into slot
 * `rank` of every mailbox; combine = wait for the G records of this exchange in
 * the own mailbox and reduce them like tal_argmax_combine.                      */
int tal_peer_publish_record(const tal_argmax_result* rec, void* const* mailboxes, int G, int rank,
                            void* stream);
int tal_peer_combine_records(void* mailbox, int G, tal_argmax_result* out, void* stream);
/* broadcast of the observed row.  bounds_host[G+1]: rank g holds global rows
 * [bounds[g], bounds[g+1]) (multiples of 4; bounds[G] = N).  Every rank signals
 * every peer that its row buffer is free, stores the piece of row idx it holds
 * into every mailbox's kbuf once that peer is ready, and signals completion;
 * full storage: the owner of row idx holds the whole row; lower != 0 (lower-block
 * storage, see tal_rank1_update_lower): the owner holds the columns below the
 * diagonal block's end, the owners of the later rows hold the rest as their
 * column idx.  If V != null (V is [q][m_pad][ldv] fp64, the incremental-
 * conditioning state of candidates [cand0, cand0 + n_cand)) the rank whose
 * candidate range contains idx also stores the column V[:, :, idx - cand0] into
 * every mailbox's pcol (GP i's column at pcol + i * pcol_stride doubles; pcol_stride >= m_pad
 * is the row length of the mailbox's pcol area, m_pad the number of rows in use).  tal_peer_wait_row blocks the stream until the G pieces of this
 * exchange have landed and writes scal[i] = mean[i][idx].                       */
int tal_peer_push_row(void* const* mailboxes, int G, int rank, const long long* bounds_host, const void* cov,
                      long long cov_stride_q, long long ld, long long N, int q, int lower,
                      const long long* idx_dev, long long idx_host, long long kbuf_off,
                      const double* V, long long v_stride_q, long long ldv, long long m_pad,
                      long long pcol_off, long long pcol_stride, long long cand0, long long n_cand, int dtype,
                      void* stream);
int tal_peer_wait_row(void* mailbox, int G, const long long* idx_dev, long long idx_host, const double* mean,
                      long long N, int q, double* scal, void* stream);
int tal_peer_error(const void* mailbox, int* error_host, void* stream); /* synchronises */

/* ---- one BO step as one host call -------------------------------------------
 * The per-step launch sequence of the greedy loop (lib/algorithms/__init__.py:77-94
 * with the directed-ITL rule over a resident conditioning state) issued from a
 * context that the host fills once: every field is 8 bytes wide (pointers, long
 * long, double), device pointers unless noted.  The O(N) work of a step is fused
 * into three kernels (csrc/fused.cu: observed row + pending correction + step
 * vectors; the read-only pass over the conditioning factor incl. the column
 * gather; appended rows + cs + score + masked arg-max [+ peer publish]); they
 * compute every value with the same sequence of rounded operations as the
 * stand-alone entry points above.  The counters `pending`, `rows`, `passes`,
 * `n_ev`, `sel_state`, `scal_ok` are advanced here and read back by the host.
 *   tal_ctx_select   makes `combined` hold the selection for the current state:
 *                    nothing to launch when the previous tal_ctx_observe already
 *                    produced it (auto_select), otherwise F = directed-ITL score of
 *                    this rank's candidates from (var, cs, rho) and the masked
 *                    arg-max (sharded: peer all-gather + combine).
 *   tal_ctx_observe  row idx (own rows, or the peer exchange incl. the conditioning
 *                    column; the arg-max records published by the previous step are
 *                    combined inside the same kernel), pending-row correction, step
 *                    vectors, the append-form conditioning update of every GP (if
 *                    cond_valid) and, with auto_select != 0, the selection of the
 *                    next candidate; then the covariance update: queued, and applied
 *                    to the matrix in one pass every `update_block` observations
 *                    (update_block = 1: at once).  y_dev == null: the observation
 *                    value is not known yet: everything that does not depend on it
 *                    is applied and tal_ctx_observe_y completes mean, l, u later.
 *                    Returns TAL_CTX_REBASE after enqueueing the observation WITHOUT
 *                    the conditioning update when the factor is full: the caller
 *                    re-conditions from scratch before the next score.
 *   tal_ctx_step     select + observe with y taken from y_table[i][idx].
 *   tal_ctx_flush    applies the pending updates now.
 *   tal_ctx_timing_begin / _read: CUDA events around every pass over the matrix
 *                    (read synchronises the stream; updates_out[i] = number of
 *                    rank-1 updates that pass applied): the kernel times bench.py reports. */
#define TAL_CTX_REBASE 6
#define TAL_CTX_MAX_EVENTS 64
typedef struct tal_step_ctx {
  /* model */
  void* cov; long long cov_stride_q, ld, row0, n_loc, N, q, dtype, lower;
  double *mean, *var, *l, *u, *rho, *scal;
  void *kbuf, *wbuf;
  double beta[TAL_MAX_Q];            /* host values */
  /* deferred covariance updates */
  void *Kp, *Wp; long long pend_stride, pend_slots, update_block, pending, passes;
  /* incremental conditioning, append form (Wb null: none) */
  double* Wb; long long ldw, cap_rows, m_pad, rows, ncols, cand0, n_cand, n_split, cond_valid;
  double *cs, *app_scratch, *pcol;
  /* selection */
  double* F; const unsigned char *safe, *sample; long long first_constraint;
  void* argmax_scratch; tal_argmax_result *rec, *combined;
  /* observations for tal_ctx_step */
  const double* y_table; long long y_stride;
  /* row-sharded model over peer memory (G == 1: unused) */
  void* const* mailboxes; void* mailbox; long long G, rank;
  long long bounds[TAL_PEER_MAX + 1]; /* host values */
  long long kbuf_off, pcol_off; const double* pcol_peer; long long pcol_stride;
  /* pass timing (host handles) */
  void* ev_start[TAL_CTX_MAX_EVENTS]; void* ev_stop[TAL_CTX_MAX_EVENTS]; long long n_ev, timing;
  long long ev_updates[TAL_CTX_MAX_EVENTS];
  /* fused step */
  long long arith;                   /* != 0: one-rounding update arithmetic (TAL_ARITH_FMA) */
  long long auto_select;             /* != 0: tal_ctx_observe also selects the next candidate */
  long long sel_state;               /* 0 none, 1 record published to the peers, 2 `combined` is current */
  long long scal_ok;                 /* scal holds mean[:, combined.idx] */
  const unsigned char *sel_safe, *sel_sample; long long sel_fc;  /* masks the selection was made under */
  void* fz_scratch;                  /* tal_fused_scratch_bytes(q, ncols, n_split) */
  /* fused_observe and fused_partial do not depend on each other (the scoring pass needs the column of the
   * conditioning factor, not the observed row): the scoring pass runs on a side stream between two events.
   * Library-owned handles, created on first use (zero-initialise); overlap == 0 keeps one stream. */
  void *side_stream, *ev_fork, *ev_join; long long overlap;
} tal_step_ctx;
/* diagnostic (TAL_BREAKDOWN=1): mean microseconds per phase of the recorded fused steps -- [0] peer push + record
 * combine, [1] fused_observe (+ row wait), [2] fused_partial, [3] fused_finish (+ arg-max, publish), [4] covariance
 * pass (amortised over the steps), [5] gap to the next step; *n_host = steps seen.  Synchronises the device. */
int tal_ctx_breakdown(double* out_us_host, long long* n_host);
long long tal_step_ctx_bytes(void);
long long tal_fused_scratch_bytes(int q, long long ncols, int n_split);
int tal_ctx_select(tal_step_ctx* ctx, void* stream);
int tal_ctx_observe(tal_step_ctx* ctx, const long long* idx_dev, long long idx_host, const double* y_dev,
                    long long y_stride, int y_gather, void* stream);
int tal_ctx_observe_y(tal_step_ctx* ctx, const long long* idx_dev, long long idx_host, const double* y_dev,
                      long long y_stride, int y_gather, void* stream);
int tal_ctx_step(tal_step_ctx* ctx, void* stream);
int tal_ctx_flush(tal_step_ctx* ctx, void* stream);
int tal_ctx_timing_begin(tal_step_ctx* ctx);
int tal_ctx_timing_read(tal_step_ctx* ctx, double* ms_out_host, long long* updates_out_host,
                        long long* n_out_host, void* stream);

/* ---- fp64 peak probes (roofline denominators of the tensor path) --------
 * kind 0: DMMA.8x8x4, kind 1: DFMA; `blocks` CTAs x 256 threads x iters x 16
 * independent ops.  flops = blocks*iters*16*(kind==0 ? 8*512 : 256*2).     */
int tal_probe_fp64(int kind, int blocks, int iters, double* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TAL_B200_H */
