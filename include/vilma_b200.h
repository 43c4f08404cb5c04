/* vilma_b200 -- C ABI of the B200-native `vilma fit` hot path (libvilma_b200.so).
 *
 * The reference (jeffspence/vilma v0.0.16) is pure Python and has no FFI: the de-facto
 * contract of its hot path is the VIScheme/MultiPopVI API (variational_inference.py:27-889)
 * over BlockDiagonalMatrix.dot (matrix_structures.py:389-408).  This header is what a
 * maintainer of the reference would bind (ctypes) to replace, one for
 * one, the array operations of that path.  Each entry point cites the reference code it
 * replaces (paths relative to /root/reference/src/vilma/).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; vb_last_error() gives the text;
 *    no C++ exception crosses the boundary; nothing falls back to the CPU.
 *  - plain pointers and sizes only.  `*_host` pointers are host memory, `*_dev` device
 *    memory (fp64 unless noted).  The library never frees caller memory.
 *  - one CUDA stream per context (given at creation; 0 = legacy default stream).  Calls are
 *    asynchronous on that stream unless they return data through a `*_host` pointer.
 *  - device layouts: vi_mu [K][P][M], vi_delta [K][M] (the TRANSPOSE of the reference's
 *    [M][K]), per-SNP vectors [P][M]; M is the number of SNPs owned by this rank.
 */
#ifndef VILMA_B200_H
#define VILMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vb_ctx vb_ctx;
typedef struct vb_ld vb_ld; /* one cohort's block-diagonal LD operator, resident in HBM */

#define VB_ABI_VERSION 1
#define VB_MAX_POPS 6 /* cohorts compiled into the per-SNP kernels */

/* ---- context ------------------------------------------------------------------------ */
int vb_abi_version(void);
/* SHA-256 (hex) of the sources this library was compiled from (vilma_b200/_build.py) */
const char* vb_source_hash(void);
const char* vb_last_error(void);
/* process-wide options (kernel selection; the defaults are the measured best):
 *   "ld_symmetric" (default 1): store dense blocks of n <= vb_ld_sym_nmax() symmetric-packed
 *                  (read when an LD operator is created)
 *   "ld_factor_once" (default 1): factor blocks of n <= vb_ld_fac_nmax() are stored as U sqrt(s) and read once
 *                  per mat-vec (8 n r bytes; needs s >= 0); 0 = always V' = diag(s) U^T and U (16 n r bytes)
 *   "shard_zero_copy" (default 1): vb_fit_set_params_shard reads page-locked sources in place (0: always staged)
 *   "ld_fused_finish" (default 0): 1 = the symmetric kernel also finishes every block whose last group it flushed
 *   "snp_three_pass" (default 1): one- and two-cohort updates with K < 32 use the three-pass softmax kernel
 *   "snp3_park" (default 1): that kernel keeps logits / weights / mu' in shared memory when they fit
 *   "snp_tile" (default -1 = automatic: P >= 3 or K >= 32): the K-split tile kernel; 0 = never,
 *                  W in {1,2,4,8,16} = always, with W warps per 32-SNP tile
 *   "snp_ann_slots" (default 1): fused annotation sums in per-thread shared-memory slots (A*K <= 16) */
int vb_set_option(const char* name, int64_t value);
/* host-only: launch plan of the K-split tile kernel for a problem shape (W = 0: thread-per-SNP kernels) */
int vb_debug_tile_plan(int P, int K, int64_t M, int akf, int num_sms, int* W, int* grid, int64_t* smem_bytes);
/* largest dense block (rows) that is stored symmetric-packed; larger ones are stored in full */
int64_t vb_ld_sym_nmax(void);
/* largest factor block (rows) that is stored in the read-once form (0: option "ld_factor_once" is off) */
int64_t vb_ld_fac_nmax(void);
/* device: CUDA ordinal; stream: cudaStream_t (may be NULL). */
int vb_ctx_create(int device, void* stream, vb_ctx** out);
int vb_ctx_destroy(vb_ctx* ctx);
int vb_ctx_sync(vb_ctx* ctx);
/* number of kernels launched by this context since creation (bench.py "gpu_launches") */
int64_t vb_ctx_launch_count(const vb_ctx* ctx);
/* per-kernel timing with CUDA events on the context's stream (bench.py roofline leg):
 * enable / reset, then read {total ms, launches} for the 4 categories 0 = LD mat-vec kernel,
 * 1 = fused per-SNP kernel, 2 = mat-vec finish kernel (incl. the final reduction / rank exchange in
 * its last CTA), 3 = bookkeeping kernels (annotation sums, convergence partials).  At most 32768
 * launches per category are timed between reads. */
int vb_ctx_profile(vb_ctx* ctx, int enable);
int vb_ctx_profile_read(vb_ctx* ctx, double* total_ms4, int64_t* count4);

/* ---- LD operator ---------------------------------------------------------------------
 * Replaces BlockDiagonalMatrix(matrices, perm, missing) and its .dot
 * (matrix_structures.py:277-331, :389-408) over LowRankMatrix.dot (:148-152).
 *
 * vb_ld_create     declare the blocks: n[b] rows, rank[b] columns of the factor (rank[b] < 0:
 *                  block b is given dense).  M = SNPs on this rank.
 * vb_ld_set_dense  block b as a dense symmetric n x n matrix, row stride `ld` doubles.
 * vb_ld_set_factor block b as U (n x r, row-major, row stride r) and s (r):  R_b = U diag(s) U^T
 *                  (the reference's u, s; its v is u^T and D must be zero on this path).  Blocks of
 *                  n <= vb_ld_fac_nmax() need s >= 0 (true of every block the reference builds from a
 *                  matrix: it drops eigenvalues <= 1e-12 max, matrix_structures.py:18) and fail otherwise.
 * vb_ld_finalize   perm[j] = SNP index (0..M-1, this rank's numbering) of block-order
 *                  position j, for j < sum_b n[b]; SNPs not listed are "missing" (zero rows).
 * vb_ld_dot        y = R x in SNP order (x, y device vectors of length M).
 * vb_ld_bytes      algorithmic bytes one mat-vec reads from the LD store: 4 n (n+1) per
 *                  symmetric-packed dense block, 8 n^2 per full dense block, 8 n r per read-once
 *                  factor block (n padded to even), 16 n r per two-pass factor block.
 */
int vb_ld_create(vb_ctx* ctx, int64_t M, int64_t nblocks, const int64_t* n, const int64_t* rank,
                 vb_ld** out);
int vb_ld_destroy(vb_ld* ld);
int vb_ld_set_dense(vb_ld* ld, int64_t b, const double* R, int64_t ldr, int on_device);
int vb_ld_set_factor(vb_ld* ld, int64_t b, const double* U, const double* s, int on_device);
int vb_ld_finalize(vb_ld* ld, const int64_t* perm_host, int64_t nperm);
int vb_ld_dot(vb_ld* ld, const double* x_dev, double* y_dev);
int64_t vb_ld_bytes(const vb_ld* ld);

/* ---- set-up on the device ------------------------------------------------------------
 * SURVEY 8(f) rows 1-2.  For DENSE blocks whose eigenvalues all exceed 1e-12 max -- where the
 * reference's LowRankMatrix(X, t=1) keeps every eigenpair (matrix_structures.py:15-28, :119), so that
 * its operator is X, its pseudo-inverse X^-1 and its rank n -- the per-block eigendecomposition
 * (:17), pseudo-inverse product (inverse_dot, :159-196) and Woodbury ridge solve (ridge_inverse_dot,
 * :349-387) of VIScheme.__init__ (variational_inference.py:236-252) reduce to two Cholesky
 * factorisations, done here by one CTA per block.
 *   n_host[b]          block sizes (<= vb_setup_nmax());  R_dev: the blocks back to back, each n x n
 *                      row-major and exactly symmetric;  W_dev: workspace of the same size
 *   z_dev, reg_dev     block-order vectors: z-scores and the ridge diagonal se^2 / prior
 *   mle_dev            X^-1 z            rmle_dev   X (X^-1 z)          ridge_dev  (X + diag(reg))^-1 (X mle)
 *   chi_dev[b]         z_b . mle_b       lam_dev[2b], [2b+1]  lambda_min estimate, ||X||_inf
 *   status_dev[b]      0 = done; 1 = the block is not safely full rank (a pivot or the inverse-iteration
 *                      estimate of lambda_min fell below the floor): its outputs are undefined and the
 *                      caller must take the exact eigen path for it. */
int64_t vb_setup_nmax(void);
int vb_setup_dense(vb_ctx* ctx, int64_t nblocks, const int64_t* n_host, const double* R_dev,
                   double* W_dev, const double* z_dev, const double* reg_dev, double* mle_dev,
                   double* rmle_dev, double* ridge_dev, double* chi_dev, double* lam_dev,
                   int32_t* status_dev);

/* ---- fit state ----------------------------------------------------------------------
 * vb_fit_create       allocate the device state for K components, P cohorts, M SNPs, A annotations
 *                     over the P finalized LD operators `lds` (same context, same M).
 * vb_fit_set_snp_data adj_marginal_effects, std_errs, scaled_ld_diags, scalings [P][M] and the
 *                     annotation label per SNP (variational_inference.py:205-259).
 * vb_fit_set_mixture  mixture_prec [K][P][P] and log_det [K]            (:622-626)
 * vb_fit_set_hyper    hyper_delta [A][K] used by the KL term (fast_delta_kl, numerics.py:132-141)
 * vb_fit_set_delta_grad  nat_grad_vi_delta as an [A][K-1] table: fast_vi_delta_grad's value
 *                     for annotation a (numerics.py:149-164; variational_inference.py:694, :706, :844)
 * vb_fit_set_tau      error_scaling [P]; implies _set_vi_sigma()        (:712-738)
 * vb_fit_set_params / vb_fit_get_params   accepted state, host arrays in the REFERENCE layouts
 *                     vi_mu [K][P][M], vi_delta [M][K] (transposed to [K][M] on the device).
 */
int vb_fit_create(vb_ctx* ctx, int K, int P, int64_t M, int A, vb_ld* const* lds);
int vb_fit_destroy(vb_ctx* ctx);
/* fuse_ann != 0: evaluations also return the per-annotation sums of delta (see below); used on
 * multi-GPU runs where a separate pass + reduction per hyper step costs more than it saves */
int vb_fit_set_fusion(vb_ctx* ctx, int fuse_ann);
int vb_fit_set_snp_data(vb_ctx* ctx, const double* adj_host, const double* se_host,
                        const double* sld_host, const double* scalings_host,
                        const int32_t* ann_host);
int vb_fit_set_mixture(vb_ctx* ctx, const double* prec_host, const double* logdet_host);
int vb_fit_set_hyper(vb_ctx* ctx, const double* hyper_host);
int vb_fit_set_delta_grad(vb_ctx* ctx, const double* table_host);
int vb_fit_set_tau(vb_ctx* ctx, const double* tau_host);
int vb_fit_set_params(vb_ctx* ctx, const double* vi_mu_host, const double* vi_delta_mk_host);
int vb_fit_get_params(vb_ctx* ctx, double* vi_mu_host, double* vi_delta_mk_host);
/* the same with device buffers (multi-GPU: shards are gathered / scattered on the device) */
int vb_fit_set_params_dev(vb_ctx* ctx, const double* vi_mu_dev, const double* vi_delta_mk_dev);
int vb_fit_get_params_dev(vb_ctx* ctx, double* vi_mu_dev, double* vi_delta_mk_dev);
/* Sharded transfers (multi-GPU): `snps_host[M]` = global index of each SNP this rank owns, of M_total.
 * The *_shard calls move only this rank's SNPs between the device state and GLOBAL host arrays in the
 * reference layouts (vi_mu [K,P,M_total], vi_delta [M_total,K]); they replace gathering
 * `variational_inference.py:340-394`'s parameter tuple on every rank.  set: any host memory -- arrays the
 * GPU can address are gathered by it directly (zero-copy reads; option "shard_zero_copy", default 1), others
 * have their runs of consecutive SNPs cut out into page-locked staging by a few host threads.  get: the GPU scatters
 * into the arrays itself, so they must be memory it can address -- page-locked allocations or memory
 * registered with vb_host_register, e.g. one node-shared mapping of which every rank fills in its own
 * part; it fails (no fallback) on plain pageable memory: ask vb_host_accessible. */
int vb_fit_set_shard(vb_ctx* ctx, const int64_t* snps_host, int64_t M_total);
int vb_fit_set_params_shard(vb_ctx* ctx, const double* vi_mu_global_host, const double* vi_delta_mk_global_host);
int vb_fit_get_params_shard(vb_ctx* ctx, double* vi_mu_global_host, double* vi_delta_mk_global_host);
int vb_host_accessible(const void* host);            /* 1 if device kernels can address it */
int vb_host_register(void* host, int64_t bytes);     /* cudaHostRegister(portable | mapped) */
int vb_host_unregister(void* host);

/* ---- evaluations: each fills stats_dev[0 .. 3P+3) for ONE parameter state; stats_dev must hold
 * 3P+3+58 doubles: with vb_fit_set_fusion(ctx, 1) and A*K <= 48, entries [3P+3, 3P+3+A*K) receive
 * the per-annotation sums of that state's delta (index a*K+k) ------------------------------
 *   [0,P)   A_p = sum_i pm adj      [P,2P)  C_p = sum_i sld pv      [2P,3P) B_p = sum_i z (R z)
 *   3P KL_delta  3P+1 KL_quad  3P+2 KL_sigma
 * from which  loglik = sum_p [(-(C_p+B_p)/2 + A_p - chi_p/2)/tau_p - rank_p log(tau_p)/2]
 * (numerics.py:31-46) and beta_KL = KL_delta + KL_quad + KL_sigma (variational_inference.py:873-885).
 *
 * vb_fit_eval           the accepted state as it stands           (_log_likelihood :452-470, _beta_KL)
 * vb_fit_beta_trial     accepted -> trial at step size `step`     (_update_beta loop body :778-787,
 *                                                                  _nat_grad_beta :804-823)
 * vb_fit_refresh_delta  trial := (accepted mu, delta recomputed from the current hyper/tau)
 *                                                                 (_nat_to_not_vi_delta :632-641)
 * vb_fit_accept         trial becomes the accepted state.
 */
int vb_fit_eval(vb_ctx* ctx, double* stats_dev);
int vb_fit_beta_trial(vb_ctx* ctx, double step, double* stats_dev);
int vb_fit_refresh_delta(vb_ctx* ctx, double* stats_dev);
int vb_fit_accept(vb_ctx* ctx);

/* sum_annotations of the accepted delta -> out_dev [A][K]   (numerics.py:118-129) */
int vb_fit_sum_annotations(vb_ctx* ctx, double* out_dev);
/* accepted state's posterior mean / marginal variance, in model units (NOT multiplied by
 * scalings) -> host [P][M]  (_posterior_mean, _posterior_marginal_variance :753-760) */
int vb_fit_posterior(vb_ctx* ctx, double* pm_host, double* pv_host);
/* convergence bookkeeping on the real posterior mean (:376-377 allclose, :292-331 _dump_info):
 * out_dev[0..10) = {#violations, sum|d|, sum d^2, sum|d_ckpt|, sum d_ckpt^2,
 *                   max|new|, max rel, max abs, max rel ckpt, max abs ckpt};
 * the first five are sums over SNPs (all-reduce SUM), the last five maxima (MAX).
 * Then prev := new.  vb_fit_pm_mark(which): 0 -> prev := current, 1 -> ckpt := current. */
int vb_fit_pm_diff(vb_ctx* ctx, double atol, double rtol, double* out_dev);
int vb_fit_pm_mark(vb_ctx* ctx, int which);
/* MultiPopVI._initialize on the device (variational_inference.py:643-700), for problems whose
 * [K,P,P,M] host intermediates do not fit: vb_fit_init_delta(fake_mu_host [P][M]) leaves the first
 * delta (:660-667) in the current state; the caller forms hyper_delta from its annotation sums
 * (vb_fit_sum_annotations; :668-674) and sets it; vb_fit_init_mu then writes mu (:675-678);
 * vb_fit_refresh_delta + vb_fit_accept complete the start (:679). */
int vb_fit_init_delta(vb_ctx* ctx, const double* fake_mu_host);
int vb_fit_init_mu(vb_ctx* ctx);
/* vi_sigma[k0:k1][P][P][M] in the reference layout -> host  (:712-724) */
int vb_fit_vi_sigma(vb_ctx* ctx, int k0, int k1, double* out_host);

/* ---- native control loop --------------------------------------------------------------
 * One outer iteration of optimize() without leaving C++: VIScheme._nat_grad_step
 * (variational_inference.py:419-450: beta line search :762-802, hyper step :825-860, tau step
 * :441-448/:472-486) followed by the convergence bookkeeping (:376-377), with the reference's
 * thresholds.  Multi-GPU: vb_nccl_unique_id (rank 0) + vb_comm_init (every rank) create a NCCL
 * communicator; the statistics of every evaluated state are then summed with ncclAllReduce.
 * vb_fit_iteration returns 0, 1 (error) or 2 ("Encountered a numerical error.", :793/:797). */
typedef struct vb_step_io {
    double L[5];               /* in/out */
    double line_search_rate;   /* in  (2.0 in optimize()) */
    double running_elbo_delta; /* in  (ignored unless has_running) */
    double obj;                /* in: objective of the accepted state; out: after the iteration */
    double elbo_delta;         /* out: accumulated change of the ELBO */
    double atol, rtol;         /* in: tolerances of the posterior-mean convergence test */
    double diff[10];           /* out: see vb_fit_pm_diff */
    int32_t has_running;       /* in */
    int32_t trials;            /* out: line-search trials executed */
    int32_t evals;             /* out: parameter states evaluated */
    int32_t do_diff;           /* in: run the convergence bookkeeping */
    int32_t speculate;         /* in: queue the next iteration's first beta trial behind this one's
                                *     last evaluation (discarded if the caller stops or touches the state) */
    int32_t rejects;                /* out: trials of this iteration whose step size was rejected (L doubled) */
} vb_step_io;
int vb_nccl_unique_id(char* out128);
/* optional faster rendezvous for one node: every rank creates a mailbox (vb_xr_create returns its
 * 64-byte CUDA IPC handle), the handles are exchanged by the caller and opened with vb_xr_open
 * (nranks x 64 bytes, rank order).  The last CTA of every evaluation then sums the statistics over
 * the ranks through NVLink peer stores and publishes them to mapped pinned memory the host polls:
 * no NCCL launch, copy or stream synchronisation per evaluation.  With nranks == 1 it just replaces
 * the copy + synchronisation by the polled flag. */
int vb_xr_create(vb_ctx* ctx, char* handle_out64);
int vb_xr_open(vb_ctx* ctx, int nranks, int rank, const char* handles);
int vb_comm_init(vb_ctx* ctx, int nranks, int rank, const char* id128);
int vb_fit_set_constants(vb_ctx* ctx, const double* chi_stat_host, const double* ld_ranks_host,
                         const double* annotation_counts_host, const double* log_det_host,
                         int scale_se);
/* tau_io [P], hyper_io [A][K], stats_io [3P+3] (statistics of the accepted state) are host in/out */
int vb_fit_iteration(vb_ctx* ctx, vb_step_io* io, double* tau_io, double* hyper_io,
                     double* stats_io);
/* host-side time of the native loop so far: {s enqueueing evaluations, s waiting for their
 * statistics, number of rendezvous, speculative trials used + 1e-6 * wasted} */
int vb_fit_timing(vb_ctx* ctx, double* out4);

#ifdef __cplusplus
}
#endif
#endif /* VILMA_B200_H */
