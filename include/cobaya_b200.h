/*
 * cobaya_b200.h -- C ABI of the B200 ensemble-MCMC engine (libcobaya_b200.so).
 *
 * The reference (CobayaSampler/cobaya v3.6.2) is pure Python: its plugin
 * boundary for this path is the Python class cobaya.samplers.mcmc.MCMC
 * (cobaya/samplers/mcmc/mcmc.py:47) built on cobaya.sampler.CovmatSampler
 * (cobaya/sampler.py:467).  There is no FFI in the reference; this header is
 * the FFI a maintainer would bind from that class with ctypes (see
 * INTEGRATION.md).  Each entry point names the reference code it replaces.
 *
 * Conventions: extern "C"; plain pointers and sizes; return 0 = ok, negative =
 * error (message via cb2_last_error): -1 bad argument / call order, -2 CUDA error,
 * -3 no device or library (there is no CPU fallback), -4 not supported for this model,
 * -5 not enough rows yet, -6 NVRTC not available, -7 run-time compile error (external
 * functions); the caller owns every HOST buffer, the
 * library owns all DEVICE memory behind the opaque handle; one host thread per
 * handle; all work is enqueued on the handle's CUDA stream, cb2_sync blocks.
 * All floating-point data is IEEE binary64; matrices are row-major.
 */
#ifndef COBAYA_B200_H
#define COBAYA_B200_H

#ifndef __CUDACC_RTC__
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb2_engine cb2_engine;

#define CB2_MAX_BLOCKS 16
#define CB2_MAX_LIKES 8
#define CB2_MAX_MODES 32

/* flags returned per chain by cb2_get_flags */
#define CB2_FLAG_STUCK 1u        /* mcmc.py:717-743 "chain has been stuck" */
#define CB2_FLAG_ROWS_FULL 2u    /* per-chain sample capacity exhausted */
#define CB2_FLAG_INTERNAL 4u     /* engine invariant violated (basis/tape window), a
                                    non-finite start point in cb2_set_state, or a NaN
                                    returned by an external function */

/* moment modes of cb2_moments */
#define CB2_MOMENTS_HALVES 0     /* multi-chain rule, mcmc.py:785-793 */
#define CB2_MOMENTS_SINGLE_SPLIT 1 /* single-chain rule, mcmc.py:795-822 */

int cb2_abi_version(void);
/* message of the last failing call on this handle (or of cb2_create if h==NULL) */
const char *cb2_last_error(const cb2_engine *h);

/* One engine = one GPU = `n_chains` lock-step chains with global ids
 * chain_id0 .. chain_id0+n_chains-1 (Philox subsequence = global id, replacing
 * Sampler._set_rng's SeedSequence.spawn per MPI rank, cobaya/sampler.py:369-384). */
int cb2_create(int device, int64_t n_chains, int32_t D, uint64_t seed,
               uint64_t chain_id0, cb2_engine **out);
int cb2_destroy(cb2_engine *h);

/* Prior.logps_internal inputs (cobaya/prior.py:514-533,733-763): per parameter
 * kind (CB2_PRIOR_*), bounds, loc/scale, periodic flag
 * (prior.py:500-513, reduce_periodic :658-676) and the precomputed
 * -sum(log(upper-lower)) over uniform parameters (prior.py:526-532). */
int cb2_set_prior(cb2_engine *h, const int32_t *kind, const double *lower,
                  const double *upper, const double *loc, const double *scale,
                  const int32_t *periodic, double uniform_logp);

/* 1-D prior kinds.  0/1 are the reference's own fast paths (uniform: prior.py:526-532;
 * normal: _fast_norm_logpdf, tools.py:720-729); the others are the scipy.stats
 * distributions the reference reaches through pdf.logpdf (prior.py:520-525,
 * tools.py:611-718), evaluated as  log_norm + f((x-loc)/scale; a, b). */
#define CB2_PRIOR_UNIFORM 0
#define CB2_PRIOR_NORMAL 1
#define CB2_PRIOR_TRUNCNORM 2  /* a, b: truncation in standard units; f = -z^2/2 */
#define CB2_PRIOR_HALFNORM 3   /* f = -z^2/2 */
#define CB2_PRIOR_EXPON 4      /* f = -z */
#define CB2_PRIOR_BETA 5       /* f = (a-1) log z + (b-1) log(1-z) */
#define CB2_PRIOR_GAMMA 6      /* f = (a-1) log z - z */
#define CB2_PRIOR_LOGNORM 7    /* a = s: f = -log z - log(z)^2 / (2 s^2) */
#define CB2_PRIOR_CAUCHY 8     /* f = -log(1+z^2) */
#define CB2_PRIOR_LAPLACE 9    /* f = -|z| */
#define CB2_PRIOR_LOGUNIFORM 10 /* a, b: support in standard units; f = -log z */

/* Shape parameters a[D], b[D] and log-normalisation log_norm[D] (includes -log(scale))
 * of the parameters whose kind is >= 2; entries of other parameters are ignored.
 * Must follow cb2_set_prior (which resets them). */
int cb2_set_prior_shapes(cb2_engine *h, const double *a, const double *b,
                         const double *log_norm);

int cb2_clear_likelihoods(cb2_engine *h);
/* GaussianMixture (cobaya/likelihoods/gaussian_mixture/gaussian_mixture.py:45-163):
 * idx[dim] = positions of its input_params in the sampled vector; means[m*dim];
 * linv[m*dim*dim] = inverse_cholesky(cov_k) (functions.py:81-89); logdet[m];
 * weights[m] normalised; derived!=0 emits the dim*m whitened derived parameters
 * (gaussian_mixture.py:146-156) into the sample rows. */
int cb2_add_gaussian_mixture(cb2_engine *h, int32_t dim, const int32_t *idx,
                             int32_t n_modes, const double *means, const double *linv,
                             const double *logdet, const double *weights,
                             int32_t derived);
/* Built-in external-likelihood stand-in for BASELINE.json configs[3]:
 * logp = -scale * sum_i [100 (x_{i+1}-x_i^2)^2 + (1-x_i)^2]. */
int cb2_add_rosenbrock(cb2_engine *h, int32_t dim, const int32_t *idx, double scale);
/* External likelihood functions (LikelihoodExternalFunction, cobaya/likelihood.py:150-255: the
 * reference accepts any Python callable) as DEVICE FUNCTORS: `cuda_source` is CUDA C++ defining
 *     extern "C" __device__ double <fn_name>(const double *p, int n)
 * which returns log L of the point whose input parameters (sampled indices `idx`, in that
 * order) are p[0..n).  The source is compiled at run time with NVRTC (loaded with dlopen; the
 * library does not link it) for sm_100a into an evaluation kernel, one thread per chain; a
 * Metropolis proposal then runs as propose / evaluate / accept launches on the engine's stream
 * (csrc/kernels_ext.cuh), with the same proposals as every other step kernel.  CB2_EXT_DIM is
 * predefined as `dim`.  Returns -6 if NVRTC cannot be loaded, -7 on a compile error (the
 * compiler log is cb2_last_error).  With dragging the step is split at every posterior
 * evaluation (k_extd_*). */
int cb2_add_external_likelihood(cb2_engine *h, int32_t dim, const int32_t *idx,
                                const char *cuda_source, const char *fn_name);
/* External priors (cobaya/prior.py:537-577,765-772: any Python callable under `prior:`) through
 * the same route: the function returns a log-prior term of its parameters; it is evaluated where
 * the internal prior is finite, added to the log-prior, and reported in its own
 * minuslogprior__<name> column after minuslogprior__0 (the row gets one column wider per external
 * prior).  Call after cb2_set_prior (which removes them) and before cb2_set_state. */
int cb2_add_external_prior(cb2_engine *h, int32_t dim, const int32_t *idx,
                           const char *cuda_source, const char *fn_name);
/* Compile check of such a source without an engine or a GPU; the compiler log goes to `log`. */
int cb2_check_external_source(const char *cuda_source, const char *fn_name, int32_t dim,
                              char *log, int64_t log_cap);
/* The same check for the FUSED form: when a model has external likelihood functions (and no
 * external prior), the engine recompiles its general step kernel (csrc/kernels_general.cuh, a
 * copy of which is embedded in the library) with the functions inlined, and a whole window --
 * Metropolis or dragging -- is one launch; CB2_EXT_FUSED=0 keeps the split launches. */
int cb2_check_external_fused(const char *cuda_source, const char *fn_name, int32_t dim,
                             char *log, int64_t log_cap);
/* Instrumentation: out = { windows run on the fused (run-time compiled) step kernel, proposals
 * replayed as a CUDA graph on the split route, fused kernel loaded (0/1), fused compile
 * failed (0/1) }. */
int cb2_ext_route_counts(cb2_engine *h, int64_t out[4]);
/* `one` (cobaya/likelihoods/one/one.py:26-28): a likelihood with no input parameters
 * whose log-value is a constant (0 for `one`); keeps its chi2__<name> column. */
int cb2_add_constant(cb2_engine *h, double value);

/* BlockedProposer.__init__ (cobaya/samplers/mcmc/proposal.py:96-201): blocks in
 * ascending speed given by their sizes, i_of_j (sorted index -> sampler index),
 * integer oversampling factors; drag!=0 selects get_new_sample_dragging
 * (mcmc.py:564-668) with slow blocks 0..i_last_slow and drag_interp_steps. */
int cb2_set_blocking(cb2_engine *h, int32_t n_blocks, const int32_t *block_sizes,
                     const int32_t *oversampling, const int32_t *i_of_j, int32_t drag,
                     int32_t i_last_slow_block, int32_t drag_interp_steps);

/* BlockedProposer.set_covariance output (proposal.py:226-260): T[D*D] = S L' in
 * block-sorted coordinates (lower triangular); transform[b] = T[j_b:, j_b:j_b+n_b].
 * May be called between cb2_advance calls (covariance learning, mcmc.py:1023). */
int cb2_set_proposal(cb2_engine *h, const double *T, double proposal_scale);

/* mcmc.yaml options used on the device: temperature (:24), burn_in (:6, absolute
 * accepted steps), max_tries (:9, absolute), output_thin (mcmc.py:377-389),
 * rows_cap = stored rows per chain the engine must be able to hold. */
int cb2_set_options(cb2_engine *h, double temperature, int64_t burn_in,
                    int64_t max_tries, int32_t output_thin, int64_t rows_cap);

/* Initial points x0[n_chains*D] (MCMC.initialize, mcmc.py:215-222): uploads,
 * evaluates the posterior on the device, weight=1, resets counters and samples. */
int cb2_set_state(cb2_engine *h, const double *x0);
/* current points (OneSamplePoint): any output pointer may be NULL */
int cb2_get_state(cb2_engine *h, double *x, double *logpost, int64_t *weight,
                  int64_t *n_rows, int64_t *n_accepted, uint32_t *flags);

/* Resuming (replaces "last point of the chain file + .checkpoint", cobaya/samplers/mcmc/
 * mcmc.py:187-214, for an ensemble): the per-chain sampler state -- current point with its
 * log-posterior pieces, weight, burn-in and thinning counters, visit counters of every
 * proposer, proposal counter -- as one opaque host buffer of cb2_snapshot_size() bytes.
 * cb2_import_state replaces cb2_set_state on an engine configured identically (same model,
 * blocking, options, seed, chain ids) and is followed by cb2_load_rows for every chain
 * whose stored rows are needed by later checkpoints; the run then continues bit-for-bit. */
int64_t cb2_snapshot_size(cb2_engine *h);
int cb2_export_state(cb2_engine *h, void *buf, int64_t nbytes);
int cb2_import_state(cb2_engine *h, const void *buf, int64_t nbytes);
int cb2_load_rows(cb2_engine *h, int64_t chain, int64_t n, const double *rows);
/* the same for chains [chain_begin, chain_end) at once: rows chain-major, counts[c] each */
int cb2_load_rows_bulk(cb2_engine *h, int64_t chain_begin, int64_t chain_end,
                       const int64_t *counts, const double *rows);

/* Model.logposterior (cobaya/model.py:579-678) for n points X[n*D]:
 * logpost[n], logprior[n], loglikes[n*n_like], derived[n*n_derived] (NULL ok). */
int cb2_logpost(cb2_engine *h, const double *X, int64_t n, double *logpost,
                double *logprior, double *loglikes, double *derived);

/* MCMC.run inner loop (mcmc.py:470-472): every chain makes n_proposals more
 * proposals (get_new_sample_metropolis / _dragging + process_accept_or_reject). */
int cb2_advance(cb2_engine *h, int64_t n_proposals);
int cb2_sync(cb2_engine *h);

/* out[0..7] = min rows, max rows, sum rows, #stuck chains, #rows-full chains,
 * #chains with CB2_FLAG_INTERNAL, sum accepted, sum of current weights */
int cb2_summary(cb2_engine *h, int64_t out[8]);

/* Per-GPU sufficient statistics of check_convergence_and_learn_proposal
 * (mcmc.py:785-822): for every chain (mode HALVES: rows [n/2:], N_c=n; mode
 * SINGLE_SPLIT with `split`: segments of one chain, mcmc.py:795-813) the weighted
 * mean m_c, ddof=0 weighted covariance C_c (collection.py:893-981) and acceptance
 * a_c; summed into out[2D^2+D+3] = { M, sum N_c, sum N_c a_c, sum (m_c-shift)[D],
 * sum (m_c-shift)(m_c-shift)^T [D*D], sum N_c C_c [D*D] }.  `shift` (host, may be
 * NULL = 0) must be identical on all ranks.  `dev_out` is a DEVICE pointer (the
 * buffer NCCL all-reduces); `host_out` a host pointer; either may be NULL. */
int cb2_moments(cb2_engine *h, int32_t mode, int32_t split, const double *shift,
                double *dev_out, double *host_out);

/* The D x D part of the checkpoint on the device (D <= 64), replacing the root-only block of
 * check_convergence_and_learn_proposal (mcmc.py:856-889: numpy cholesky, scipy dtrtri, numpy
 * eigvalsh) and BlockedProposer.set_covariance (proposal.py:226-260, tools.py:761-788):
 * from the all-reduced sums of cb2_moments at DEVICE pointer `dev_sums` (NULL = the engine's
 * own buffer of the last cb2_moments call; the shift is that call's) one kernel forms W, B,
 * R-1 = max eig(L^-1 (B/dd^T) L^-T) (parallel-order Jacobi) and the candidate transform
 * T = S chol(corr(W)) in block-sorted coordinates, all kept on the device.
 * out_host[8 + D] = { M, sum N_c, acceptance, R-1 (NaN: Cholesky of W/dd^T failed,
 * mcmc.py:872-887), candidate valid (1/0), chol ok (1/0), Jacobi sweeps, 0, mean[D] }.
 * Returns -4 for D > 64 (use cb2_moments + host algebra + cb2_set_proposal). */
int cb2_checkpoint_device(cb2_engine *h, const double *dev_sums, double *out_host);
/* W (mean of the chains' covariances, sampler order, row-major D x D) of that call -> host */
int cb2_checkpoint_cov(cb2_engine *h, double *W_out);
/* mcmc.py:1023 `self.proposer.set_covariance(...)` without the host: the candidate transform
 * of the last cb2_checkpoint_device becomes the proposal; the step kernels' constant blocks
 * (fragment-ordered T, G = L^-1 P T, column-major T) are rewritten by a device kernel.
 * -4 when the model's constant blocks need the host packer (streamed path, several
 * components): fall back to cb2_set_proposal. */
int cb2_adopt_proposal(cb2_engine *h);
/* the proposal transform in use (row-major D x D, block-sorted coordinates) */
int cb2_get_proposal(cb2_engine *h, double *T_out);

/* Model.measure_and_set_speeds (cobaya/model.py:1543-1592) for the device components:
 * evaluations per second of every likelihood alone (speeds_out[n_like], in the order they were
 * added) at the n host points X[n*D], timed with CUDA events over `repeats` passes; they feed
 * the reference's automatic blocking (Model.get_param_blocking_for_sampler). */
int cb2_measure_speeds(cb2_engine *h, const double *X, int64_t n, int32_t repeats,
                       double *speeds_out);

/* R-1 of the confidence-interval bounds (mcmc.py:918-1002): per (virtual) chain the raw
 * weighted sample quantiles at limfrac and 1-limfrac of every sampled parameter (GetDist
 * `MCSamples.confidence`, mcmc.py:926-929 with limfrac = Rminus1_cl_level/2), summed into
 * out[1+4D] = { M, sum (low-shift)[D], sum (low-shift)^2[D], sum (up-shift)[D],
 * sum (up-shift)^2[D] } for the NCCL all-reduce.  mode/split/shift as in cb2_moments. */
int cb2_bounds(cb2_engine *h, int32_t mode, int32_t split, double limfrac,
               const double *shift, double *dev_out, double *host_out);

/* SampleCollection rows (collection.py:154-159,519-542) of one chain:
 * out[n*width], width = cb2_row_width. Returns number of rows copied (>=0). */
int64_t cb2_copy_rows(cb2_engine *h, int64_t chain, int64_t row_begin, int64_t n,
                      double *out);
/* Bulk transfer of sample rows.  The reference appends every accepted point to a host-side
 * SampleCollection (cobaya/collection.py:402-427,519-571, dumped by out_update :1287-1315);
 * the engine stores rows in HBM and hands them over in bulk: the rows
 * [first[c], n_rows[c]) of chains chain_begin <= c < chain_end (first == NULL: from row 0;
 * first is indexed from chain_begin) are compacted chain-major on the device and copied
 * with one transfer.  counts[c - chain_begin] (NULL ok) = rows of chain c.  out == NULL:
 * size query.  Returns the total number of rows (>= 0) or a negative error; fails if the
 * total exceeds max_rows. */
int64_t cb2_copy_rows_bulk(cb2_engine *h, int64_t chain_begin, int64_t chain_end,
                           const int64_t *first, int64_t *counts, double *out,
                           int64_t max_rows);
/* Asynchronous drain inside the run loop (the engine keeps a per-chain cursor: rows
 * before it have been handed over).  cb2_drain_start compacts the rows added since the
 * previous drain and enqueues ONE device-to-host copy of them (then of counts[n_chains],
 * NULL ok) on a copy stream, so that the following cb2_advance runs under the transfer;
 * it returns the number of rows being copied.  `out` and `counts` should be page-locked
 * (cb2_host_alloc) and must not be read before cb2_drain_wait returns; two drains may be
 * in flight (staging double buffer).  cb2_drain_reset sets the cursor (NULL: row 0). */
int64_t cb2_drain_start(cb2_engine *h, double *out, int64_t max_rows, int64_t *counts);
int cb2_drain_wait(cb2_engine *h);
int cb2_drain_reset(cb2_engine *h, const int64_t *first);
/* page-locked host memory for the transfers above (plain pointers; NULL on failure) */
void *cb2_host_alloc(int64_t bytes);
int cb2_host_free(void *p);
/* Re-layout of the row store for a larger per-chain capacity; stored rows are kept.  Call
 * it before a chain can run out of room: cb2_advance(n) adds at most n rows per chain. */
int cb2_grow_rows(cb2_engine *h, int64_t new_rows_cap);
/* device memory: free and total bytes of the engine's GPU, bytes held by the row store */
int cb2_mem_info(cb2_engine *h, int64_t *free_bytes, int64_t *total_bytes,
                 int64_t *rows_bytes);
int32_t cb2_row_width(const cb2_engine *h);
int32_t cb2_n_derived(const cb2_engine *h);

/* random_SO_N (cobaya/functions.py:21-60) for (local chain, block, epoch):
 * R[n_b*n_b] row-major -- test/diagnostic entry point. */
int cb2_debug_basis(cb2_engine *h, int64_t chain, int32_t block, uint32_t epoch,
                    double *R);

/* instrumentation: kernels launched so far; device-side timing of a region on the
 * engine's stream (CUDA events) */
int64_t cb2_launch_count(const cb2_engine *h);
int cb2_timer_start(cb2_engine *h);
int cb2_timer_stop(cb2_engine *h, float *ms);
/* per-kernel-class device time: when profiling is on every launch is bracketed by CUDA
 * events on the engine's stream; ms[k]/n[k] accumulate over launches of class
 * k = 0 cycler tapes, 1 Haar bases, 2 step kernel, 3 moments, 4 normals of the Haar bases
 * (stream2 when they are generated under the previous window's step kernel). */
int cb2_set_profiling(cb2_engine *h, int32_t on);
int cb2_kernel_times(cb2_engine *h, double ms[5], int64_t n[5], int32_t reset);
/* which step kernel the last cb2_advance used: 0 = general warp-per-chain,
 * 1 = DMMA (mma.sync.m8n8k4.f64) register-resident, 2 = DMMA producer/consumer,
 * 3 = streamed (batched DMMA / cuBLAS products of every proposal direction + accept chain) */
int cb2_last_step_kernel(const cb2_engine *h);
const char *cb2_debug_message(const cb2_engine *h); /* why a faster kernel was not used */
/* windows run by each step kernel since the last reset: out[0..3] indexed as
 * cb2_last_step_kernel; out[4] = windows whose producer/consumer launch was refused (ran on
 * the next DMMA kernel), out[5] = windows whose streamed buffers did not fit in device
 * memory (ran on the general kernel), out[6] = of out[2], the windows on the variant with
 * the products split over the SM sub-partitions (k_step_pc2), out[7] = windows that were run
 * chunk by chunk over the chains because the per-window buffers of all chains did not fit in
 * device memory (out[0..6] then count one per chunk).  Nothing changes kernels silently. */
int cb2_window_counts(cb2_engine *h, int64_t out[8], int32_t reset);
/* kernel experiments: cycle counters of one producer and one consumer warp of k_step_pc2 in
 * libraries built with -DCB2_PC2_TIMING (zeros otherwise) */
int cb2_debug_counters(cb2_engine *h, int64_t out[16], int32_t reset);
/* 0 auto, 1 force the general kernels, 2 fast kernels without the producer/consumer one;
 * +4: Householder sweep of the Haar bases with DFMA instead of the tensor pipe (n <= 64);
 * +8: producer/consumer kernel with one producer warp per tile (k_step_pc) where the
 *     split-product variant (k_step_pc2) would be chosen */
int cb2_set_kernel_policy(cb2_engine *h, int32_t policy);

#ifdef __cplusplus
}
#endif
#endif
