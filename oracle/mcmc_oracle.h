/*
 * oracle/mcmc_oracle.h -- TEST INFRASTRUCTURE ONLY (parity oracle).
 *
 * Plain-C, scalar, CPU restatement of the reference's adaptive Metropolis hot
 * path (cobaya/samplers/mcmc/mcmc.py, proposal.py, cobaya/functions.py,
 * cobaya/prior.py, cobaya/likelihoods/gaussian_mixture/gaussian_mixture.py,
 * cobaya/collection.py).  Every function in mcmc_oracle.c cites the reference
 * file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (cobaya_b200/) never does.
 *
 * Randomness: the reference draws from numpy's PCG64 Generator, whose
 * ziggurat/rejection samplers cannot be reproduced draw-for-draw on a GPU
 * ("parity unpinned" for the RNG stream itself).  The
 * oracle therefore consumes the SAME counter-based Philox4x32-10 streams as
 * the CUDA engine (layout in DESIGN.md section 4), and is pinned against the
 * unmodified reference by running the reference's own MCMC/BlockedProposer
 * classes with a Philox-backed random_state (oracle/make_golden.py ->
 * tests/golden/).
 */
#ifndef MCMC_ORACLE_H
#define MCMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BLOCKS 16

/* likelihood component kinds */
#define ORC_LIKE_GAUSSIAN_MIXTURE 0
#define ORC_LIKE_ROSENBROCK 1
#define ORC_LIKE_CONSTANT 2 /* likelihoods/one/one.py:26-28: logp = scale (0 for `one`) */
#define ORC_LIKE_EXTERNAL 3 /* LikelihoodExternalFunction (likelihood.py:150-255): a host
                               function of the inputs (the tests compile the SAME source the
                               engine hands to NVRTC as host code) */

typedef struct {
    int32_t kind;      /* ORC_LIKE_* */
    int32_t dim;       /* number of input params */
    int32_t n_modes;   /* gaussian_mixture only */
    int32_t derived;   /* gaussian_mixture: emit dim*n_modes whitened derived params */
    const int32_t *idx;   /* [dim] indices into the sampled vector */
    const double *means;  /* [n_modes*dim] */
    const double *linv;   /* [n_modes*dim*dim] row-major lower-tri inverse Cholesky */
    const double *logdet; /* [n_modes] log|Sigma_k| */
    const double *weights;/* [n_modes] (already normalised) */
    double scale;         /* rosenbrock: logp = -scale * sum(...) */
    double (*fn)(const double *p, int n); /* external: log L of the gathered inputs */
} orc_like;

/* scipy.stats distributions reached through pdf.logpdf (cobaya/prior.py:520-525);
 * numbering shared with include/cobaya_b200.h */
enum { ORC_PRIOR_UNIFORM = 0, ORC_PRIOR_NORMAL = 1, ORC_PRIOR_TRUNCNORM = 2,
       ORC_PRIOR_HALFNORM = 3, ORC_PRIOR_EXPON = 4, ORC_PRIOR_BETA = 5, ORC_PRIOR_GAMMA = 6,
       ORC_PRIOR_LOGNORM = 7, ORC_PRIOR_CAUCHY = 8, ORC_PRIOR_LAPLACE = 9,
       ORC_PRIOR_LOGUNIFORM = 10 };

typedef struct {
    int32_t D;
    /* prior (cobaya/prior.py:514-533,733-763) */
    const int32_t *prior_kind; /* [D] 0 uniform, 1 normal, 2.. scipy.stats kinds (ORC_PRIOR_*) */
    const double *lower;       /* [D] */
    const double *upper;       /* [D] */
    const double *loc;         /* [D] non-uniform kinds */
    const double *pscale;      /* [D] non-uniform kinds */
    const double *pa;          /* [D] scipy shape parameter a (kinds >= 2), may be NULL */
    const double *pb;          /* [D] scipy shape parameter b (kinds >= 2), may be NULL */
    const int32_t *periodic;   /* [D] 0/1 */
    double uniform_logp;       /* -sum(log(upper-lower)) over uniform params */
    /* likelihoods */
    int32_t n_like;
    const orc_like *likes;
    /* blocking (cobaya/samplers/mcmc/proposal.py:96-201) */
    int32_t n_blocks;
    int32_t block_size[ORC_MAX_BLOCKS];
    int32_t oversampling[ORC_MAX_BLOCKS];
    const int32_t *i_of_j;     /* [D] sorted index -> sampler index */
    /* dragging (mcmc.py:334-371): drag=0 -> plain Metropolis */
    int32_t drag;
    int32_t i_last_slow_block;
    int32_t drag_interp_steps;
    /* proposal */
    const double *T;           /* [D*D] row-major, sorted coords: sigma_j * L'_{jk} */
    double proposal_scale;
    /* options */
    double temperature;
    int64_t max_tries;         /* already in absolute units */
    int32_t output_thin;
    /* external priors (cobaya/prior.py:537-577,765-772): host functions of some sampled
     * parameters, evaluated where the internal prior is finite; each has its own
     * minuslogprior__<name> column after minuslogprior__0.  Metropolis only. */
    int32_t n_ext_prior;
    const orc_like *ext_priors; /* kind ORC_LIKE_EXTERNAL entries: idx, dim, fn */
} orc_model;

#define ORC_MAX_EXT_PRIORS 8

typedef struct orc_chain orc_chain;

/* row width: weight, minuslogpost, D sampled, n_derived, minuslogprior,
 * minuslogprior__0, n_ext_prior minuslogprior__x, chi2, n_like chi2__x
 * (collection.py:154-159) */
int32_t orc_row_width(const orc_model *m);
int32_t orc_n_derived(const orc_model *m);

/* log-posterior of one point. Returns logpost; fills logprior, loglikes[n_like]
 * (only if logprior is finite, else NaN) and derived[n_derived]. */
double orc_logpost(const orc_model *m, const double *x, double *logprior,
                   double *loglikes, double *derived);
/* the same with the prior components: pl[0] internal, pl[1..n_ext_prior] external (may be NULL) */
double orc_logpost_ex(const orc_model *m, const double *x, double *logprior,
                      double *loglikes, double *derived, double *pl);

orc_chain *orc_chain_new(const orc_model *m, uint64_t seed, uint64_t chain_id,
                         const double *x0, int64_t burn_in);
void orc_chain_free(orc_chain *c);
/* the proposal matrix of the model may be replaced between calls (learning) */
void orc_chain_set_model(orc_chain *c, const orc_model *m);

/* Advance n_proposals proposals. Stored rows appended to rows[(*n_rows)...].
 * Returns 0 ok, 1 = stuck (max_tries exceeded, mcmc.py:717-743), 2 = rows_cap hit. */
int orc_chain_advance(orc_chain *c, int64_t n_proposals, double *rows,
                      int64_t rows_cap, int64_t *n_rows);

/* read back state */
void orc_chain_get(const orc_chain *c, double *x, double *logpost, int64_t *weight,
                   int64_t *n_steps, int64_t *n_accepted, int64_t *burn_in_left);

/* unit-level entry points used by the golden-vector tests */
void orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t out[4]);
/* Haar SO(n) exactly as functions.py:21-60, normals from the BASIS stream */
void orc_random_SO_N(int32_t n, uint64_t seed, uint64_t chain_id, int32_t block,
                     uint32_t epoch, double *R /* [n*n] row-major */);
void orc_basis_normals(int32_t n, uint64_t seed, uint64_t chain_id, int32_t block,
                       uint32_t epoch, double *xx /* [(n+2)(n-1)/2 (+1)] */);
void orc_so_n_from_normals(int32_t n, double *xx, double *R);
void orc_permutation(int32_t len, uint64_t seed, uint64_t chain_id, int32_t which,
                     uint32_t cycle, const int32_t *sorted, int32_t *out);
/* radial + sign draws of one proposal (proposal.py:71-93) */
void orc_radial(int32_t n_block, uint64_t seed, uint64_t chain_id, uint64_t t,
                uint32_t sub, double *r, double *sign);
double orc_accept_exp(uint64_t seed, uint64_t chain_id, uint64_t t, uint32_t sub);

/* ensemble driver (OpenMP over chains) used by the cpu_baseline leg */
int orc_ensemble_advance(orc_chain **chains, int64_t n_chains, int64_t n_proposals,
                         double *rows, int64_t rows_cap, int64_t *n_rows,
                         int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
