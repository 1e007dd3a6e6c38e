/*
 * oracle/mcmc_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle).
 * See mcmc_oracle.h for the role of this file.  Reference citations are
 * relative to the reference tree (cobaya v3.6.2).
 */
#include "mcmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TAG_STEP 0u
#define TAG_ACCEPT 1u
#define TAG_BASIS 2u
#define TAG_CYCLER 3u
#define TAG_STEP2 4u

#define LOG_2PI 1.8378770664093453 /* log(2*pi) */

/* ---------------- Philox4x32-10 (Salmon et al. 2011), counter layout in DESIGN.md */
void orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t out[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox(uint64_t seed, uint32_t c0, uint32_t c1, uint64_t chain, uint32_t tag,
                   uint32_t out[4]) {
    orc_philox4x32((uint32_t)seed, (uint32_t)(seed >> 32), c0, c1, (uint32_t)chain, tag,
                   out);
}

/* 52-bit uniform strictly inside (0,1): (k + 0.5) * 2^-52 is exact in binary64 */
static double u52(uint32_t a, uint32_t b) {
    uint64_t k = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
    return ((double)k + 0.5) * 2.220446049250313e-16; /* 2^-52 */
}

/* ---------------- numpy Generator.permutation stand-in (proposal.py:54) ---------- */
void orc_permutation(int32_t len, uint64_t seed, uint64_t chain_id, int32_t which,
                     uint32_t cycle, const int32_t *sorted, int32_t *out) {
    for (int32_t i = 0; i < len; ++i) out[i] = sorted[i];
    for (int32_t i = len - 1; i >= 1; --i) {
        uint32_t w[4];
        philox(seed, (uint32_t)i, cycle, chain_id, TAG_CYCLER | ((uint32_t)which << 8), w);
        uint64_t r = ((uint64_t)w[0] << 32) | w[1];
        uint64_t j = (uint64_t)(((unsigned __int128)r * (uint64_t)(i + 1)) >> 64);
        int32_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
}

/* ---------------- standard normals for random_SO_N (functions.py:36) ------------ */
void orc_basis_normals(int32_t n, uint64_t seed, uint64_t chain_id, int32_t block,
                       uint32_t epoch, double *xx) {
    int32_t nn = (n + 2) * (n - 1) / 2;
    int32_t npairs = (nn + 1) / 2;
    for (int32_t p = 0; p < npairs; ++p) {
        uint32_t w[4];
        philox(seed, (uint32_t)p, epoch, chain_id, TAG_BASIS | ((uint32_t)block << 8), w);
        double u1 = u52(w[0], w[1]), u2 = u52(w[2], w[3]);
        double rad = sqrt(-2.0 * log(u1));
        double ang = 2.0 * M_PI * u2; /* engine uses sincospi(2*u2): same value to ~1ulp */
        (void)ang;
        double s, c;
        /* sincospi(2*u2) without relying on glibc extensions: reduce exactly */
        double t = 2.0 * u2; /* in (0,2) */
        /* exact octant reduction: t = q*0.5 + f, f in [-0.25,0.25] */
        double q = nearbyint(t * 2.0);
        double f = t - q * 0.5; /* exact */
        double sf = sin(M_PI * f), cf = cos(M_PI * f);
        int qi = ((int)q) & 3;
        switch (qi) {
            case 0: s = sf; c = cf; break;
            case 1: s = cf; c = -sf; break;
            case 2: s = -sf; c = -cf; break;
            default: s = -cf; c = sf; break;
        }
        xx[2 * p] = rad * c;
        xx[2 * p + 1] = rad * s; /* caller provides room for an odd trailing element */
    }
}

/* ---------------- _rvs (functions.py:44-60), statement by statement ------------- */
void orc_so_n_from_normals(int32_t dim, double *xx, double *H) {
    double *Dv = (double *)malloc(sizeof(double) * dim);
    double *tmp = (double *)malloc(sizeof(double) * dim);
    for (int i = 0; i < dim * dim; ++i) H[i] = 0.0;
    for (int i = 0; i < dim; ++i) H[i * dim + i] = 1.0; /* H = np.eye(dim), :34 */
    int ix = 0;
    for (int n = 0; n < dim - 1; ++n) {                 /* :48 */
        double *x = xx + ix;                            /* :49 (a view: xx is modified) */
        int len = dim - n;
        ix += len;                                      /* :50 */
        double norm2 = 0.0;
        for (int i = 0; i < len; ++i) norm2 += x[i] * x[i]; /* :51 */
        double x0 = x[0];                               /* :52 */
        Dv[n] = (x[0] != 0.0) ? (x[0] > 0 ? 1.0 : -1.0) : 1.0; /* :53 */
        x[0] += Dv[n] * sqrt(norm2);                    /* :54 */
        double sc = sqrt((norm2 - x0 * x0 + x[0] * x[0]) / 2.0); /* :55 */
        for (int i = 0; i < len; ++i) x[i] /= sc;
        for (int r = 0; r < dim; ++r) {                 /* :57 tmp = H[:, n:] @ x */
            double a = 0.0;
            for (int i = 0; i < len; ++i) a += H[r * dim + n + i] * x[i];
            tmp[r] = a;
        }
        for (int r = 0; r < dim; ++r)                   /* :58 H[:, n:] -= outer(tmp,x) */
            for (int i = 0; i < len; ++i) H[r * dim + n + i] -= tmp[r] * x[i];
    }
    double prod = 1.0;
    for (int n = 0; n < dim - 1; ++n) prod *= Dv[n];
    Dv[dim - 1] = (((dim - 1) & 1) ? -1.0 : 1.0) * prod; /* :59 */
    for (int r = 0; r < dim; ++r)                       /* :60 H = (D * H.T).T */
        for (int c = 0; c < dim; ++c) H[r * dim + c] *= Dv[r];
    free(Dv);
    free(tmp);
}

void orc_random_SO_N(int32_t n, uint64_t seed, uint64_t chain_id, int32_t block,
                     uint32_t epoch, double *R) {
    int32_t nn = (n + 2) * (n - 1) / 2;
    double *xx = (double *)malloc(sizeof(double) * (nn + 2));
    orc_basis_normals(n, seed, chain_id, block, epoch, xx);
    orc_so_n_from_normals(n, xx, R);
    free(xx);
}

/* ---------------- propose_r / RandProposer1D (proposal.py:71-93) ---------------- */
void orc_radial(int32_t n_block, uint64_t seed, uint64_t chain_id, uint64_t t,
                uint32_t sub, double *r, double *sign) {
    uint32_t w[4];
    philox(seed, (uint32_t)t, (uint32_t)(t >> 32), chain_id, TAG_STEP | (sub << 8), w);
    double u_mix = ((double)w[0] + 0.5) * 2.3283064365386963e-10; /* 2^-32 */
    double u_r = u52(w[2], w[3]);
    *sign = (w[1] & 1u) ? 1.0 : -1.0;                   /* :90 integers(2) */
    if (u_mix < 0.33) {                                 /* :79 */
        *r = -log(u_r);                                 /* :80 standard_exponential */
    } else if (n_block >= 2) {
        *r = sqrt(-2.0 * log(u_r));                     /* :82 sqrt(chisquare(2)) */
    } else {
        uint32_t w2[4];                                 /* :82 sqrt(chisquare(1)) = |z| */
        philox(seed, (uint32_t)t, (uint32_t)(t >> 32), chain_id, TAG_STEP2 | (sub << 8),
               w2);
        double u2 = u52(w2[0], w2[1]);
        double tt = 2.0 * u2;
        double q = nearbyint(tt * 2.0);
        double f = tt - q * 0.5;
        double sf = sin(M_PI * f), cf = cos(M_PI * f);
        double c;
        switch (((int)q) & 3) {
            case 0: c = cf; break;
            case 1: c = -sf; break;
            case 2: c = -cf; break;
            default: c = sf; break;
        }
        *r = fabs(sqrt(-2.0 * log(u_r)) * c);
    }
}

double orc_accept_exp(uint64_t seed, uint64_t chain_id, uint64_t t, uint32_t sub) {
    uint32_t w[4];
    philox(seed, (uint32_t)t, (uint32_t)(t >> 32), chain_id, TAG_ACCEPT | (sub << 8), w);
    return -log(u52(w[0], w[1]));                       /* mcmc.py:683 */
}

/* ---------------- model helpers ------------------------------------------------- */
int32_t orc_n_derived(const orc_model *m) {
    int32_t nd = 0;
    for (int l = 0; l < m->n_like; ++l)
        if (m->likes[l].kind == ORC_LIKE_GAUSSIAN_MIXTURE && m->likes[l].derived)
            nd += m->likes[l].dim * m->likes[l].n_modes;
    return nd;
}

int32_t orc_row_width(const orc_model *m) {
    return 2 + m->D + orc_n_derived(m) + 2 + m->n_ext_prior + 1 + m->n_like; /* collection.py:154-159 */
}

/* log(Phi(b) - Phi(a)) for a < b (scipy.stats truncnorm's normalisation, restated from
 * the published definition: the standard normal mass of [a, b], taken on the tail
 * side where the subtraction does not cancel). */
static double std_norm_cdf(double x) { return 0.5 * erfc(-x / sqrt(2.0)); }
static double log_gauss_mass(double a, double b) {
    if (b <= 0) return log(std_norm_cdf(b) - std_norm_cdf(a));
    if (a >= 0) return log(std_norm_cdf(-a) - std_norm_cdf(-b));
    return log1p(-std_norm_cdf(a) - std_norm_cdf(-b));
}

/* scipy.stats <dist>(loc, scale, a, b).logpdf(x) for the distributions of
 * ORC_PRIOR_*: rv_continuous.logpdf = _logpdf((x-loc)/scale, shapes) - log(scale).
 * Called inside the support only (bounds are checked first, prior.py:752). */
static double scipy_logpdf_1d(int kind, double x, double loc, double scale, double a, double b) {
    double z = (x - loc) / scale, ls = log(scale);
    switch (kind) {
        case ORC_PRIOR_TRUNCNORM: return -z * z / 2 - LOG_2PI / 2 - log_gauss_mass(a, b) - ls;
        case ORC_PRIOR_HALFNORM: return 0.5 * log(2.0 / M_PI) - z * z / 2 - ls;
        case ORC_PRIOR_EXPON: return -z - ls;
        case ORC_PRIOR_BETA: {
            double v = -(lgamma(a) + lgamma(b) - lgamma(a + b));
            if (a != 1.0) v += (a - 1.0) * log(z);
            if (b != 1.0) v += (b - 1.0) * log1p(-z);
            return v - ls;
        }
        case ORC_PRIOR_GAMMA: return (a != 1.0 ? (a - 1.0) * log(z) : 0.0) - z - lgamma(a) - ls;
        case ORC_PRIOR_LOGNORM: {
            if (!(z > 0)) return -INFINITY;
            double l = log(z);
            return -l * l / (2 * a * a) - log(a * z * sqrt(2 * M_PI)) - ls;
        }
        case ORC_PRIOR_CAUCHY: return -log(M_PI) - log1p(z * z) - ls;
        case ORC_PRIOR_LAPLACE: return log(0.5) - fabs(z) - ls;
        case ORC_PRIOR_LOGUNIFORM:
            if (!(z > 0)) return -INFINITY;
            return -log(z) - log(log(b) - log(a)) - ls;
        default: return 0.0;
    }
}

/* Prior.logps_internal (prior.py:733-763) with _fast_norm_logpdf (tools.py:720-729) */
static double logprior_internal(const orc_model *m, const double *x) {
    for (int i = 0; i < m->D; ++i)
        if (!(x[i] <= m->upper[i]) || !(x[i] >= m->lower[i])) return -INFINITY;
    double s = 0.0;
    for (int i = 0; i < m->D; ++i) {
        if (m->prior_kind[i] == 1) {
            double m_log_scale = -log(m->pscale[i]) - LOG_2PI / 2;
            double z = (x[i] - m->loc[i]) / m->pscale[i];
            s += m_log_scale - z * z / 2;
        } else if (m->prior_kind[i] >= 2) {
            s += scipy_logpdf_1d(m->prior_kind[i], x[i], m->loc[i], m->pscale[i],
                                 m->pa ? m->pa[i] : 0.0, m->pb ? m->pb[i] : 0.0);
        }
    }
    return m->uniform_logp + s;
}

/* GaussianMixture.logp (gaussian_mixture.py:138-163); Gaussian density in the
 * Cholesky form  -1/2 (d log 2pi + log|S| + |L^-1 (x-mu)|^2)   */
static double like_gaussian_mixture(const orc_like *L, const double *x, double *derived) {
    int d = L->dim, nm = L->n_modes;
    double lp[64];
    double *y = (double *)malloc(sizeof(double) * d);
    double *z = (double *)malloc(sizeof(double) * d);
    for (int k = 0; k < nm; ++k) {
        for (int i = 0; i < d; ++i) z[i] = x[L->idx[i]] - L->means[k * d + i];
        const double *Li = L->linv + (size_t)k * d * d;
        double q = 0.0;
        for (int i = 0; i < d; ++i) {
            double a = 0.0;
            for (int j = 0; j <= i; ++j) a += Li[i * d + j] * z[j]; /* :148 */
            y[i] = a;
            q += a * a;
            if (L->derived && derived) derived[k * d + i] = a;       /* :149-156 */
        }
        lp[k] = -0.5 * (d * LOG_2PI + L->logdet[k] + q);
    }
    free(y);
    free(z);
    if (nm == 1) return lp[0];                                       /* :158-159 */
    double mx = lp[0];                                               /* :161 logsumexp */
    for (int k = 1; k < nm; ++k) if (lp[k] > mx) mx = lp[k];
    if (mx == -INFINITY) return -INFINITY;
    double s = 0.0;
    for (int k = 0; k < nm; ++k) s += L->weights[k] * exp(lp[k] - mx);
    return log(s) + mx;
}

/* builder-defined Rosenbrock (the reference has none):
 * logp = -scale * sum_{i<d-1} [100 (x_{i+1}-x_i^2)^2 + (1-x_i)^2] */
static double like_rosenbrock(const orc_like *L, const double *x) {
    double s = 0.0;
    for (int i = 0; i + 1 < L->dim; ++i) {
        double a = x[L->idx[i]], b = x[L->idx[i + 1]];
        double t1 = b - a * a, t2 = 1.0 - a;
        s += 100.0 * t1 * t1 + t2 * t2;
    }
    return -L->scale * s;
}

/* Model.logposterior (model.py:579-678): prior first; -inf prior skips likelihoods */
double orc_logpost(const orc_model *m, const double *x, double *logprior,
                   double *loglikes, double *derived) {
    return orc_logpost_ex(m, x, logprior, loglikes, derived, NULL);
}

double orc_logpost_ex(const orc_model *m, const double *x, double *logprior,
                      double *loglikes, double *derived, double *pl) {
    for (int i = 0; i < m->D; ++i)
        if (!isfinite(x[i])) { /* model.py:623-630 non-finite -> -inf */
            *logprior = -INFINITY;
            for (int l = 0; l < m->n_like; ++l) loglikes[l] = NAN;
            return -INFINITY;
        }
    double lp = logprior_internal(m, x);
    if (pl) pl[0] = lp;
    if (lp != -INFINITY) {
        /* Prior.logps (prior.py:700-720): external priors after a finite internal one */
        double ext = 0.0;
        for (int k = 0; k < m->n_ext_prior; ++k) {
            const orc_like *P = &m->ext_priors[k];
            double p[256];
            for (int i = 0; i < P->dim && i < 256; ++i) p[i] = x[P->idx[i]];
            const double v = P->fn(p, P->dim);
            if (pl) pl[1 + k] = v;
            ext += v;
        }
        if (m->n_ext_prior) lp += ext;
    }
    *logprior = lp;
    if (lp == -INFINITY) {
        for (int l = 0; l < m->n_like; ++l) loglikes[l] = NAN;
        return -INFINITY;
    }
    double total = lp;
    int doff = 0;
    for (int l = 0; l < m->n_like; ++l) {
        const orc_like *L = &m->likes[l];
        double v;
        if (L->kind == ORC_LIKE_GAUSSIAN_MIXTURE) {
            v = like_gaussian_mixture(L, x, derived ? derived + doff : NULL);
            if (L->derived) doff += L->dim * L->n_modes;
        } else if (L->kind == ORC_LIKE_CONSTANT) {
            v = L->scale; /* one.logp_one (likelihoods/one/one.py:26-28) */
        } else if (L->kind == ORC_LIKE_EXTERNAL) {
            /* LikelihoodExternalFunction.logp (likelihood.py:226-255): the callable on the
             * likelihood's own input parameters */
            double p[256];
            for (int i = 0; i < L->dim && i < 256; ++i) p[i] = x[L->idx[i]];
            v = L->fn(p, L->dim);
        } else {
            v = like_rosenbrock(L, x);
        }
        loglikes[l] = v;
        total += v;
    }
    return total;
}

/* ---------------- per-chain state ------------------------------------------------ */
typedef struct {
    int32_t n;          /* length */
    int32_t loop_index; /* -1 initially (proposal.py:30) */
    int64_t cycle;      /* number of permutations drawn - 1 */
    int32_t *sorted;
    int32_t *indices;
    int32_t which;
} cycler_t;

typedef struct {
    int32_t n;
    int32_t loop_index; /* -1 */
    int64_t epoch;      /* number of bases drawn - 1 */
    double *R;          /* [n*n] */
} dirprop_t;

struct orc_chain {
    const orc_model *m;
    uint64_t seed, id;
    int32_t D, n_like, n_der, width;
    int32_t j_start[ORC_MAX_BLOCKS];
    double *x, *trial, *der, *trial_der, *loglikes, *trial_loglikes;
    double logpost, logprior;
    double pl[1 + ORC_MAX_EXT_PRIORS], trial_pl[1 + ORC_MAX_EXT_PRIORS]; /* prior components */
    int64_t weight, prior_rej, burn_in_left, added_weight;
    int64_t n_steps, n_accepted;
    cycler_t cyc_main, cyc_slow, cyc_fast;
    dirprop_t prop[ORC_MAX_BLOCKS];
    /* scratch for dragging */
    double *s_pt, *e_pt, *ps_pt, *pe_pt, *delta, *e_der, *pe_der, *e_ll, *pe_ll, *tmp_ll;
};

static void cycler_init(cycler_t *c, const int32_t *sorted, int32_t n, int32_t which) {
    c->n = n; c->loop_index = -1; c->cycle = -1; c->which = which;
    c->sorted = (int32_t *)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    c->indices = (int32_t *)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) c->sorted[i] = c->indices[i] = sorted[i];
}

/* CyclicIndexRandomizer.next (proposal.py:46-55) */
static int32_t cycler_next(cycler_t *c, uint64_t seed, uint64_t id) {
    c->loop_index = (c->loop_index + 1) % c->n;
    if (c->loop_index == 0 && c->n > 2) {
        c->cycle += 1;
        orc_permutation(c->n, seed, id, c->which, (uint32_t)c->cycle, c->sorted,
                        c->indices);
    }
    return c->indices[c->loop_index];
}

orc_chain *orc_chain_new(const orc_model *m, uint64_t seed, uint64_t chain_id,
                         const double *x0, int64_t burn_in) {
    orc_chain *c = (orc_chain *)calloc(1, sizeof(orc_chain));
    c->m = m; c->seed = seed; c->id = chain_id;
    c->D = m->D; c->n_like = m->n_like; c->n_der = orc_n_derived(m);
    c->width = orc_row_width(m);
    int D = m->D, nd = c->n_der > 0 ? c->n_der : 1, nl = m->n_like > 0 ? m->n_like : 1;
    c->x = (double *)calloc(D, sizeof(double));
    c->trial = (double *)calloc(D, sizeof(double));
    c->der = (double *)calloc(nd, sizeof(double));
    c->trial_der = (double *)calloc(nd, sizeof(double));
    c->loglikes = (double *)calloc(nl, sizeof(double));
    c->trial_loglikes = (double *)calloc(nl, sizeof(double));
    c->s_pt = (double *)calloc(D, sizeof(double));
    c->e_pt = (double *)calloc(D, sizeof(double));
    c->ps_pt = (double *)calloc(D, sizeof(double));
    c->pe_pt = (double *)calloc(D, sizeof(double));
    c->delta = (double *)calloc(D, sizeof(double));
    c->e_der = (double *)calloc(nd, sizeof(double));
    c->pe_der = (double *)calloc(nd, sizeof(double));
    c->e_ll = (double *)calloc(nl, sizeof(double));
    c->pe_ll = (double *)calloc(nl, sizeof(double));
    c->tmp_ll = (double *)calloc(nl, sizeof(double));
    memcpy(c->x, x0, sizeof(double) * D);
    c->logpost = orc_logpost_ex(m, c->x, &c->logprior, c->loglikes, c->der, c->pl);
    c->weight = 1;                              /* OneSamplePoint.add, collection.py:1355 */
    c->prior_rej = 0;
    c->burn_in_left = burn_in * m->output_thin + 1; /* mcmc.py:265 */
    c->added_weight = 0;
    /* BlockedProposer.__init__ (proposal.py:153-201) */
    int js = 0, total = 0;
    for (int b = 0; b < m->n_blocks; ++b) {
        c->j_start[b] = js;
        js += m->block_size[b];
        total += m->block_size[b] * m->oversampling[b];
        c->prop[b].n = m->block_size[b];
        c->prop[b].loop_index = -1;
        c->prop[b].epoch = -1;
        c->prop[b].R = (double *)calloc((size_t)m->block_size[b] * m->block_size[b],
                                        sizeof(double));
    }
    int32_t *rep = (int32_t *)malloc(sizeof(int32_t) * (total > 0 ? total : 1));
    int p = 0;
    for (int b = 0; b < m->n_blocks; ++b)       /* np.repeat(block_indices, o*n) :193 */
        for (int r = 0; r < m->block_size[b] * m->oversampling[b]; ++r) rep[p++] = b;
    cycler_init(&c->cyc_main, rep, total, 0);
    free(rep);
    int32_t *ibj = (int32_t *)malloc(sizeof(int32_t) * D);
    p = 0;
    for (int b = 0; b < m->n_blocks; ++b)
        for (int r = 0; r < m->block_size[b]; ++r) ibj[p++] = b; /* :198 */
    int last_slow = m->drag ? m->i_last_slow_block : m->n_blocks - 1; /* :138-140 */
    int n_slow = 0;
    for (int b = 0; b <= last_slow; ++b) n_slow += m->block_size[b];
    cycler_init(&c->cyc_slow, ibj, n_slow, 1);                   /* :199 */
    cycler_init(&c->cyc_fast, ibj + n_slow, D - n_slow, 2);      /* :200 */
    free(ibj);
    return c;
}

void orc_chain_set_model(orc_chain *c, const orc_model *m) { c->m = m; }

void orc_chain_free(orc_chain *c) {
    if (!c) return;
    free(c->x); free(c->trial); free(c->der); free(c->trial_der);
    free(c->loglikes); free(c->trial_loglikes);
    free(c->s_pt); free(c->e_pt); free(c->ps_pt); free(c->pe_pt); free(c->delta);
    free(c->e_der); free(c->pe_der); free(c->e_ll); free(c->pe_ll); free(c->tmp_ll);
    free(c->cyc_main.sorted); free(c->cyc_main.indices);
    free(c->cyc_slow.sorted); free(c->cyc_slow.indices);
    free(c->cyc_fast.sorted); free(c->cyc_fast.indices);
    for (int b = 0; b < c->m->n_blocks; ++b) free(c->prop[b].R);
    free(c);
}

/* BlockedProposer.get_block_proposal (proposal.py:222-224) applied to P */
static void block_proposal(orc_chain *c, double *P, int b, uint32_t sub) {
    const orc_model *m = c->m;
    dirprop_t *bp = &c->prop[b];
    int n = bp->n, D = c->D, j0 = c->j_start[b];
    double r, sign;
    double vec[512];
    double *v = (n <= 512) ? vec : (double *)malloc(sizeof(double) * n);
    if (n > 1) {                                 /* RandDirectionProposer.propose_vec :59 */
        bp->loop_index = (bp->loop_index + 1) % n;
        if (bp->loop_index == 0) {
            bp->epoch += 1;
            orc_random_SO_N(n, c->seed, c->id, b, (uint32_t)bp->epoch, bp->R); /* :68 */
        }
        orc_radial(n, c->seed, c->id, (uint64_t)c->n_steps, sub, &r, &sign);
        for (int i = 0; i < n; ++i)              /* :69 R[:,k] * r * scale */
            v[i] = bp->R[i * n + bp->loop_index] * r * m->proposal_scale;
    } else {                                     /* RandProposer1D.propose_vec :86-93 */
        orc_radial(1, c->seed, c->id, (uint64_t)c->n_steps, sub, &r, &sign);
        v[0] = (sign > 0) ? r * m->proposal_scale : -r * m->proposal_scale;
    }
    /* P[par_blocks[b]] += transform[b].dot(vec): rows j>=j0 of T[:, j0:j0+n] */
    for (int j = j0; j < D; ++j) {
        double a = 0.0;
        int kmax = (j - j0 + 1 < n) ? (j - j0 + 1) : n; /* lower-triangular zeros skipped */
        for (int k = 0; k < kmax; ++k) a += m->T[(size_t)j * D + j0 + k] * v[k];
        P[m->i_of_j[j]] += a;
    }
    if (v != vec) free(v);
}

/* Prior.reduce_periodic (prior.py:658-676) */
static void reduce_periodic(const orc_model *m, double *x) {
    for (int i = 0; i < m->D; ++i)
        if (m->periodic[i]) {
            double a = m->lower[i], b = m->upper[i];
            double q = (x[i] - a) / (b - a);
            q = q - floor(q); /* python float % 1 */
            x[i] = q * (b - a) + a;
        }
}

/* MCMC.metropolis_accept (mcmc.py:670-683) */
static int metropolis_accept(orc_chain *c, double lt, double lc, uint32_t sub) {
    if (lt == -INFINITY) return 0;
    if (lt > lc) return 1;
    double ratio = (lc - lt) / c->m->temperature;
    return orc_accept_exp(c->seed, c->id, (uint64_t)c->n_steps, sub) > ratio;
}

/* collection.py:519-542 row layout */
static void write_row(const orc_chain *c, double *row, double weight) {
    const orc_model *m = c->m;
    int k = 0;
    row[k++] = weight;
    row[k++] = -(c->logpost / m->temperature);
    for (int i = 0; i < c->D; ++i) row[k++] = c->x[i];
    for (int i = 0; i < c->n_der; ++i) row[k++] = c->der[i];
    row[k++] = -c->logprior;
    row[k++] = m->n_ext_prior ? -c->pl[0] : -c->logprior; /* minuslogprior__0 */
    for (int e = 0; e < m->n_ext_prior; ++e) row[k++] = -c->pl[1 + e];
    double ll = 0.0;
    for (int l = 0; l < c->n_like; ++l) ll += c->loglikes[l];
    row[k++] = -2 * ll;
    for (int l = 0; l < c->n_like; ++l) row[k++] = -2 * c->loglikes[l];
}

/* MCMC.process_accept_or_reject (mcmc.py:685-748) +
 * OneSamplePoint.add_to_collection (collection.py:1366-1383).
 * returns 0 ok, 1 stuck, 2 rows full */
static int process(orc_chain *c, int accept, const double *trial, double t_logpost,
                   double t_logprior, const double *t_ll, const double *t_der,
                   double *rows, int64_t cap, int64_t *n_rows) {
    const orc_model *m = c->m;
    if (accept) {
        if (c->burn_in_left <= 0) {
            int64_t w;
            int store = 1;
            if (m->output_thin > 1) {
                c->added_weight += c->weight;
                if (c->added_weight >= m->output_thin) {
                    w = c->added_weight / m->output_thin;
                    c->added_weight %= m->output_thin;
                } else { store = 0; w = 0; }
            } else w = c->weight;
            if (store) {
                if (rows) { /* rows == NULL: count only (cpu_baseline timing leg) */
                    if (*n_rows >= cap) return 2;
                    write_row(c, rows + (size_t)(*n_rows) * c->width, (double)w);
                }
                *n_rows += 1;
            }
        } else c->burn_in_left -= 1;
        memcpy(c->x, trial, sizeof(double) * c->D);
        c->logpost = t_logpost; c->logprior = t_logprior;
        if (m->n_ext_prior) memcpy(c->pl, c->trial_pl, sizeof(c->pl));
        memcpy(c->loglikes, t_ll, sizeof(double) * c->n_like);
        if (c->n_der) memcpy(c->der, t_der, sizeof(double) * c->n_der);
        c->weight = 1;
        c->prior_rej = 0;
        c->n_accepted += 1;
    } else {
        c->weight += 1;
        if (t_logprior == -INFINITY) c->prior_rej += 1;
        int64_t sgn = (c->burn_in_left > 0) - (c->burn_in_left < 0);
        int64_t max_now = m->max_tries * (1 + 9 * sgn);
        if (c->weight - c->prior_rej > max_now) return 1;
    }
    return 0;
}

/* MCMC.get_new_sample_metropolis (mcmc.py:545-562) */
static int step_metropolis(orc_chain *c, double *rows, int64_t cap, int64_t *n_rows) {
    const orc_model *m = c->m;
    memcpy(c->trial, c->x, sizeof(double) * c->D);                 /* :556 */
    int b = cycler_next(&c->cyc_main, c->seed, c->id);             /* proposal.py:207 */
    block_proposal(c, c->trial, b, 0);                             /* :557 */
    reduce_periodic(m, c->trial);                                  /* :558 */
    double lp;
    double lpost = orc_logpost_ex(m, c->trial, &lp, c->trial_loglikes, c->trial_der,
                                  c->trial_pl);
    int acc = metropolis_accept(c, lpost, c->logpost, 0);          /* :560 */
    return process(c, acc, c->trial, lpost, lp, c->trial_loglikes, c->trial_der, rows,
                   cap, n_rows);
}

/* MCMC.get_new_sample_dragging (mcmc.py:564-668) */
static int step_dragging(orc_chain *c, double *rows, int64_t cap, int64_t *n_rows) {
    const orc_model *m = c->m;
    int D = c->D, nd = c->n_der, nl = c->n_like;
    int nds = m->drag_interp_steps;
    memcpy(c->s_pt, c->x, sizeof(double) * D);                     /* :579 */
    double s_lp = c->logpost;                                      /* :580 */
    memcpy(c->e_pt, c->x, sizeof(double) * D);                     /* :581 */
    int b = cycler_next(&c->cyc_slow, c->seed, c->id);             /* proposal.py:216 */
    block_proposal(c, c->e_pt, b, 0);                              /* :582 */
    reduce_periodic(m, c->e_pt);                                   /* :583 */
    double e_prior;
    double e_lp = orc_logpost(m, c->e_pt, &e_prior, c->e_ll, c->e_der); /* :589 */
    if (e_lp == -INFINITY) { c->weight += 1; return 0; }           /* :590-592 */
    double s_acc = s_lp, e_acc = e_lp;                             /* :595-596 */
    for (int i = 1; i <= nds; ++i) {                               /* :603 */
        for (int k = 0; k < D; ++k) c->delta[k] = 0.0;             /* :606 */
        int bf = cycler_next(&c->cyc_fast, c->seed, c->id);        /* proposal.py:220 */
        block_proposal(c, c->delta, bf, (uint32_t)i);              /* :607 */
        reduce_periodic(m, c->delta);                              /* :608 (quirk) */
        for (int k = 0; k < D; ++k) c->ps_pt[k] = c->s_pt[k] + c->delta[k]; /* :610 */
        double ps_prior;
        double ps_lp = orc_logpost(m, c->ps_pt, &ps_prior, c->tmp_ll, NULL); /* :616 */
        if (ps_lp != -INFINITY) {
            for (int k = 0; k < D; ++k) c->pe_pt[k] = c->e_pt[k] + c->delta[k]; /* :622 */
            double pe_prior;
            double pe_lp = orc_logpost(m, c->pe_pt, &pe_prior, c->pe_ll, c->pe_der);
            if (pe_lp != -INFINITY) {
                double frac = (double)i / (double)(1 + nds);       /* :630 */
                double p_int = (1 - frac) * ps_lp + frac * pe_lp;  /* :631-633 */
                double c_int = (1 - frac) * s_lp + frac * e_lp;    /* :634-636 */
                if (metropolis_accept(c, p_int, c_int, (uint32_t)i)) { /* :637 */
                    memcpy(c->s_pt, c->ps_pt, sizeof(double) * D); /* :642-645 */
                    s_lp = ps_lp;
                    memcpy(c->e_pt, c->pe_pt, sizeof(double) * D);
                    e_lp = pe_lp; e_prior = pe_prior;
                    memcpy(c->e_ll, c->pe_ll, sizeof(double) * nl);
                    if (nd) memcpy(c->e_der, c->pe_der, sizeof(double) * nd);
                }
            }
        }
        s_acc += s_lp;                                             /* :655-656 */
        e_acc += e_lp;
    }
    double navg = 1 + nds;                                         /* :658 */
    int acc = metropolis_accept(c, e_acc / navg, s_acc / navg, 0); /* :659-661 */
    return process(c, acc, c->e_pt, e_lp, e_prior, c->e_ll, c->e_der, rows, cap, n_rows);
}

int orc_chain_advance(orc_chain *c, int64_t n_proposals, double *rows, int64_t rows_cap,
                      int64_t *n_rows) {
    for (int64_t s = 0; s < n_proposals; ++s) {
        int rc = c->m->drag ? step_dragging(c, rows, rows_cap, n_rows)
                            : step_metropolis(c, rows, rows_cap, n_rows);
        c->n_steps += 1;                                           /* mcmc.py:472 */
        if (rc) return rc;
    }
    return 0;
}

void orc_chain_get(const orc_chain *c, double *x, double *logpost, int64_t *weight,
                   int64_t *n_steps, int64_t *n_accepted, int64_t *burn_in_left) {
    if (x) memcpy(x, c->x, sizeof(double) * c->D);
    if (logpost) *logpost = c->logpost;
    if (weight) *weight = c->weight;
    if (n_steps) *n_steps = c->n_steps;
    if (n_accepted) *n_accepted = c->n_accepted;
    if (burn_in_left) *burn_in_left = c->burn_in_left;
}

int orc_ensemble_advance(orc_chain **chains, int64_t n_chains, int64_t n_proposals,
                         double *rows, int64_t rows_cap, int64_t *n_rows,
                         int32_t n_threads) {
    int rc_all = 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t i = 0; i < n_chains; ++i) {
        orc_chain *c = chains[i];
        int rc = orc_chain_advance(c, n_proposals,
                                   rows ? rows + (size_t)i * rows_cap * c->width : NULL,
                                   rows ? rows_cap : 0, &n_rows[i]);
        if (rc == 1) {
#ifdef _OPENMP
#pragma omp critical
#endif
            rc_all = 1;
        }
    }
    return rc_all;
}
