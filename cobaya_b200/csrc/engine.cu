// engine.cu -- host side of libcobaya_b200.so: the C ABI declared in
// include/cobaya_b200.h, device memory management, and the launch plan that turns
// "advance every chain by n proposals" (MCMC.run, cobaya/samplers/mcmc/mcmc.py:470-472)
// into windows of {cycler tapes, Haar bases, step kernel}.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels_general.cuh"
#include "kernels_moments.cuh"
#include "kernels_fast.cuh"

#define CB2_ABI_VERSION 1

static thread_local std::string g_create_error;

struct LikeHost {
    LikeDev d;
    std::vector<int32_t> idx;
    std::vector<double> means, linvT, c0, w;
};

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct cb2_engine {
    int device = 0;
    int64_t n_chains = 0;
    int32_t D = 0;
    uint64_t seed = 0, chain_id0 = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;
    int sm_count = 148;
    // ---- host model
    bool have_shapes = false;
    bool have_prior = false, have_blocking = false, have_proposal = false, have_state = false;
    std::vector<int32_t> prior_kind, periodic;
    std::vector<double> lower, upper, loc, pscale, pa, pb, pcn;
    double uniform_logp = 0.0;
    std::vector<LikeHost> likes;
    int32_t n_der = 0;
    int32_t n_blocks = 0;
    int32_t bsize[CB2_MAX_BLOCKS] = {0}, oversamp[CB2_MAX_BLOCKS] = {0},
            jstart[CB2_MAX_BLOCKS] = {0};
    std::vector<int32_t> i_of_j;
    int32_t drag = 0, last_slow = 0, drag_steps = 0, n_slow = 0, n_fast = 0;
    std::vector<double> TT, Trow;
    double proposal_scale = 2.4;
    double temperature = 1.0;
    int64_t burn_in = 0, max_tries = (int64_t)1 << 62, rows_cap = 0;
    int32_t output_thin = 1;
    bool model_dirty = true;
    // ---- device model
    DevBuf<int32_t> d_prior_kind, d_periodic, d_ipool, d_i_of_j;
    DevBuf<double> d_lower, d_upper, d_loc, d_pscale, d_pa, d_pb, d_pcn, d_dpool, d_TT;
    DevBuf<double> d_fastpack;  // fragment-ordered matrices for the DMMA kernel
    FastPackDesc fast_desc;
    bool fast_ready = false;
    ModelDev M;
    // ---- chain state
    DevBuf<double> d_x, d_logpost, d_logprior, d_ll, d_der, d_rows;
    DevBuf<int64_t> d_weight, d_prior_rej, d_burn_left, d_added_w, d_n_rows, d_n_acc, d_vis;
    DevBuf<uint32_t> d_flags;
    ChainState S;
    int64_t steps_done = 0;
    // ---- window resources
    std::vector<uint8_t> ms_main, ms_slow, ms_fast;  // sorted multisets
    DevBuf<uint8_t> d_ms_main, d_ms_slow, d_ms_fast, d_tape_main, d_tape_slow, d_tape_fast,
        d_perm_scratch;
    DevBuf<double> d_basis[CB2_MAX_BLOCKS];
    DevBuf<double> d_basis_scratch;
    DevBuf<double2> d_draws;
    DevBuf<int2> d_plan;
    // ---- moments
    DevBuf<MomentTask> d_tasks;
    DevBuf<double> d_means, d_sw, d_partials, d_mom_out, d_shift, d_bounds;
    DevBuf<int64_t> d_summary;
    // ---- misc
    DevBuf<double> d_tmp;
    int64_t launches = 0;
    int last_kernel = 0, policy = 0;
    char pc_error[256] = {0};
    // ---- per-kernel-class device timing (CUDA events on the launching stream)
    struct ProfRec { cudaEvent_t a, b; int kind; };
    bool profiling = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_n[4] = {0, 0, 0, 0};
    cudaEvent_t take_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void prof_begin(int kind) {
        if (!profiling) return;
        ProfRec r; r.a = take_event(); r.b = take_event(); r.kind = kind;
        cudaEventRecord(r.a, stream);
        prof.push_back(r);
    }
    void prof_end() {
        if (!profiling || prof.empty()) return;
        cudaEventRecord(prof.back().b, stream);
    }
};
enum { PROF_TAPE = 0, PROF_BASIS = 1, PROF_STEP = 2, PROF_MOMENTS = 3 };

#define FAIL(h, code, ...)                                 \
    do {                                                   \
        char _b[512];                                      \
        snprintf(_b, sizeof(_b), __VA_ARGS__);             \
        (h)->err = _b;                                     \
        return (code);                                     \
    } while (0)

#define CK(h, call)                                                                   \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess)                                                        \
            FAIL(h, -2, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, \
                 __LINE__, #call);                                                    \
    } while (0)

template <typename T>
static int upload(cb2_engine *h, DevBuf<T> &b, const std::vector<T> &v) {
    CK(h, b.ensure(v.size()));
    if (!v.empty())
        CK(h, cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice,
                              h->stream));
    return 0;
}

extern "C" int cb2_abi_version(void) { return CB2_ABI_VERSION; }

extern "C" const char *cb2_last_error(const cb2_engine *h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int cb2_create(int device, int64_t n_chains, int32_t D, uint64_t seed,
                          uint64_t chain_id0, cb2_engine **out) {
    if (!out) return -1;
    *out = nullptr;
    if (n_chains <= 0 || D <= 0) {
        g_create_error = "cb2_create: n_chains and D must be positive";
        return -1;
    }
    if (chain_id0 + (uint64_t)n_chains > 0xFFFFFFFFull) {
        g_create_error = "cb2_create: global chain ids must fit in 32 bits";
        return -1;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("cb2_create: no CUDA device available (") +
                         cudaGetErrorString(e) + "); the engine has no CPU fallback";
        return -3;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "cb2_create: invalid device index";
        return -1;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return -2;
    }
    cb2_engine *h = new cb2_engine();
    h->device = device;
    h->n_chains = n_chains;
    h->D = D;
    h->seed = seed;
    h->chain_id0 = chain_id0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
        g_create_error = "cb2_create: could not create stream/events";
        delete h;
        return -2;
    }
    {  // logarithm table of the basis kernels (common.cuh log_tab), once per engine/device
        double tab[CB2_LOGTAB_DOUBLES];
        for (int j = 0; j < CB2_LOGTAB_N; ++j) {
            const long double c = 0.75L + (long double)j / 128.0L;
            tab[3 * j] = (double)c;
            tab[3 * j + 1] = (double)(1.0L / c);
            tab[3 * j + 2] = (j == 32) ? 0.0 : (double)logl(c);
        }
        if (cudaMemcpyToSymbol(g_logtab, tab, sizeof(tab)) != cudaSuccess) {
            g_create_error = "cb2_create: could not upload the logarithm table";
            delete h;
            return -2;
        }
    }
    // default blocking: one block with every parameter
    h->n_blocks = 1;
    h->bsize[0] = D;
    h->oversamp[0] = 1;
    h->jstart[0] = 0;
    h->i_of_j.resize(D);
    for (int i = 0; i < D; ++i) h->i_of_j[i] = i;
    h->n_slow = D;
    h->n_fast = 0;
    h->have_blocking = true;
    *out = h;
    return 0;
}

extern "C" int cb2_destroy(cb2_engine *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    h->d_prior_kind.release(); h->d_periodic.release(); h->d_ipool.release();
    h->d_i_of_j.release(); h->d_lower.release(); h->d_upper.release(); h->d_loc.release();
    h->d_pscale.release(); h->d_pa.release(); h->d_pb.release(); h->d_pcn.release();
    h->d_dpool.release(); h->d_TT.release(); h->d_fastpack.release();
    h->d_x.release(); h->d_logpost.release(); h->d_logprior.release(); h->d_ll.release();
    h->d_der.release(); h->d_rows.release(); h->d_weight.release(); h->d_prior_rej.release();
    h->d_burn_left.release(); h->d_added_w.release(); h->d_n_rows.release();
    h->d_n_acc.release(); h->d_vis.release(); h->d_flags.release();
    h->d_ms_main.release(); h->d_ms_slow.release(); h->d_ms_fast.release();
    h->d_tape_main.release(); h->d_tape_slow.release(); h->d_tape_fast.release();
    h->d_perm_scratch.release();
    for (int b = 0; b < CB2_MAX_BLOCKS; ++b) h->d_basis[b].release();
    h->d_basis_scratch.release(); h->d_draws.release(); h->d_plan.release(); h->d_tasks.release(); h->d_means.release();
    h->d_sw.release(); h->d_bounds.release(); h->d_partials.release(); h->d_mom_out.release(); h->d_shift.release();
    h->d_summary.release(); h->d_tmp.release();
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

extern "C" int cb2_set_prior(cb2_engine *h, const int32_t *kind, const double *lower,
                             const double *upper, const double *loc, const double *scale,
                             const int32_t *periodic, double uniform_logp) {
    if (!h) return -1;
    const int D = h->D;
    for (int i = 0; i < D; ++i) {
        if (kind[i] < 0 || kind[i] > CB2_PRIOR_LOGUNIFORM)
            FAIL(h, -1, "cb2_set_prior: unknown prior kind %d", kind[i]);
        if (kind[i] != 0 && !(scale[i] > 0)) FAIL(h, -1, "cb2_set_prior: prior scale must be > 0");
        if (periodic[i] && !(std::isfinite(lower[i]) && std::isfinite(upper[i])))
            FAIL(h, -1, "cb2_set_prior: periodic parameter %d is not bounded", i);
    }
    h->prior_kind.assign(kind, kind + D);
    h->lower.assign(lower, lower + D);
    h->upper.assign(upper, upper + D);
    h->loc.assign(loc, loc + D);
    h->pscale.assign(scale, scale + D);
    h->periodic.assign(periodic, periodic + D);
    h->uniform_logp = uniform_logp;
    h->pa.assign(D, 0.0); h->pb.assign(D, 0.0); h->pcn.assign(D, 0.0);
    h->have_prior = true;
    h->have_shapes = false;
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_set_prior_shapes(cb2_engine *h, const double *a, const double *b,
                                    const double *log_norm) {
    if (!h) return -1;
    if (!h->have_prior) FAIL(h, -1, "cb2_set_prior_shapes: call cb2_set_prior first");
    const int D = h->D;
    for (int i = 0; i < D; ++i)
        if (h->prior_kind[i] >= 2 && !std::isfinite(log_norm[i]))
            FAIL(h, -1, "cb2_set_prior_shapes: non-finite normalisation for parameter %d", i);
    h->pa.assign(a, a + D);
    h->pb.assign(b, b + D);
    h->pcn.assign(log_norm, log_norm + D);
    h->have_shapes = true;
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_clear_likelihoods(cb2_engine *h) {
    if (!h) return -1;
    h->likes.clear();
    h->n_der = 0;
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_add_gaussian_mixture(cb2_engine *h, int32_t dim, const int32_t *idx,
                                        int32_t n_modes, const double *means,
                                        const double *linv, const double *logdet,
                                        const double *weights, int32_t derived) {
    if (!h) return -1;
    if ((int)h->likes.size() >= CB2_MAX_LIKES) FAIL(h, -1, "too many likelihoods (max %d)", CB2_MAX_LIKES);
    if (dim <= 0 || dim > h->D) FAIL(h, -1, "gaussian_mixture: bad dimension %d", dim);
    if (n_modes <= 0 || n_modes > CB2_MAX_MODES) FAIL(h, -1, "gaussian_mixture: 1..%d modes supported", CB2_MAX_MODES);
    LikeHost L;
    L.d.kind = 0;
    L.d.dim = dim;
    L.d.n_modes = n_modes;
    L.d.derived = derived ? 1 : 0;
    L.d.scale = 0.0;
    L.idx.assign(idx, idx + dim);
    for (int i = 0; i < dim; ++i)
        if (idx[i] < 0 || idx[i] >= h->D) FAIL(h, -1, "gaussian_mixture: parameter index out of range");
    L.means.assign(means, means + (size_t)n_modes * dim);
    L.linvT.resize((size_t)n_modes * dim * dim);
    for (int k = 0; k < n_modes; ++k)
        for (int i = 0; i < dim; ++i)
            for (int j = 0; j < dim; ++j)
                L.linvT[(size_t)k * dim * dim + (size_t)j * dim + i] =
                    (j <= i) ? linv[(size_t)k * dim * dim + (size_t)i * dim + j] : 0.0;
    L.c0.resize(n_modes);
    for (int k = 0; k < n_modes; ++k) L.c0[k] = dim * CB2_LOG_2PI + logdet[k];
    L.w.assign(weights, weights + n_modes);
    L.d.der_off = h->n_der;
    if (derived) h->n_der += dim * n_modes;
    h->likes.push_back(L);
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_add_rosenbrock(cb2_engine *h, int32_t dim, const int32_t *idx, double scale) {
    if (!h) return -1;
    if ((int)h->likes.size() >= CB2_MAX_LIKES) FAIL(h, -1, "too many likelihoods (max %d)", CB2_MAX_LIKES);
    if (dim < 2 || dim > h->D) FAIL(h, -1, "rosenbrock: bad dimension %d", dim);
    LikeHost L;
    L.d.kind = 1;
    L.d.dim = dim;
    L.d.n_modes = 0;
    L.d.derived = 0;
    L.d.scale = scale;
    L.d.der_off = h->n_der;
    L.idx.assign(idx, idx + dim);
    for (int i = 0; i < dim; ++i)
        if (idx[i] < 0 || idx[i] >= h->D) FAIL(h, -1, "rosenbrock: parameter index out of range");
    h->likes.push_back(L);
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_add_constant(cb2_engine *h, double value) {
    if (!h) return -1;
    if ((int)h->likes.size() >= CB2_MAX_LIKES) FAIL(h, -1, "too many likelihoods (max %d)", CB2_MAX_LIKES);
    if (!std::isfinite(value)) FAIL(h, -1, "constant likelihood: value must be finite");
    LikeHost L;
    L.d.kind = 2;
    L.d.dim = 0;
    L.d.n_modes = 0;
    L.d.derived = 0;
    L.d.scale = value;
    L.d.der_off = h->n_der;
    h->likes.push_back(L);
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_set_blocking(cb2_engine *h, int32_t n_blocks, const int32_t *block_sizes,
                                const int32_t *oversampling, const int32_t *i_of_j,
                                int32_t drag, int32_t i_last_slow_block,
                                int32_t drag_interp_steps) {
    if (!h) return -1;
    if (h->have_state && h->steps_done > 0)
        FAIL(h, -1, "cb2_set_blocking: blocking cannot change after sampling started");
    if (n_blocks < 1 || n_blocks > CB2_MAX_BLOCKS) FAIL(h, -1, "1..%d blocks supported", CB2_MAX_BLOCKS);
    int tot = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (block_sizes[b] < 1) FAIL(h, -1, "empty parameter block");
        if (oversampling[b] < 1) FAIL(h, -1, "Oversampling factors must be integer >= 1");
        tot += block_sizes[b];
    }
    if (tot != h->D) FAIL(h, -1, "The blocks do not contain all the parameter indices.");
    std::vector<int> seen(h->D, 0);
    for (int j = 0; j < h->D; ++j) {
        if (i_of_j[j] < 0 || i_of_j[j] >= h->D || seen[i_of_j[j]]++)
            FAIL(h, -1, "The blocks do not contain all the parameter indices.");
    }
    if (drag) {
        if (n_blocks < 2 || i_last_slow_block < 0 || i_last_slow_block >= n_blocks - 1)
            FAIL(h, -1, "dragging needs at least one slow and one fast block");
        if (drag_interp_steps < 1) FAIL(h, -1, "drag_interp_steps must be >= 1");
    }
    h->n_blocks = n_blocks;
    int js = 0;
    for (int b = 0; b < n_blocks; ++b) {
        h->bsize[b] = block_sizes[b];
        h->oversamp[b] = oversampling[b];
        h->jstart[b] = js;
        js += block_sizes[b];
    }
    h->i_of_j.assign(i_of_j, i_of_j + h->D);
    h->drag = drag ? 1 : 0;
    h->last_slow = drag ? i_last_slow_block : n_blocks - 1;
    h->drag_steps = drag ? drag_interp_steps : 0;
    h->n_slow = 0;
    for (int b = 0; b <= h->last_slow; ++b) h->n_slow += h->bsize[b];
    h->n_fast = h->D - h->n_slow;
    // cycler multisets (proposal.py:193-201)
    h->ms_main.clear(); h->ms_slow.clear(); h->ms_fast.clear();
    for (int b = 0; b < n_blocks; ++b)
        for (int r = 0; r < h->bsize[b] * h->oversamp[b]; ++r) h->ms_main.push_back((uint8_t)b);
    for (int b = 0; b < n_blocks; ++b)
        for (int r = 0; r < h->bsize[b]; ++r)
            (b <= h->last_slow ? h->ms_slow : h->ms_fast).push_back((uint8_t)b);
    if (h->ms_main.size() > 60000) FAIL(h, -1, "cycle length too large");
    h->have_blocking = true;
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_set_proposal(cb2_engine *h, const double *T, double proposal_scale) {
    if (!h) return -1;
    const int D = h->D;
    h->TT.resize((size_t)D * D);
    h->Trow.assign(T, T + (size_t)D * D);
    for (int j = 0; j < D; ++j)
        for (int k = 0; k < D; ++k) {
            double v = T[(size_t)j * D + k];
            if (!std::isfinite(v)) FAIL(h, -1, "cb2_set_proposal: non-finite transform");
            h->TT[(size_t)k * D + j] = (k <= j) ? v : 0.0;
        }
    h->proposal_scale = proposal_scale;
    h->have_proposal = true;
    h->model_dirty = true;
    return 0;
}

extern "C" int cb2_set_options(cb2_engine *h, double temperature, int64_t burn_in,
                               int64_t max_tries, int32_t output_thin, int64_t rows_cap) {
    if (!h) return -1;
    if (!(temperature > 0)) FAIL(h, -1, "temperature must be positive");
    if (output_thin < 1) FAIL(h, -1, "output_thin must be >= 1");
    if (rows_cap < 1) FAIL(h, -1, "rows_cap must be >= 1");
    if (h->have_state && rows_cap != h->rows_cap)
        FAIL(h, -1, "rows_cap cannot change after cb2_set_state");
    h->temperature = temperature;
    h->burn_in = burn_in;
    h->max_tries = std::min<int64_t>(max_tries, (int64_t)1 << 59);  // x10 in burn-in fits int64
    h->output_thin = output_thin;
    h->rows_cap = rows_cap;
    h->model_dirty = true;
    return 0;
}

static int row_width(const cb2_engine *h) {
    return 2 + h->D + h->n_der + 2 + 1 + (int)h->likes.size();
}

extern "C" int32_t cb2_row_width(const cb2_engine *h) { return h ? row_width(h) : -1; }
extern "C" int32_t cb2_n_derived(const cb2_engine *h) { return h ? h->n_der : -1; }

static int pack_fast(cb2_engine *h);  // fills d_fastpack (kernels_fast.cuh layout)

// upload the model if it changed; fill h->M
static int build_model(cb2_engine *h) {
    if (!h->model_dirty) return 0;
    if (!h->have_prior) FAIL(h, -1, "prior not set (cb2_set_prior)");
    if (h->likes.empty()) FAIL(h, -1, "no likelihood set");
    if (!h->have_proposal) FAIL(h, -1, "proposal transform not set (cb2_set_proposal)");
    CK(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = upload(h, h->d_prior_kind, h->prior_kind))) return rc;
    if ((rc = upload(h, h->d_periodic, h->periodic))) return rc;
    if ((rc = upload(h, h->d_lower, h->lower))) return rc;
    if ((rc = upload(h, h->d_upper, h->upper))) return rc;
    if ((rc = upload(h, h->d_loc, h->loc))) return rc;
    if ((rc = upload(h, h->d_pscale, h->pscale))) return rc;
    for (int i = 0; i < h->D; ++i)
        if (h->prior_kind[i] >= 2 && !h->have_shapes)
            FAIL(h, -1, "prior kind %d needs cb2_set_prior_shapes", h->prior_kind[i]);
    if ((rc = upload(h, h->d_pa, h->pa))) return rc;
    if ((rc = upload(h, h->d_pb, h->pb))) return rc;
    if ((rc = upload(h, h->d_pcn, h->pcn))) return rc;
    if ((rc = upload(h, h->d_i_of_j, h->i_of_j))) return rc;
    if ((rc = upload(h, h->d_TT, h->TT))) return rc;
    std::vector<double> dpool;
    std::vector<int32_t> ipool;
    ModelDev &M = h->M;
    memset(&M, 0, sizeof(M));
    M.D = h->D;
    M.n_like = (int)h->likes.size();
    M.n_der = h->n_der;
    M.width = row_width(h);
    for (size_t l = 0; l < h->likes.size(); ++l) {
        LikeHost &L = h->likes[l];
        L.d.idx_off = (int)ipool.size();
        ipool.insert(ipool.end(), L.idx.begin(), L.idx.end());
        L.d.means_off = (int)dpool.size();
        dpool.insert(dpool.end(), L.means.begin(), L.means.end());
        L.d.linvT_off = (int)dpool.size();
        dpool.insert(dpool.end(), L.linvT.begin(), L.linvT.end());
        L.d.c0_off = (int)dpool.size();
        dpool.insert(dpool.end(), L.c0.begin(), L.c0.end());
        L.d.w_off = (int)dpool.size();
        dpool.insert(dpool.end(), L.w.begin(), L.w.end());
        M.likes[l] = L.d;
    }
    if ((rc = upload(h, h->d_dpool, dpool))) return rc;
    if ((rc = upload(h, h->d_ipool, ipool))) return rc;
    M.prior_kind = h->d_prior_kind.p;
    M.lower = h->d_lower.p; M.upper = h->d_upper.p; M.loc = h->d_loc.p; M.pscale = h->d_pscale.p;
    M.periodic = h->d_periodic.p;
    M.pa = h->d_pa.p; M.pb = h->d_pb.p; M.pcn = h->d_pcn.p;
    M.any_periodic = 0; M.any_normal = 0; M.any_generic = 0;
    for (int i = 0; i < h->D; ++i) {
        M.any_periodic |= h->periodic[i] ? 1 : 0;
        M.any_normal |= h->prior_kind[i] != 0 ? 1 : 0;
        M.any_generic |= h->prior_kind[i] >= 2 ? 1 : 0;
    }
    M.uniform_logp = h->uniform_logp;
    M.dpool = h->d_dpool.p;
    M.ipool = h->d_ipool.p;
    M.n_blocks = h->n_blocks;
    for (int b = 0; b < h->n_blocks; ++b) {
        M.bsize[b] = h->bsize[b];
        M.jstart[b] = h->jstart[b];
        M.oversamp[b] = h->oversamp[b];
    }
    M.i_of_j = h->d_i_of_j.p;
    M.drag = h->drag; M.last_slow = h->last_slow; M.drag_steps = h->drag_steps;
    M.n_slow = h->n_slow; M.n_fast = h->n_fast;
    M.TT = h->d_TT.p;
    M.proposal_scale = h->proposal_scale;
    M.temperature = h->temperature;
    M.max_tries = h->max_tries;
    M.output_thin = h->output_thin;
    M.key0 = (uint32_t)h->seed;
    M.key1 = (uint32_t)(h->seed >> 32);
    M.chain_id0 = h->chain_id0;
    if ((rc = upload(h, h->d_ms_main, h->ms_main))) return rc;
    if ((rc = upload(h, h->d_ms_slow, h->ms_slow))) return rc;
    if ((rc = upload(h, h->d_ms_fast, h->ms_fast))) return rc;
    if ((rc = pack_fast(h))) return rc;
    h->model_dirty = false;
    return 0;
}

// ---------------------------------------------------------------- shared-memory plans
static StepSmem plan_step_smem(const cb2_engine *h) {
    StepSmem L;
    const int D = h->D, ND = std::max(h->n_der, 1), NL = CB2_MAX_LIKES;
    int maxn = 1, maxdl = 1;
    for (int b = 0; b < h->n_blocks; ++b) maxn = std::max(maxn, h->bsize[b]);
    for (auto &lk : h->likes) maxdl = std::max(maxdl, lk.d.dim);
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 1) & ~1; return r; };
    L.x = take(D); L.trial = take(D); L.v = take(maxn); L.z = take(maxdl);
    L.der = take(ND); L.tder = take(ND); L.ll = take(NL); L.tll = take(NL);
    L.lpk = take(CB2_MAX_MODES); L.vis = take(2 * (CB2_MAX_BLOCKS + 1));
    if (h->drag) {
        L.e_pt = take(D); L.ps = take(D); L.pe = take(D); L.delta = take(D);
        L.e_der = take(ND); L.pe_der = take(ND); L.e_ll = take(NL); L.pe_ll = take(NL);
        L.tmp_ll = take(NL);
    } else {
        L.e_pt = L.ps = L.pe = L.delta = L.e_der = L.pe_der = L.e_ll = L.pe_ll = L.tmp_ll = 0;
    }
    L.total = o;
    return L;
}

// ---------------------------------------------------------------- init / state
__global__ void k_init_state(ModelDev M, ChainState S, StepSmem L, int64_t n_chains,
                             int64_t burn_left0) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    double *base = sm + (size_t)wid * L.total;
    double *x = base + L.x, *z = base + L.z, *der = base + L.der, *ll = base + L.ll,
           *lpk = base + L.lpk;
    const int D = M.D;
    for (int i = lane; i < D; i += 32) x[i] = S.x[chain * D + i];
    for (int i = lane; i < M.n_der; i += 32) der[i] = 0.0;
    if (lane < CB2_MAX_LIKES) ll[lane] = 0.0;
    __syncwarp();
    double lp;
    double v = warp_logpost(M, x, lp, ll, der, z, lpk, lane);
    __syncwarp();
    for (int i = lane; i < M.n_der; i += 32) S.der[chain * M.n_der + i] = der[i];
    for (int i = lane; i < M.n_like; i += 32) S.ll[chain * M.n_like + i] = ll[i];
    for (int i = lane; i < M.n_blocks + 1; i += 32) S.vis[chain * (M.n_blocks + 1) + i] = 0;
    if (lane == 0) {
        S.logpost[chain] = v;
        S.logprior[chain] = lp;
        S.weight[chain] = 1;        // OneSamplePoint.add (collection.py:1355)
        S.prior_rej[chain] = 0;
        S.burn_left[chain] = burn_left0;  // mcmc.py:265
        S.added_w[chain] = 0;
        S.n_rows[chain] = 0;
        S.n_acc[chain] = 0;
        S.flags[chain] = isfinite(v) ? 0u : CB2_FLAG_INTERNAL;
    }
}

static int step_launch_dims(cb2_engine *h, const StepSmem &L, int &warps, size_t &bytes) {
    size_t per = (size_t)L.total * sizeof(double);
    if (per > 200 * 1024) FAIL(h, -1, "model too large for the general step kernel (D=%d)", h->D);
    warps = (int)std::min<size_t>(8, (200 * 1024) / per);
    // prefer more CTAs over wide CTAs when chains are few
    while (warps > 1 && (h->n_chains + warps - 1) / warps < 2 * h->sm_count) warps >>= 1;
    if (warps < 1) warps = 1;
    bytes = per * warps;
    return 0;
}

static void fill_state_ptrs(cb2_engine *h) {
    ChainState &S = h->S;
    S.x = h->d_x.p; S.logpost = h->d_logpost.p; S.logprior = h->d_logprior.p;
    S.ll = h->d_ll.p; S.der = h->d_der.p; S.weight = h->d_weight.p;
    S.prior_rej = h->d_prior_rej.p; S.burn_left = h->d_burn_left.p;
    S.added_w = h->d_added_w.p; S.n_rows = h->d_n_rows.p; S.n_acc = h->d_n_acc.p;
    S.vis = h->d_vis.p; S.flags = h->d_flags.p; S.rows = h->d_rows.p; S.cap = h->rows_cap;
}

static int alloc_state(cb2_engine *h) {
    const int64_t C = h->n_chains;
    const int D = h->D, NL = (int)h->likes.size(), ND = std::max(h->n_der, 1);
    CK(h, h->d_x.ensure((size_t)C * D)); CK(h, h->d_logpost.ensure(C));
    CK(h, h->d_logprior.ensure(C)); CK(h, h->d_ll.ensure((size_t)C * NL));
    CK(h, h->d_der.ensure((size_t)C * ND)); CK(h, h->d_weight.ensure(C));
    CK(h, h->d_prior_rej.ensure(C)); CK(h, h->d_burn_left.ensure(C));
    CK(h, h->d_added_w.ensure(C)); CK(h, h->d_n_rows.ensure(C)); CK(h, h->d_n_acc.ensure(C));
    CK(h, h->d_vis.ensure((size_t)C * (h->n_blocks + 1))); CK(h, h->d_flags.ensure(C));
    size_t nrows = (size_t)C * (size_t)h->rows_cap * (size_t)row_width(h);
    cudaError_t e = h->d_rows.ensure(nrows);
    if (e != cudaSuccess)
        FAIL(h, -2, "cannot allocate %.2f GB for sample rows (%lld chains x %lld rows x %d): %s",
             nrows * 8.0 / 1e9, (long long)C, (long long)h->rows_cap, row_width(h),
             cudaGetErrorString(e));
    fill_state_ptrs(h);
    return 0;
}

extern "C" int cb2_set_state(cb2_engine *h, const double *x0) {
    if (!h) return -1;
    if (h->rows_cap < 1) FAIL(h, -1, "cb2_set_options must be called before cb2_set_state");
    CK(h, cudaSetDevice(h->device));
    int rc = build_model(h);
    if (rc) return rc;
    const int64_t C = h->n_chains;
    const int D = h->D;
    if ((rc = alloc_state(h))) return rc;
    CK(h, cudaMemcpyAsync(h->d_x.p, x0, (size_t)C * D * sizeof(double), cudaMemcpyHostToDevice,
                          h->stream));
    StepSmem L = plan_step_smem(h);
    int warps; size_t bytes;
    if ((rc = step_launch_dims(h, L, warps, bytes))) return rc;
    CK(h, cudaFuncSetAttribute(k_init_state, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int grid = (int)((C + warps - 1) / warps);
    k_init_state<<<grid, warps * 32, bytes, h->stream>>>(
        h->M, h->S, L, C, h->burn_in * h->output_thin + 1);
    h->launches++;
    CK(h, cudaGetLastError());
    h->steps_done = 0;
    h->have_state = true;
    // the reference requires a finite starting posterior (model.py:707-754)
    int64_t out[8];
    rc = cb2_summary(h, out);
    if (rc) return rc;
    if (out[5] != 0)
        FAIL(h, -4, "%lld initial points have a non-finite log-posterior", (long long)out[5]);
    return 0;
}

extern "C" int cb2_get_state(cb2_engine *h, double *x, double *logpost, int64_t *weight,
                             int64_t *n_rows, int64_t *n_accepted, uint32_t *flags) {
    if (!h || !h->have_state) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    const int64_t C = h->n_chains;
    if (x) CK(h, cudaMemcpy(x, h->d_x.p, (size_t)C * h->D * 8, cudaMemcpyDeviceToHost));
    if (logpost) CK(h, cudaMemcpy(logpost, h->d_logpost.p, C * 8, cudaMemcpyDeviceToHost));
    if (weight) CK(h, cudaMemcpy(weight, h->d_weight.p, C * 8, cudaMemcpyDeviceToHost));
    if (n_rows) CK(h, cudaMemcpy(n_rows, h->d_n_rows.p, C * 8, cudaMemcpyDeviceToHost));
    if (n_accepted) CK(h, cudaMemcpy(n_accepted, h->d_n_acc.p, C * 8, cudaMemcpyDeviceToHost));
    if (flags) CK(h, cudaMemcpy(flags, h->d_flags.p, C * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---- snapshot: everything per-chain that the kernels carry from one proposal to the next.
// Cycler tapes, Haar bases and draws are functions of (seed, chain id, counters) and are
// regenerated, so a restored engine continues bit-for-bit.
struct SnapHeader {
    uint64_t magic;
    int64_t version, C, D, n_blocks, NL, ND, width, steps_done;
    uint64_t seed, chain_id0;
};
static const uint64_t CB2_SNAP_MAGIC = 0x0031504e53324243ull;  // "CB2SNP1"

struct SnapSection { void *dev; size_t bytes; };
static std::vector<SnapSection> snap_sections(cb2_engine *h) {
    const size_t C = (size_t)h->n_chains, D = (size_t)h->D, NL = h->likes.size(),
                 ND = (size_t)h->n_der, NV = (size_t)h->n_blocks + 1;
    return {
        {h->d_x.p, C * D * 8}, {h->d_logpost.p, C * 8}, {h->d_logprior.p, C * 8},
        {h->d_ll.p, C * NL * 8}, {h->d_der.p, C * ND * 8}, {h->d_weight.p, C * 8},
        {h->d_prior_rej.p, C * 8}, {h->d_burn_left.p, C * 8}, {h->d_added_w.p, C * 8},
        {h->d_n_rows.p, C * 8}, {h->d_n_acc.p, C * 8}, {h->d_vis.p, C * NV * 8},
        {h->d_flags.p, ((C * 4 + 7) / 8) * 8},
    };
}

extern "C" int64_t cb2_snapshot_size(cb2_engine *h) {
    if (!h) return -1;
    size_t tot = sizeof(SnapHeader);
    for (const auto &sec : snap_sections(h)) tot += sec.bytes;
    return (int64_t)tot;
}

extern "C" int cb2_export_state(cb2_engine *h, void *buf, int64_t nbytes) {
    if (!h || !h->have_state) return -1;
    if (nbytes < cb2_snapshot_size(h)) FAIL(h, -1, "cb2_export_state: buffer too small");
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    SnapHeader hd = {CB2_SNAP_MAGIC, 1, h->n_chains, h->D, h->n_blocks, (int64_t)h->likes.size(),
                     h->n_der, row_width(h), h->steps_done, h->seed, h->chain_id0};
    char *p = (char *)buf;
    memcpy(p, &hd, sizeof(hd));
    p += sizeof(hd);
    for (const auto &sec : snap_sections(h)) {
        const size_t real = (sec.dev == (void *)h->d_flags.p) ? (size_t)h->n_chains * 4 : sec.bytes;
        if (real) CK(h, cudaMemcpy(p, sec.dev, real, cudaMemcpyDeviceToHost));
        if (real < sec.bytes) memset(p + real, 0, sec.bytes - real);
        p += sec.bytes;
    }
    return 0;
}

extern "C" int cb2_import_state(cb2_engine *h, const void *buf, int64_t nbytes) {
    if (!h) return -1;
    if (h->rows_cap < 1) FAIL(h, -1, "cb2_set_options must be called before cb2_import_state");
    CK(h, cudaSetDevice(h->device));
    int rc = build_model(h);
    if (rc) return rc;
    if (nbytes < (int64_t)sizeof(SnapHeader)) FAIL(h, -1, "cb2_import_state: truncated snapshot");
    SnapHeader hd;
    memcpy(&hd, buf, sizeof(hd));
    if (hd.magic != CB2_SNAP_MAGIC || hd.version != 1)
        FAIL(h, -1, "cb2_import_state: not a snapshot of this engine version");
    if (hd.C != h->n_chains || hd.D != h->D || hd.n_blocks != h->n_blocks ||
        hd.NL != (int64_t)h->likes.size() || hd.ND != h->n_der || hd.width != row_width(h))
        FAIL(h, -1, "cb2_import_state: snapshot of a different model/blocking "
                    "(chains %lld/%lld, D %lld/%d, blocks %lld/%d)",
             (long long)hd.C, (long long)h->n_chains, (long long)hd.D, h->D,
             (long long)hd.n_blocks, h->n_blocks);
    if (hd.seed != h->seed || hd.chain_id0 != h->chain_id0)
        FAIL(h, -1, "cb2_import_state: snapshot was taken with another seed / chain id range");
    if ((rc = alloc_state(h))) return rc;
    if (nbytes < cb2_snapshot_size(h)) FAIL(h, -1, "cb2_import_state: truncated snapshot");
    const char *p = (const char *)buf + sizeof(hd);
    for (const auto &sec : snap_sections(h)) {
        const size_t real = (sec.dev == (void *)h->d_flags.p) ? (size_t)h->n_chains * 4 : sec.bytes;
        if (real) CK(h, cudaMemcpy(sec.dev, p, real, cudaMemcpyHostToDevice));
        p += sec.bytes;
    }
    h->steps_done = hd.steps_done;
    h->have_state = true;
    return 0;
}

extern "C" int cb2_load_rows(cb2_engine *h, int64_t chain, int64_t n, const double *rows) {
    if (!h || !h->have_state) return -1;
    if (chain < 0 || chain >= h->n_chains) FAIL(h, -1, "chain index out of range");
    if (n < 0 || n > h->rows_cap) FAIL(h, -1, "cb2_load_rows: %lld rows exceed rows_cap %lld",
                                       (long long)n, (long long)h->rows_cap);
    CK(h, cudaSetDevice(h->device));
    const int W = row_width(h);
    if (n > 0)
        CK(h, cudaMemcpyAsync(h->d_rows.p + (size_t)chain * h->rows_cap * W, rows,
                              (size_t)n * W * 8, cudaMemcpyHostToDevice, h->stream));
    return 0;
}

extern "C" int cb2_logpost(cb2_engine *h, const double *X, int64_t n, double *logpost,
                           double *logprior, double *loglikes, double *derived) {
    if (!h) return -1;
    if (!h->have_proposal) {  // the posterior does not depend on the proposal
        std::vector<double> I((size_t)h->D * h->D, 0.0);
        for (int i = 0; i < h->D; ++i) I[(size_t)i * h->D + i] = 1.0;
        cb2_set_proposal(h, I.data(), h->proposal_scale);
        h->have_proposal = true;
    }
    CK(h, cudaSetDevice(h->device));
    int rc = build_model(h);
    if (rc) return rc;
    const int D = h->D, NL = (int)h->likes.size(), ND = h->n_der;
    size_t tot = (size_t)n * (D + 2 + NL + std::max(ND, 1));
    CK(h, h->d_tmp.ensure(tot));
    double *dX = h->d_tmp.p, *dlp = dX + (size_t)n * D, *dpr = dlp + n, *dll = dpr + n,
           *dder = dll + (size_t)n * NL;
    CK(h, cudaMemcpyAsync(dX, X, (size_t)n * D * 8, cudaMemcpyHostToDevice, h->stream));
    int per_warp = 2 * D + CB2_MAX_LIKES + std::max(ND, 1) + CB2_MAX_MODES;
    per_warp = (per_warp + 1) & ~1;
    int warps = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / ((size_t)per_warp * 8)));
    size_t bytes = (size_t)per_warp * 8 * warps;
    CK(h, cudaFuncSetAttribute(k_logpost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    int grid = (int)((n + warps - 1) / warps);
    k_logpost<<<grid, warps * 32, bytes, h->stream>>>(h->M, dX, n, dlp, dpr, dll,
                                                       ND ? dder : nullptr, per_warp);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(logpost, dlp, n * 8, cudaMemcpyDeviceToHost));
    if (logprior) CK(h, cudaMemcpy(logprior, dpr, n * 8, cudaMemcpyDeviceToHost));
    if (loglikes) CK(h, cudaMemcpy(loglikes, dll, (size_t)n * NL * 8, cudaMemcpyDeviceToHost));
    if (derived && ND) CK(h, cudaMemcpy(derived, dder, (size_t)n * ND * 8, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------- windows
static bool single_valued(const std::vector<uint8_t> &v) {
    for (size_t i = 1; i < v.size(); ++i)
        if (v[i] != v[0]) return false;
    return true;
}

static int launch_tape(cb2_engine *h, int which, const std::vector<uint8_t> &ms,
                       DevBuf<uint8_t> &d_ms, DevBuf<uint8_t> &d_tape, int64_t i0,
                       const int64_t *per_chain, int stride, int len) {
    const int64_t C = h->n_chains;
    CK(h, d_tape.ensure((size_t)C * len));
    CK(h, h->d_perm_scratch.ensure((size_t)C * ms.size()));
    int bs = 128, grid = (int)((C + bs - 1) / bs);
    h->prof_begin(PROF_TAPE);
    k_cycler_tape<<<grid, bs, 0, h->stream>>>(h->M, which, d_ms.p, (int)ms.size(), i0, per_chain,
                                             stride, len, C, d_tape.p, h->d_perm_scratch.p);
    h->prof_end();
    h->launches++;
    CK(h, cudaGetLastError());
    return 0;
}

// generate cnt epochs of the Haar basis of block b for every chain, starting at each
// chain's current epoch (vis[b]/n_b)
static int launch_basis_impl(cb2_engine *h, int b, int cnt);
static int launch_basis(cb2_engine *h, int b, int cnt) {
    h->prof_begin(PROF_BASIS);
    int rc = launch_basis_impl(h, b, cnt);
    h->prof_end();
    return rc;
}
static int launch_basis_impl(cb2_engine *h, int b, int cnt) {
    const int n = h->bsize[b];
    const int64_t C = h->n_chains;
    const size_t tasks = (size_t)C * cnt;
    CK(h, h->d_basis[b].ensure(tasks * (size_t)n * n));
    const int nn = (n + 2) * (n - 1) / 2, ldh = n | 1, nn_pad = (nn + 2) & ~1;
    const size_t per_task = (size_t)nn_pad + (size_t)n * ldh + n;
    const size_t bytes = per_task * 8;
    const int use_global = bytes > 200 * 1024;
    int threads = std::min(256, std::max(32, ((n + 31) / 32) * 32));
    if (fast_basis_supported(n) && h->policy != 1) {
        int rc = launch_basis_fast(h->stream, h->M.key0, h->M.key1, h->chain_id0, b, n,
                                   h->d_vis.p, h->n_blocks + 1, cnt, h->d_basis[b].p, C);
        if (rc == 0) {
            h->launches++;
            CK(h, cudaGetLastError());
            return 0;
        }
    }
    if (!use_global) {
        CK(h, cudaFuncSetAttribute(k_basis_general, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)bytes));
        k_basis_general<<<(unsigned)tasks, threads, bytes, h->stream>>>(
            h->M.key0, h->M.key1, h->chain_id0, b, n, h->d_vis.p, h->n_blocks + 1, 0u, cnt,
            h->d_basis[b].p, nullptr, 0, 0, 0);
        h->launches++;
        CK(h, cudaGetLastError());
    } else {
        const size_t batch = std::max<size_t>(1, std::min<size_t>(tasks, ((size_t)2 << 30) / bytes));
        CK(h, h->d_basis_scratch.ensure(batch * per_task));
        for (size_t t0 = 0; t0 < tasks; t0 += batch) {
            size_t nb = std::min(batch, tasks - t0);
            k_basis_general<<<(unsigned)nb, threads, 0, h->stream>>>(
                h->M.key0, h->M.key1, h->chain_id0, b, n, h->d_vis.p, h->n_blocks + 1, 0u, cnt,
                h->d_basis[b].p, h->d_basis_scratch.p, 1, (int64_t)t0, 0);
            h->launches++;
            CK(h, cudaGetLastError());
        }
    }
    return 0;
}

static int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

extern "C" int cb2_advance(cb2_engine *h, int64_t n_proposals) {
    if (!h) return -1;
    if (!h->have_state) FAIL(h, -1, "cb2_set_state must be called before cb2_advance");
    if (n_proposals < 0) FAIL(h, -1, "negative number of proposals");
    CK(h, cudaSetDevice(h->device));
    int rc = build_model(h);
    if (rc) return rc;
    fill_state_ptrs(h);
    const int64_t C = h->n_chains;
    StepSmem L = plan_step_smem(h);
    int warps; size_t bytes;
    if ((rc = step_launch_dims(h, L, warps, bytes))) return rc;
    const bool fast_ok = (h->policy == 0 || h->policy == 2) && h->fast_ready;
    int64_t remaining = n_proposals;
    while (remaining > 0) {
        WindowDev W;
        memset(&W, 0, sizeof(W));
        int w;
        if (!h->drag) {
            const int64_t Lc = (int64_t)h->ms_main.size();
            const int64_t pos = h->steps_done % Lc;
            w = (int)std::min<int64_t>(remaining, Lc - pos);
            if (single_valued(h->ms_main)) {
                W.tape_main = nullptr;
                W.const_main = h->ms_main[0];
            } else {
                if ((rc = launch_tape(h, 0, h->ms_main, h->d_ms_main, h->d_tape_main,
                                      h->steps_done, nullptr, 0, w))) return rc;
                W.tape_main = h->d_tape_main.p;
                W.len_main = w;
                W.base_main = h->steps_done;
            }
            for (int b = 0; b < h->n_blocks; ++b) {
                const int n = h->bsize[b];
                if (n < 2) continue;
                int64_t maxv = (h->n_blocks == 1) ? w : std::min<int64_t>(w, (int64_t)h->oversamp[b] * n);
                int cnt = ceil_div(maxv, n) + (pos == 0 ? 0 : 1);
                if ((rc = launch_basis(h, b, cnt))) return rc;
                W.basis[b] = h->d_basis[b].p;
                W.cnt[b] = cnt;
            }
        } else {
            const int64_t Ls = h->n_slow;
            const int64_t pos = h->steps_done % Ls;
            w = (int)std::min<int64_t>(remaining, Ls - pos);
            const int nds = h->drag_steps;
            if (single_valued(h->ms_slow)) {
                W.tape_slow = nullptr;
                W.const_slow = h->ms_slow[0];
            } else {
                if ((rc = launch_tape(h, 1, h->ms_slow, h->d_ms_slow, h->d_tape_slow,
                                      h->steps_done, nullptr, 0, w))) return rc;
                W.tape_slow = h->d_tape_slow.p;
                W.len_slow = w;
                W.base_slow = h->steps_done;
            }
            if (single_valued(h->ms_fast)) {
                W.tape_fast = nullptr;
                W.const_fast = h->ms_fast[0];
            } else {
                if ((rc = launch_tape(h, 2, h->ms_fast, h->d_ms_fast, h->d_tape_fast, 0,
                                      h->d_vis.p + h->n_blocks, h->n_blocks + 1, w * nds)))
                    return rc;
                W.tape_fast = h->d_tape_fast.p;
                W.len_fast = w * nds;
            }
            for (int b = 0; b < h->n_blocks; ++b) {
                const int n = h->bsize[b];
                if (n < 2) continue;
                int cnt;
                if (b <= h->last_slow) {
                    cnt = ceil_div(std::min<int64_t>(w, n), n) + 1;
                } else {
                    int64_t fv = (int64_t)w * nds;
                    int64_t byblock = ceil_div(fv, n);
                    int64_t bycycle = ceil_div(fv, h->n_fast) + 1;
                    cnt = (int)std::min<int64_t>(byblock, bycycle) + 1;
                }
                if ((rc = launch_basis(h, b, cnt))) return rc;
                W.basis[b] = h->d_basis[b].p;
                W.cnt[b] = cnt;
            }
        }
        if (fast_ok) {
            CK(h, h->d_draws.ensure((size_t)C * w));
            CK(h, h->d_plan.ensure((size_t)C * w));
            h->prof_begin(PROF_TAPE);
            if (launch_draws(h->stream, h->M, h->S, W, C, (uint64_t)h->steps_done, w,
                             h->d_draws.p, h->d_plan.p))
                FAIL(h, -2, "draws/plan kernel launch failed");
            h->prof_end();
            h->launches += 2;
        }
        h->prof_begin(PROF_STEP);
        if (fast_ok) {
            rc = -2;
            if (h->policy != 2 && pc_step_supported(h->M, h->fast_desc)) {
                rc = launch_step_pc(h->stream, h->M, h->S, W, h->d_fastpack.p, h->fast_desc,
                                    h->d_draws.p, h->d_plan.p, C, (uint64_t)h->steps_done, w,
                                    h->sm_count);
                if (rc == 0) h->last_kernel = 2;
            }
            if (rc <= -1000) {  // launch problem: report it, then use the single-role kernel
                snprintf(h->pc_error, sizeof(h->pc_error), "k_step_pc launch: %s (%d)",
                         cudaGetErrorString((cudaError_t)((-rc) % 1000)), rc);
                rc = -2;
            }
            if (rc == -2) {  // unsupported by / too large for the producer-consumer kernel
                rc = launch_step_fast(h->stream, h->M, h->S, W, h->d_fastpack.p, h->fast_desc,
                                      h->d_draws.p, h->d_plan.p, C, (uint64_t)h->steps_done,
                                      w, h->sm_count);
                if (rc == 0) h->last_kernel = 1;
            }
            if (rc) FAIL(h, -2, "fast step kernel launch failed (%d)", rc);
        } else {
            CK(h, cudaFuncSetAttribute(k_step_general, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bytes));
            int grid = (int)((C + warps - 1) / warps);
            k_step_general<<<grid, warps * 32, bytes, h->stream>>>(h->M, h->S, W, L, C,
                                                                   (uint64_t)h->steps_done, w);
            h->last_kernel = 0;
        }
        h->prof_end();
        h->launches++;
        CK(h, cudaGetLastError());
        h->steps_done += w;
        remaining -= w;
    }
    return 0;
}

extern "C" int cb2_sync(cb2_engine *h) {
    if (!h) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb2_summary(cb2_engine *h, int64_t out[8]) {
    if (!h || !h->have_state) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, h->d_summary.ensure(8));
    k_summary<<<1, 256, 0, h->stream>>>(h->d_n_rows.p, h->d_n_acc.p, h->d_weight.p, h->d_flags.p,
                                        h->n_chains, h->d_summary.p);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    int64_t tmp[8];
    CK(h, cudaMemcpy(tmp, h->d_summary.p, sizeof(tmp), cudaMemcpyDeviceToHost));
    // tmp[5] = chains with the internal-error flag
    out[0] = tmp[0]; out[1] = tmp[1]; out[2] = tmp[2]; out[3] = tmp[3]; out[4] = tmp[4];
    out[5] = tmp[5]; out[6] = tmp[6]; out[7] = tmp[7];
    return 0;
}

template <int PER>
static void launch_accumulate(cb2_engine *h, int grid, int nt, int64_t n_tasks) {
    size_t sm = (size_t)2 * h->D * 8;
    k_task_accumulate<PER><<<grid, nt, sm, h->stream>>>(
        h->d_rows.p, h->rows_cap, row_width(h), h->D, h->d_tasks.p, n_tasks, h->d_means.p,
        h->d_sw.p, h->d_shift.p, h->d_partials.p);
}

extern "C" int cb2_moments(cb2_engine *h, int32_t mode, int32_t split, const double *shift,
                           double *dev_out, double *host_out) {
    if (!h || !h->have_state) return -1;
    CK(h, cudaSetDevice(h->device));
    const int D = h->D, DD = D * D, W = row_width(h);
    const int len = 3 + D + 2 * DD;
    const int64_t C = h->n_chains;
    int64_t n_tasks = 0;
    double single_acc = 0.0;
    std::vector<double> sh(D, 0.0);
    if (shift) sh.assign(shift, shift + D);
    int rc;
    if ((rc = upload(h, h->d_shift, sh))) return rc;
    if (mode == CB2_MOMENTS_HALVES) {
        n_tasks = C;
        CK(h, h->d_tasks.ensure(n_tasks));
        k_tasks_halves<<<(int)((C + 127) / 128), 128, 0, h->stream>>>(h->d_n_rows.p, C, h->d_tasks.p);
        h->launches++;
    } else if (mode == CB2_MOMENTS_SINGLE_SPLIT) {
        if (C != 1) FAIL(h, -1, "single-split moments need exactly one chain on this engine");
        if (split < 1) FAIL(h, -1, "Rminus1_single_split must be >= 1");
        CK(h, cudaStreamSynchronize(h->stream));
        int64_t n = 0;
        CK(h, cudaMemcpy(&n, h->d_n_rows.p, 8, cudaMemcpyDeviceToHost));
        const int m = 1 + split;                      // mcmc.py:796
        const int64_t cut = n / m;                    // :797
        if (cut < 2) FAIL(h, -5, "Not enough points in chain to check convergence.");
        std::vector<MomentTask> tasks;
        for (int i = 1; i < m; ++i) {                 // :801 ranges, python slice [a:b]
            MomentTask t;
            t.chain = 0; t.first = i * cut; t.last = (i + 1) * cut - 1; t.N = (double)cut;
            tasks.push_back(t);
        }
        MomentTask ta;                                // :799 get_acceptance_rate(cut)
        ta.chain = 0; ta.first = cut; ta.last = n; ta.N = 0;
        tasks.push_back(ta);
        n_tasks = m - 1;
        CK(h, h->d_tasks.ensure(tasks.size()));
        CK(h, cudaMemcpyAsync(h->d_tasks.p, tasks.data(), tasks.size() * sizeof(MomentTask),
                              cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
    } else {
        FAIL(h, -1, "unknown moments mode %d", mode);
    }
    h->prof_begin(PROF_MOMENTS);
    const int64_t n_mean_tasks = n_tasks + (mode == CB2_MOMENTS_SINGLE_SPLIT ? 1 : 0);
    CK(h, h->d_means.ensure((size_t)n_mean_tasks * D));
    CK(h, h->d_sw.ensure(n_mean_tasks));
    int grid = (int)std::min<int64_t>(n_tasks, 8 * h->sm_count);
    CK(h, h->d_partials.ensure((size_t)grid * len));
    CK(h, h->d_mom_out.ensure(len));
    const bool dmma_ok = (D <= 64) && (h->policy != 1);
    if (dmma_ok) {
        // proposal-covariance SYRK on the FP64 tensor pipe (one pass over the rows)
        const int NT = (D + 7) / 8;
#define CB2_MM(N)                                                                          \
    case N:                                                                                \
        k_task_moments_dmma<N><<<grid, N * 32, 0, h->stream>>>(                            \
            h->d_rows.p, h->rows_cap, W, D, h->d_tasks.p, n_tasks, h->d_shift.p,           \
            h->d_partials.p, nullptr);                                                     \
        break;
        switch (NT) { CB2_MM(1) CB2_MM(2) CB2_MM(3) CB2_MM(4) CB2_MM(5) CB2_MM(6) CB2_MM(7) CB2_MM(8) }
#undef CB2_MM
        h->launches++;
        if (mode == CB2_MOMENTS_SINGLE_SPLIT) {  // sw of the acceptance task only
            k_task_means<<<1, 32, 0, h->stream>>>(h->d_rows.p, h->rows_cap, W, D,
                                                  h->d_tasks.p + n_tasks, 1,
                                                  h->d_means.p + (size_t)n_tasks * D,
                                                  h->d_sw.p + n_tasks);
            h->launches++;
        }
    } else {
        int warps = 4, gridm = (int)((n_mean_tasks + warps - 1) / warps);
        k_task_means<<<gridm, warps * 32, 0, h->stream>>>(h->d_rows.p, h->rows_cap, W, D,
                                                          h->d_tasks.p, n_mean_tasks,
                                                          h->d_means.p, h->d_sw.p);
        h->launches++;
        int nt = 256;
        int per = (DD + nt - 1) / nt;
        if (per <= 1) launch_accumulate<1>(h, grid, nt, n_tasks);
        else if (per <= 4) launch_accumulate<4>(h, grid, nt, n_tasks);
        else if (per <= 16) launch_accumulate<16>(h, grid, nt, n_tasks);
        else launch_accumulate<0>(h, grid, nt, n_tasks);
        h->launches++;
    }
    CK(h, cudaGetLastError());
    double *out = dev_out ? dev_out : h->d_mom_out.p;
    k_reduce_partials<<<(len + 255) / 256, 256, 0, h->stream>>>(h->d_partials.p, grid, len, out);
    h->prof_end();
    h->launches++;
    CK(h, cudaGetLastError());
    if (mode == CB2_MOMENTS_SINGLE_SPLIT) {
        // acceptance over rows [cut:] (mcmc.py:799): patch sum N*a = (sum N) * a
        CK(h, cudaStreamSynchronize(h->stream));
        double sw = 0.0;
        CK(h, cudaMemcpy(&sw, h->d_sw.p + n_tasks, 8, cudaMemcpyDeviceToHost));
        int64_t n = 0;
        CK(h, cudaMemcpy(&n, h->d_n_rows.p, 8, cudaMemcpyDeviceToHost));
        const int64_t cut = n / (1 + split);
        single_acc = (double)(n - cut) / sw;
        double s1 = 0.0;
        CK(h, cudaMemcpy(&s1, out + 1, 8, cudaMemcpyDeviceToHost));
        double s2 = s1 * single_acc;
        CK(h, cudaMemcpy(out + 2, &s2, 8, cudaMemcpyHostToDevice));
    }
    if (host_out) {
        CK(h, cudaStreamSynchronize(h->stream));
        CK(h, cudaMemcpy(host_out, out, (size_t)len * 8, cudaMemcpyDeviceToHost));
    }
    return 0;
}

// build the task list of a checkpoint on the device (shared by cb2_moments / cb2_bounds)
static int build_tasks(cb2_engine *h, int32_t mode, int32_t split, int64_t &n_tasks) {
    const int64_t C = h->n_chains;
    if (mode == CB2_MOMENTS_HALVES) {
        n_tasks = C;
        CK(h, h->d_tasks.ensure(n_tasks + 1));
        k_tasks_halves<<<(int)((C + 127) / 128), 128, 0, h->stream>>>(h->d_n_rows.p, C, h->d_tasks.p);
        h->launches++;
        return 0;
    }
    if (mode != CB2_MOMENTS_SINGLE_SPLIT) FAIL(h, -1, "unknown moments mode %d", mode);
    if (C != 1) FAIL(h, -1, "single-split statistics need exactly one chain on this engine");
    if (split < 1) FAIL(h, -1, "Rminus1_single_split must be >= 1");
    CK(h, cudaStreamSynchronize(h->stream));
    int64_t n = 0;
    CK(h, cudaMemcpy(&n, h->d_n_rows.p, 8, cudaMemcpyDeviceToHost));
    const int m = 1 + split;
    const int64_t cut = n / m;
    if (cut < 2) FAIL(h, -5, "Not enough points in chain to check convergence.");
    std::vector<MomentTask> tasks;
    for (int i = 1; i < m; ++i) {
        MomentTask t;
        t.chain = 0; t.first = i * cut; t.last = (i + 1) * cut - 1; t.N = (double)cut;
        tasks.push_back(t);
    }
    n_tasks = m - 1;
    CK(h, h->d_tasks.ensure(tasks.size() + 1));
    CK(h, cudaMemcpyAsync(h->d_tasks.p, tasks.data(), tasks.size() * sizeof(MomentTask),
                          cudaMemcpyHostToDevice, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int cb2_bounds(cb2_engine *h, int32_t mode, int32_t split, double limfrac,
                          const double *shift, double *dev_out, double *host_out) {
    if (!h || !h->have_state) return -1;
    CK(h, cudaSetDevice(h->device));
    if (!(limfrac > 0.0 && limfrac < 1.0)) FAIL(h, -1, "limfrac must be in (0,1)");
    const int D = h->D, W = row_width(h);
    const int len = 1 + 4 * D;
    int64_t n_tasks = 0;
    int rc = build_tasks(h, mode, split, n_tasks);
    if (rc) return rc;
    std::vector<double> sh(D, 0.0);
    if (shift) sh.assign(shift, shift + D);
    if ((rc = upload(h, h->d_shift, sh))) return rc;
    // longest window -> power-of-two sort size (capped; longer windows are thinned)
    int64_t sum8[8];
    if ((rc = cb2_summary(h, sum8))) return rc;
    int64_t maxwin = (mode == CB2_MOMENTS_HALVES) ? (sum8[1] - sum8[1] / 2) : sum8[1];
    int n_pow2 = 64;
    while (n_pow2 < maxwin && n_pow2 < 8192) n_pow2 <<= 1;
    CK(h, h->d_bounds.ensure((size_t)n_tasks * D * 2));
    CK(h, h->d_mom_out.ensure(std::max(len, 3 + D + 2 * D * D)));
    const size_t smem = (size_t)2 * n_pow2 * sizeof(double);
    CK(h, cudaFuncSetAttribute(k_task_bounds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->prof_begin(PROF_MOMENTS);
    k_task_bounds<<<(unsigned)(n_tasks * D), 256, smem, h->stream>>>(
        h->d_rows.p, h->rows_cap, W, D, h->d_tasks.p, n_pow2, limfrac, h->d_bounds.p);
    h->launches++;
    double *out = dev_out ? dev_out : h->d_mom_out.p;
    k_reduce_bounds<<<(2 * D + 127) / 128, 128, 0, h->stream>>>(h->d_bounds.p, n_tasks, D,
                                                               h->d_shift.p, out);
    h->prof_end();
    h->launches++;
    CK(h, cudaGetLastError());
    if (host_out) {
        CK(h, cudaStreamSynchronize(h->stream));
        CK(h, cudaMemcpy(host_out, out, (size_t)len * 8, cudaMemcpyDeviceToHost));
    }
    return 0;
}

extern "C" int64_t cb2_copy_rows(cb2_engine *h, int64_t chain, int64_t row_begin, int64_t n,
                                 double *out) {
    if (!h || !h->have_state) return -1;
    if (chain < 0 || chain >= h->n_chains) FAIL(h, -1, "chain index out of range");
    if (cudaSetDevice(h->device) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) return -2;
    int64_t have = 0;
    if (cudaMemcpy(&have, h->d_n_rows.p + chain, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
    if (row_begin < 0) row_begin = 0;
    int64_t cnt = std::max<int64_t>(0, std::min<int64_t>(n, have - row_begin));
    if (cnt > 0) {
        const int W = row_width(h);
        const double *src = h->d_rows.p + ((size_t)chain * h->rows_cap + row_begin) * W;
        if (cudaMemcpy(out, src, (size_t)cnt * W * 8, cudaMemcpyDeviceToHost) != cudaSuccess) {
            h->err = "cb2_copy_rows: cudaMemcpy failed";
            return -2;
        }
    }
    return cnt;
}

extern "C" int cb2_debug_basis(cb2_engine *h, int64_t chain, int32_t block, uint32_t epoch,
                               double *R) {
    if (!h) return -1;
    if (block < 0 || block >= h->n_blocks) FAIL(h, -1, "block out of range");
    if (chain < 0 || chain >= h->n_chains) FAIL(h, -1, "chain out of range");
    CK(h, cudaSetDevice(h->device));
    const int n = h->bsize[block];
    if (n < 2) FAIL(h, -1, "block of size 1 has no basis");
    const int nn = (n + 2) * (n - 1) / 2, ldh = n | 1, nn_pad = (nn + 2) & ~1;
    const size_t per_task = (size_t)nn_pad + (size_t)n * ldh + n;
    const size_t bytes = per_task * 8;
    const int use_global = bytes > 200 * 1024;
    int threads = std::min(256, std::max(32, ((n + 31) / 32) * 32));
    DevBuf<double> out;
    CK(h, out.ensure((size_t)n * n));
    const uint32_t k0 = (uint32_t)h->seed, k1 = (uint32_t)(h->seed >> 32);
    if (fast_basis_supported(n) && h->policy != 1) {
        int rc = launch_basis_fast_one(h->stream, k0, k1, h->chain_id0 + (uint64_t)chain, block,
                                       n, epoch, out.p);
        if (rc) FAIL(h, -2, "fast basis kernel launch failed");
    } else if (!use_global) {
        CK(h, cudaFuncSetAttribute(k_basis_general, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)bytes));
        k_basis_general<<<1, threads, bytes, h->stream>>>(k0, k1, h->chain_id0, block, n, nullptr, 0,
                                                          epoch, 1, out.p, nullptr, 0, chain, chain);
    } else {
        CK(h, h->d_basis_scratch.ensure(per_task));
        k_basis_general<<<1, threads, 0, h->stream>>>(k0, k1, h->chain_id0, block, n, nullptr, 0,
                                                      epoch, 1, out.p, h->d_basis_scratch.p, 1,
                                                      chain, chain);
    }
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaStreamSynchronize(h->stream));
    std::vector<double> Rt((size_t)n * n);
    CK(h, cudaMemcpy(Rt.data(), out.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost));
    out.release();
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < n; ++k) R[(size_t)i * n + k] = Rt[(size_t)k * n + i];
    return 0;
}

extern "C" int64_t cb2_launch_count(const cb2_engine *h) { return h ? h->launches : -1; }

extern "C" int cb2_timer_start(cb2_engine *h) {
    if (!h) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventRecord(h->ev0, h->stream));
    return 0;
}

extern "C" int cb2_timer_stop(cb2_engine *h, float *ms) {
    if (!h) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaEventRecord(h->ev1, h->stream));
    CK(h, cudaEventSynchronize(h->ev1));
    CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int cb2_last_step_kernel(const cb2_engine *h) { return h ? h->last_kernel : -1; }
extern "C" const char *cb2_debug_message(const cb2_engine *h) { return h ? h->pc_error : ""; }

extern "C" int cb2_set_profiling(cb2_engine *h, int32_t on) {
    if (!h) return -1;
    h->profiling = on != 0;
    return 0;
}

extern "C" int cb2_kernel_times(cb2_engine *h, double ms[4], int64_t n[4], int32_t reset) {
    if (!h) return -1;
    CK(h, cudaSetDevice(h->device));
    CK(h, cudaStreamSynchronize(h->stream));
    for (auto &r : h->prof) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            h->prof_ms[r.kind] += t;
            h->prof_n[r.kind] += 1;
        }
        h->ev_pool.push_back(r.a);
        h->ev_pool.push_back(r.b);
    }
    h->prof.clear();
    for (int k = 0; k < 4; ++k) {
        ms[k] = h->prof_ms[k];
        n[k] = h->prof_n[k];
        if (reset) { h->prof_ms[k] = 0; h->prof_n[k] = 0; }
    }
    return 0;
}

extern "C" int cb2_set_kernel_policy(cb2_engine *h, int32_t policy) {
    if (!h) return -1;
    h->policy = policy & 3;
    g_basis_wy = (policy & 4) ? 0 : 1;  // +4: Householder sweep with DFMA (k_basis_fast)
    return 0;
}

#include "fast_host.inl"
