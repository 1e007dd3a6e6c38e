// kernels_drag.cuh -- k_step_drag<NT>: MCMC.get_new_sample_dragging (cobaya/samplers/mcmc/
// mcmc.py:564-668) with the chain state in registers, for D <= 32 (sm_100a).
//
// The general kernel (k_step_general) gives a chain a whole warp and keeps its vectors in
// shared memory; a dragged proposal is 2 n_drag + 1 posterior evaluations of an O(D) or
// O(D^2) target plus n_drag + 1 block proposals, so the warp spends its time in shuffles and
// shared-memory round trips.  Here, as in k_step_fast, a warp carries 8 chains: quad q is a
// chain, lane r holds elements 8n+2r, 8n+2r+1 of every D-vector (block-sorted coordinates),
// the block proposals T v (proposal.py:224) are m8n8k4 DMMA tiles with the 8 chains as M, the
// Gaussian-mixture likelihood (gaussian_mixture.py:138-163) another DMMA product per mode, and
// the Rosenbrock stand-in of BASELINE configs[3] is evaluated in the fragment layout with two
// shuffles per tile.  Chains of a warp diverge only in data (which fast block, whether the
// slow proposal fell outside the prior): everything is computed for the 8 chains and masked.
//
// Supported: one likelihood (gaussian_mixture <= 4 modes without derived parameters, or the
// built-in Rosenbrock over parameters in sampler order), any blocking, uniform / normal /
// scipy-family 1-D priors, no periodic parameter; D <= 32.  Everything else stays on
// k_step_general.  Draws, visit counters and the row layout are those of the oracle.
#pragma once
#include "kernels_fast.cuh"

struct DragPackExtra {
    int like_kind;      // 0 gaussian mixture (fragments at FastPackDesc.off_A), 1 Rosenbrock
    int like_dim;       // Rosenbrock: number of coupled parameters (sorted coordinates 0..dim-1)
    double like_scale;  // Rosenbrock: logp = -scale * sum(...)
};

template <int NT>
struct DragEval {
    double prior, like, post;
};

// log-posterior of the 8 points of a warp in fragment layout: Prior.logps_internal
// (prior.py:733-763) + the likelihood; returns per-chain values (uniform within a quad)
template <int NT>
__device__ __forceinline__ DragEval<NT> drag_logpost(const ModelDev &M, const FastPackDesc &P,
                                                     const DragPackExtra &X,
                                                     const double *__restrict__ pack, int lane,
                                                     const double (&xt)[NT][2], uint32_t m_norm) {
    constexpr int DP = NT * 8;
    const int r = lane & 3;
    const double *lower = pack + P.off_lower, *upper = pack + P.off_upper;
    const int *pflag = reinterpret_cast<const int *>(pack + P.off_flags);
    bool bad = false;
    double ps = 0.0;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const double2 lo2 = *reinterpret_cast<const double2 *>(lower + 8 * n + 2 * r);
        const double2 up2 = *reinterpret_cast<const double2 *>(upper + 8 * n + 2 * r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double xv = xt[n][h];
            const double lo = h ? lo2.y : lo2.x, up = h ? up2.y : up2.x;
            if (!(xv <= up) || !(xv >= lo) || !isfinite(xv)) bad = true;
            if (M.any_normal && ((m_norm >> (2 * n + h)) & 1u)) {
                const int j = 8 * n + 2 * r + h;
                const double zz = (xv - pack[P.off_loc + j]) / pack[P.off_isc + j];
                if (M.any_generic)
                    ps += pack[P.off_mls + j] +
                          prior1d_shape(pflag[j] >> 8, zz, pack[P.off_pa + j], pack[P.off_pb + j]);
                else
                    ps += pack[P.off_mls + j] - zz * zz / 2;
            }
        }
    }
    bad = __shfl_xor_sync(0xffffffffu, (int)bad, 1) | (int)bad;
    bad = __shfl_xor_sync(0xffffffffu, (int)bad, 2) | (int)bad;
    if (M.any_normal) ps = quad_sum(ps);
    DragEval<NT> out;
    out.prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
    double t_like = 0.0;
    if (X.like_kind == 0) {
        double lp0 = 0.0, lp1 = 0.0, lp2 = 0.0, lp3 = 0.0;
        const double *Af = pack + P.off_A;
        for (int km = 0; km < P.n_modes; ++km) {
            const double *mu = pack + P.off_mu + km * DP;
            double z[NT][2], y[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const double2 m2 = *reinterpret_cast<const double2 *>(mu + 8 * n + 2 * r);
                z[n][0] = xt[n][0] - m2.x;
                z[n][1] = xt[n][1] - m2.y;
                y[n][0] = 0.0;
                y[n][1] = 0.0;
            }
            const double *Ak = Af + (size_t)km * P.blocks_A * 64;
            if (P.tri_like) warp_matvec8<NT, true>(Ak, lane, z, y);
            else warp_matvec8<NT, false>(Ak, lane, z, y);
            double qsum = 0.0;
#pragma unroll
            for (int n = 0; n < NT; ++n) qsum += y[n][0] * y[n][0] + y[n][1] * y[n][1];
            qsum = quad_sum(qsum);
            const double lp = -0.5 * (pack[P.off_c0 + km] + qsum);
            if (km == 0) lp0 = lp; else if (km == 1) lp1 = lp; else if (km == 2) lp2 = lp; else lp3 = lp;
        }
        if (P.n_modes == 1) t_like = lp0;
        else {
            const int nm = P.n_modes;
            double mx = lp0;
            if (nm > 1) mx = fmax(mx, lp1);
            if (nm > 2) mx = fmax(mx, lp2);
            if (nm > 3) mx = fmax(mx, lp3);
            if (mx == -CUDART_INF) t_like = -CUDART_INF;
            else {
                double acc = pack[P.off_w] * exp(lp0 - mx);
                if (nm > 1) acc += pack[P.off_w + 1] * exp(lp1 - mx);
                if (nm > 2) acc += pack[P.off_w + 2] * exp(lp2 - mx);
                if (nm > 3) acc += pack[P.off_w + 3] * exp(lp3 - mx);
                t_like = log(acc) + mx;
            }
        }
    } else {
        // Rosenbrock: -scale * sum_{i < dim-1} [100 (x_{i+1} - x_i^2)^2 + (1 - x_i)^2].  Lane r
        // holds a = x[8n+2r], b = x[8n+2r+1]; the successor of b is the `a` of lane r+1 of the
        // same tile, or of lane 0 of the next tile (two shuffles per tile).
        double acc = 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const double a = xt[n][0], b = xt[n][1];
            const double nxt_same = __shfl_down_sync(0xffffffffu, a, 1);
            const double a_next_tile = (n + 1 < NT) ? xt[(n + 1 < NT) ? n + 1 : n][0] : 0.0;
            const double nxt_tile = __shfl_sync(0xffffffffu, a_next_tile, lane & ~3);
            const double c = (r == 3) ? nxt_tile : nxt_same;
            const int i0 = 8 * n + 2 * r;
            if (i0 + 1 < X.like_dim) {
                const double t1 = b - a * a, t2 = 1.0 - a;
                acc += 100.0 * t1 * t1 + t2 * t2;
            }
            if (i0 + 2 < X.like_dim) {
                const double t1 = c - b * b, t2 = 1.0 - b;
                acc += 100.0 * t1 * t1 + t2 * t2;
            }
        }
        acc = quad_sum(acc);
        t_like = -X.like_scale * acc;
    }
    out.like = t_like;
    out.post = bad ? -CUDART_INF : (out.prior + t_like);
    return out;
}

// Issued one fast step ahead in the dragging loop: the loads of the direction are in flight
// while the two posterior evaluations of the current step run.  `v` comes back RAW (unit
// direction, or nothing for a 1-parameter block); drag_block_finish applies the radius.
struct DragPending {
    double f;     // radius x proposal_scale (signed for a 1-parameter block)
    int jsel;     // sorted index of a 1-parameter block's coordinate, -1 for n >= 2
    bool ok;
};

template <int NT>
__device__ __forceinline__ DragPending drag_block_issue(const ModelDev &M, const WindowDev &W,
                                                        int64_t chain, uint64_t gid, uint64_t t,
                                                        uint32_t sub, int b, long long *vis_q,
                                                        const long long *e0_q, bool adv, int lane,
                                                        bool vec, double (&v)[NT][2]) {
    const int r = lane & 3;
    const int n = M.bsize[b];
    double rad, sign;
    DragPending out;
    out.ok = true;
    out.jsel = -1;
    if (n >= 2) {
        const long long vb = vis_q[b];
        const long long e = vb / n;
        const int k = (int)(vb % n);
        long long slot = e - e0_q[b];
        if (slot < 0 || slot >= W.cnt[b]) { out.ok = false; slot = 0; }
        int2 pl;
        pl.x = (int)((chain * W.cnt[b] + slot) * n + k);
        pl.y = b;
        fetch_direction<NT>(M, W, pl, r, vec, v);
        draw_radial(M, gid, t, sub, n, rad, sign);
        out.f = rad * M.proposal_scale;
    } else {
        draw_radial(M, gid, t, sub, n, rad, sign);
        out.f = (sign > 0) ? rad * M.proposal_scale : -(rad * M.proposal_scale);
        out.jsel = M.jstart[b];
    }
    __syncwarp();
    if (adv && r == 0) vis_q[b] += 1;
    __syncwarp();
    return out;
}

template <int NT>
__device__ __forceinline__ void drag_block_finish(const DragPending &pd, int lane,
                                                  double (&v)[NT][2]) {
    const int r = lane & 3;
    if (pd.jsel < 0) {
#pragma unroll
        for (int nn = 0; nn < NT; ++nn) { v[nn][0] = v[nn][0] * pd.f; v[nn][1] = v[nn][1] * pd.f; }
    } else {
#pragma unroll
        for (int nn = 0; nn < NT; ++nn)
#pragma unroll
            for (int h = 0; h < 2; ++h) v[nn][h] = (8 * nn + 2 * r + h == pd.jsel) ? pd.f : 0.0;
    }
}

// direction (x radius x scale) of a block proposal of every chain of the warp in fragment
// layout, from the chain's own visit counter (RandDirectionProposer, proposal.py:58-82;
// RandProposer1D :85-93); advances the counter of the chains in `adv`.
template <int NT>
__device__ __forceinline__ bool drag_block_vector(const ModelDev &M, const WindowDev &W,
                                                  int64_t chain, uint64_t gid, uint64_t t,
                                                  uint32_t sub, int b, long long *vis_q,
                                                  const long long *e0_q, bool adv, int lane,
                                                  bool vec, double (&v)[NT][2]) {
    const DragPending pd = drag_block_issue<NT>(M, W, chain, gid, t, sub, b, vis_q, e0_q, adv,
                                                lane, vec, v);
    drag_block_finish<NT>(pd, lane, v);
    return pd.ok;
}

template <int NT>
__global__ void __launch_bounds__(128, 1)
k_step_drag(ModelDev M, ChainState S, WindowDev W, const double *__restrict__ gpack,
            FastPackDesc P, DragPackExtra X, int64_t n_chains, uint64_t t0, int n_steps) {
    extern __shared__ __align__(16) double fsm[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ long long s_vis[4][8][CB2_MAX_BLOCKS + 1], s_e0[4][8][CB2_MAX_BLOCKS + 1];
    double *pack = fsm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    // ---- stage the constant block with a TMA bulk copy (global -> shared, mbarrier)
    const uint32_t bytes = (uint32_t)P.total * 8u;
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (tid == 0) {
        mbar_init(mbar_a, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"((uint32_t)__cvta_generic_to_shared(pack)), "l"(gpack), "r"(bytes), "r"(mbar_a)
            : "memory");
    }
    mbar_wait(mbar_a, 0);

    const int q = lane >> 2, r = lane & 3;
    const int64_t tile = blockIdx.x * (int64_t)nwarps + wid;
    const int64_t chain_raw = tile * 8 + q;
    const bool active = chain_raw < n_chains;
    const int64_t chain = active ? chain_raw : (n_chains - 1);
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, NB = M.n_blocks, NV = NB + 1;
    const double *Tf = pack + P.off_T;
    const int *iofj = reinterpret_cast<const int *>(pack + P.off_iofj);
    const int *pflag = reinterpret_cast<const int *>(pack + P.off_flags);
    const bool vec = P.vec_ok != 0;
    long long *vis_q = s_vis[wid][q];
    long long *e0_q = s_e0[wid][q];
    if (r == 0)
        for (int i = 0; i < NV; ++i) {
            const long long vv = S.vis[chain * NV + i];
            vis_q[i] = vv;
            e0_q[i] = (i < NB) ? vv / M.bsize[i] : vv;
        }
    __syncwarp();

    double xs[NT][2];
    uint32_t m_norm = 0;
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 8 * n + 2 * r + h;
            const int i = iofj[j];
            xs[n][h] = (i >= 0) ? S.x[chain * D + i] : 0.0;
            m_norm |= (uint32_t)(pflag[j] & 1) << (2 * n + h);
        }
    double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike = S.ll[chain];
    long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
              burn_left = S.burn_left[chain], added_w = S.added_w[chain],
              n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
    uint32_t flags = S.flags[chain];
    const int nds = M.drag_steps;

    for (int s = 0; s < n_steps; ++s) {
        const uint64_t t = t0 + (uint64_t)s;
        // ---- slow proposal on the end point (mcmc.py:579-589)
        double sp[NT][2], ep[NT][2], v[NT][2], dl[NT][2];
        const int b = W.tape_slow ? W.tape_slow[chain * W.len_slow + (int64_t)(t - W.base_slow)]
                                  : W.const_slow;
        if (!drag_block_vector<NT>(M, W, chain, gid, t, 0u, b, vis_q, e0_q, true, lane, vec, v))
            flags |= CB2_FLAG_INTERNAL;
#pragma unroll
        for (int n = 0; n < NT; ++n) { dl[n][0] = 0.0; dl[n][1] = 0.0; }
        warp_matvec8<NT, true>(Tf, lane, v, dl);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            sp[n][0] = xs[n][0]; sp[n][1] = xs[n][1];
            ep[n][0] = xs[n][0] + dl[n][0]; ep[n][1] = xs[n][1] + dl[n][1];
        }
        double s_lp = logpost;
        const DragEval<NT> e0v = drag_logpost<NT>(M, P, X, pack, lane, ep, m_norm);
        double e_lp = e0v.post, e_prior = e0v.prior, e_ll = e0v.like;
        const bool live = e_lp != -CUDART_INF;   // :590-592: else weight += 1, no fast steps
        double s_acc = s_lp, e_acc = e_lp;
        // the fast cycler and the direction of step i + 1 are issued while step i is evaluated
        // (`live` is fixed for the whole slow step, so the visit counters advance the same way)
        double vn[NT][2];
        DragPending pend;
        double exn = 0.0;
        auto issue_fast = [&](int i) {
            int bf;
            const long long fi = vis_q[NB];
            const long long rel = fi - e0_q[NB];
            const bool inr = rel >= 0 && rel < W.len_fast;
            bf = W.tape_fast ? (inr ? W.tape_fast[chain * W.len_fast + rel] : M.last_slow + 1)
                             : W.const_fast;
            if (live && W.tape_fast && !inr) flags |= CB2_FLAG_INTERNAL;
            __syncwarp();
            if (live && r == 0) vis_q[NB] = fi + 1;
            __syncwarp();
            pend = drag_block_issue<NT>(M, W, chain, gid, t, (uint32_t)i, bf, vis_q, e0_q, live,
                                        lane, vec, vn);
            exn = draw_accept_exp(M, gid, t, (uint32_t)i);   // off the accept decision's chain
        };
        if (nds >= 1) issue_fast(1);
        for (int i = 1; i <= nds; ++i) {          // :603
#pragma unroll
            for (int n = 0; n < NT; ++n) { v[n][0] = vn[n][0]; v[n][1] = vn[n][1]; }
            drag_block_finish<NT>(pend, lane, v);
            if (!pend.ok && live) flags |= CB2_FLAG_INTERNAL;
            const double ex_i = exn;
            if (i < nds) issue_fast(i + 1);
#pragma unroll
            for (int n = 0; n < NT; ++n) { dl[n][0] = 0.0; dl[n][1] = 0.0; }
            warp_matvec8<NT, true>(Tf, lane, v, dl);   // delta on a zero vector (:606-608)
            double ps[NT][2], pe[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                ps[n][0] = sp[n][0] + dl[n][0]; ps[n][1] = sp[n][1] + dl[n][1];
                pe[n][0] = ep[n][0] + dl[n][0]; pe[n][1] = ep[n][1] + dl[n][1];
            }
            const DragEval<NT> a = drag_logpost<NT>(M, P, X, pack, lane, ps, m_norm);   // :610-620
            const DragEval<NT> c = drag_logpost<NT>(M, P, X, pack, lane, pe, m_norm);   // :622-627
            if (live && a.post != -CUDART_INF && c.post != -CUDART_INF) {
                const double frac = (double)i / (double)(1 + nds);            // :630
                const double p_int = (1 - frac) * a.post + frac * c.post;
                const double c_int = (1 - frac) * s_lp + frac * e_lp;
                bool ad;
                if (p_int == -CUDART_INF) ad = false;
                else if (p_int > c_int) ad = true;
                else ad = ex_i > (c_int - p_int) / M.temperature;
                if (ad) {                                                    // :640-645
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        sp[n][0] = ps[n][0]; sp[n][1] = ps[n][1];
                        ep[n][0] = pe[n][0]; ep[n][1] = pe[n][1];
                    }
                    s_lp = a.post; e_lp = c.post; e_prior = c.prior; e_ll = c.like;
                }
            }
            s_acc += s_lp;                                                   // :655-656
            e_acc += e_lp;
        }
        if (!live) {
            weight += 1;
            continue;
        }
        const double navg = (double)(1 + nds);                               // :658
        const double lt = e_acc / navg, lc = s_acc / navg;
        bool acc;
        if (lt == -CUDART_INF) acc = false;
        else if (lt > lc) acc = true;
        else acc = draw_accept_exp(M, gid, t, 0u) > (lc - lt) / M.temperature;
        // ---- process_accept_or_reject (mcmc.py:685-748)
        if (acc) {
            if (burn_left <= 0) {
                long long wst = weight;
                bool store = true;
                if (M.output_thin > 1) {
                    added_w += weight;
                    if (added_w >= M.output_thin) {
                        wst = added_w / M.output_thin;
                        added_w %= M.output_thin;
                    } else store = false;
                }
                if (store) {
                    if (n_rows >= S.cap) flags |= CB2_FLAG_ROWS_FULL;
                    else {
                        if (active) {
                            double *row = S.rows + ((size_t)chain * S.cap + n_rows) * M.width;
                            if (r == 0) {
                                row[0] = (double)wst;
                                row[1] = -(logpost / M.temperature);
                            } else if (r == 1) {
                                row[2 + D] = -logprior;
                                row[3 + D] = -logprior;
                            } else if (r == 2) {
                                row[4 + D] = -2 * loglike;
                                row[5 + D] = -2 * loglike;
                            }
#pragma unroll
                            for (int n = 0; n < NT; ++n)
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    const int i = iofj[8 * n + 2 * r + h];
                                    if (i >= 0) row[2 + i] = xs[n][h];
                                }
                        }
                        n_rows += 1;
                    }
                }
            } else burn_left -= 1;
#pragma unroll
            for (int n = 0; n < NT; ++n) { xs[n][0] = ep[n][0]; xs[n][1] = ep[n][1]; }
            logpost = e_lp; logprior = e_prior; loglike = e_ll;
            weight = 1; prior_rej = 0; n_acc += 1;
        } else {
            weight += 1;
            if (e_prior == -CUDART_INF) prior_rej += 1;
            const long long sgn = (burn_left > 0) - (burn_left < 0);
            if (weight - prior_rej > M.max_tries * (1 + 9 * sgn)) flags |= CB2_FLAG_STUCK;
        }
    }
    // ---- write the state back
    __syncwarp();
    if (active) {
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = iofj[8 * n + 2 * r + h];
                if (i >= 0) S.x[chain * D + i] = xs[n][h];
            }
        if (r == 0) {
            S.logpost[chain] = logpost; S.logprior[chain] = logprior; S.ll[chain] = loglike;
            S.weight[chain] = weight; S.prior_rej[chain] = prior_rej;
            S.burn_left[chain] = burn_left; S.added_w[chain] = added_w;
            S.n_rows[chain] = n_rows; S.n_acc[chain] = n_acc; S.flags[chain] = flags;
            for (int i = 0; i < NV; ++i) S.vis[chain * NV + i] = vis_q[i];
        }
    }
}

static inline bool drag_step_supported(const ModelDev &M, int NT) {
    return M.drag && NT <= 4 && !M.any_periodic;
}

template <int NT>
static int launch_step_drag_t(cudaStream_t st, const ModelDev &M, const ChainState &S,
                              const WindowDev &W, const double *gpack, const FastPackDesc &P,
                              const DragPackExtra &X, int64_t n_chains, uint64_t t0, int n_steps,
                              int sm_count) {
    const int64_t tiles = (n_chains + 7) / 8;
    int wpc = (int)((tiles + sm_count - 1) / sm_count);
    if (wpc < 1) wpc = 1;
    if (wpc > 4) wpc = 4;
    const int grid = (int)((tiles + wpc - 1) / wpc);
    const size_t smem = (size_t)P.total * 8;
    if (smem > 200 * 1024) return -2;
    if (cudaFuncSetAttribute(k_step_drag<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return -1;
    k_step_drag<NT><<<grid, wpc * 32, smem, st>>>(M, S, W, gpack, P, X, n_chains, t0, n_steps);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static inline int launch_step_drag(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                   const WindowDev &W, const double *gpack, const FastPackDesc &P,
                                   const DragPackExtra &X, int64_t n_chains, uint64_t t0,
                                   int n_steps, int sm_count) {
    switch (P.NT) {
        case 1: return launch_step_drag_t<1>(st, M, S, W, gpack, P, X, n_chains, t0, n_steps, sm_count);
        case 2: return launch_step_drag_t<2>(st, M, S, W, gpack, P, X, n_chains, t0, n_steps, sm_count);
        case 3: return launch_step_drag_t<3>(st, M, S, W, gpack, P, X, n_chains, t0, n_steps, sm_count);
        case 4: return launch_step_drag_t<4>(st, M, S, W, gpack, P, X, n_chains, t0, n_steps, sm_count);
    }
    return -2;
}
