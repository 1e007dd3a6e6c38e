// kernels_ext.cuh -- the Metropolis step split around EXTERNAL likelihood functions.
//
// The reference accepts any Python callable as a likelihood (LikelihoodExternalFunction,
// cobaya/likelihood.py:150-255).  The engine's counterpart is a device functor: CUDA source of
//     extern "C" __device__ double NAME(const double *p, int n)
// compiled at run time with NVRTC (ext_functor.inl) into its own kernel, one thread per chain.
// A foreign kernel cannot be inlined into the step kernel, so a proposal becomes three launches
// on the engine's stream:
//     k_ext_propose   trial point, prior and built-in likelihoods of every chain -> stash
//     <user kernel>   log-likelihood of every stashed trial point               -> ext[]
//     k_ext_accept    accept / reject + bookkeeping (mcmc.py:545-562,685-748)
// with the cycler tapes, Haar bases and Philox draws of the fused kernels unchanged, so a chain
// makes the same proposals as on any other step kernel.
#pragma once
#include "kernels_general.cuh"

struct ExtStash {
    double *x;       // [C*D]   trial points (sampler order; the user kernels read this)
    double *lp;      // [C]     log-posterior without the external terms
    double *prior;   // [C]
    double *ll;      // [C*NL]
    double *der;     // [C*ND]
    double *ext;     // [C*n_ext] values of the external functions
    int64_t *e0;     // [C*(NB+1)] epochs of the blocks at window start
};

#define CB2_MAX_EXT (2 * CB2_MAX_LIKES)
struct ExtSlots {
    int32_t n;                  // external functions: likelihoods and priors
    int32_t n_ep;               // of which priors
    int32_t slot[CB2_MAX_EXT];  // likelihood index of function k, or -(1 + j) for external prior j
};

__global__ void __launch_bounds__(256)
k_ext_propose(ModelDev M, ChainState S, WindowDev W, StepSmem L, ExtStash E, int64_t n_chains,
              uint64_t t_arg, const uint64_t *__restrict__ t_dev) {
    const uint64_t t = t_dev ? *t_dev : t_arg;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like;
    double *base = sm + (size_t)wid * L.total;
    double *trial = base + L.trial, *v = base + L.v, *z = base + L.z;
    double *tder = base + L.tder, *tll = base + L.tll, *lpk = base + L.lpk;
    const int NV = M.n_blocks + 1;
    int64_t *vis = reinterpret_cast<int64_t *>(base + L.vis);
    int64_t *e0s = vis + NV;
    for (int i = lane; i < D; i += 32) trial[i] = S.x[chain * D + i];       // mcmc.py:556
    for (int i = lane; i < NV; i += 32) {
        const int64_t vv = S.vis[chain * NV + i];
        vis[i] = vv;
        e0s[i] = E.e0[chain * NV + i];   // epochs at window start (k_ext_window_begin)
    }
    for (int i = lane; i < ND; i += 32) tder[i] = 0.0;
    __syncwarp();
    const int b = W.tape_main ? W.tape_main[chain * W.len_main + (int64_t)(t - W.base_main)]
                              : W.const_main;
    const bool ok = warp_block_proposal(M, W, chain, gid, t, 0, b, trial, v, vis, e0s, lane);
    warp_reduce_periodic(M, trial, lane);                                     // :558
    double tprior;
    const double tl = warp_logpost(M, trial, tprior, tll, tder, z, lpk, lane);  // :559
    __syncwarp();
    for (int i = lane; i < D; i += 32) E.x[chain * D + i] = trial[i];
    for (int i = lane; i < ND; i += 32) E.der[chain * ND + i] = tder[i];
    for (int i = lane; i < NL; i += 32) E.ll[chain * NL + i] = tll[i];
    for (int i = lane; i < NV; i += 32) S.vis[chain * NV + i] = vis[i];
    if (lane == 0) {
        E.lp[chain] = tl;
        E.prior[chain] = tprior;
        if (!ok) S.flags[chain] |= CB2_FLAG_INTERNAL;
    }
}

__global__ void __launch_bounds__(256)
k_ext_accept(ModelDev M, ChainState S, StepSmem L, ExtStash E, ExtSlots X, int64_t n_chains,
             uint64_t t_arg, const uint64_t *__restrict__ t_dev) {
    const uint64_t t = t_dev ? *t_dev : t_arg;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like;
    double *base = sm + (size_t)wid * L.total;
    double *x = base + L.x, *trial = base + L.trial;
    double *der = base + L.der, *tder = base + L.tder, *ll = base + L.ll, *tll = base + L.tll;
    for (int i = lane; i < D; i += 32) { x[i] = S.x[chain * D + i]; trial[i] = E.x[chain * D + i]; }
    for (int i = lane; i < ND; i += 32) { der[i] = S.der[chain * ND + i]; tder[i] = E.der[chain * ND + i]; }
    for (int i = lane; i < NL; i += 32) { ll[i] = S.ll[chain * NL + i]; tll[i] = E.ll[chain * NL + i]; }
    ChainRegs R;
    R.logpost = S.logpost[chain]; R.logprior = S.logprior[chain];
    R.weight = S.weight[chain]; R.prior_rej = S.prior_rej[chain];
    R.burn_left = S.burn_left[chain]; R.added_w = S.added_w[chain];
    R.n_rows = S.n_rows[chain]; R.n_acc = S.n_acc[chain]; R.flags = S.flags[chain];
    __syncwarp();
    double tprior = E.prior[chain];
    double tl = E.lp[chain];
    const int NP1 = X.n_ep + 1;
    if (tprior != -CUDART_INF) {
        // external priors after the internal one (prior.py:700-720)
        double like_part = tl - tprior, esum = 0.0;
        for (int k = 0; k < X.n; ++k) {
            if (X.slot[k] >= 0) continue;
            double e = E.ext[chain * X.n + k];
            if (e != e) { R.flags |= CB2_FLAG_INTERNAL; e = -CUDART_INF; }
            esum += e;
        }
        if (X.n_ep) { tprior += esum; tl = tprior + like_part; }
    }
    if (tprior == -CUDART_INF) tl = -CUDART_INF;
    if (tl != -CUDART_INF) {
        // the likelihoods are only evaluated where the prior is finite (model.py:640-678)
        for (int k = 0; k < X.n; ++k) {
            if (X.slot[k] < 0) continue;
            double e = E.ext[chain * X.n + k];
            if (e != e) {                       // NaN: the reference raises; flag the chain
                R.flags |= CB2_FLAG_INTERNAL;
                e = -CUDART_INF;
            }
            if (lane == 0) tll[X.slot[k]] = e;
            tl += e;
        }
    }
    __syncwarp();
    bool acc = metropolis_accept(M, gid, t, 0, tl, R.logpost);               // :560
    acc = __shfl_sync(FULLMASK, (int)acc, 0);
    warp_process(M, S, chain, R, acc, x, der, ll, trial, tder, tll, tl, tprior, lane,
                 X.n_ep ? S.pl + chain * NP1 : nullptr);                     // :561
    __syncwarp();
    if (acc && X.n_ep && lane == 0) {   // prior components of the new current point
        S.pl[chain * NP1] = E.prior[chain];
        for (int k = 0; k < X.n; ++k)
            if (X.slot[k] < 0) S.pl[chain * NP1 - X.slot[k]] = E.ext[chain * X.n + k];
    }
    for (int i = lane; i < D; i += 32) S.x[chain * D + i] = x[i];
    for (int i = lane; i < ND; i += 32) S.der[chain * ND + i] = der[i];
    for (int i = lane; i < NL; i += 32) S.ll[chain * NL + i] = ll[i];
    if (lane == 0) {
        S.logpost[chain] = R.logpost; S.logprior[chain] = R.logprior;
        S.weight[chain] = R.weight; S.prior_rej[chain] = R.prior_rej;
        S.burn_left[chain] = R.burn_left; S.added_w[chain] = R.added_w;
        S.n_rows[chain] = R.n_rows; S.n_acc[chain] = R.n_acc; S.flags[chain] = R.flags;
    }
}

// logpost[i] += sum_k ext[i][k], ll[i][slot_k] = ext[i][k] for points with a finite prior
// (cb2_set_state, cb2_logpost).  flags != nullptr: the start-point check of k_init_state.
__global__ void k_ext_add(ExtSlots X, int64_t n, int NL, const double *__restrict__ ext,
                          double *__restrict__ logpost, double *__restrict__ logprior,
                          double *__restrict__ ll, uint32_t *__restrict__ flags,
                          double *__restrict__ pl) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lp = logpost[i], pr = logprior[i];
    const int NP1 = X.n_ep + 1;
    if (pl) pl[i * NP1] = pr;
    if (pr != -CUDART_INF) {
        const double like_part = lp - pr;
        double esum = 0.0;
        for (int k = 0; k < X.n; ++k)
            if (X.slot[k] < 0) {
                const double e = ext[i * X.n + k];
                if (pl) pl[i * NP1 - X.slot[k]] = e;
                esum += e;
            }
        if (X.n_ep) { pr += esum; lp = pr + like_part; logprior[i] = pr; }
    }
    if (pr == -CUDART_INF || pr != pr) lp = -CUDART_INF;
    if (lp != -CUDART_INF) {
        for (int k = 0; k < X.n; ++k)
            if (X.slot[k] >= 0) {
                const double e = ext[i * X.n + k];
                ll[i * NL + X.slot[k]] = e;
                lp += e;
            }
    }
    logpost[i] = lp;
    if (flags) flags[i] = isfinite(lp) ? 0u : CB2_FLAG_INTERNAL;
}

// =========================================================================================
// Dragging (mcmc.py:564-668) with external likelihood functions: BASELINE configs[3] as it is
// stated ("Rosenbrock external-likelihood, dragging").  A slow proposal needs 1 + 2 n_drag
// posterior evaluations, each fast step depending on the accept decision of the one before,
// so the step is split at every evaluation:
//   k_extd_begin         slow proposal on the end point              -> 1 point per chain
//   <user kernel>
//   k_extd_mid(1)        e_lp, `live`, running sums; fast proposal 1   -> 2 points per chain
//   <user kernel>
//   k_extd_mid(i)        decision of fast step i-1; fast proposal i    (i = 2 .. n_drag)
//   <user kernel>
//   k_extd_last          decision of fast step n_drag; final accept on the averaged
//                        log-posteriors, bookkeeping, row
// following the drag branch of k_step_general statement by statement (same Philox draws,
// cyclers and Haar bases).  External PRIORS are not supported together with dragging.
// =========================================================================================
struct DragStash {
    double *s_pt, *e_pt;   // [C*D] current start / end point
    double *pts;           // [2*C*D] evaluation points: slow step [0, C), fast step ps | pe
    double *lp, *prior;    // [2*C]   partial log-posterior / prior of the evaluation points
    double *ll, *der;      // [2*C*NL], [2*C*ND]
    double *s_lp, *e_lp, *e_prior, *s_acc, *e_acc;   // [C]
    double *e_ll, *e_der;  // [C*NL], [C*ND]
    int32_t *live;         // [C]
    double *ext;           // [2*C*n_ext]
    int64_t *e0;           // [C*(NB+1)]
};

// log-posterior of an evaluation point from its partial value and the external terms;
// ll_row (global, optional) receives the external log-likelihoods
__device__ __forceinline__ double ext_like_total(const ExtSlots &X, const double *ext_row,
                                                 double partial_lp, double prior, double *ll_row,
                                                 bool &nan_seen) {
    if (prior == -CUDART_INF || partial_lp == -CUDART_INF) return -CUDART_INF;
    double tl = partial_lp;
    for (int k = 0; k < X.n; ++k) {
        if (X.slot[k] < 0) continue;
        double e = ext_row[k];
        if (e != e) { nan_seen = true; e = -CUDART_INF; }
        if (ll_row) ll_row[X.slot[k]] = e;
        tl += e;
    }
    return tl;
}

__global__ void __launch_bounds__(256)
k_extd_begin(ModelDev M, ChainState S, WindowDev W, StepSmem L, DragStash E, int64_t n_chains,
             uint64_t t_arg, const uint64_t *__restrict__ t_dev) {
    const uint64_t t = t_dev ? *t_dev : t_arg;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like, NV = M.n_blocks + 1;
    double *base = sm + (size_t)wid * L.total;
    double *e_pt = base + L.e_pt, *v = base + L.v, *z = base + L.z, *lpk = base + L.lpk;
    double *e_der = base + L.e_der, *e_ll = base + L.e_ll;
    int64_t *vis = reinterpret_cast<int64_t *>(base + L.vis);
    int64_t *e0s = vis + NV;
    for (int i = lane; i < D; i += 32) {                                  // mcmc.py:579-581
        const double xi = S.x[chain * D + i];
        e_pt[i] = xi;
        E.s_pt[chain * D + i] = xi;
    }
    for (int i = lane; i < NV; i += 32) {
        const int64_t vv = S.vis[chain * NV + i];
        vis[i] = vv;
        e0s[i] = E.e0[chain * NV + i];
    }
    for (int i = lane; i < ND; i += 32) e_der[i] = 0.0;
    __syncwarp();
    const int b = W.tape_slow ? W.tape_slow[chain * W.len_slow + (int64_t)(t - W.base_slow)]
                              : W.const_slow;
    const bool ok = warp_block_proposal(M, W, chain, gid, t, 0, b, e_pt, v, vis, e0s, lane);  // :582
    warp_reduce_periodic(M, e_pt, lane);                                   // :583
    double e_prior;
    const double e_lp = warp_logpost(M, e_pt, e_prior, e_ll, e_der, z, lpk, lane);  // :589
    __syncwarp();
    for (int i = lane; i < D; i += 32) {
        E.e_pt[chain * D + i] = e_pt[i];
        E.pts[chain * D + i] = e_pt[i];
    }
    for (int i = lane; i < ND; i += 32) E.der[chain * ND + i] = e_der[i];
    for (int i = lane; i < NL; i += 32) E.ll[chain * NL + i] = e_ll[i];
    for (int i = lane; i < NV; i += 32) S.vis[chain * NV + i] = vis[i];
    if (lane == 0) {
        E.lp[chain] = e_lp;
        E.prior[chain] = e_prior;
        E.s_lp[chain] = S.logpost[chain];                                  // :580
        if (!ok) S.flags[chain] |= CB2_FLAG_INTERNAL;
    }
}

// e_lp of the slow proposal, `live`, running sums (one thread per chain)
__device__ __forceinline__ bool extd_finish_body(const ModelDev &M, const ChainState &S,
                                                 const DragStash &E, const ExtSlots &X,
                                                 int64_t chain) {
    const int NL = M.n_like, ND = M.n_der;
    bool nan_seen = false;
    const double e_lp = ext_like_total(X, E.ext + chain * X.n, E.lp[chain], E.prior[chain],
                                       E.ll + chain * NL, nan_seen);
    if (nan_seen) S.flags[chain] |= CB2_FLAG_INTERNAL;
    const bool live = e_lp != -CUDART_INF;                                 // :590-592
    E.live[chain] = live ? 1 : 0;
    if (!live) {
        S.weight[chain] += 1;   // the fast cycler is not advanced on this path
        return false;
    }
    E.e_lp[chain] = e_lp;
    E.e_prior[chain] = E.prior[chain];
    for (int i = 0; i < NL; ++i) E.e_ll[chain * NL + i] = E.ll[chain * NL + i];
    for (int i = 0; i < ND; ++i) E.e_der[chain * ND + i] = E.der[chain * ND + i];
    E.s_acc[chain] = E.s_lp[chain];                                        // :595-596
    E.e_acc[chain] = e_lp;
    return true;
}

__device__ __forceinline__ void extd_fast_propose_body(const ModelDev &M, const ChainState &S,
                                                       const WindowDev &W, const StepSmem &L,
                                                       const DragStash &E, int64_t n_chains,
                                                       int64_t chain, uint64_t t, int i_step,
                                                       double *sm, int lane, int wid) {
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like, NV = M.n_blocks + 1;
    const int64_t C = n_chains;
    double *base = sm + (size_t)wid * L.total;
    double *ps = base + L.ps, *pe = base + L.pe, *delta = base + L.delta, *v = base + L.v,
           *z = base + L.z, *lpk = base + L.lpk, *pe_der = base + L.pe_der,
           *pe_ll = base + L.pe_ll, *tmp_ll = base + L.tmp_ll;
    int64_t *vis = reinterpret_cast<int64_t *>(base + L.vis);
    int64_t *e0s = vis + NV;
    for (int k = lane; k < NV; k += 32) {
        vis[k] = S.vis[chain * NV + k];
        e0s[k] = E.e0[chain * NV + k];
    }
    for (int k = lane; k < D; k += 32) delta[k] = 0.0;                     // :606
    for (int k = lane; k < ND; k += 32) pe_der[k] = 0.0;
    __syncwarp();
    int bf;
    {
        const int64_t fi = vis[M.n_blocks];  // fast-cycler visit counter
        const int64_t rel = fi - e0s[M.n_blocks];
        const bool inr = rel >= 0 && rel < W.len_fast;
        bf = W.tape_fast ? (inr ? W.tape_fast[chain * W.len_fast + rel] : 0) : W.const_fast;
        if (W.tape_fast && !inr && lane == 0) S.flags[chain] |= CB2_FLAG_INTERNAL;
        __syncwarp();
        if (lane == 0) vis[M.n_blocks] = fi + 1;
    }
    if (!warp_block_proposal(M, W, chain, gid, t, (uint32_t)i_step, bf, delta, v, vis, e0s, lane)
        && lane == 0)
        S.flags[chain] |= CB2_FLAG_INTERNAL;                               // :607
    warp_reduce_periodic(M, delta, lane);                                  // :608
    for (int k = lane; k < D; k += 32) {
        ps[k] = E.s_pt[chain * D + k] + delta[k];                          // :610
        pe[k] = E.e_pt[chain * D + k] + delta[k];                          // :622
    }
    __syncwarp();
    double ps_prior, pe_prior;
    const double ps_lp = warp_logpost(M, ps, ps_prior, tmp_ll, nullptr, z, lpk, lane);
    __syncwarp();
    for (int k = lane; k < NL; k += 32) E.ll[chain * NL + k] = tmp_ll[k];
    const double pe_lp = warp_logpost(M, pe, pe_prior, pe_ll, pe_der, z, lpk, lane);
    __syncwarp();
    for (int k = lane; k < D; k += 32) {
        E.pts[chain * D + k] = ps[k];
        E.pts[(C + chain) * D + k] = pe[k];
    }
    for (int k = lane; k < NL; k += 32) E.ll[(C + chain) * NL + k] = pe_ll[k];
    for (int k = lane; k < ND; k += 32) E.der[(C + chain) * ND + k] = pe_der[k];
    for (int k = lane; k < NV; k += 32) S.vis[chain * NV + k] = vis[k];
    if (lane == 0) {
        E.lp[chain] = ps_lp; E.prior[chain] = ps_prior;
        E.lp[C + chain] = pe_lp; E.prior[C + chain] = pe_prior;
    }
}

__device__ __forceinline__ void extd_fast_accept_body(const ModelDev &M, const ChainState &S,
                                                      const DragStash &E, const ExtSlots &X,
                                                      int64_t n_chains, int64_t chain,
                                                      uint64_t t, int i_step, int lane) {
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like;
    const int64_t C = n_chains;
    bool nan_seen = false;
    double ps_lp = 0.0, pe_lp = 0.0;
    if (lane == 0) {
        ps_lp = ext_like_total(X, E.ext + chain * X.n, E.lp[chain], E.prior[chain], nullptr,
                               nan_seen);
        pe_lp = ext_like_total(X, E.ext + (C + chain) * X.n, E.lp[C + chain], E.prior[C + chain],
                               E.ll + (C + chain) * NL, nan_seen);
        if (nan_seen) S.flags[chain] |= CB2_FLAG_INTERNAL;
    }
    __syncwarp();   // the external log-likelihoods lane 0 wrote into E.ll are read by all lanes
    ps_lp = __shfl_sync(FULLMASK, ps_lp, 0);
    pe_lp = __shfl_sync(FULLMASK, pe_lp, 0);
    double s_lp = E.s_lp[chain], e_lp = E.e_lp[chain];
    if (ps_lp != -CUDART_INF && pe_lp != -CUDART_INF) {                    // :621,:628
        const double frac = (double)i_step / (double)(1 + M.drag_steps);   // :630
        const double p_int = (1 - frac) * ps_lp + frac * pe_lp;
        const double c_int = (1 - frac) * s_lp + frac * e_lp;
        bool ad = metropolis_accept(M, gid, t, (uint32_t)i_step, p_int, c_int);
        ad = __shfl_sync(FULLMASK, (int)ad, 0);
        if (ad) {                                                          // :640-645
            for (int k = lane; k < D; k += 32) {
                E.s_pt[chain * D + k] = E.pts[chain * D + k];
                E.e_pt[chain * D + k] = E.pts[(C + chain) * D + k];
            }
            for (int k = lane; k < ND; k += 32) E.e_der[chain * ND + k] = E.der[(C + chain) * ND + k];
            for (int k = lane; k < NL; k += 32) E.e_ll[chain * NL + k] = E.ll[(C + chain) * NL + k];
            s_lp = ps_lp; e_lp = pe_lp;
            if (lane == 0) {
                E.s_lp[chain] = s_lp; E.e_lp[chain] = e_lp;
                E.e_prior[chain] = E.prior[C + chain];
            }
        }
    }
    if (lane == 0) {                                                       // :655-656
        E.s_acc[chain] += s_lp;
        E.e_acc[chain] += e_lp;
    }
}

__device__ __forceinline__ void extd_end_body(const ModelDev &M, const ChainState &S,
                                              const StepSmem &L, const DragStash &E,
                                              int64_t chain, uint64_t t, double *sm, int lane,
                                              int wid) {
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like;
    double *base = sm + (size_t)wid * L.total;
    double *x = base + L.x, *e_pt = base + L.e_pt;
    double *der = base + L.der, *e_der = base + L.e_der, *ll = base + L.ll, *e_ll = base + L.e_ll;
    for (int i = lane; i < D; i += 32) { x[i] = S.x[chain * D + i]; e_pt[i] = E.e_pt[chain * D + i]; }
    for (int i = lane; i < ND; i += 32) { der[i] = S.der[chain * ND + i]; e_der[i] = E.e_der[chain * ND + i]; }
    for (int i = lane; i < NL; i += 32) { ll[i] = S.ll[chain * NL + i]; e_ll[i] = E.e_ll[chain * NL + i]; }
    ChainRegs R;
    R.logpost = S.logpost[chain]; R.logprior = S.logprior[chain];
    R.weight = S.weight[chain]; R.prior_rej = S.prior_rej[chain];
    R.burn_left = S.burn_left[chain]; R.added_w = S.added_w[chain];
    R.n_rows = S.n_rows[chain]; R.n_acc = S.n_acc[chain]; R.flags = S.flags[chain];
    __syncwarp();
    const double navg = (double)(1 + M.drag_steps);                        // :658
    bool acc = metropolis_accept(M, gid, t, 0, E.e_acc[chain] / navg, E.s_acc[chain] / navg);
    acc = __shfl_sync(FULLMASK, (int)acc, 0);
    warp_process(M, S, chain, R, acc, x, der, ll, e_pt, e_der, e_ll, E.e_lp[chain],
                 E.e_prior[chain], lane);                                  // :666
    __syncwarp();
    for (int i = lane; i < D; i += 32) S.x[chain * D + i] = x[i];
    for (int i = lane; i < ND; i += 32) S.der[chain * ND + i] = der[i];
    for (int i = lane; i < NL; i += 32) S.ll[chain * NL + i] = ll[i];
    if (lane == 0) {
        S.logpost[chain] = R.logpost; S.logprior[chain] = R.logprior;
        S.weight[chain] = R.weight; S.prior_rej[chain] = R.prior_rej;
        S.burn_left[chain] = R.burn_left; S.added_w[chain] = R.added_w;
        S.n_rows[chain] = R.n_rows; S.n_acc[chain] = R.n_acc; S.flags[chain] = R.flags;
    }
}


// The launches of the fast steps, merged pairwise: the decision of fast step i-1 (or, for
// i = 1, the evaluation of the slow proposal) and the proposal of fast step i need the same
// warp and the same stash, so they share a kernel; the last one closes the step.
//   k_extd_begin, U, k_extd_mid(1), U, k_extd_mid(2), U, ..., k_extd_mid(n_drag), U, k_extd_last
__global__ void __launch_bounds__(256)
k_extd_mid(ModelDev M, ChainState S, WindowDev W, StepSmem L, DragStash E, ExtSlots X,
           int64_t n_chains, uint64_t t_arg, const uint64_t *__restrict__ t_dev, int i_step) {
    const uint64_t t = t_dev ? *t_dev : t_arg;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    if (i_step == 1) {
        int live = 0;
        if (lane == 0) live = extd_finish_body(M, S, E, X, chain) ? 1 : 0;
        live = __shfl_sync(FULLMASK, live, 0);
        __syncwarp();
        if (!live) return;
    } else {
        if (!E.live[chain]) return;
        extd_fast_accept_body(M, S, E, X, n_chains, chain, t, i_step - 1, lane);
        __syncwarp();
    }
    extd_fast_propose_body(M, S, W, L, E, n_chains, chain, t, i_step, sm, lane, wid);
}

__global__ void __launch_bounds__(256)
k_extd_last(ModelDev M, ChainState S, StepSmem L, DragStash E, ExtSlots X, int64_t n_chains,
            uint64_t t_arg, const uint64_t *__restrict__ t_dev) {
    const uint64_t t = t_dev ? *t_dev : t_arg;
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    if (M.drag_steps >= 1) {
        if (!E.live[chain]) return;
        extd_fast_accept_body(M, S, E, X, n_chains, chain, t, M.drag_steps, lane);
    } else {
        int live = 0;
        if (lane == 0) live = extd_finish_body(M, S, E, X, chain) ? 1 : 0;
        live = __shfl_sync(FULLMASK, live, 0);
        if (!live) return;
    }
    __syncwarp();
    extd_end_body(M, S, L, E, chain, t, sm, lane, wid);
}

// epochs of every block at window start (the bases of a window are numbered from them) and the
// device-side proposal counter of the captured step graph
__global__ void k_ext_window_begin(ModelDev M, ChainState S, int64_t *__restrict__ e0,
                                   int64_t n_chains, uint64_t *__restrict__ t_dev, uint64_t t0) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int NV = M.n_blocks + 1;
    if (i == 0 && t_dev) *t_dev = t0;
    if (i >= n_chains * NV) return;
    const int b = (int)(i % NV);
    const int64_t vv = S.vis[i];
    e0[i] = (b < M.n_blocks) ? vv / M.bsize[b] : vv;
}
__global__ void k_ext_next_step(uint64_t *__restrict__ t_dev) { *t_dev += 1; }
