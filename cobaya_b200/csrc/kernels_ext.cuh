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
              uint64_t t, int first) {
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
        if (first) {
            const int64_t e = (i < M.n_blocks) ? vv / M.bsize[i] : vv;
            e0s[i] = e;
            E.e0[chain * NV + i] = e;
        } else {
            e0s[i] = E.e0[chain * NV + i];
        }
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
             uint64_t t) {
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
