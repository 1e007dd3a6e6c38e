// kernels_general.cuh -- the general (any D / blocks / modes / dragging) CUDA path:
// one warp per chain, chain state staged in shared memory.  Each device function
// cites the reference code it implements (cobaya v3.6.2).
#pragma once
#include "common.cuh"

// CB2_FLAG_INTERNAL: include/cobaya_b200.h
#define FULLMASK 0xffffffffu

#ifndef CB2_NVRTC_USER   // host-launched kernels the run-time compiled unit does not need
// ------------------------------------------------------------------ cycler tapes
// CyclicIndexRandomizer.next (proposal.py:46-55) for visits [i0, i0+len) of one
// cycler: tape[chain*len + v] = indices[(i0+v) % n] of cycle (i0+v)/n, where the
// permutation of a cycle is a Fisher-Yates shuffle of the sorted multiset driven by
// the CYCLER stream (stands in for Generator.permutation, proposal.py:54).
// i0 is either uniform (`i0_uniform`) or read per chain from `i0_per_chain[chain*stride]`
// (the fast cycler of the dragging sampler advances a chain-dependent number of times,
// mcmc.py:590-592).
__global__ void k_cycler_tape(ModelDev M, int which, const uint8_t *__restrict__ sorted,
                              int n, int64_t i0_uniform,
                              const int64_t *__restrict__ i0_per_chain, int i0_stride,
                              int len, int64_t n_chains, uint8_t *__restrict__ tape,
                              uint8_t *__restrict__ scratch) {
    int64_t chain = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const int64_t i0 = i0_per_chain ? i0_per_chain[chain * i0_stride] : i0_uniform;
    uint64_t gid = M.chain_id0 + (uint64_t)chain;
    uint8_t *perm = scratch + chain * (int64_t)n;
    uint8_t *out = tape + chain * (int64_t)len;
    int64_t cur = -1;
    for (int v = 0; v < len; ++v) {
        int64_t i = i0 + v;
        int64_t cyc = i / n;
        int pos = (int)(i % n);
        if (n <= 2) {  // proposal.py:43-44: fixed order
            out[v] = sorted[pos];
            continue;
        }
        if (cyc != cur) {
            for (int a = 0; a < n; ++a) perm[a] = sorted[a];
            for (int a = n - 1; a >= 1; --a) {
                u32x4 w = philox4x32_10(M.key0, M.key1, (uint32_t)a, (uint32_t)cyc,
                                        (uint32_t)gid,
                                        CB2_TAG_CYCLER | ((uint32_t)which << 8));
                uint64_t r = ((uint64_t)w.x << 32) | w.y;
                uint32_t j = (uint32_t)__umul64hi(r, (uint64_t)(a + 1));
                uint8_t tmp = perm[a];
                perm[a] = perm[j];
                perm[j] = tmp;
            }
            cur = cyc;
        }
        out[v] = perm[pos];
    }
}

// ------------------------------------------------------------------ random_SO_N
// Haar-random SO(n) per (chain, block, epoch): cobaya/functions.py:21-60.
// One CTA per task.  Householder vectors x_m are prepared as in _rvs (:49-55); the
// product H = G_0 G_1 ... G_{n-2} (G_m = I - x_m x_m^T on coordinates m..n-1) is
// accumulated column by column from the innermost factor (LAPACK dorg2r order)
// instead of the reference's left-to-right sweep (:57-58): same matrix, ~1/3 fewer
// flops, and every thread owns whole columns so no inter-thread reduction is needed.
// Output layout: Rt[k*n + i] = R[i][k] = D[i]*H[i][k]  (direction k contiguous).
// Epochs generated for a chain: e0_c .. e0_c+cnt-1 with e0_c = vis_c[block]/n, the epoch
// the chain's next visit to this block belongs to (vis == NULL: e0 = `e0_fixed`).
__global__ void k_basis_general(uint32_t key0, uint32_t key1, uint64_t chain_id0, int block,
                                int n, const int64_t *__restrict__ vis, int vis_stride,
                                uint32_t e0_fixed, int cnt, double *__restrict__ store,
                                double *__restrict__ gscratch, int use_global,
                                int64_t task0, int64_t store_task0) {
    extern __shared__ double sm[];
    const int64_t task = task0 + blockIdx.x;  // chain*cnt + ei
    const int64_t chain = task / cnt;
    const uint32_t e0 = vis ? (uint32_t)(vis[chain * vis_stride + block] / n) : e0_fixed;
    const uint32_t epoch = e0 + (uint32_t)(task % cnt);
    const uint64_t gid = chain_id0 + (uint64_t)chain;
    const int nn = (n + 2) * (n - 1) / 2;
    const int ldh = n | 1;
    const int nn_pad = (nn + 2) & ~1;
    double *xx, *H, *Dv;
    if (use_global) {
        double *base = gscratch + (size_t)blockIdx.x * (size_t)(nn_pad + (size_t)n * ldh + n);
        xx = base;
        H = base + nn_pad;
        Dv = H + (size_t)n * ldh;
    } else {
        xx = sm;
        H = sm + nn_pad;
        Dv = H + (size_t)n * ldh;
    }
    const int tid = threadIdx.x, nt = blockDim.x;
    // 1. standard normals (functions.py:36)
    for (int p = tid; p < (nn + 1) / 2; p += nt) {
        double z0, z1;
        draw_normal_pair(key0, key1, gid, block, epoch, (uint32_t)p, z0, z1);
        xx[2 * p] = z0;
        xx[2 * p + 1] = z1;
    }
    __syncthreads();
    // 2. Householder vectors (functions.py:49-55); vector m lives at offset ix(m)
    for (int m = tid; m < n - 1; m += nt) {
        int ix = m * n - (m * (m - 1)) / 2;
        int len = n - m;
        double *x = xx + ix;
        double norm2 = 0.0;
        for (int i = 0; i < len; ++i) norm2 += x[i] * x[i];
        double x0 = x[0];
        double d = (x0 != 0.0) ? (x0 > 0 ? 1.0 : -1.0) : 1.0;
        double x0n = x0 + d * sqrt(norm2);
        x[0] = x0n;
        double sc = sqrt((norm2 - x0 * x0 + x0n * x0n) / 2.0);
        for (int i = 0; i < len; ++i) x[i] /= sc;
        Dv[m] = d;
    }
    __syncthreads();
    if (tid == 0) {  // functions.py:59
        double prod = 1.0;
        for (int m = 0; m < n - 1; ++m) prod *= Dv[m];
        Dv[n - 1] = (((n - 1) & 1) ? -1.0 : 1.0) * prod;
    }
    // 3. columns of H
    for (int j = tid; j < n; j += nt) {
        for (int i = 0; i < n; ++i) H[(size_t)i * ldh + j] = (i == j) ? 1.0 : 0.0;
        int mtop = (j < n - 2) ? j : n - 2;
        for (int m = mtop; m >= 0; --m) {
            const double *x = xx + (m * n - (m * (m - 1)) / 2);
            int len = n - m;
            double w = 0.0;
            for (int i = 0; i < len; ++i) w += x[i] * H[(size_t)(m + i) * ldh + j];
            for (int i = 0; i < len; ++i) H[(size_t)(m + i) * ldh + j] -= x[i] * w;
        }
    }
    __syncthreads();
    // 4. R = diag(D) H (functions.py:60), stored transposed
    double *out = store + (size_t)(task - store_task0) * (size_t)n * n;
    for (int e = tid; e < n * n; e += nt) {
        int k = e / n, i = e % n;
        out[e] = Dv[i] * H[(size_t)i * ldh + k];
    }
}

#endif  // CB2_NVRTC_USER
// ------------------------------------------------------------------ log-posterior
// Model.logposterior (model.py:579-678) for the recognised model set, evaluated by
// one warp on a point staged in shared memory.  Returns logpost (warp-uniform).
//   prior : Prior.logps_internal (prior.py:733-763) + _fast_norm_logpdf (tools.py:720)
//   like  : GaussianMixture.logp (gaussian_mixture.py:138-163) in the Cholesky form
//           -1/2 (d log 2pi + log|S_k| + |L_k^-1 (x - mu_k)|^2), logsumexp over modes;
//           derived = L_k^-1 (x - mu_k) (:146-156)
#ifdef CB2_NVRTC_USER
__device__ double cb2_user_like(int user_id, const double *p, int n);  // generated (ext_functor.inl)
#endif
// `only` >= 0: that likelihood component alone, -2: none (cb2_measure_speeds); -1: all.
__device__ __forceinline__ double warp_logpost(const ModelDev &M, const double *xs,
                                               double &lprior, double *ll, double *der,
                                               double *z, double *lpk, int lane,
                                               int only = -1, bool skip_user = false) {
    const int D = M.D;
    bool bad = false;
    for (int i = lane; i < D; i += 32) {
        double xi = xs[i];
        if (!(xi <= M.upper[i]) || !(xi >= M.lower[i]) || !isfinite(xi)) bad = true;
    }
    if (__any_sync(FULLMASK, bad)) {
        lprior = -CUDART_INF;
        return -CUDART_INF;
    }
    double s = 0.0;
    if (M.any_normal) {
        for (int i = lane; i < D; i += 32)
            {
                const int kd = M.prior_kind[i];
                if (kd == CB2_PRIOR_NORMAL) {
                    double sc = M.pscale[i];
                    double zz = (xs[i] - M.loc[i]) / sc;
                    s += (-log(sc) - CB2_LOG_2PI / 2) - zz * zz / 2;
                } else if (kd >= 2) {
                    double zz = (xs[i] - M.loc[i]) / M.pscale[i];
                    s += M.pcn[i] + prior1d_shape(kd, zz, M.pa[i], M.pb[i]);
                }
            }
        s = warp_sum(s);
    }
    lprior = M.uniform_logp + s;
    double total = lprior;
    for (int l = 0; l < M.n_like; ++l) {
        if (only != -1 && l != only) continue;
        const LikeDev &L = M.likes[l];
        const int d = L.dim;
        const int32_t *idx = M.ipool + L.idx_off;
        double val;
        if (L.kind == 0) {
            const int nm = L.n_modes;
            for (int k = 0; k < nm; ++k) {
                const double *mu = M.dpool + L.means_off + (size_t)k * d;
                __syncwarp();
                for (int i = lane; i < d; i += 32) z[i] = xs[idx[i]] - mu[i];
                __syncwarp();
                const double *col = M.dpool + L.linvT_off + (size_t)k * d * d;
                double q = 0.0;
                for (int i = lane; i < d; i += 32) {
                    double a = 0.0;
                    for (int j = 0; j <= i; ++j) a += col[(size_t)j * d + i] * z[j];
                    q += a * a;
                    if (der != nullptr && L.derived) der[L.der_off + k * d + i] = a;
                }
                q = warp_sum(q);
                double lp_k = -0.5 * (M.dpool[L.c0_off + k] + q);
                if (nm == 1) val = lp_k;
                else if (lane == 0) lpk[k] = lp_k;
            }
            if (nm > 1) {
                __syncwarp();
                double mx = lpk[0];
                for (int k = 1; k < nm; ++k) mx = fmax(mx, lpk[k]);
                if (mx == -CUDART_INF) val = -CUDART_INF;
                else {
                    double acc = 0.0;
                    for (int k = 0; k < nm; ++k)
                        acc += M.dpool[L.w_off + k] * exp(lpk[k] - mx);
                    val = log(acc) + mx;
                }
            }
        } else if (L.kind == 2) {
            val = L.scale;  // `one` (likelihoods/one/one.py:26-28)
        } else if (L.kind == 3 && skip_user) {
            val = 0.0;      // added by user_pair_eval (dragging: two points at once)
        } else if (L.kind == 3) {
#ifdef CB2_NVRTC_USER
            // run-time compiled kernel: the user's function is part of this translation unit
            // (ext_functor.inl); its inputs are gathered into the warp's scratch vector
            __syncwarp();
            for (int i = lane; i < d; i += 32) z[i] = xs[idx[i]];
            __syncwarp();
            val = 0.0;
            if (lane == 0) val = cb2_user_like(L.n_modes, z, d);
            val = __shfl_sync(FULLMASK, val, 0);
            if (val != val) {   // NaN (the reference raises): the slot keeps it for
                if (lane == 0) ll[l] = val;   // user_nan_check, the point is rejected
                total += -CUDART_INF;
                continue;
            }
#else
            val = 0.0;      // external function: added by its own kernel (kernels_ext.cuh)
#endif
        } else {
            double acc = 0.0;
            for (int i = lane; i + 1 < d; i += 32) {
                double a = xs[idx[i]], b = xs[idx[i + 1]];
                double t1 = b - a * a, t2 = 1.0 - a;
                acc += 100.0 * t1 * t1 + t2 * t2;
            }
            acc = warp_sum(acc);
            val = -L.scale * acc;
        }
        if (lane == 0) ll[l] = val;
        total += val;
    }
    __syncwarp();
    return total;
}

// a NaN from a user function inlined by run-time compilation: flag the chain (the split route
// does the same in k_ext_accept)
__device__ __forceinline__ void user_nan_check(const ModelDev &M, const double *ll,
                                               uint32_t &flags) {
#ifdef CB2_NVRTC_USER
    for (int l = 0; l < M.n_like; ++l)
        if (M.likes[l].kind == 3 && ll[l] != ll[l]) flags |= CB2_FLAG_INTERNAL;
#endif
}

#ifdef CB2_NVRTC_USER
// Dragging evaluates the moved start and end point of a fast step together: the user's
// functions of both points run side by side on lanes 0 and 1 (warp_logpost was called with
// skip_user for both).  lp_a / lp_b: log-posteriors without the external terms (-inf: the
// point is out and is not evaluated); za / zb: scratch vectors of max(dim) doubles.
__device__ __forceinline__ void user_pair_eval(const ModelDev &M, const double *pa,
                                               const double *pb, double *za, double *zb,
                                               double *lla, double *llb, double &lp_a,
                                               double &lp_b, uint32_t &flags, int lane) {
    const bool ok_a = lp_a != -CUDART_INF, ok_b = lp_b != -CUDART_INF;
    for (int l = 0; l < M.n_like; ++l) {
        const LikeDev &L = M.likes[l];
        if (L.kind != 3) continue;
        const int d = L.dim;
        const int32_t *idx = M.ipool + L.idx_off;
        __syncwarp();
        for (int i = lane; i < d; i += 32) { za[i] = pa[idx[i]]; zb[i] = pb[idx[i]]; }
        __syncwarp();
        double val = 0.0;
        if (lane == 0 && ok_a) val = cb2_user_like(L.n_modes, za, d);
        if (lane == 1 && ok_b) val = cb2_user_like(L.n_modes, zb, d);
        const double va = __shfl_sync(FULLMASK, val, 0), vb = __shfl_sync(FULLMASK, val, 1);
        if (ok_a) {
            if (va != va) { flags |= CB2_FLAG_INTERNAL; lp_a = -CUDART_INF; }
            else { if (lane == 0) lla[l] = va; if (lp_a != -CUDART_INF) lp_a += va; }
        }
        if (ok_b) {
            if (vb != vb) { flags |= CB2_FLAG_INTERNAL; lp_b = -CUDART_INF; }
            else { if (lane == 0) llb[l] = vb; if (lp_b != -CUDART_INF) lp_b += vb; }
        }
    }
    __syncwarp();
}
#endif

#ifndef CB2_NVRTC_USER
// parity entry point (cb2_logpost): one warp per point
__global__ void k_logpost(ModelDev M, const double *__restrict__ X, int64_t n,
                          double *__restrict__ logpost, double *__restrict__ logprior,
                          double *__restrict__ loglikes, double *__restrict__ derived,
                          int per_warp, int only = -1) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t p = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (p >= n) return;
    double *base = sm + (size_t)wid * per_warp;
    const int D = M.D;
    double *xs = base, *z = xs + D, *ll = z + D, *der = ll + CB2_MAX_LIKES,
           *lpk = der + (M.n_der > 0 ? M.n_der : 1);
    for (int i = lane; i < D; i += 32) xs[i] = X[p * D + i];
    for (int i = lane; i < M.n_der; i += 32) der[i] = CUDART_NAN;
    if (lane < CB2_MAX_LIKES) ll[lane] = CUDART_NAN;
    __syncwarp();
    double lp;
    double v = warp_logpost(M, xs, lp, ll, der, z, lpk, lane, only);
    __syncwarp();
    if (lane == 0) {
        logpost[p] = v;
        logprior[p] = lp;
    }
    for (int l = lane; l < M.n_like; l += 32) loglikes[p * M.n_like + l] = ll[l];
    if (derived)
        for (int i = lane; i < M.n_der; i += 32) derived[p * M.n_der + i] = der[i];
}

#endif  // CB2_NVRTC_USER

// ------------------------------------------------------------------ chain state
struct ChainState {
    double *x;         // [chains*D]
    double *logpost;   // [chains]
    double *logprior;  // [chains]
    double *ll;        // [chains*n_like]
    double *der;       // [chains*n_der]
    int64_t *weight, *prior_rej, *burn_left, *added_w, *n_rows, *n_acc;
    int64_t *vis;      // [chains*n_blocks] visits per block proposer
    double *pl;        // [chains*(n_ep+1)] internal + external prior components (n_ep > 0)
    uint32_t *flags;
    double *rows;      // [chains*cap*width]
    int64_t cap;
};

struct WindowDev {
    // tapes: block index per visit (nullptr when the cycler has a single distinct block).
    // main/slow cyclers are visited once per proposal: index = t - base; the fast cycler
    // is indexed from the chain's own visit counter at window start.
    const uint8_t *tape_main; int len_main; int64_t base_main;  // Metropolis
    const uint8_t *tape_slow; int len_slow; int64_t base_slow;  // dragging
    const uint8_t *tape_fast; int len_fast;
    uint8_t const_main, const_slow, const_fast;
    // bases per block: per chain `cnt` epochs starting at vis_start[b]/n_b
    const double *basis[CB2_MAX_BLOCKS];
    int32_t cnt[CB2_MAX_BLOCKS];
};

struct StepSmem {  // offsets in doubles within a warp's shared-memory slab
    int x, trial, v, z, der, tder, ll, tll, lpk, vis;
    int e_pt, ps, pe, delta, e_der, pe_der, e_ll, pe_ll, tmp_ll;
    int total;
};

// BlockedProposer.get_block_proposal (proposal.py:222-224) applied to the vector P
// (shared memory): P[i_of_j[j]] += sum_i T[j][j0+i] * vec_i,  vec = R[:,k] * r * scale
// (RandDirectionProposer.propose_vec :59-69) or +-r*scale (RandProposer1D :86-93).
__device__ __forceinline__ bool warp_block_proposal(const ModelDev &M, const WindowDev &W,
                                                    int64_t chain, uint64_t gid, uint64_t t,
                                                    uint32_t sub, int b, double *P, double *v,
                                                    int64_t *vis, const int64_t *e0s,
                                                    int lane) {
    const int n = M.bsize[b], j0 = M.jstart[b], D = M.D;
    double r, sign;
    draw_radial(M, gid, t, sub, n, r, sign);
    bool ok = true;
    __syncwarp();
    if (n >= 2) {
        int64_t vb = vis[b];
        int64_t e = vb / n;
        int k = (int)(vb % n);
        int64_t slot = e - e0s[b];
        if (slot < 0 || slot >= W.cnt[b]) {
            ok = false;
            slot = 0;
        }
        const double *Rk = W.basis[b] + (((size_t)chain * W.cnt[b] + slot) * n + k) * n;
        for (int i = lane; i < n; i += 32) v[i] = Rk[i] * r * M.proposal_scale;
    } else if (lane == 0) {
        v[0] = (sign > 0) ? r * M.proposal_scale : -(r * M.proposal_scale);
    }
    __syncwarp();
    if (lane == 0) vis[b] += 1;
    for (int j = j0 + lane; j < D; j += 32) {
        int kmax = min(n, j - j0 + 1);
        double a = 0.0;
        for (int i = 0; i < kmax; ++i) a += M.TT[(size_t)(j0 + i) * D + j] * v[i];
        P[M.i_of_j[j]] += a;
    }
    __syncwarp();
    return ok;
}

// Prior.reduce_periodic (prior.py:658-676)
__device__ __forceinline__ void warp_reduce_periodic(const ModelDev &M, double *x, int lane) {
    if (!M.any_periodic) return;
    for (int i = lane; i < M.D; i += 32)
        if (M.periodic[i]) {
            double a = M.lower[i], b = M.upper[i];
            double q = (x[i] - a) / (b - a);
            q = q - floor(q);
            x[i] = q * (b - a) + a;
        }
    __syncwarp();
}

// MCMC.metropolis_accept (mcmc.py:670-683)
__device__ __forceinline__ bool metropolis_accept(const ModelDev &M, uint64_t gid, uint64_t t,
                                                  uint32_t sub, double lt, double lc) {
    if (lt == -CUDART_INF) return false;
    if (lt > lc) return true;
    double ratio = (lc - lt) / M.temperature;
    return draw_accept_exp(M, gid, t, sub) > ratio;
}

struct ChainRegs {
    double logpost, logprior;
    int64_t weight, prior_rej, burn_left, added_w, n_rows, n_acc;
    uint32_t flags;
};

// MCMC.process_accept_or_reject (mcmc.py:685-748) + OneSamplePoint.add_to_collection
// (collection.py:1366-1383) + SampleCollection._cache_add_row (collection.py:519-542).
// `cur_*` is the current point (shared memory), `new_*` the trial to install on accept.
__device__ __forceinline__ void warp_process(const ModelDev &M, const ChainState &S,
                                             int64_t chain, ChainRegs &R, bool accept,
                                             double *cur_x, double *cur_der, double *cur_ll,
                                             const double *new_x, const double *new_der,
                                             const double *new_ll, double new_logpost,
                                             double new_logprior, int lane,
                                             const double *cur_pl = nullptr) {
    // cur_pl (models with external priors): {internal prior, external priors...} of the
    // current point, the minuslogprior__0 / minuslogprior__<name> columns of its row
    const int D = M.D, ND = M.n_der, NL = M.n_like, NE = M.n_ep;
    if (accept) {
        if (R.burn_left <= 0) {
            int64_t w = R.weight;
            bool store = true;
            if (M.output_thin > 1) {
                R.added_w += R.weight;
                if (R.added_w >= M.output_thin) {
                    w = R.added_w / M.output_thin;
                    R.added_w %= M.output_thin;
                } else store = false;
            }
            if (store) {
                if (R.n_rows >= S.cap) {
                    R.flags |= CB2_FLAG_ROWS_FULL;
                } else {
                    double *row = S.rows + ((size_t)chain * S.cap + R.n_rows) * M.width;
                    double llsum = 0.0;
                    for (int l = 0; l < NL; ++l) llsum += cur_ll[l];
                    for (int e = lane; e < M.width; e += 32) {
                        double val;
                        if (e == 0) val = (double)w;
                        else if (e == 1) val = -(R.logpost / M.temperature);
                        else if (e < 2 + D) val = cur_x[e - 2];
                        else if (e < 2 + D + ND) val = cur_der[e - 2 - D];
                        else if (e < 2 + D + ND + 1) val = -R.logprior;
                        else if (e < 2 + D + ND + 2) val = NE ? -cur_pl[0] : -R.logprior;
                        else if (e < 2 + D + ND + 2 + NE) val = -cur_pl[1 + e - (2 + D + ND + 2)];
                        else if (e == 2 + D + ND + 2 + NE) val = -2 * llsum;
                        else val = -2 * cur_ll[e - (2 + D + ND + 3 + NE)];
                        row[e] = val;
                    }
                    R.n_rows += 1;
                }
            }
        } else {
            R.burn_left -= 1;
        }
        __syncwarp();
        for (int i = lane; i < D; i += 32) cur_x[i] = new_x[i];
        for (int i = lane; i < ND; i += 32) cur_der[i] = new_der[i];
        for (int i = lane; i < NL; i += 32) cur_ll[i] = new_ll[i];
        R.logpost = new_logpost;
        R.logprior = new_logprior;
        R.weight = 1;
        R.prior_rej = 0;
        R.n_acc += 1;
        __syncwarp();
    } else {
        R.weight += 1;
        if (new_logprior == -CUDART_INF) R.prior_rej += 1;
        int64_t sgn = (R.burn_left > 0) - (R.burn_left < 0);
        int64_t max_now = M.max_tries * (1 + 9 * sgn);
        if (R.weight - R.prior_rej > max_now) R.flags |= CB2_FLAG_STUCK;
    }
}

// The hot loop (MCMC.run, mcmc.py:470-472): every chain makes n_steps proposals,
// proposal counters t0 .. t0+n_steps-1.  One warp per chain.
__device__ __forceinline__ void step_general_body(const ModelDev &M, const ChainState &S,
                                                  const WindowDev &W, const StepSmem &L,
                                                  int64_t n_chains, uint64_t t0, int n_steps) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t chain = blockIdx.x * (int64_t)(blockDim.x >> 5) + wid;
    if (chain >= n_chains) return;
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int D = M.D, ND = M.n_der, NL = M.n_like;
    double *base = sm + (size_t)wid * L.total;
    double *x = base + L.x, *trial = base + L.trial, *v = base + L.v, *z = base + L.z;
    double *der = base + L.der, *tder = base + L.tder, *ll = base + L.ll,
           *tll = base + L.tll, *lpk = base + L.lpk;
    const int NV = M.n_blocks + 1;  // per-block visit counters + fast-cycler visits
    int64_t *vis = reinterpret_cast<int64_t *>(base + L.vis);
    int64_t *e0s = vis + NV;        // epoch of each block at window start

    for (int i = lane; i < D; i += 32) x[i] = S.x[chain * D + i];
    for (int i = lane; i < ND; i += 32) der[i] = S.der[chain * ND + i];
    for (int i = lane; i < NL; i += 32) ll[i] = S.ll[chain * NL + i];
    for (int i = lane; i < NV; i += 32) {
        int64_t vv = S.vis[chain * NV + i];
        vis[i] = vv;
        e0s[i] = (i < M.n_blocks) ? vv / M.bsize[i] : vv;
    }
    ChainRegs R;
    R.logpost = S.logpost[chain]; R.logprior = S.logprior[chain];
    R.weight = S.weight[chain]; R.prior_rej = S.prior_rej[chain];
    R.burn_left = S.burn_left[chain]; R.added_w = S.added_w[chain];
    R.n_rows = S.n_rows[chain]; R.n_acc = S.n_acc[chain]; R.flags = S.flags[chain];
    __syncwarp();

    if (!M.drag) {
        // ---- MCMC.get_new_sample_metropolis (mcmc.py:545-562)
        for (int s = 0; s < n_steps; ++s) {
            const uint64_t t = t0 + (uint64_t)s;
            int b = W.tape_main
                        ? W.tape_main[chain * W.len_main + (int64_t)(t - W.base_main)]
                        : W.const_main;
            for (int i = lane; i < D; i += 32) trial[i] = x[i];      // :556
            __syncwarp();
            if (!warp_block_proposal(M, W, chain, gid, t, 0, b, trial, v, vis, e0s, lane))
                R.flags |= CB2_FLAG_INTERNAL;                          // :557
            warp_reduce_periodic(M, trial, lane);                      // :558
            double tprior;
            double tl = warp_logpost(M, trial, tprior, tll, tder, z, lpk, lane);  // :559
            if (tprior != -CUDART_INF) user_nan_check(M, tll, R.flags);
            bool acc = metropolis_accept(M, gid, t, 0, tl, R.logpost);  // :560
            acc = __shfl_sync(FULLMASK, (int)acc, 0);
            warp_process(M, S, chain, R, acc, x, der, ll, trial, tder, tll, tl, tprior,
                         lane);                                        // :561
        }
    } else {
        // ---- MCMC.get_new_sample_dragging (mcmc.py:564-668)
        double *e_pt = base + L.e_pt, *ps = base + L.ps, *pe = base + L.pe,
               *delta = base + L.delta, *e_der = base + L.e_der, *pe_der = base + L.pe_der,
               *e_ll = base + L.e_ll, *pe_ll = base + L.pe_ll, *tmp_ll = base + L.tmp_ll;
        double *s_pt = trial;  // "current_start_point"
        const int nds = M.drag_steps;
        for (int s = 0; s < n_steps; ++s) {
            const uint64_t t = t0 + (uint64_t)s;
            for (int i = lane; i < D; i += 32) { s_pt[i] = x[i]; e_pt[i] = x[i]; }  // :579-581
            __syncwarp();
            double s_lp = R.logpost;                                   // :580
            int b = W.tape_slow
                        ? W.tape_slow[chain * W.len_slow + (int64_t)(t - W.base_slow)]
                        : W.const_slow;
            if (!warp_block_proposal(M, W, chain, gid, t, 0, b, e_pt, v, vis, e0s, lane))
                R.flags |= CB2_FLAG_INTERNAL;                          // :582
            warp_reduce_periodic(M, e_pt, lane);                       // :583
            double e_prior;
            double e_lp = warp_logpost(M, e_pt, e_prior, e_ll, e_der, z, lpk, lane);  // :589
            if (e_prior != -CUDART_INF) user_nan_check(M, e_ll, R.flags);
            if (e_lp == -CUDART_INF) {                                 // :590-592
                R.weight += 1;  // the fast cycler is not advanced on this path
                continue;
            }
            double s_acc = s_lp, e_acc = e_lp;                         // :595-596
            for (int i = 1; i <= nds; ++i) {                           // :603
                for (int k = lane; k < D; k += 32) delta[k] = 0.0;     // :606
                __syncwarp();
                int bf;
                {
                    int64_t fi = vis[M.n_blocks];  // fast-cycler visit counter
                    int64_t rel = fi - e0s[M.n_blocks];
                    bool inr = rel >= 0 && rel < W.len_fast;
                    bf = W.tape_fast ? (inr ? W.tape_fast[chain * W.len_fast + rel] : 0)
                                     : W.const_fast;
                    if (W.tape_fast && !inr) R.flags |= CB2_FLAG_INTERNAL;
                    __syncwarp();
                    if (lane == 0) vis[M.n_blocks] = fi + 1;
                }
                if (!warp_block_proposal(M, W, chain, gid, t, (uint32_t)i, bf, delta, v, vis,
                                         e0s, lane))
                    R.flags |= CB2_FLAG_INTERNAL;                      // :607
                warp_reduce_periodic(M, delta, lane);                  // :608
                for (int k = lane; k < D; k += 32) ps[k] = s_pt[k] + delta[k];  // :610
                __syncwarp();
                double ps_prior;
#ifdef CB2_NVRTC_USER
                // built-in parts of both points, then the user's functions of the two points
                // side by side (lanes 0 and 1); delta is free once ps and pe exist
                for (int k = lane; k < D; k += 32) pe[k] = e_pt[k] + delta[k];      // :622
                __syncwarp();
                double pe_prior;
                double ps_lp = warp_logpost(M, ps, ps_prior, tmp_ll, nullptr, z, lpk, lane, -1, true);
                double pe_lp = -CUDART_INF;
                pe_prior = -CUDART_INF;
                if (ps_lp != -CUDART_INF)
                    pe_lp = warp_logpost(M, pe, pe_prior, pe_ll, pe_der, z, lpk, lane, -1, true);
                user_pair_eval(M, ps, pe, z, delta, tmp_ll, pe_ll, ps_lp, pe_lp, R.flags, lane);
                if (ps_lp != -CUDART_INF) {                            // :621
                    if (pe_lp != -CUDART_INF) {                        // :628
#else
                double ps_lp = warp_logpost(M, ps, ps_prior, tmp_ll, nullptr, z, lpk, lane);
                if (ps_lp != -CUDART_INF) {                            // :621
                    for (int k = lane; k < D; k += 32) pe[k] = e_pt[k] + delta[k];  // :622
                    __syncwarp();
                    double pe_prior;
                    double pe_lp = warp_logpost(M, pe, pe_prior, pe_ll, pe_der, z, lpk, lane);
                    if (pe_lp != -CUDART_INF) {                        // :628
#endif
                        double frac = (double)i / (double)(1 + nds);   // :630
                        double p_int = (1 - frac) * ps_lp + frac * pe_lp;
                        double c_int = (1 - frac) * s_lp + frac * e_lp;
                        bool ad = metropolis_accept(M, gid, t, (uint32_t)i, p_int, c_int);
                        ad = __shfl_sync(FULLMASK, (int)ad, 0);
                        if (ad) {                                      // :640-645
                            for (int k = lane; k < D; k += 32) { s_pt[k] = ps[k]; e_pt[k] = pe[k]; }
                            for (int k = lane; k < ND; k += 32) e_der[k] = pe_der[k];
                            for (int k = lane; k < NL; k += 32) e_ll[k] = pe_ll[k];
                            s_lp = ps_lp; e_lp = pe_lp; e_prior = pe_prior;
                            __syncwarp();
                        }
                    }
                }
                s_acc += s_lp;                                         // :655-656
                e_acc += e_lp;
            }
            double navg = (double)(1 + nds);                           // :658
            bool acc = metropolis_accept(M, gid, t, 0, e_acc / navg, s_acc / navg);
            acc = __shfl_sync(FULLMASK, (int)acc, 0);
            warp_process(M, S, chain, R, acc, x, der, ll, e_pt, e_der, e_ll, e_lp, e_prior,
                         lane);                                        // :666
        }
    }
    __syncwarp();
    for (int i = lane; i < D; i += 32) S.x[chain * D + i] = x[i];
    for (int i = lane; i < ND; i += 32) S.der[chain * ND + i] = der[i];
    for (int i = lane; i < NL; i += 32) S.ll[chain * NL + i] = ll[i];
    for (int i = lane; i < NV; i += 32) S.vis[chain * NV + i] = vis[i];
    if (lane == 0) {
        S.logpost[chain] = R.logpost; S.logprior[chain] = R.logprior;
        S.weight[chain] = R.weight; S.prior_rej[chain] = R.prior_rej;
        S.burn_left[chain] = R.burn_left; S.added_w[chain] = R.added_w;
        S.n_rows[chain] = R.n_rows; S.n_acc[chain] = R.n_acc; S.flags[chain] = R.flags;
    }
}

#ifndef CB2_NVRTC_USER
__global__ void __launch_bounds__(256)
k_step_general(ModelDev M, ChainState S, WindowDev W, StepSmem L, int64_t n_chains,
               uint64_t t0, int n_steps) {
    step_general_body(M, S, W, L, n_chains, t0, n_steps);
}
#endif
