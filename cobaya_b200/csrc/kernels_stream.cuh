// kernels_stream.cuh -- the "streamed" Metropolis path: 64 < D <= 512, and D <= 64 for models
// the register-resident kernels do not take (several gaussian_mixture components) (sm_100a).
//
// By linearity of the proposal map and of the Gaussian's whitening map, everything on the
// proposal side of a Metropolis step (mcmc.py:545-562) is independent of the chain state:
//     delta_k = T (D o H)[:, k]          (proposal.py:224, unit radius)
//     w_k^m   = L_m^-1 P delta_k          (gaussian_mixture.py:148, per mixture mode m; several
//                                          components over disjoint parameters are embedded
//                                          block-diagonally in one D x D matrix per mode)
// for every direction k of every Haar basis of a window.  They are formed as batched matrix
// products on the FP64 tensor pipe, 8 directions per warp as the M dimension of m8n8k4 DMMA
// tiles (k_stream_products), and streamed once through HBM.  The accept chain itself
// (k_stream_accept, one warp per chain, lanes over coordinates) only needs
//     x' = x + r delta_k,   y'_m = y_m + r w_k^m,   |y'_m|^2,
// bounds and 1-D priors of x', the Metropolis test (mcmc.py:670-683) and the bookkeeping /
// row store (mcmc.py:685-748): O(D) work and 8 D (1 + modes) bytes per proposal, HBM-bound.
// y_m = L_m^-1 P (x - mu_m) is recomputed from x at every window start (k_stream_whiten).
// Unlike the register-resident kernels of kernels_fast.cuh nothing here is limited by the
// registers one warp has for a D-vector, so the same three kernels cover any block layout
// and up to 4 mixture modes.
#pragma once
#include "kernels_fast.cuh"

#define CB2_STREAM_MAX_D 512   /* 64 < D <= 128: DMMA products; above: plain GEMMs (cuBLAS) */
#define CB2_STREAM_FRAG_MAX_D 128
#define CB2_STREAM_MAX_MODES 4
#ifndef CB2_STREAM_STAGE
#define CB2_STREAM_STAGE 4   /* look-ahead (steps) of the accept kernel's cp.async staging; 0: registers */
#endif
#define CB2_STREAM_MAX_LIKES 3

struct StreamPackDesc {
    int NT, DP;       // DP = 8 NT >= D (block-sorted coordinates, zero padded)
    int n_modes;
    int tri_like;     // L^-1 P lower-triangular in sorted coordinates
    int blocks_T, blocks_A;
    int off_T;        // fragment-ordered T (lower-triangular blocks), see warp_matvec8
    int off_A;        // fragment-ordered L_m^-1 P per mode
    int off_mu;       // [modes][DP]
    int n_like;       // Gaussian-mixture components over disjoint parameter sets (<= 3)
    int like_modes[CB2_STREAM_MAX_LIKES];
    int off_c0, off_w;                                    // [like][CB2_STREAM_MAX_MODES]
    int off_likeof;   // [DP] int32: component owning whitened coordinate a (-1: padding)
    int off_lower, off_upper, off_loc, off_mls, off_isc, off_pa, off_pb;  // [DP]
    int off_kind;     // [DP] int32 pairs {prior kind, sampler index i_of_j (or -1)}
    int off_d1;       // [n_blocks][DP]        delta of a 1-parameter block: T[:, j0]
    int off_w1;       // [n_blocks][modes][DP] its whitened images
    int iofj_identity;
    int total;
};

// one or more gaussian_mixture components (no derived parameters) whose input parameters
// are disjoint and together cover all D sampled parameters (checked in pack_stream)
static inline bool stream_step_supported(const ModelDev &M, size_t n_likes) {
    if (M.drag || M.D > CB2_STREAM_MAX_D || M.any_periodic) return false;
    if (n_likes < 1 || n_likes > CB2_STREAM_MAX_LIKES) return false;
    int dims = 0;
    for (size_t l = 0; l < n_likes; ++l) {
        const LikeDev &L = M.likes[l];
        if (L.kind != 0 || L.derived || L.n_modes > CB2_STREAM_MAX_MODES) return false;
        dims += L.dim;
    }
    // above 128 parameters: one component with one mode (the accept kernel keeps 6 D-vectors
    // of a chain in the registers of one warp)
    if (M.D > CB2_STREAM_FRAG_MAX_D && (n_likes != 1 || M.likes[0].n_modes != 1)) return false;
    return dims == M.D;
}

// out[nt] += sum over column blocks m in [m_lo, m_hi] of Mat(nt, m) a[m]; B fragments come
// from global memory (shared by every warp of the GPU: L1/L2 resident), 4 tiles per batch.
template <int NT, bool TRI>
__device__ __forceinline__ void warp_matvec8_g(const double *__restrict__ frag, int lane,
                                               const double (&a)[NT][2], double (&out)[NT][2],
                                               int m_lo, int m_hi) {
    const double2 *f2 = reinterpret_cast<const double2 *>(frag) + lane;
#pragma unroll
    for (int m = 0; m < NT; ++m) {
        if (m < m_lo || m > m_hi) continue;  // warp-uniform
#pragma unroll
        for (int n0 = TRI ? (m & ~3) : 0; n0 < NT; n0 += 4) {
            double2 b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int nt = n0 + u;
                if (nt < NT && (!TRI || nt >= m)) {
                    const int blk = TRI ? (nt * (nt + 1)) / 2 + m : nt * NT + m;
                    b[u] = __ldg(f2 + blk * 32);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int nt = n0 + u;
                if (nt < NT && (!TRI || nt >= m)) dmma8x8x4(out[nt][0], out[nt][1], a[m][0], b[u].x);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int nt = n0 + u;
                if (nt < NT && (!TRI || nt >= m)) dmma8x8x4(out[nt][0], out[nt][1], a[m][1], b[u].y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_stream_products: one warp per 8 directions of one stored basis (task = chain x epoch).
//   basis : [task][k][n_b]      (k_basis_wy / k_basis_general store, R[:, k] contiguous)
//   delta : [task][k][DP]       T (R[:, k] placed at sorted coordinates j0 .. j0 + n_b)
//   wout  : [task][k][modes][DP]
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256, 1)
k_stream_products(const double *__restrict__ pack, StreamPackDesc P, const double *__restrict__ basis,
                  int n_b, int j0, int64_t n_tasks, double *__restrict__ delta,
                  double *__restrict__ wout) {
    constexpr int DP = NT * 8;
    const int lane = threadIdx.x & 31, q = lane >> 2, r = lane & 3;
    const int kgroups = (n_b + 7) >> 3;
    const int64_t wt = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wt >= n_tasks * kgroups) return;
    const int64_t task = wt / kgroups;
    const int k = (int)(wt % kgroups) * 8 + q;
    const bool kvalid = k < n_b;
    const double *Rk = basis + ((size_t)task * n_b + (kvalid ? k : 0)) * (size_t)n_b;
    const bool vec = ((n_b | j0) & 1) == 0;
    double un[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const int j = 8 * n + 2 * r - j0;
        un[n][0] = 0.0;
        un[n][1] = 0.0;
        if (kvalid) {
            if (vec) {
                if (j >= 0 && j < n_b) {
                    const double2 v2 = __ldg(reinterpret_cast<const double2 *>(Rk + j));
                    un[n][0] = v2.x;
                    un[n][1] = v2.y;
                }
            } else {
                if (j >= 0 && j < n_b) un[n][0] = __ldg(Rk + j);
                if (j + 1 >= 0 && j + 1 < n_b) un[n][1] = __ldg(Rk + j + 1);
            }
        }
    }
    const int m_lo = j0 >> 3, m_hi = (j0 + n_b - 1) >> 3;
    double dl[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { dl[n][0] = 0.0; dl[n][1] = 0.0; }
    warp_matvec8_g<NT, true>(pack + P.off_T, lane, un, dl, m_lo, m_hi);
    const size_t row = (size_t)task * n_b + k;
    if (kvalid) {
        double2 *o = reinterpret_cast<double2 *>(delta + row * DP) + r;
#pragma unroll
        for (int n = 0; n < NT; ++n) o[4 * n] = make_double2(dl[n][0], dl[n][1]);
    }
    for (int km = 0; km < P.n_modes; ++km) {
        double wv[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) { wv[n][0] = 0.0; wv[n][1] = 0.0; }
        const double *Ak = pack + P.off_A + (size_t)km * P.blocks_A * 64;
        // delta vanishes above the block (T is lower triangular in sorted coordinates)
        if (P.tri_like) warp_matvec8_g<NT, true>(Ak, lane, dl, wv, m_lo, NT - 1);
        else warp_matvec8_g<NT, false>(Ak, lane, dl, wv, m_lo, NT - 1);
        if (kvalid) {
            double2 *o = reinterpret_cast<double2 *>(wout + (row * P.n_modes + km) * DP) + r;
#pragma unroll
            for (int n = 0; n < NT; ++n) o[4 * n] = make_double2(wv[n][0], wv[n][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_stream_products1: the triangular-L^-1 case with the column sweeps of T and of the first
// mode's L^-1 P interleaved (further modes are swept from the delta kept in registers).
// After column m of T the tile delta[m] is final: it is stored and immediately feeds column m
// of L^-1 P, after which w[m] is final too.  Direction tiles are loaded two
// columns ahead instead of all at once, so a warp keeps 2 (not 3) D-vectors of accumulators
// live and 12 instead of 8 warps fit an SM.
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(128, 3)
k_stream_products1(const double *__restrict__ pack, StreamPackDesc P,
                   const double *__restrict__ basis, int n_b, int j0, int64_t n_tasks,
                   double *__restrict__ delta, double *__restrict__ wout) {
    constexpr int DP = NT * 8;
    const int lane = threadIdx.x & 31, q = lane >> 2, r = lane & 3;
    const int kgroups = (n_b + 7) >> 3;
    const int64_t wt = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wt >= n_tasks * kgroups) return;
    const int64_t task = wt / kgroups;
    const int k = (int)(wt % kgroups) * 8 + q;
    const bool kvalid = k < n_b;
    const double *Rk = basis + ((size_t)task * n_b + (kvalid ? k : 0)) * (size_t)n_b;
    const bool vec = ((n_b | j0) & 1) == 0;
    const int m_lo = j0 >> 3, m_hi = (j0 + n_b - 1) >> 3;
    auto load_un = [&](int m, double &u0, double &u1) {
        const int j = 8 * m + 2 * r - j0;
        u0 = 0.0;
        u1 = 0.0;
        if (kvalid && m >= m_lo && m <= m_hi) {
            if (vec) {
                if (j >= 0 && j < n_b) {
                    const double2 v2 = ldg_f64x2_early(reinterpret_cast<const double2 *>(Rk + j));
                    u0 = v2.x;
                    u1 = v2.y;
                }
            } else {
                if (j >= 0 && j < n_b) u0 = ldg_f64_early(Rk + j);
                if (j + 1 >= 0 && j + 1 < n_b) u1 = ldg_f64_early(Rk + j + 1);
            }
        }
    };
    const double2 *fT = reinterpret_cast<const double2 *>(pack + P.off_T) + lane;
    const double2 *fA = reinterpret_cast<const double2 *>(pack + P.off_A) + lane;
    const size_t row = (size_t)task * n_b + (kvalid ? k : 0);
    const int nmod = P.n_modes;
    double2 *od = reinterpret_cast<double2 *>(delta + row * DP) + r;
    double2 *ow = reinterpret_cast<double2 *>(wout + row * nmod * DP) + r;
    double dl[NT][2], wv[NT][2], un[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { dl[n][0] = dl[n][1] = wv[n][0] = wv[n][1] = 0.0; }
    load_un(0, un[0][0], un[0][1]);
    if (NT > 1) load_un(1, un[1][0], un[1][1]);
#pragma unroll
    for (int m = 0; m < NT; ++m) {
        if (m + 2 < NT) load_un(m + 2, un[m + 2][0], un[m + 2][1]);
        if (m >= m_lo && m <= m_hi) {  // warp-uniform
#pragma unroll
            for (int n0 = (m & ~3); n0 < NT; n0 += 4) {
                double2 b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) b[u] = __ldg(fT + ((nt * (nt + 1)) / 2 + m) * 32);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) dmma8x8x4(dl[nt][0], dl[nt][1], un[m][0], b[u].x);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) dmma8x8x4(dl[nt][0], dl[nt][1], un[m][1], b[u].y);
                }
            }
        }
        if (kvalid) od[4 * m] = make_double2(dl[m][0], dl[m][1]);
        if (m >= m_lo) {  // delta vanishes above the block
#pragma unroll
            for (int n0 = (m & ~3); n0 < NT; n0 += 4) {
                double2 b[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) b[u] = __ldg(fA + ((nt * (nt + 1)) / 2 + m) * 32);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) dmma8x8x4(wv[nt][0], wv[nt][1], dl[m][0], b[u].x);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int nt = n0 + u;
                    if (nt < NT && nt >= m) dmma8x8x4(wv[nt][0], wv[nt][1], dl[m][1], b[u].y);
                }
            }
        }
        if (kvalid) ow[4 * m] = make_double2(wv[m][0], wv[m][1]);
    }
    // further mixture modes: column sweeps of L_m^-1 P over the delta kept in registers
    for (int km = 1; km < nmod; ++km) {
        const double2 *fAk = fA + (size_t)km * P.blocks_A * 32;
        double2 *owk = ow + (size_t)km * (DP / 2);
#pragma unroll
        for (int n = 0; n < NT; ++n) { wv[n][0] = 0.0; wv[n][1] = 0.0; }
#pragma unroll
        for (int m = 0; m < NT; ++m) {
            if (m >= m_lo) {
#pragma unroll
                for (int n0 = (m & ~3); n0 < NT; n0 += 4) {
                    double2 b[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nt = n0 + u;
                        if (nt < NT && nt >= m) b[u] = __ldg(fAk + ((nt * (nt + 1)) / 2 + m) * 32);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nt = n0 + u;
                        if (nt < NT && nt >= m) dmma8x8x4(wv[nt][0], wv[nt][1], dl[m][0], b[u].x);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int nt = n0 + u;
                        if (nt < NT && nt >= m) dmma8x8x4(wv[nt][0], wv[nt][1], dl[m][1], b[u].y);
                    }
                }
            }
            if (kvalid) owk[4 * m] = make_double2(wv[m][0], wv[m][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_stream_whiten: y_m = L_m^-1 P (x - mu_m) of every chain at window start, 8 chains per warp
// on the tensor pipe.  ys : [chain][modes][DP]
// ---------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(256, 1)
k_stream_whiten(const double *__restrict__ pack, StreamPackDesc P, const double *__restrict__ x,
                int D, int64_t n_chains, double *__restrict__ ys) {
    constexpr int DP = NT * 8;
    const int lane = threadIdx.x & 31, q = lane >> 2, r = lane & 3;
    const int64_t tile = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile * 8 >= n_chains) return;
    const int64_t chain_raw = tile * 8 + q;
    const bool active = chain_raw < n_chains;
    const int64_t chain = active ? chain_raw : n_chains - 1;
    const int2 *kind = reinterpret_cast<const int2 *>(pack + P.off_kind);
    double xs[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = __ldg(&kind[8 * n + 2 * r + h]).y;
            xs[n][h] = (i >= 0) ? x[chain * D + i] : 0.0;
        }
    for (int km = 0; km < P.n_modes; ++km) {
        const double *mu = pack + P.off_mu + km * DP;
        double z[NT][2], y[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const double2 m2 = __ldg(reinterpret_cast<const double2 *>(mu + 8 * n + 2 * r));
            z[n][0] = xs[n][0] - m2.x;
            z[n][1] = xs[n][1] - m2.y;
            y[n][0] = 0.0;
            y[n][1] = 0.0;
        }
        const double *Ak = pack + P.off_A + (size_t)km * P.blocks_A * 64;
        if (P.tri_like) warp_matvec8_g<NT, true>(Ak, lane, z, y, 0, NT - 1);
        else warp_matvec8_g<NT, false>(Ak, lane, z, y, 0, NT - 1);
        if (active) {
            double2 *o = reinterpret_cast<double2 *>(ys + ((size_t)chain * P.n_modes + km) * DP) + r;
#pragma unroll
            for (int n = 0; n < NT; ++n) o[4 * n] = make_double2(y[n][0], y[n][1]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// k_stream_accept: the serial accept chain of a window, one warp per chain.  Lane l owns the
// sorted coordinates l + 32 c, c < NC.
// ---------------------------------------------------------------------------------------
struct StreamWindow {
    const double *delta[CB2_MAX_BLOCKS];  // per block (n_b >= 2): [task][k][DP]
    const double *wv[CB2_MAX_BLOCKS];     // [task][k][modes][DP]
};

__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (capping the registers for 5 CTAs/SM was measured: the spills cost what the occupancy gains)
template <int NC, int NM, int NL>
__global__ void __launch_bounds__(128, (NM >= 3) ? 3 : 1)
k_stream_accept(ModelDev M, ChainState S, StreamWindow SW, const double *__restrict__ pack,
                StreamPackDesc P, const double2 *__restrict__ draws,
                const int2 *__restrict__ plan, const double *__restrict__ ys_in,
                int64_t n_chains, int n_steps) {
    const int lane = threadIdx.x & 31;
    const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= n_chains) return;
    const int D = M.D, DP = P.DP;
    const int2 *kind = reinterpret_cast<const int2 *>(pack + P.off_kind);
    const int *likeof = reinterpret_cast<const int *>(pack + P.off_likeof);
    double xs[NC], ys[NM][NC], lo[NC], up[NC];
    int iof[NC], knd[NC], lko[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int j = lane + 32 * c;
        iof[c] = -1; knd[c] = 0; xs[c] = 0.0; lo[c] = -CUDART_INF; up[c] = CUDART_INF;
        lko[c] = (NL > 1 && j < DP) ? __ldg(&likeof[j]) : 0;
        if (j < DP) {
            const int2 kk = __ldg(&kind[j]);
            knd[c] = kk.x; iof[c] = kk.y;
            lo[c] = __ldg(pack + P.off_lower + j);
            up[c] = __ldg(pack + P.off_upper + j);
            if (kk.y >= 0) xs[c] = S.x[chain * D + kk.y];
        }
#pragma unroll
        for (int m = 0; m < NM; ++m)
            ys[m][c] = (j < DP && m < P.n_modes) ? ys_in[((size_t)chain * P.n_modes + m) * DP + j] : 0.0;
    }
    double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) loglike[l] = S.ll[chain * NL + l];
    long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
              burn_left = S.burn_left[chain], added_w = S.added_w[chain],
              n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
    uint32_t flags = 0;
    const double2 *my_draws = draws + chain * (int64_t)n_steps;
    const int2 *my_plan = plan + chain * (int64_t)n_steps;
    const bool temp_one = M.temperature == 1.0;

#if CB2_STREAM_STAGE
    // The proposal-side vectors are staged through shared memory with cp.async, CB2_STREAM_STAGE
    // steps ahead of their use: every lane copies exactly the elements it will read itself, so
    // no cross-lane synchronisation is needed, only cp.async.wait_group.  HBM latency (the
    // chain is serial) is hidden by the look-ahead instead of by occupancy.
    constexpr int KS = CB2_STREAM_STAGE;
    constexpr int VS = (1 + NM) * NC;                     // doubles per lane and step
    extern __shared__ __align__(16) double stage_sm[];
    double *my_stage = stage_sm + (size_t)(threadIdx.x >> 5) * KS * VS * 32 + lane;
    int2 pl_q = __ldg(my_plan);                           // plan entry of the next step to issue
    auto issue = [&](int t) {
        if (t < n_steps) {
            const int2 pl = pl_q;
            if (t + 1 < n_steps) pl_q = __ldg(my_plan + t + 1);
            const int b = pl.y;
            const double *dp, *wp;
            if (M.bsize[b] >= 2) {
                dp = SW.delta[b] + (size_t)pl.x * DP;
                wp = SW.wv[b] + (size_t)pl.x * P.n_modes * DP;
            } else {
                dp = pack + P.off_d1 + (size_t)b * DP;
                wp = pack + P.off_w1 + (size_t)b * P.n_modes * DP;
            }
            double *slot = my_stage + (size_t)(t % KS) * VS * 32;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int j = lane + 32 * c;
                if (j < DP) {
                    const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(slot + (size_t)c * 32);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d0), "l"(dp + j));
#pragma unroll
                    for (int m = 0; m < NM; ++m)
                        if (m < P.n_modes) {
                            const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(
                                slot + (size_t)((1 + m) * NC + c) * 32);
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d1),
                                         "l"(wp + (size_t)m * DP + j));
                        }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");  // one group per step, even if empty
    };
#pragma unroll
    for (int t = 0; t < KS; ++t) issue(t);
    double2 dr = __ldg(my_draws);
    for (int s = 0; s < n_steps; ++s) {
        const double rs = dr.x * M.proposal_scale, e_acc = dr.y;
        if (s + 1 < n_steps) dr = __ldg(my_draws + s + 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(KS - 1) : "memory");
        double dl[NC], wl[NM][NC];
        {
            const double *slot = my_stage + (size_t)(s % KS) * VS * 32;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int j = lane + 32 * c;
                dl[c] = (j < DP) ? slot[(size_t)c * 32] : 0.0;
#pragma unroll
                for (int m = 0; m < NM; ++m)
                    wl[m][c] = (j < DP && m < P.n_modes) ? slot[(size_t)((1 + m) * NC + c) * 32] : 0.0;
            }
        }
        issue(s + KS);  // refills the slot just read
#else
    // the proposal-side vectors of step s + 1 are loaded during step s
    double dn[NC], wn[NM][NC];
    double2 dr;
    auto fetch = [&](int s) {
        const int2 pl = __ldg(my_plan + s);
        dr = __ldg(my_draws + s);
        const int b = pl.y;
        const double *dp, *wp;
        if (M.bsize[b] >= 2) {
            dp = SW.delta[b] + (size_t)pl.x * DP;
            wp = SW.wv[b] + (size_t)pl.x * P.n_modes * DP;
        } else {
            dp = pack + P.off_d1 + (size_t)b * DP;
            wp = pack + P.off_w1 + (size_t)b * P.n_modes * DP;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = lane + 32 * c;
            dn[c] = (j < DP) ? __ldg(dp + j) : 0.0;
#pragma unroll
            for (int m = 0; m < NM; ++m)
                wn[m][c] = (j < DP && m < P.n_modes) ? __ldg(wp + (size_t)m * DP + j) : 0.0;
        }
    };
    fetch(0);
    for (int s = 0; s < n_steps; ++s) {
        const double rs = dr.x * M.proposal_scale, e_acc = dr.y;
        double dl[NC], wl[NM][NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            dl[c] = dn[c];
#pragma unroll
            for (int m = 0; m < NM; ++m) wl[m][c] = wn[m][c];
        }
        if (s + 1 < n_steps) fetch(s + 1);
#endif
        // ---- trial point, bounds and priors (prior.py:733-763)
        bool bad = false;
        double ps = 0.0, qs[NL][NM];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int m = 0; m < NM; ++m) qs[l][m] = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double xt = fma(rs, dl[c], xs[c]);
            if (!(xt <= up[c]) || !(xt >= lo[c]) || !isfinite(xt)) bad = true;
            if (M.any_normal && knd[c] != 0) {
                const int j = lane + 32 * c;
                const double zz = (xt - __ldg(pack + P.off_loc + j)) / __ldg(pack + P.off_isc + j);
                if (knd[c] == 1) ps += __ldg(pack + P.off_mls + j) - zz * zz / 2;
                else ps += __ldg(pack + P.off_mls + j) +
                           prior1d_shape(knd[c], zz, __ldg(pack + P.off_pa + j),
                                         __ldg(pack + P.off_pb + j));
            }
            dl[c] = xt;  // keep the trial coordinates
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const double yt = fma(rs, wl[m][c], ys[m][c]);
                const double y2 = yt * yt;
                wl[m][c] = yt;
                // whitened coordinate a = lane + 32 c belongs to one component
#pragma unroll
                for (int l = 0; l < NL; ++l)
                    if (NL == 1 || lko[c] == l) qs[l][m] += y2;
            }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (M.any_normal) ps = warp_sum_all(ps);
        const double t_prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
        // ---- GaussianMixture.logp (gaussian_mixture.py:138-163), one term per component.
        // The log-sum-exp terms are spread over the lanes (lane 4 l + m owns mode m of
        // component l): one exp and one log per step instead of one per mode and component.
        double t_ll[NL], t_like = 0.0;
        {
            double lp[NL][NM], mx[NL];
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const int nm = P.like_modes[l];
                mx[l] = -CUDART_INF;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    lp[l][m] = (m < nm)
                                   ? -0.5 * (__ldg(pack + P.off_c0 + l * CB2_STREAM_MAX_MODES + m) +
                                             warp_sum_all(qs[l][m]))
                                   : -CUDART_INF;
                    mx[l] = fmax(mx[l], lp[l][m]);
                }
            }
            if (NM == 1) {
#pragma unroll
                for (int l = 0; l < NL; ++l) t_ll[l] = lp[l][0];
            } else {
                const int l_me = lane >> 2, m_me = lane & 3;
                double my_lp = -CUDART_INF, my_mx = 0.0;
#pragma unroll
                for (int l = 0; l < NL; ++l)
#pragma unroll
                    for (int m = 0; m < NM; ++m)
                        if (l_me == l && m_me == m) { my_lp = lp[l][m]; my_mx = mx[l]; }
                double e = 0.0;
                if (l_me < NL && m_me < NM && my_lp != -CUDART_INF)
                    e = __ldg(pack + P.off_w + l_me * CB2_STREAM_MAX_MODES + m_me) *
                        exp(my_lp - my_mx);
                e += __shfl_xor_sync(0xffffffffu, e, 1);
                e += __shfl_xor_sync(0xffffffffu, e, 2);
                const double ll_me = log(e) + my_mx;
#pragma unroll
                for (int l = 0; l < NL; ++l) {
                    const double v = __shfl_sync(0xffffffffu, ll_me, 4 * l);
                    // a single-mode component keeps its exact value (no exp/log round trip)
                    t_ll[l] = (P.like_modes[l] == 1) ? lp[l][0]
                                                     : (mx[l] == -CUDART_INF ? -CUDART_INF : v);
                }
            }
#pragma unroll
            for (int l = 0; l < NL; ++l) t_like += t_ll[l];
        }
        const double t_post = bad ? -CUDART_INF : (t_prior + t_like);
        // ---- metropolis_accept (mcmc.py:670-683)
        bool acc;
        if (t_post == -CUDART_INF) acc = false;
        else if (t_post > logpost) acc = true;
        else {
            const double dlp = logpost - t_post;
            acc = e_acc > (temp_one ? dlp : dlp / M.temperature);
        }
        // ---- process_accept_or_reject (mcmc.py:685-748); the whole warp is one chain
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
                        double *row = S.rows + ((size_t)chain * S.cap + n_rows) * M.width;
                        if (lane == 0) {
                            row[0] = (double)wst;
                            row[1] = temp_one ? -logpost : -(logpost / M.temperature);
                        } else if (lane == 1) {
                            row[2 + D] = -logprior;
                            row[3 + D] = -logprior;
                        } else if (lane == 2) {
                            double tot = 0.0;
#pragma unroll
                            for (int l = 0; l < NL; ++l) {
                                tot += loglike[l];
                                row[5 + D + l] = -2 * loglike[l];
                            }
                            row[4 + D] = -2 * tot;
                        }
#pragma unroll
                        for (int c = 0; c < NC; ++c)
                            if (iof[c] >= 0) row[2 + iof[c]] = xs[c];
                        n_rows += 1;
                    }
                }
            } else burn_left -= 1;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                xs[c] = dl[c];
#pragma unroll
                for (int m = 0; m < NM; ++m) ys[m][c] = wl[m][c];
            }
            logpost = t_post; logprior = t_prior;
#pragma unroll
            for (int l = 0; l < NL; ++l) loglike[l] = t_ll[l];
            weight = 1; prior_rej = 0; n_acc += 1;
        } else {
            weight += 1;
            if (t_prior == -CUDART_INF) prior_rej += 1;
            const long long sgn = (burn_left > 0) - (burn_left < 0);
            if (weight - prior_rej > M.max_tries * (1 + 9 * sgn)) flags |= CB2_FLAG_STUCK;
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (iof[c] >= 0) S.x[chain * D + iof[c]] = xs[c];
    if (lane == 0) {
        S.logpost[chain] = logpost; S.logprior[chain] = logprior;
#pragma unroll
        for (int l = 0; l < NL; ++l) S.ll[chain * NL + l] = loglike[l];
        S.weight[chain] = weight; S.prior_rej[chain] = prior_rej;
        S.burn_left[chain] = burn_left; S.added_w[chain] = added_w;
        S.n_rows[chain] = n_rows; S.n_acc[chain] = n_acc;
        if (flags) atomicOr(&S.flags[chain], flags);
    }
}

// ---------------------------------------------------------------------------------------
// 128 < D <= 512: the proposal-side products are plain D x D x (directions) GEMMs with one
// shared matrix and go to k_rowgemm (kernels_gemm.cuh, FP64 tensor pipe); only the glue is here.
// k_stream_center: z[chain][j] = x_sorted[j] - mu[j] (input of the whitening GEMM).
// k_stream_accept_big<NC>: k_stream_accept for one single-mode component with the per-coordinate
// constants read from the constant block instead of registers (NC up to 16 coordinates per lane).
// ---------------------------------------------------------------------------------------
__global__ void k_stream_center(const double *__restrict__ pack, StreamPackDesc P,
                                const double *__restrict__ x, int D, int64_t n_chains,
                                double *__restrict__ z) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_chains * P.DP) return;
    const int64_t chain = e / P.DP;
    const int j = (int)(e % P.DP);
    const int i = __ldg(reinterpret_cast<const int2 *>(pack + P.off_kind) + j).y;
    z[e] = (i >= 0) ? x[chain * D + i] - __ldg(pack + P.off_mu + j) : 0.0;
}

template <int NC>
__global__ void __launch_bounds__(128)
k_stream_accept_big(ModelDev M, ChainState S, StreamWindow SW, const double *__restrict__ pack,
                    StreamPackDesc P, const double2 *__restrict__ draws,
                    const int2 *__restrict__ plan, const double *__restrict__ ys_in,
                    int64_t n_chains, int n_steps) {
    const int lane = threadIdx.x & 31;
    const int64_t chain = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= n_chains) return;
    const int D = M.D, DP = P.DP;
    const int2 *kind = reinterpret_cast<const int2 *>(pack + P.off_kind);
    const double *lower = pack + P.off_lower, *upper = pack + P.off_upper;
    double xs[NC], ys[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int j = lane + 32 * c;
        xs[c] = 0.0;
        ys[c] = 0.0;
        if (j < DP) {
            const int i = __ldg(&kind[j]).y;
            if (i >= 0) xs[c] = S.x[chain * D + i];
            ys[c] = ys_in[(size_t)chain * DP + j];
        }
    }
    double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike = S.ll[chain];
    long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
              burn_left = S.burn_left[chain], added_w = S.added_w[chain],
              n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
    uint32_t flags = 0;
    const double2 *my_draws = draws + chain * (int64_t)n_steps;
    const int2 *my_plan = plan + chain * (int64_t)n_steps;
    const bool temp_one = M.temperature == 1.0;
    const double c0 = __ldg(pack + P.off_c0);
    double dn[NC], wn[NC];
    double2 dr;
    auto fetch = [&](int s) {
        const int2 pl = __ldg(my_plan + s);
        dr = __ldg(my_draws + s);
        const int b = pl.y;
        const double *dp, *wp;
        if (M.bsize[b] >= 2) {
            dp = SW.delta[b] + (size_t)pl.x * DP;
            wp = SW.wv[b] + (size_t)pl.x * DP;
        } else {
            dp = pack + P.off_d1 + (size_t)b * DP;
            wp = pack + P.off_w1 + (size_t)b * DP;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = lane + 32 * c;
            dn[c] = (j < DP) ? __ldg(dp + j) : 0.0;
            wn[c] = (j < DP) ? __ldg(wp + j) : 0.0;
        }
    };
    fetch(0);
    for (int s = 0; s < n_steps; ++s) {
        const double rs = dr.x * M.proposal_scale, e_acc = dr.y;
        // trial point first (it replaces the fetched vectors), then the next fetch
        double xt[NC], yt[NC];
        bool bad = false;
        double ps = 0.0, qs = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int j = lane + 32 * c;
            const double t = fma(rs, dn[c], xs[c]);
            xt[c] = t;
            if (j < DP) {
                if (!(t <= __ldg(upper + j)) || !(t >= __ldg(lower + j)) || !isfinite(t)) bad = true;
                if (M.any_normal) {
                    const int kd = __ldg(&kind[j]).x;
                    if (kd != 0) {
                        const double zz = (t - __ldg(pack + P.off_loc + j)) / __ldg(pack + P.off_isc + j);
                        if (kd == 1) ps += __ldg(pack + P.off_mls + j) - zz * zz / 2;
                        else ps += __ldg(pack + P.off_mls + j) +
                                   prior1d_shape(kd, zz, __ldg(pack + P.off_pa + j),
                                                 __ldg(pack + P.off_pb + j));
                    }
                }
            }
            const double y = fma(rs, wn[c], ys[c]);
            yt[c] = y;
            qs = fma(y, y, qs);
        }
        if (s + 1 < n_steps) fetch(s + 1);
        bad = __any_sync(0xffffffffu, bad);
        if (M.any_normal) ps = warp_sum_all(ps);
        const double t_prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
        const double t_like = -0.5 * (c0 + warp_sum_all(qs));
        const double t_post = bad ? -CUDART_INF : (t_prior + t_like);
        bool acc;
        if (t_post == -CUDART_INF) acc = false;
        else if (t_post > logpost) acc = true;
        else {
            const double dlp = logpost - t_post;
            acc = e_acc > (temp_one ? dlp : dlp / M.temperature);
        }
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
                        double *row = S.rows + ((size_t)chain * S.cap + n_rows) * M.width;
                        if (lane == 0) {
                            row[0] = (double)wst;
                            row[1] = temp_one ? -logpost : -(logpost / M.temperature);
                        } else if (lane == 1) {
                            row[2 + D] = -logprior;
                            row[3 + D] = -logprior;
                        } else if (lane == 2) {
                            row[4 + D] = -2 * loglike;
                            row[5 + D] = -2 * loglike;
                        }
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const int j = lane + 32 * c;
                            if (j < DP) {
                                const int i = __ldg(&kind[j]).y;
                                if (i >= 0) row[2 + i] = xs[c];
                            }
                        }
                        n_rows += 1;
                    }
                }
            } else burn_left -= 1;
#pragma unroll
            for (int c = 0; c < NC; ++c) { xs[c] = xt[c]; ys[c] = yt[c]; }
            logpost = t_post; logprior = t_prior; loglike = t_like;
            weight = 1; prior_rej = 0; n_acc += 1;
        } else {
            weight += 1;
            if (t_prior == -CUDART_INF) prior_rej += 1;
            const long long sgn = (burn_left > 0) - (burn_left < 0);
            if (weight - prior_rej > M.max_tries * (1 + 9 * sgn)) flags |= CB2_FLAG_STUCK;
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int j = lane + 32 * c;
        if (j < DP) {
            const int i = __ldg(&kind[j]).y;
            if (i >= 0) S.x[chain * D + i] = xs[c];
        }
    }
    if (lane == 0) {
        S.logpost[chain] = logpost; S.logprior[chain] = logprior; S.ll[chain] = loglike;
        S.weight[chain] = weight; S.prior_rej[chain] = prior_rej;
        S.burn_left[chain] = burn_left; S.added_w[chain] = added_w;
        S.n_rows[chain] = n_rows; S.n_acc[chain] = n_acc;
        if (flags) atomicOr(&S.flags[chain], flags);
    }
}

static int launch_stream_accept_big(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                    const StreamWindow &SW, const double *pack,
                                    const StreamPackDesc &P, const double2 *draws,
                                    const int2 *plan, const double *ys, int64_t n_chains,
                                    int n_steps) {
    const unsigned grid = (unsigned)((n_chains + 3) / 4);
    const int NC = (P.DP + 31) / 32;
#define CB2_SAB(C_)                                                                          \
    if (NC <= C_) {                                                                          \
        k_stream_accept_big<C_><<<grid, 128, 0, st>>>(M, S, SW, pack, P, draws, plan, ys,    \
                                                      n_chains, n_steps);                    \
        return cudaGetLastError() == cudaSuccess ? 0 : -2;                                   \
    }
    CB2_SAB(6) CB2_SAB(8) CB2_SAB(12) CB2_SAB(16)
#undef CB2_SAB
    return -1;
}

// ---------------------------------------------------------------- launchers
static int launch_stream_products(cudaStream_t st, const double *pack, const StreamPackDesc &P,
                                  const double *basis, int n_b, int j0, int64_t n_tasks,
                                  double *delta, double *wout) {
    const int64_t wts = n_tasks * ((n_b + 7) / 8);
    const unsigned grid = (unsigned)((wts + 7) / 8);
    const bool one = P.tri_like != 0;  // interleaved sweeps (triangular L^-1 P), 12 warps/SM
    const unsigned grid1 = (unsigned)((wts + 3) / 4);
#define CB2_SP(N)                                                                            \
    case N:                                                                                  \
        if (one)                                                                             \
            k_stream_products1<N><<<grid1, 128, 0, st>>>(pack, P, basis, n_b, j0, n_tasks,   \
                                                         delta, wout);                       \
        else                                                                                 \
            k_stream_products<N><<<grid, 256, 0, st>>>(pack, P, basis, n_b, j0, n_tasks,     \
                                                       delta, wout);                         \
        break;
    switch (P.NT) {
        CB2_SP(1) CB2_SP(2) CB2_SP(3) CB2_SP(4) CB2_SP(5) CB2_SP(6) CB2_SP(7) CB2_SP(8)
        CB2_SP(9) CB2_SP(10) CB2_SP(11) CB2_SP(12) CB2_SP(13) CB2_SP(14) CB2_SP(15) CB2_SP(16)
        default: return -1;
    }
#undef CB2_SP
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

static int launch_stream_whiten(cudaStream_t st, const double *pack, const StreamPackDesc &P,
                                const double *x, int D, int64_t n_chains, double *ys) {
    const int64_t tiles = (n_chains + 7) / 8;
    const unsigned grid = (unsigned)((tiles + 7) / 8);
#define CB2_SW(N)                                                                        \
    case N:                                                                              \
        k_stream_whiten<N><<<grid, 256, 0, st>>>(pack, P, x, D, n_chains, ys);           \
        break;
    switch (P.NT) {
        CB2_SW(1) CB2_SW(2) CB2_SW(3) CB2_SW(4) CB2_SW(5) CB2_SW(6) CB2_SW(7) CB2_SW(8)
        CB2_SW(9) CB2_SW(10) CB2_SW(11) CB2_SW(12) CB2_SW(13) CB2_SW(14) CB2_SW(15) CB2_SW(16)
        default: return -1;
    }
#undef CB2_SW
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

static int launch_stream_accept(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                const StreamWindow &SW, const double *pack,
                                const StreamPackDesc &P, const double2 *draws, const int2 *plan,
                                const double *ys, int64_t n_chains, int n_steps) {
    const unsigned grid = (unsigned)((n_chains + 3) / 4);
    const int NC = (P.DP + 31) / 32;
    const int NM = P.n_modes;  // 1..4
    const int NL = P.n_like;
#define CB2_SA(C_, M_, L_)                                                                  \
    if (NC == C_ && NM == M_ && NL == L_) {                                                 \
        const size_t sm = (size_t)CB2_STREAM_STAGE * (1 + M_) * C_ * 32 * 8 * 4;            \
        if (sm > 48 * 1024 &&                                                               \
            cudaFuncSetAttribute(k_stream_accept<C_, M_, L_>,                               \
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) !=   \
                cudaSuccess)                                                                \
            return -3;                                                                      \
        k_stream_accept<C_, M_, L_><<<grid, 128, sm, st>>>(M, S, SW, pack, P, draws, plan,  \
                                                           ys, n_chains, n_steps);          \
        return cudaGetLastError() == cudaSuccess ? 0 : -2;                                  \
    }
#define CB2_SA_L(C_, M_) CB2_SA(C_, M_, 1) CB2_SA(C_, M_, 2) CB2_SA(C_, M_, 3)
    CB2_SA_L(1, 1) CB2_SA_L(1, 2) CB2_SA_L(1, 3) CB2_SA_L(1, 4)
    CB2_SA_L(2, 1) CB2_SA_L(2, 2) CB2_SA_L(2, 3) CB2_SA_L(2, 4)
    CB2_SA_L(3, 1) CB2_SA_L(3, 2) CB2_SA_L(3, 3) CB2_SA_L(3, 4)
    CB2_SA_L(4, 1) CB2_SA_L(4, 2) CB2_SA_L(4, 3) CB2_SA_L(4, 4)
#undef CB2_SA_L
#undef CB2_SA
    return -1;
}
