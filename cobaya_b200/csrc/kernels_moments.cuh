// kernels_moments.cuh -- per-GPU sufficient statistics of the convergence check
// (MCMC.check_convergence_and_learn_proposal, mcmc.py:773-889) and small reductions.
#pragma once
#include "common.cuh"

// A "task" is one (virtual) chain of the R-1 computation: rows [first, last) of a
// stored chain, counted with weight N in the mean of covariances (mcmc.py:791-822).
struct MomentTask {
    int64_t chain, first, last;
    double N;
};

// mode HALVES: task c = rows [n_c/2, n_c) of chain c, N = n_c   (mcmc.py:787-793)
__global__ void k_tasks_halves(const int64_t *__restrict__ n_rows, int64_t n_chains,
                               MomentTask *__restrict__ tasks) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= n_chains) return;
    int64_t n = n_rows[c];
    MomentTask t;
    t.chain = c;
    t.first = n / 2;  // int(self.n() / 2)
    t.last = n;
    t.N = (double)n;
    tasks[c] = t;
}

// Per task: Sw = sum w, mean m = sum w x / Sw over rows [first,last)
// (SampleCollection.mean, collection.py:893-935).  One warp per task.
__global__ void k_task_means(const double *__restrict__ rows, int64_t cap, int width, int D,
                             const MomentTask *__restrict__ tasks, int64_t n_tasks,
                             double *__restrict__ means, double *__restrict__ sw_out) {
    const int lane = threadIdx.x & 31;
    const int64_t t = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (t >= n_tasks) return;
    const MomentTask T = tasks[t];
    const double *base = rows + (size_t)T.chain * cap * width;
    double sw = 0.0;
    for (int64_t r = T.first; r < T.last; ++r) sw += base[(size_t)r * width];
    for (int d = lane; d < D; d += 32) {
        double acc = 0.0;
        for (int64_t r = T.first; r < T.last; ++r) {
            const double *row = base + (size_t)r * width;
            acc += row[0] * row[2 + d];
        }
        means[t * D + d] = acc / sw;
    }
    if (lane == 0) sw_out[t] = sw;
}

// Accumulate, over the tasks assigned to this CTA, the partial sums
//   P[0]=M, P[1]=sum N, P[2]=sum N*a, P[3..3+D)=sum (m-shift),
//   P[3+D..+D^2) = sum (m-shift)(m-shift)^T,  P[3+D+D^2..) = sum N * C
// with C = sum_r w_r (x_r-m)(x_r-m)^T / Sw  (np.cov(ddof=0, fweights), collection.py:968)
// and a = (#rows)/Sw (get_acceptance_rate, mcmc.py:311-318).
// Each thread owns a fixed set of (i,j) entries -> deterministic summation order.
// PER > 0: entries live in registers (PER per thread); PER == 0: in the global partial.
template <int PER>
__global__ void __launch_bounds__(256)
k_task_accumulate(const double *__restrict__ rows, int64_t cap, int width, int D,
                  const MomentTask *__restrict__ tasks, int64_t n_tasks,
                  const double *__restrict__ means, const double *__restrict__ sw_in,
                  const double *__restrict__ shift, double *__restrict__ partials) {
    extern __shared__ double sm[];
    double *xc = sm;       // [D] centred row
    double *ms = sm + D;   // [D] mean - shift
    const int tid = threadIdx.x, nt = blockDim.x;
    const int DD = D * D;
    const int per = (DD + nt - 1) / nt;
    double *P = partials + (size_t)blockIdx.x * (size_t)(3 + D + 2 * DD);
    for (int e = tid; e < 3 + D + 2 * DD; e += nt) P[e] = 0.0;
    __syncthreads();
    constexpr int NR = PER > 0 ? PER : 1;
    double cacc[NR], macc[NR];
    int ei[NR], ej[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        cacc[k] = 0.0;
        macc[k] = 0.0;
        int e = tid + k * nt;
        ei[k] = (e < DD) ? e / D : 0;
        ej[k] = (e < DD) ? e % D : 0;
    }
    double accM = 0.0, accN = 0.0, accNa = 0.0;
    for (int64_t t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const MomentTask T = tasks[t];
        const double sw = sw_in[t];
        const double *base = rows + (size_t)T.chain * cap * width;
        __syncthreads();
        for (int d = tid; d < D; d += nt) ms[d] = means[t * D + d] - (shift ? shift[d] : 0.0);
        __syncthreads();
        for (int d = tid; d < D; d += nt) P[3 + d] += ms[d];
        if (PER > 0) {
#pragma unroll
            for (int k = 0; k < NR; ++k) macc[k] += ms[ei[k]] * ms[ej[k]];
        } else {
            for (int k = 0; k < per; ++k) {
                int e = tid + k * nt;
                if (e < DD) P[3 + D + e] += ms[e / D] * ms[e % D];
            }
        }
        const double f = T.N / sw;
        for (int64_t r = T.first; r < T.last; ++r) {
            const double *row = base + (size_t)r * width;
            __syncthreads();
            for (int d = tid; d < D; d += nt) xc[d] = row[2 + d] - means[t * D + d];
            __syncthreads();
            const double wf = row[0] * f;
            if (PER > 0) {
#pragma unroll
                for (int k = 0; k < NR; ++k) cacc[k] += wf * xc[ei[k]] * xc[ej[k]];
            } else {
                for (int k = 0; k < per; ++k) {
                    int e = tid + k * nt;
                    if (e < DD) P[3 + D + DD + e] += wf * xc[e / D] * xc[e % D];
                }
            }
        }
        if (tid == 0) {
            accM += 1.0;
            accN += T.N;
            accNa += T.N * ((double)(T.last - T.first) / sw);
        }
    }
    __syncthreads();
    if (PER > 0) {
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            int e = tid + k * nt;
            if (e < DD) {
                P[3 + D + e] = macc[k];
                P[3 + D + DD + e] = cacc[k];
            }
        }
    }
    if (tid == 0) {
        P[0] = accM;
        P[1] = accN;
        P[2] = accNa;
    }
}

// out[e] = sum over CTAs (fixed order) of partials[cta][e]
__global__ void k_reduce_partials(const double *__restrict__ partials, int n_parts, int len,
                                  double *__restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= len) return;
    double acc = 0.0;
    for (int p = 0; p < n_parts; ++p) acc += partials[(size_t)p * len + e];
    out[e] = acc;
}

// out[0..7] = min rows, max rows, sum rows, #stuck, #rows-full, (unused), sum accepted,
// sum of current weights.  Single CTA.
__global__ void k_summary(const int64_t *__restrict__ n_rows, const int64_t *__restrict__ n_acc,
                          const int64_t *__restrict__ weight, const uint32_t *__restrict__ flags,
                          int64_t n_chains, int64_t *__restrict__ out) {
    __shared__ long long s_min[256], s_max[256], s_sum[256], s_stuck[256], s_full[256],
        s_acc[256], s_w[256], s_int[256];
    int tid = threadIdx.x;
    long long mn = INT64_MAX, mx = INT64_MIN, sm_ = 0, st = 0, fu = 0, ac = 0, ww = 0, in = 0;
    for (int64_t c = tid; c < n_chains; c += blockDim.x) {
        long long n = n_rows[c];
        mn = n < mn ? n : mn;
        mx = n > mx ? n : mx;
        sm_ += n;
        st += (flags[c] & 1u) ? 1 : 0;
        fu += (flags[c] & 2u) ? 1 : 0;
        in += (flags[c] & 4u) ? 1 : 0;
        ac += n_acc[c];
        ww += weight[c];
    }
    s_min[tid] = mn; s_max[tid] = mx; s_sum[tid] = sm_; s_stuck[tid] = st; s_full[tid] = fu;
    s_acc[tid] = ac; s_w[tid] = ww; s_int[tid] = in;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (tid < o) {
            s_min[tid] = s_min[tid + o] < s_min[tid] ? s_min[tid + o] : s_min[tid];
            s_max[tid] = s_max[tid + o] > s_max[tid] ? s_max[tid + o] : s_max[tid];
            s_sum[tid] += s_sum[tid + o]; s_stuck[tid] += s_stuck[tid + o];
            s_full[tid] += s_full[tid + o]; s_acc[tid] += s_acc[tid + o];
            s_w[tid] += s_w[tid + o]; s_int[tid] += s_int[tid + o];
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[0] = s_min[0]; out[1] = s_max[0]; out[2] = s_sum[0]; out[3] = s_stuck[0];
        out[4] = s_full[0]; out[5] = s_int[0]; out[6] = s_acc[0]; out[7] = s_w[0];
    }
}

// =====================================================================================
// DMMA checkpoint statistics (D <= 64): the proposal-covariance SYRK on the FP64 tensor
// pipe.  One pass over rows [first,last) of every task: with xr = x - ref (ref = first row
// of the window, removes the cancellation of an uncentred sum)
//     sw = sum w,  S1 = sum w xr,  S2 = sum w xr xr^T   (m8n8k4 tiles, K = rows)
//     m = ref + S1/sw,  C = S2/sw - (S1/sw)(S1/sw)^T          (== np.cov(ddof=0, fweights))
// then the same CTA-level partial sums as k_task_accumulate.  NT warps per CTA, warp w owns
// output rows 8w..8w+7 (all NT column tiles) in MMA C-fragment layout.
// =====================================================================================
#define CB2_MOM_BATCH 32
#define CB2_MOM_PER_THREAD 8   // CB2_MOM_BATCH*DP / (32*NT)

__device__ __forceinline__ void mom_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

template <int NT>
__global__ void __launch_bounds__(NT * 32, (NT >= 7) ? 2 : 1)
k_task_moments_dmma(const double *__restrict__ rows, int64_t cap, int width, int D,
                    const MomentTask *__restrict__ tasks, int64_t n_tasks,
                    const double *__restrict__ shift, double *__restrict__ partials,
                    double *__restrict__ sw_out) {
    constexpr int DP = NT * 8;
    constexpr int LDX = DP + 4;                    // padded row stride (bank spread)
    __shared__ double xt[CB2_MOM_BATCH][LDX];      // xr of the current batch
    __shared__ double wt[CB2_MOM_BATCH];           // weights
    __shared__ double refv[DP], s1v[DP], mrel[DP], msv[DP];
    __shared__ double sw_s;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5;
    const int q = lane >> 2, r = lane & 3;
    const int DD = D * D;
    double *P = partials + (size_t)blockIdx.x * (size_t)(3 + D + 2 * DD);
    double sc[NT][2], smm[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) { sc[n][0] = sc[n][1] = smm[n][0] = smm[n][1] = 0.0; }
    double accM = 0.0, accN = 0.0, accNa = 0.0, accm = 0.0;  // accm: thread d<D holds sum ms[d]
    for (int64_t t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const MomentTask T = tasks[t];
        const double *base = rows + (size_t)T.chain * cap * width;
        double s2[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) { s2[n][0] = 0.0; s2[n][1] = 0.0; }
        double s1 = 0.0, swl = 0.0;
        __syncthreads();
        if (tid < DP) refv[tid] = (tid < D) ? base[(size_t)T.first * width + 2 + tid] : 0.0;
        __syncthreads();
        // software pipeline: the rows of batch b+1 are fetched into registers while the
        // tensor pipe works on batch b (DRAM latency hidden behind the MMAs)
        double vals[CB2_MOM_PER_THREAD], wnext = 0.0;
        auto fetch = [&](int64_t r0) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
#pragma unroll
            for (int u = 0; u < CB2_MOM_PER_THREAD; ++u) {
                const int e = tid + u * (NT * 32);
                const int k = e / DP, d = e % DP;
                vals[u] = (k < nb && d < D) ? base[(size_t)(r0 + k) * width + 2 + d] : 0.0;
            }
            if (tid < CB2_MOM_BATCH) wnext = (tid < nb) ? base[(size_t)(r0 + tid) * width] : 0.0;
        };
        if (T.first < T.last) fetch(T.first);
        for (int64_t r0 = T.first; r0 < T.last; r0 += CB2_MOM_BATCH) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < CB2_MOM_PER_THREAD; ++u) {
                const int e = tid + u * (NT * 32);
                const int k = e / DP, d = e % DP;
                xt[k][d] = (k < nb && d < D) ? vals[u] - refv[d] : 0.0;
            }
            if (tid < CB2_MOM_BATCH) wt[tid] = wnext;
            __syncthreads();
            if (r0 + CB2_MOM_BATCH < T.last) fetch(r0 + CB2_MOM_BATCH);
            if (tid < DP) {
#pragma unroll 8
                for (int k = 0; k < CB2_MOM_BATCH; ++k) s1 = fma(wt[k], xt[k][tid], s1);
            }
            if (tid == 0)
                for (int k = 0; k < CB2_MOM_BATCH; ++k) swl += wt[k];
            // S2[8w+q][8n+2r+{0,1}] += sum_k (w_k xr_k[8w+q]) * xr_k[8n+..]
#pragma unroll
            for (int kk = 0; kk < CB2_MOM_BATCH / 4; ++kk) {
                const int k = 4 * kk + r;
                const double a = wt[k] * xt[k][8 * wid + q];
#pragma unroll
                for (int n = 0; n < NT; ++n) mom_dmma(s2[n][0], s2[n][1], a, xt[k][8 * n + q]);
            }
        }
        __syncthreads();
        if (tid == 0) sw_s = swl;
        if (tid < DP) s1v[tid] = s1;
        __syncthreads();
        const double sw = sw_s;
        if (tid < DP) {
            const double mr = s1v[tid] / sw;
            mrel[tid] = mr;
            const double ms = (tid < D) ? (refv[tid] + mr) - (shift ? shift[tid] : 0.0) : 0.0;
            msv[tid] = ms;
            accm += ms;
        }
        __syncthreads();
        const double f = T.N / sw;
        const int i = 8 * wid + q;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 8 * n + 2 * r + h;
                sc[n][h] += f * s2[n][h] - T.N * (mrel[i] * mrel[j]);
                smm[n][h] += msv[i] * msv[j];
            }
        if (tid == 0) {
            accM += 1.0;
            accN += T.N;
            accNa += T.N * ((double)(T.last - T.first) / sw);
            if (sw_out) sw_out[t] = sw;
        }
    }
    __syncthreads();
    {
        const int i = 8 * wid + q;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 8 * n + 2 * r + h;
                if (i < D && j < D) {
                    P[3 + D + i * D + j] = smm[n][h];
                    P[3 + D + DD + i * D + j] = sc[n][h];
                }
            }
    }
    if (tid < D) P[3 + tid] = accm;
    if (tid == 0) {
        P[0] = accM;
        P[1] = accN;
        P[2] = accNa;
    }
}

// =====================================================================================
// R-1 of the confidence-interval bounds (mcmc.py:918-1002)
// =====================================================================================
// The reference asks GetDist for `MCSamples.confidence(i, limfrac, upper)` of every chain:
// the raw weighted sample quantile -- sort the values, cumulate the weights, take the first
// sample whose cumulative weight reaches limfrac*norm (lower) or (1-limfrac)*norm (upper)
// (getdist/chains.py `confidence`, GetDist>=1.3.1; not vendored: parity unpinned).
// One CTA per (task, parameter): bitonic sort of (value, weight) in shared memory, inclusive
// scan of the weights, two binary searches.  Windows longer than the sort size go to
// k_task_bounds_select (below).
__global__ void __launch_bounds__(256)
k_task_bounds(const double *__restrict__ rows, int64_t cap, int width, int D,
              const MomentTask *__restrict__ tasks, int n_pow2, double limfrac,
              double *__restrict__ bounds /* [task][D][2] */) {
    extern __shared__ double bsm[];
    double *val = bsm;            // [n_pow2]
    double *wgt = bsm + n_pow2;   // [n_pow2]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t task = blockIdx.x / D;
    const int par = blockIdx.x % D;
    const MomentTask T = tasks[task];
    const double *base = rows + (size_t)T.chain * cap * width;
    const int64_t nrows = T.last - T.first;
    const int64_t stride = (nrows + n_pow2 - 1) / n_pow2;
    const int n = (int)((nrows + stride - 1) / stride);
    for (int e = tid; e < n_pow2; e += nt) {
        if (e < n) {
            const double *row = base + (size_t)(T.first + (int64_t)e * stride) * width;
            val[e] = row[2 + par];
            wgt[e] = row[0];
        } else {
            val[e] = CUDART_INF;
            wgt[e] = 0.0;
        }
    }
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int e = tid; e < n_pow2; e += nt) {
                const int p = e ^ j;
                if (p > e) {
                    const bool up = (e & k) == 0;
                    const double a = val[e], b = val[p];
                    if ((a > b) == up) {
                        val[e] = b; val[p] = a;
                        const double wa = wgt[e];
                        wgt[e] = wgt[p]; wgt[p] = wa;
                    }
                }
            }
            __syncthreads();
        }
    // inclusive scan of the weights (Hillis-Steele; integer-valued weights: exact)
    for (int off = 1; off < n_pow2; off <<= 1) {
        double add[32];
        int cnt = 0;
        for (int e = tid; e < n_pow2; e += nt) add[cnt++] = (e >= off) ? wgt[e - off] : 0.0;
        __syncthreads();
        cnt = 0;
        for (int e = tid; e < n_pow2; e += nt) wgt[e] += add[cnt++];
        __syncthreads();
    }
    if (tid < 2) {
        const double norm = wgt[n - 1];
        const double target = tid == 0 ? norm * limfrac : norm * (1.0 - limfrac);
        int lo = 0, hi = n;  // np.searchsorted(cumsum, target) (side='left')
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (wgt[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        if (lo > n - 1) lo = n - 1;
        bounds[((size_t)task * D + par) * 2 + tid] = val[lo];
    }
}

// Windows that do not fit the shared-memory sort: the same quantile by RADIX SELECTION on the
// order-preserving 64-bit key of the values -- 8 passes over the column, each building a
// 256-bin histogram of the weights of the rows that match the key prefix found so far, for
// the lower and the upper bound at once.  The answer is the smallest sample value v with
// W(values <= v) >= target, i.e. exactly what sorting + cumulating + searchsorted returns
// (integer-valued weights: the sums are exact in any order).  No thinning.
__device__ __forceinline__ unsigned long long f64_order_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_order_key(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256)
k_task_bounds_select(const double *__restrict__ rows, int64_t cap, int width, int D,
                     const MomentTask *__restrict__ tasks, double limfrac,
                     double *__restrict__ bounds /* [task][D][2] */) {
    __shared__ double hist[2][256];
    __shared__ double s_norm;
    __shared__ unsigned long long s_prefix[2];
    __shared__ double s_below[2];
    const int tid = threadIdx.x;
    const int64_t task = blockIdx.x / D;
    const int par = blockIdx.x % D;
    const MomentTask T = tasks[task];
    const double *base = rows + ((size_t)T.chain * cap + T.first) * width;
    const int64_t n = T.last - T.first;
    // total weight
    double wsum = 0.0;
    for (int64_t e = tid; e < n; e += 256) wsum += base[(size_t)e * width];
    wsum = warp_sum(wsum);
    if (tid == 0) { s_norm = 0.0; s_prefix[0] = s_prefix[1] = 0ull; s_below[0] = s_below[1] = 0.0; }
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd(&s_norm, wsum);
    __syncthreads();
    const double norm = s_norm;
    const double target[2] = {norm * limfrac, norm * (1.0 - limfrac)};
    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 56 - 8 * pass;
        hist[0][tid] = 0.0;
        hist[1][tid] = 0.0;
        __syncthreads();
        const unsigned long long p0 = s_prefix[0], p1 = s_prefix[1];
        // rows whose key agrees with the prefix on the bits above `shift + 8`
        const unsigned long long mask = (pass == 0) ? 0ull : (~0ull << (shift + 8));
        for (int64_t e = tid; e < n; e += 256) {
            const double *row = base + (size_t)e * width;
            const unsigned long long k = f64_order_key(row[2 + par]);
            const double w = row[0];
            const int digit = (int)((k >> shift) & 0xffull);
            if ((k & mask) == p0) atomicAdd(&hist[0][digit], w);
            if ((k & mask) == p1) atomicAdd(&hist[1][digit], w);
        }
        __syncthreads();
        if (tid < 2) {
            double below = s_below[tid];
            int d = 0;
            for (; d < 255; ++d) {
                if (below + hist[tid][d] >= target[tid]) break;
                below += hist[tid][d];
            }
            s_below[tid] = below;
            s_prefix[tid] |= (unsigned long long)d << shift;
        }
        __syncthreads();
    }
    if (tid < 2)
        bounds[((size_t)task * D + par) * 2 + tid] = f64_from_order_key(s_prefix[tid]);
}

// out = { M, sum (low-s)[D], sum (low-s)^2[D], sum (up-s)[D], sum (up-s)^2[D] }
__global__ void k_reduce_bounds(const double *__restrict__ bounds, int64_t n_tasks, int D,
                                const double *__restrict__ shift, double *__restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;  // (par, which)
    if (e == 0) out[0] = (double)n_tasks;
    if (e >= 2 * D) return;
    const int par = e >> 1, which = e & 1;
    double s1 = 0.0, s2 = 0.0;
    const double sh = shift ? shift[par] : 0.0;
    for (int64_t t = 0; t < n_tasks; ++t) {
        const double b = bounds[((size_t)t * D + par) * 2 + which] - sh;
        s1 += b;
        s2 += b * b;
    }
    out[1 + (which ? 2 * D : 0) + par] = s1;
    out[1 + (which ? 2 * D : 0) + D + par] = s2;
}

// =====================================================================================
// Symmetric variant for D <= 64 (NT even): S2 is symmetric, so only the tiles on and below
// the diagonal are accumulated (NT(NT+1)/2 instead of NT^2 MMAs per k-step) and mirrored at
// the end.  NT/2 warps per CTA; warp w owns the row tiles t0 = w and t1 = NT-1-w, i.e. NT+1
// tiles: slots 0..t0 are (t0, n = slot), slots t0+1..NT are (t1, n = slot - t0 - 1).
// =====================================================================================
template <int NT>
__global__ void __launch_bounds__(NT * 16, (NT >= 7) ? 3 : 1)
k_task_moments_dmma_sym(const double *__restrict__ rows, int64_t cap, int width, int D,
                        const MomentTask *__restrict__ tasks, int64_t n_tasks,
                        const double *__restrict__ shift, double *__restrict__ partials) {
    constexpr int DP = NT * 8;
    constexpr int LDX = DP + 4;
    constexpr int NTH = NT * 16;                   // threads
    constexpr int PER = CB2_MOM_BATCH * DP / NTH;  // staged values per thread
    constexpr int NS = NT + 1;                     // tile slots per warp
    __shared__ double xt[CB2_MOM_BATCH][LDX];
    __shared__ double wt[CB2_MOM_BATCH];
    __shared__ double refv[DP], s1v[DP], mrel[DP], msv[DP];
    __shared__ double sw_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = lane >> 2, r = lane & 3;
    const int DD = D * D;
    const int t0 = wid, t1 = NT - 1 - wid;         // t0 < t1
    double *P = partials + (size_t)blockIdx.x * (size_t)(3 + D + 2 * DD);
    double sc[NS][2], smm[NS][2];
#pragma unroll
    for (int n = 0; n < NS; ++n) { sc[n][0] = sc[n][1] = smm[n][0] = smm[n][1] = 0.0; }
    double accM = 0.0, accN = 0.0, accNa = 0.0;
    double accm[(DP + NTH - 1) / NTH];
#pragma unroll
    for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) accm[u] = 0.0;
    for (int64_t t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const MomentTask T = tasks[t];
        const double *base = rows + (size_t)T.chain * cap * width;
        double s2[NS][2];
#pragma unroll
        for (int n = 0; n < NS; ++n) { s2[n][0] = 0.0; s2[n][1] = 0.0; }
        double s1[(DP + NTH - 1) / NTH], swl = 0.0;
#pragma unroll
        for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) s1[u] = 0.0;
        __syncthreads();
        for (int d = tid; d < DP; d += NTH)
            refv[d] = (d < D) ? base[(size_t)T.first * width + 2 + d] : 0.0;
        __syncthreads();
        double vals[PER], wnext = 0.0;
        auto fetch = [&](int64_t r0) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int e = tid + u * NTH;
                const int k = e / DP, d = e % DP;
                vals[u] = (k < nb && d < D) ? base[(size_t)(r0 + k) * width + 2 + d] : 0.0;
            }
            if (tid < CB2_MOM_BATCH) wnext = (tid < nb) ? base[(size_t)(r0 + tid) * width] : 0.0;
        };
        if (T.first < T.last) fetch(T.first);
        for (int64_t r0 = T.first; r0 < T.last; r0 += CB2_MOM_BATCH) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int e = tid + u * NTH;
                const int k = e / DP, d = e % DP;
                xt[k][d] = (k < nb && d < D) ? vals[u] - refv[d] : 0.0;
            }
            if (tid < CB2_MOM_BATCH) wt[tid] = wnext;
            __syncthreads();
            if (r0 + CB2_MOM_BATCH < T.last) fetch(r0 + CB2_MOM_BATCH);
#pragma unroll
            for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) {
                const int d = tid + u * NTH;
                if (d < DP) {
#pragma unroll 8
                    for (int k = 0; k < CB2_MOM_BATCH; ++k) s1[u] = fma(wt[k], xt[k][d], s1[u]);
                }
            }
            if (tid == 0)
                for (int k = 0; k < CB2_MOM_BATCH; ++k) swl += wt[k];
#pragma unroll
            for (int kk = 0; kk < CB2_MOM_BATCH / 4; ++kk) {
                const int k = 4 * kk + r;
                const double wk = wt[k];
                const double a0 = wk * xt[k][8 * t0 + q], a1 = wk * xt[k][8 * t1 + q];
#pragma unroll
                for (int sl = 0; sl < NS; ++sl) {
                    const bool first = sl <= t0;
                    const int n = first ? sl : sl - t0 - 1;
                    mom_dmma(s2[sl][0], s2[sl][1], first ? a0 : a1, xt[k][8 * n + q]);
                }
            }
        }
        __syncthreads();
        if (tid == 0) sw_s = swl;
#pragma unroll
        for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) {
            const int d = tid + u * NTH;
            if (d < DP) s1v[d] = s1[u];
        }
        __syncthreads();
        const double sw = sw_s;
#pragma unroll
        for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) {
            const int d = tid + u * NTH;
            if (d < DP) {
                const double mr = s1v[d] / sw;
                mrel[d] = mr;
                const double ms = (d < D) ? (refv[d] + mr) - (shift ? shift[d] : 0.0) : 0.0;
                msv[d] = ms;
                accm[u] += ms;
            }
        }
        __syncthreads();
        const double f = T.N / sw;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
            const bool first = sl <= t0;
            const int i = 8 * (first ? t0 : t1) + q;
            const int n = first ? sl : sl - t0 - 1;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 8 * n + 2 * r + h;
                sc[sl][h] += f * s2[sl][h] - T.N * (mrel[i] * mrel[j]);
                smm[sl][h] += msv[i] * msv[j];
            }
        }
        if (tid == 0) {
            accM += 1.0;
            accN += T.N;
            accNa += T.N * ((double)(T.last - T.first) / sw);
        }
    }
    __syncthreads();
#pragma unroll
    for (int sl = 0; sl < NS; ++sl) {
        const bool first = sl <= t0;
        const int ti = first ? t0 : t1;
        const int i = 8 * ti + q;
        const int n = first ? sl : sl - t0 - 1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 8 * n + 2 * r + h;
            if (i < D && j < D) {
                // diagonal tiles hold both triangles already; off-diagonal ones are mirrored
                P[3 + D + i * D + j] = smm[sl][h];
                P[3 + D + DD + i * D + j] = sc[sl][h];
                if (n != ti) {
                    P[3 + D + j * D + i] = smm[sl][h];
                    P[3 + D + DD + j * D + i] = sc[sl][h];
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < (DP + NTH - 1) / NTH; ++u) {
        const int d = tid + u * NTH;
        if (d < D) P[3 + d] = accm[u];
    }
    if (tid == 0) {
        P[0] = accM;
        P[1] = accN;
        P[2] = accNa;
    }
}

// =====================================================================================
// The same SYRK for 64 < D <= 256: the D x D output is cut into 64 x 64 blocks; a CTA
// (8 warps) owns one block pair (ib >= jb; the mirror image is written at the end) for its
// share of the tasks.  grid = (task CTAs, block pairs).  Rows are staged for all D
// coordinates, so every pair re-reads its tasks' rows (the kernel is FP64-bound: 2 D^2 FLOP
// against 8 D B per row).  Scalars and the mean vector are written by pair 0.
// =====================================================================================
#define CB2_MOMB_VALS 32   // CB2_MOM_BATCH * 256 / 256

__global__ void __launch_bounds__(256, 1)
k_task_moments_dmma_blk(const double *__restrict__ rows, int64_t cap, int width, int D, int DP,
                        const MomentTask *__restrict__ tasks, int64_t n_tasks,
                        const double *__restrict__ shift, double *__restrict__ partials) {
    extern __shared__ __align__(16) double msm[];
    const int LDX = DP + 4;
    double *xt = msm;                          // [CB2_MOM_BATCH][LDX]
    double *wt = xt + CB2_MOM_BATCH * LDX;     // [CB2_MOM_BATCH]
    double *refv = wt + CB2_MOM_BATCH, *s1v = refv + DP, *mrel = s1v + DP, *msv = mrel + DP;
    __shared__ double sw_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = lane >> 2, r = lane & 3;
    const int DD = D * D;
    // block pair index -> (ib, jb), jb <= ib
    int ib = 0, jb = (int)blockIdx.y;
    while (jb > ib) { jb -= ib + 1; ++ib; }
    const bool lead = blockIdx.y == 0;
    double *P = partials + (size_t)blockIdx.x * (size_t)(3 + D + 2 * DD);
    const int per = DP >> 3;                   // values per thread and batch
    double sc[8][2], smm[8][2];
#pragma unroll
    for (int n = 0; n < 8; ++n) { sc[n][0] = sc[n][1] = smm[n][0] = smm[n][1] = 0.0; }
    double accM = 0.0, accN = 0.0, accNa = 0.0, accm = 0.0;
    for (int64_t t = blockIdx.x; t < n_tasks; t += gridDim.x) {
        const MomentTask T = tasks[t];
        const double *base = rows + (size_t)T.chain * cap * width;
        double s2[8][2];
#pragma unroll
        for (int n = 0; n < 8; ++n) { s2[n][0] = 0.0; s2[n][1] = 0.0; }
        double s1 = 0.0, swl = 0.0;
        __syncthreads();
        if (tid < DP) refv[tid] = (tid < D) ? base[(size_t)T.first * width + 2 + tid] : 0.0;
        __syncthreads();
        double vals[CB2_MOMB_VALS], wnext = 0.0;
        auto fetch = [&](int64_t r0) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
#pragma unroll
            for (int u = 0; u < CB2_MOMB_VALS; ++u) {
                if (u < per) {
                    const int e = tid + u * 256;
                    const int k = e / DP, d = e % DP;
                    vals[u] = (k < nb && d < D) ? base[(size_t)(r0 + k) * width + 2 + d] : 0.0;
                }
            }
            if (tid < CB2_MOM_BATCH) wnext = (tid < nb) ? base[(size_t)(r0 + tid) * width] : 0.0;
        };
        if (T.first < T.last) fetch(T.first);
        for (int64_t r0 = T.first; r0 < T.last; r0 += CB2_MOM_BATCH) {
            const int nb = (int)min((int64_t)CB2_MOM_BATCH, T.last - r0);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < CB2_MOMB_VALS; ++u) {
                if (u < per) {
                    const int e = tid + u * 256;
                    const int k = e / DP, d = e % DP;
                    xt[k * LDX + d] = (k < nb && d < D) ? vals[u] - refv[d] : 0.0;
                }
            }
            if (tid < CB2_MOM_BATCH) wt[tid] = wnext;
            __syncthreads();
            if (r0 + CB2_MOM_BATCH < T.last) fetch(r0 + CB2_MOM_BATCH);
            if (tid < DP) {
#pragma unroll 8
                for (int k = 0; k < CB2_MOM_BATCH; ++k) s1 = fma(wt[k], xt[k * LDX + tid], s1);
            }
            if (tid == 0)
                for (int k = 0; k < CB2_MOM_BATCH; ++k) swl += wt[k];
#pragma unroll
            for (int kk = 0; kk < CB2_MOM_BATCH / 4; ++kk) {
                const int k = 4 * kk + r;
                const double a = wt[k] * xt[k * LDX + 64 * ib + 8 * wid + q];
#pragma unroll
                for (int n = 0; n < 8; ++n)
                    mom_dmma(s2[n][0], s2[n][1], a, xt[k * LDX + 64 * jb + 8 * n + q]);
            }
        }
        __syncthreads();
        if (tid == 0) sw_s = swl;
        if (tid < DP) s1v[tid] = s1;
        __syncthreads();
        const double sw = sw_s;
        if (tid < DP) {
            const double mr = s1v[tid] / sw;
            mrel[tid] = mr;
            const double ms = (tid < D) ? (refv[tid] + mr) - (shift ? shift[tid] : 0.0) : 0.0;
            msv[tid] = ms;
            accm += ms;
        }
        __syncthreads();
        const double f = T.N / sw;
        const int i = 64 * ib + 8 * wid + q;
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 64 * jb + 8 * n + 2 * r + h;
                if (i < DP && j < DP) {
                    sc[n][h] += f * s2[n][h] - T.N * (mrel[i] * mrel[j]);
                    smm[n][h] += msv[i] * msv[j];
                }
            }
        if (tid == 0) {
            accM += 1.0;
            accN += T.N;
            accNa += T.N * ((double)(T.last - T.first) / sw);
        }
    }
    __syncthreads();
    {
        const int i = 64 * ib + 8 * wid + q;
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = 64 * jb + 8 * n + 2 * r + h;
                if (i < D && j < D) {
                    P[3 + D + i * D + j] = smm[n][h];
                    P[3 + D + DD + i * D + j] = sc[n][h];
                    if (ib != jb) {
                        P[3 + D + j * D + i] = smm[n][h];
                        P[3 + D + DD + j * D + i] = sc[n][h];
                    }
                }
            }
    }
    if (lead) {
        if (tid < D) P[3 + tid] = accm;
        if (tid == 0) {
            P[0] = accM;
            P[1] = accN;
            P[2] = accNa;
        }
    }
}
