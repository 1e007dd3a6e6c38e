// kernels_fast.cuh -- register-resident fast paths for D <= 64 (sm_100a).
//
//  * k_basis_fast<NG>  : random_SO_N (cobaya/functions.py:21-60) with one thread per ROW
//    of H held in registers; the Householder sweep is the reference's own left-to-right
//    order (functions.py:48-58), the vectors x_m are broadcast from shared memory.
//  * k_step_fast<NT>   : the propose -> logposterior -> accept -> store loop
//    (mcmc.py:545-562,670-748) for 8 chains per warp.  The two triangular matrix
//    products of a proposal -- T v (proposal.py:224) and L^-1 (x'-mu)
//    (gaussian_mixture.py:148) -- run on the FP64 tensor pipe as m8n8k4 DMMA tiles with
//    the 8 chains as the M dimension; chain state lives in registers for the whole
//    window; the shared matrices are staged into shared memory once per CTA with a TMA
//    bulk copy (cp.async.bulk + mbarrier) in MMA-fragment order.
#pragma once
#include "common.cuh"
#include "kernels_general.cuh"

// D[8x8] += A[8x4] B[4x8] in fp64 on the tensor pipe.  Lane l = 4q + r holds
// A[q][r], B[r][q], D[q][2r], D[q][2r+1].
__device__ __forceinline__ void dmma8x8x4(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// =====================================================================================
// Haar basis, n <= 64
// =====================================================================================
static inline bool fast_basis_supported(int n) { return n >= 2 && n <= 64; }
// 64 < n <= 128: compact-WY kernel with 8 warps per basis, normals from k_normals
static inline bool wy_large_supported(int n) { return n > 64 && n <= 128; }

template <int NG>
__global__ void __launch_bounds__((NG * 16 < 32) ? 32 : NG * 16, (NG >= 7) ? 4 : 1)
k_basis_fast(uint32_t key0, uint32_t key1, uint64_t chain_id0, int block, int n,
             const int64_t *__restrict__ vis, int vis_stride, uint32_t e0_fixed, int cnt,
             double *__restrict__ store, int64_t task0, int64_t store_task0) {
    // Two threads per row of H: thread (row i, half p) keeps the column pairs
    // {4g+2p, 4g+2p+1 : g = 0..NP/4-1} of its row in registers (NP/2 doubles).
    constexpr int NP = NG * 8;   // padded size
    constexpr int NK = NP / 2;   // columns per thread
    extern __shared__ __align__(16) double fsm[];
    double *X = fsm;              // [NP][NP]: X[m][c] = x_m[c-m] for c >= m, else 0
    double *Dv = fsm + NP * NP;   // [NP]
    double *inv = Dv + NP;        // [NP] 1/sc_m
    const int tid = threadIdx.x, nt = blockDim.x;
    const int64_t task = task0 + blockIdx.x;
    const int64_t chain = task / cnt;
    const uint32_t e0 = vis ? (uint32_t)(vis[chain * vis_stride + block] / n) : e0_fixed;
    const uint32_t epoch = e0 + (uint32_t)(task % cnt);
    const uint64_t gid = chain_id0 + (uint64_t)chain;
    for (int e = tid; e < NP * NP; e += nt) X[e] = 0.0;
    __syncthreads();
    // 1. standard normals (functions.py:36) scattered into the padded layout:
    //    normal q = ix(m) + i with ix(m) = m*n - m(m-1)/2 goes to X[m][m+i]
    const int nn = (n + 2) * (n - 1) / 2;
    {
        int m = 0, base = 0;  // base = ix(m); q only grows for a thread
        for (int p = tid; p < (nn + 1) / 2; p += nt) {
            double z[2];
            draw_normal_pair(key0, key1, gid, block, epoch, (uint32_t)p, z[0], z[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int q = 2 * p + h;
                if (q < nn) {
                    while (q >= base + (n - m)) {
                        base += n - m;
                        ++m;
                    }
                    X[m * NP + m + (q - base)] = z[h];
                }
            }
        }
    }
    __syncthreads();
    // 2. Householder vectors (functions.py:49-55): norms by one thread per vector, the
    //    rescaling x /= sc by all threads
    if (tid < n - 1) {
        const int m = tid, len = n - m;
        const double *x = X + m * NP + m;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int i = 0;
        for (; i + 4 <= len; i += 4) {
            s0 = fma(x[i], x[i], s0);
            s1 = fma(x[i + 1], x[i + 1], s1);
            s2 = fma(x[i + 2], x[i + 2], s2);
            s3 = fma(x[i + 3], x[i + 3], s3);
        }
        for (; i < len; ++i) s0 = fma(x[i], x[i], s0);
        const double norm2 = (s0 + s1) + (s2 + s3);
        const double x0 = x[0];
        const double d = (x0 != 0.0) ? (x0 > 0 ? 1.0 : -1.0) : 1.0;
        const double x0n = x0 + d * sqrt(norm2);
        X[m * NP + m] = x0n;
        inv[m] = 1.0 / sqrt((norm2 - x0 * x0 + x0n * x0n) / 2.0);
        Dv[m] = d;
    }
    __syncthreads();
    for (int e = tid; e < (n - 1) * NP; e += nt) X[e] *= inv[e / NP];
    if (tid == 0) {  // functions.py:59
        double prod = 1.0;
        for (int m = 0; m < n - 1; ++m) prod *= Dv[m];
        Dv[n - 1] = (((n - 1) & 1) ? -1.0 : 1.0) * prod;
    }
    __syncthreads();
    // 3. H = I; for m: H[:, m:] -= (H[:, m:] x_m) x_m^T  (functions.py:57-58)
    const int row = tid >> 1, half = tid & 1;
    double h[NK];
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int c = 4 * (k >> 1) + 2 * half + (k & 1);
        h[k] = (c == row) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int G = 0; G < NG; ++G) {
        const int m_end = min(8 * G + 8, n - 1);
        for (int m = 8 * G; m < m_end; ++m) {
            // this thread's pairs of group g >= 2G: columns 4g+2*half, +1 -> double2 index 2g+half
            const double2 *x2 = reinterpret_cast<const double2 *>(X + m * NP) + half;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
            for (int g = 2 * G; g < NP / 4; g += 2) {  // NP/4 - 2G is even
                const double2 a = x2[2 * g], b = x2[2 * g + 2];
                t0 = fma(h[2 * g], a.x, t0);
                t1 = fma(h[2 * g + 1], a.y, t1);
                t2 = fma(h[2 * g + 2], b.x, t2);
                t3 = fma(h[2 * g + 3], b.y, t3);
            }
            double tmp = (t0 + t1) + (t2 + t3);
            tmp += __shfl_xor_sync(0xffffffffu, tmp, 1);
#pragma unroll
            for (int g = 2 * G; g < NP / 4; ++g) {
                const double2 a = x2[2 * g];
                h[2 * g] = fma(-tmp, a.x, h[2 * g]);
                h[2 * g + 1] = fma(-tmp, a.y, h[2 * g + 1]);
            }
        }
    }
    // 4. R = diag(D) H (functions.py:60); stored as Rt[c*n + row] = R[row][c]
    if (row < n) {
        const double d = Dv[row];
        double *out = store + (size_t)(task - store_task0) * (size_t)n * n;
#pragma unroll
        for (int k = 0; k < NK; ++k) {
            const int c = 4 * (k >> 1) + 2 * half + (k & 1);
            if (c < n) out[(size_t)c * n + row] = d * h[k];
        }
    }
}

// -------------------------------------------------------------------------------------
// k_normals: the standard normals random_SO_N consumes (functions.py:36), one CTA per
// (chain, epoch), written in the reference's flat order (vector m occupies
// [m(2n-m+1)/2, +n-m)) to out[task * nn_pad ...].  It is integer/ALU work (Philox rounds,
// bit-assembled uniforms, table logarithm, series sincos) with no tensor-pipe use and a
// 64-register footprint, so that the engine can run it on a side stream UNDER the step
// kernel of the previous window (which leaves ~3/4 of the issue slots idle) instead of in
// front of the Householder sweep.  Same draws as draw_normal_pair_tab in the fused kernels.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 8)
k_normals(uint32_t key0, uint32_t key1, uint64_t chain_id0, int block, int n,
          const int64_t *__restrict__ vis, int vis_stride, uint32_t e0_fixed, int cnt,
          double *__restrict__ out, int nn_pad) {
    __shared__ double ltab[CB2_LOGTAB_DOUBLES];
    const int tid = threadIdx.x;
    for (int e = tid; e < CB2_LOGTAB_DOUBLES; e += 128) ltab[e] = g_logtab[e];
    const int64_t task = blockIdx.x;
    const int64_t chain = task / cnt;
    uint32_t e0 = e0_fixed;
    if (vis) {
        const int64_t v = vis[chain * vis_stride + block];
        e0 = (v >> 32) ? (uint32_t)(v / n) : (uint32_t)v / (uint32_t)n;
    }
    const uint32_t epoch = e0 + (uint32_t)(task % cnt);
    const uint64_t gid = chain_id0 + (uint64_t)chain;
    const int npairs = nn_pad >> 1;  // nn_pad = nn rounded up to even
    double2 *o2 = reinterpret_cast<double2 *>(out + (size_t)task * (size_t)nn_pad);
    __syncthreads();
    // two pairs per iteration: independent Philox / log / sincos chains for the scheduler
    int p = tid;
    for (; p + 128 < npairs; p += 256) {
        double a0, a1, b0, b1;
        draw_normal_pair_tab(key0, key1, gid, block, epoch, (uint32_t)p, ltab, a0, a1);
        draw_normal_pair_tab(key0, key1, gid, block, epoch, (uint32_t)(p + 128), ltab, b0, b1);
        o2[p] = make_double2(a0, a1);
        o2[p + 128] = make_double2(b0, b1);
    }
    if (p < npairs) {
        double a0, a1;
        draw_normal_pair_tab(key0, key1, gid, block, epoch, (uint32_t)p, ltab, a0, a1);
        o2[p] = make_double2(a0, a1);
    }
}

static inline int launch_normals(cudaStream_t st, uint32_t k0, uint32_t k1, uint64_t chain_id0,
                                 int block, int n, const int64_t *vis, int vis_stride, int cnt,
                                 double *out, int64_t n_chains) {
    const int nn = (n + 2) * (n - 1) / 2, nn_pad = (nn + 1) & ~1;
    k_normals<<<(unsigned)(n_chains * cnt), 128, 0, st>>>(k0, k1, chain_id0, block, n, vis,
                                                          vis_stride, 0u, cnt, out, nn_pad);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
static inline size_t normals_per_task(int n) {
    const int nn = (n + 2) * (n - 1) / 2;
    return (size_t)((nn + 1) & ~1);
}

// -------------------------------------------------------------------------------------
// k_basis_wy<NG>: the same Haar basis with the Householder sweep on the FP64 TENSOR pipe.
// Reflectors are grouped 8 at a time in compact-WY form (LAPACK dlarft, forward/columnwise):
//   G_{m0} ... G_{m0+7} = I - V T V^T,  V = [x_{m0} .. x_{m0+7}],  T upper triangular,
//   T_jj = tau_j = 2/|x_j|^2,  T[0:j, j] = -tau_j T[0:j,0:j] (V[:,0:j]^T v_j)
// and H <- H - ((H V) T) V^T is three small matrix products per group, issued as m8n8k4 DMMA
// tiles.  4 warps per basis, warp w owns rows 16w..16w+15 of H in C-fragment layout (lane
// (q,r): H[row q][cols 8nt+2r, +1]); with the even/odd k interleave a C fragment is directly
// an A fragment.  Same matrix as functions.py:48-58 up to rounding (~1e-15).
// -------------------------------------------------------------------------------------
// NW warps per basis: 4 for n <= 64; 8 for 64 < n <= 128 (NG = 16 groups, the normals must then
// come from k_normals).  Warp w owns the row tiles {w, NG-1-w} of N.
template <int NG, int NW>
__global__ void __launch_bounds__(NW * 32, (NG <= 8) ? 5 : 1)
k_basis_wy(uint32_t key0, uint32_t key1, uint64_t chain_id0, int block, int n,
           const int64_t *__restrict__ vis, int vis_stride, uint32_t e0_fixed, int cnt,
           double *__restrict__ store, int64_t task0, int64_t store_task0,
           const double *__restrict__ normals, int nn_pad) {
    constexpr int NP = NG * 8;
    constexpr int LDX = NP + 1;  // odd row stride: both B-fragment access patterns spread banks
    extern __shared__ __align__(16) double fsm[];
    double *X = fsm;                    // [NP][LDX]: X[m][c] = x_m[c-m] for c >= m, else 0
    double *Dv = X + NP * LDX;          // [NP]
    double *inv = Dv + NP;              // [NP]
    double *Sg = inv + NP;              // [NG][8][8] Gram (strict upper) per group
    double *Tg = Sg + NG * 64;          // [NG][8][8] T per group
    __shared__ int sign_cnt[NW];
    const int tid = threadIdx.x, nt_ = blockDim.x;
    const int64_t task = task0 + blockIdx.x;
    const int64_t chain = task / cnt;
    if (normals) {
        // normals pre-generated by k_normals (flat reference order): row m of X is the
        // contiguous run [ix(m), ix(m) + n - m).  The first batch of loads is issued before
        // the zero fill so that its latency overlaps it.
        constexpr int RPW = (NP + NW - 1) / NW;  // rows per warp
        constexpr int NH = (NP + 31) / 32;       // column chunks of 32
        constexpr int UB = (RPW * NH <= 32) ? RPW : 4;  // rows in flight per batch
        const int lane_ = tid & 31, w_ = tid >> 5;
        const double *src = normals + (size_t)(task - task0) * (size_t)nn_pad;
        double v[UB][NH];
        auto load_rows = [&](int u0) {
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int m = w_ + NW * (u0 + u);
                const int len = n - m, base = (m * (2 * n - m + 1)) >> 1;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = lane_ + 32 * h;
                    v[u][h] = (u0 + u < RPW && m < n - 1 && i < len) ? __ldg(src + base + i) : 0.0;
                }
            }
        };
        auto store_rows = [&](int u0) {
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int m = w_ + NW * (u0 + u);
                const int len = n - m;
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = lane_ + 32 * h;
                    if (u0 + u < RPW && m < n - 1 && i < len) X[m * LDX + m + i] = v[u][h];
                }
            }
        };
        load_rows(0);  // in flight during the zero fill
        for (int e = tid; e < NP * LDX + 2 * NP; e += nt_) X[e] = 0.0;  // X, Dv, inv (= tau)
        __syncthreads();
        store_rows(0);
        for (int u0 = UB; u0 < RPW; u0 += UB) {
            load_rows(u0);
            store_rows(u0);
        }
        __syncthreads();
    } else if constexpr (NG <= 8) {
    // the logarithm table lives in the (not yet used) Gram/T area; its loads are issued first
    // so that their L2 latency overlaps the zero fill
    double *ltab = Sg;
    constexpr int LT_PER = (CB2_LOGTAB_DOUBLES + 127) / 128;
    double lt[LT_PER];
#pragma unroll
    for (int u = 0; u < LT_PER; ++u) {
        const int e = tid + u * 128;
        lt[u] = (e < CB2_LOGTAB_DOUBLES) ? g_logtab[e] : 0.0;
    }
    uint32_t e0 = e0_fixed;
    if (vis) {  // visits / n: 32-bit division unless the counter is huge
        const int64_t v = vis[chain * vis_stride + block];
        e0 = (v >> 32) ? (uint32_t)(v / n) : (uint32_t)v / (uint32_t)n;
    }
    const uint32_t epoch = e0 + (uint32_t)(task % cnt);
    const uint64_t gid = chain_id0 + (uint64_t)chain;
    for (int e = tid; e < NP * LDX + 2 * NP; e += nt_) X[e] = 0.0;  // X, Dv, inv (= tau)
#pragma unroll
    for (int u = 0; u < LT_PER; ++u) {
        const int e = tid + u * 128;
        if (e < CB2_LOGTAB_DOUBLES) ltab[e] = lt[u];
    }
    __syncthreads();
    const int nn = (n + 2) * (n - 1) / 2;
    {
        // all Box-Muller pairs of this thread first (independent chains: the compiler can
        // interleave the Philox rounds / log / sincospi of several pairs), then the scatter
        constexpr int NNMAX = (NP + 2) * (NP - 1) / 2;
        constexpr int PPT = ((NNMAX + 1) / 2 + 127) / 128;  // pairs per thread
        double z[PPT][2];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int p = tid + u * 128;
            if (2 * p < nn)
                draw_normal_pair_tab(key0, key1, gid, block, epoch, (uint32_t)p, ltab,
                                     z[u][0], z[u][1]);
        }
        // normal q of the reference's flat array belongs to vector m with
        // ix(m) = m(2n-m+1)/2 <= q < ix(m+1): m from the root of the quadratic (fp32 is exact
        // for these magnitudes up to the rounding of sqrt; one fix-up step each way)
        const int s2 = 2 * n + 1;
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int p = tid + u * 128;
            const int q0 = 2 * p;
            if (q0 < nn) {
                int m = (int)(((float)s2 - sqrtf((float)(s2 * s2 - 8 * q0))) * 0.5f);
                m = min(max(m, 0), n - 2);
                int base = (m * (s2 - m)) >> 1;
                if (q0 < base) { --m; base -= n - m; }
                if (q0 >= base + (n - m)) { base += n - m; ++m; }
                X[m * LDX + m + (q0 - base)] = z[u][0];
                if (q0 + 1 < nn) {
                    if (q0 + 1 >= base + (n - m)) { base += n - m; ++m; }
                    X[m * LDX + m + (q0 + 1 - base)] = z[u][1];
                }
            }
        }
    }
    __syncthreads();
    }
    const int lane = tid & 31, w = tid >> 5, q = lane >> 2, r = lane & 3;
    // Gram S_g = V_g^T V_g of every group on the tensor pipe (A and B fragments coincide:
    // lane (q,r) holds x_{8g+q}[k = r]).  It is taken over the normals as drawn: the
    // diagonal is nu_m = |x_m|^2 (functions.py:50), from which sign, x_m[0] += D sqrt(nu) and
    // tau_m follow (functions.py:51-55; the reference's normalisation is carried by tau);
    // for i < j the modified leading element of x_j only changes
    // S_ij by (x_j[0]' - x_j[0]) x_i[j - i].  Then row i of T (dlarft) by lane i.
    int n_negative = 0;
    for (int g = w; g < NG; g += NW) {
        double s0 = 0.0, s1 = 0.0;
        const double *xq = X + (8 * g + q) * LDX + 8 * g + r;
#pragma unroll 4
        for (int kk = 0; kk < (NP - 8 * g) / 4; ++kk) {
            const double a = xq[4 * kk];
            dmma8x8x4(s0, s1, a, a);
        }
        // lane (q,r) holds S[q][2r], S[q][2r+1]; the diagonal entry j sits in lane 4j + j/2
        const double nu_a = __shfl_sync(0xffffffffu, s0, 9 * r);      // j = 2r
        const double nu_b = __shfl_sync(0xffffffffu, s1, 9 * r + 4);  // j = 2r + 1
        const int ja = 2 * r, jb = 2 * r + 1, ma = 8 * g + ja, mb = 8 * g + jb;
        const double xa = X[ma * LDX + ma], xb = X[mb * LDX + mb];
        const double da = (xa != 0.0) ? (xa > 0 ? 1.0 : -1.0) : 1.0;
        const double db = (xb != 0.0) ? (xb > 0 ? 1.0 : -1.0) : 1.0;
        const bool va = ma < n - 1, vb = mb < n - 1;
        const double ca = va ? da * sqrt(nu_a) : 0.0, cb = vb ? db * sqrt(nu_b) : 0.0;
        const double *xrow = X + (8 * g + q) * LDX + 8 * g;
        if (q < ja) s0 = fma(ca, xrow[ja], s0);
        if (q < jb) s1 = fma(cb, xrow[jb], s1);
        double *S = Sg + g * 64;
        S[q * 8 + 2 * r] = s0;
        S[q * 8 + 2 * r + 1] = s1;
        __syncwarp();  // every lane has read the leading elements it needs
        if (q == 0) {
            if (va) {
                const double x0n = xa + ca;
                X[ma * LDX + ma] = x0n;
                inv[ma] = 2.0 / (nu_a - xa * xa + x0n * x0n);
                Dv[ma] = da;
            }
            if (vb) {
                const double x0n = xb + cb;
                X[mb * LDX + mb] = x0n;
                inv[mb] = 2.0 / (nu_b - xb * xb + x0n * x0n);
                Dv[mb] = db;
            }
        }
        // functions.py:59: D[n-1] = (-1)^(n-1) prod(D[:-1]) -- a parity count
        n_negative += __popc(__ballot_sync(0xffffffffu, q == 0 && va && da < 0.0)) +
                      __popc(__ballot_sync(0xffffffffu, q == 0 && vb && db < 0.0));
        __syncwarp();
        if (lane < 8) {
            const int i = lane;
            double *T = Tg + g * 64;
            double trow[8];
            const double *tau = inv + 8 * g;  // 0 for the padding reflectors (m >= n-1)
#pragma unroll
            for (int j = 0; j < 8; ++j) trow[j] = (j == i) ? tau[i] : 0.0;
#pragma unroll
            for (int j = 1; j < 8; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int l = 0; l < 8; ++l)
                    if (l >= i && l < j) acc = fma(trow[l], S[l * 8 + j], acc);
                if (i < j) trow[j] = -tau[j] * acc;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) T[i * 8 + j] = trow[j];
        }
    }
    if (lane == 0) sign_cnt[w] = n_negative;
    __syncthreads();
    // D[n-1] (functions.py:59) stays in a register (d_last): nothing reads Dv[n-1]
    int c_all = n - 1;
#pragma unroll
    for (int i = 0; i < NW; ++i) c_all += sign_cnt[i];
    const double d_last = (c_all & 1) ? -1.0 : 1.0;
    // ---- accumulate H = Q_0 Q_1 ... Q_{NG-1} from the innermost factor (dorgqr order): with
    // N = M^T,  M <- Q_g M  is  N <- N - ((N V_g) T_g^T) V_g^T, and only rows/columns >= 8g
    // of N differ from the identity, so row tiles below 8g are skipped (888 instead of 1280
    // MMAs at n = 64).  Warp w owns the row tiles {w, NG-1-w} of N (balanced work).
    const int t0 = w, t1 = NG - 1 - w;
    const int nslot = (t0 < t1) ? 2 : ((t0 == t1) ? 1 : 0);
    double nreg[2][NG][2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int nt = 0; nt < NG; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h)
                nreg[sl][nt][h] = (8 * (sl ? t1 : t0) + q == 8 * nt + 2 * r + h) ? 1.0 : 0.0;
    if (nslot > 0) {
#pragma unroll
        for (int gg = 0; gg < NG; ++gg) {
            const int g = NG - 1 - gg;
            if (8 * g >= n - 1) continue;  // no reflectors in this group
            const double *Xg = X + (8 * g) * LDX;
            const double *T = Tg + g * 64;
            const bool act0 = t0 >= g, act1 = (nslot > 1) && (t1 >= g);
            if (!act0 && !act1) continue;  // warp-uniform
            // Y = N[:, 8g:] V
            double y[2][2][2];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                for (int pa = 0; pa < 2; ++pa) { y[sl][pa][0] = 0.0; y[sl][pa][1] = 0.0; }
#pragma unroll
            for (int nt = g; nt < NG; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double bv = Xg[q * LDX + 8 * nt + 2 * r + h];
                    if (act0) dmma8x8x4(y[0][(nt + h) & 1][0], y[0][(nt + h) & 1][1], nreg[0][nt][h], bv);
                    if (act1) dmma8x8x4(y[1][(nt + h) & 1][0], y[1][(nt + h) & 1][1], nreg[1][nt][h], bv);
                }
            // Z = Y T^T
            double z[2][2];
            const double tb0 = T[q * 8 + 2 * r], tb1 = T[q * 8 + 2 * r + 1];
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                z[sl][0] = 0.0; z[sl][1] = 0.0;
                if (sl == 0 ? act0 : act1) {
                    const double ya = y[sl][0][0] + y[sl][1][0], yb = y[sl][0][1] + y[sl][1][1];
                    dmma8x8x4(z[sl][0], z[sl][1], ya, tb0);
                    dmma8x8x4(z[sl][0], z[sl][1], yb, tb1);
                }
            }
            // N[:, 8g:] -= Z V^T
#pragma unroll
            for (int nt = g; nt < NG; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double bv = Xg[(2 * r + h) * LDX + 8 * nt + q];
                    if (act0) dmma8x8x4(nreg[0][nt][0], nreg[0][nt][1], -z[0][h], bv);
                    if (act1) dmma8x8x4(nreg[1][nt][0], nreg[1][nt][1], -z[1][h], bv);
                }
        }
        // R[i][c] = D[i] H[i][c] = D[i] N[c][i]  (functions.py:60), stored as Rt[c*n + i]
        double *out = store + (size_t)(task - store_task0) * (size_t)n * n;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            if (sl >= nslot) break;
            const int c = 8 * (sl ? t1 : t0) + q;
            if (c < n) {
                if ((n & 1) == 0) {  // 16-byte stores: i even, rows of n doubles stay aligned
                    // D[i] = +-1: the product is a sign flip, done on the integer pipe (a
                    // vector FP64 instruction issued among DMMAs of the co-resident CTAs costs
                    // the FP64 pipe about as much as a DMMA; tools/fp64_mix.cu)
#pragma unroll
                    for (int nt = 0; nt < NG; ++nt) {
                        const int i = 8 * nt + 2 * r;
                        if (i < n) {
                            const double2 d2 = *reinterpret_cast<const double2 *>(Dv + i);
                            const double db = (i + 1 == n - 1) ? d_last : d2.y;
                            const unsigned long long sa =
                                (unsigned long long)__double_as_longlong(d2.x) & 0x8000000000000000ull;
                            const unsigned long long sb =
                                (unsigned long long)__double_as_longlong(db) & 0x8000000000000000ull;
                            *reinterpret_cast<double2 *>(out + (size_t)c * n + i) = make_double2(
                                __longlong_as_double((long long)((unsigned long long)
                                    __double_as_longlong(nreg[sl][nt][0]) ^ sa)),
                                __longlong_as_double((long long)((unsigned long long)
                                    __double_as_longlong(nreg[sl][nt][1]) ^ sb)));
                        }
                    }
                } else {
#pragma unroll
                    for (int nt = 0; nt < NG; ++nt)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i = 8 * nt + 2 * r + h;
                            if (i < n)
                                out[(size_t)c * n + i] =
                                    (i == n - 1 ? d_last : Dv[i]) * nreg[sl][nt][h];
                        }
                }
            }
        }
    }
}

// -------------------------------------------------------------------------------------
// k_basis_wy_big: the same compact-WY backward accumulation for 128 < n <= 512 (n % 8 == 0).
// N = H^T (n x n) no longer fits the registers of a CTA: it lives in the basis store itself
// (global memory, L2 where it fits) and every panel of 16 reflectors streams the rows/columns
// >= 16g of it twice (Y = N V_g; N -= (Y T_g^T) V_g^T) as m8n8k4 A/C fragments.  One CTA of 8
// warps per basis; the 16 Householder vectors of the current panel sit in shared memory.
// 4/3 n^3 FLOP and ~n^3/2 B of N traffic per basis (n = 512: 67 MB): bandwidth-bound.
// Normals from k_normals.  Output as everywhere: store[c*n + i] = D[i] N[c][i] = R[i][c].
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 2)
k_basis_wy_big(int n, const double *__restrict__ normals, int nn_pad, double *__restrict__ store,
               int64_t n_tasks) {
    // panels of 16 reflectors (two 8-wide halves a, b): half as many passes over N as with 8
    extern __shared__ __align__(16) double bsm[];
    const int LDX = n + 1;
    double *X = bsm;               // [16][LDX]: X[j][c] = x_{16g+j}[c - (16g+j)]
    double *Dv = X + 16 * LDX;     // [n] signs D (functions.py:51,59)
    double *Sg = Dv + n;           // [256] Gram of the panel
    double *Tg = Sg + 256;         // [256] T of the panel (upper triangular, row-major)
    double *tau = Tg + 256;        // [16]
    __shared__ int n_negative;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, q = lane >> 2, r = lane & 3;
    const int64_t task = blockIdx.x;
    if (task >= n_tasks) return;
    double *N = store + (size_t)task * (size_t)n * n;
    const double *src = normals + (size_t)task * (size_t)nn_pad;
    const int NT = n >> 3;                 // 8-wide tiles per side
    const int NGr = (n - 1 + 15) >> 4;     // panels that hold reflectors (m < n-1)
    for (int e = tid; e < n * (n >> 1); e += 256) {  // N = I
        const int row = e / (n >> 1), c2 = (e % (n >> 1)) * 2;
        reinterpret_cast<double2 *>(N + (size_t)row * n)[c2 >> 1] =
            make_double2(row == c2 ? 1.0 : 0.0, row == c2 + 1 ? 1.0 : 0.0);
    }
    if (tid == 0) n_negative = 0;
    for (int e = tid; e < n; e += 256) Dv[e] = 1.0;
    __syncthreads();
    for (int g = NGr - 1; g >= 0; --g) {
        const int c_lo = 16 * g;
        // ---- the 16 vectors of the panel (zero where a vector has not started / m >= n-1)
        for (int e = tid; e < 16 * LDX; e += 256) {
            const int j = e / LDX, c = e % LDX, m = c_lo + j;
            double v = 0.0;
            if (m < n - 1 && c >= m && c < n) v = __ldg(src + ((m * (2 * n - m + 1)) >> 1) + (c - m));
            X[e] = v;
        }
        __syncthreads();
        // ---- Gram over the normals as drawn: one entry per thread
        {
            const int i = tid >> 4, j = tid & 15;
            const double *xi = X + i * LDX, *xj = X + j * LDX;
            double a0 = 0.0, a1 = 0.0;
            int c = c_lo;
            for (; c + 1 < n; c += 2) {
                a0 = fma(xi[c], xj[c], a0);
                a1 = fma(xi[c + 1], xj[c + 1], a1);
            }
            if (c < n) a0 = fma(xi[c], xj[c], a0);
            Sg[tid] = a0 + a1;
        }
        __syncthreads();
        // ---- norms, signs, x0 += D |x|, tau (functions.py:50-55); T by 16 lanes (dlarft)
        if (w == 0) {
            double ca = 0.0;
            int neg = 0;
            if (lane < 16) {
                const int m = c_lo + lane;
                const bool valid = m < n - 1;
                const double nu = Sg[lane * 17], x0 = X[lane * LDX + m];
                const double d = (x0 != 0.0) ? (x0 > 0 ? 1.0 : -1.0) : 1.0;
                ca = valid ? d * sqrt(nu) : 0.0;
                const double x0n = x0 + ca;
                tau[lane] = valid ? 2.0 / (nu - x0 * x0 + x0n * x0n) : 0.0;
                if (valid) { Dv[m] = d; neg = d < 0.0; }
            }
            neg = __popc(__ballot_sync(0xffffffffu, neg != 0));
            if (lane == 0) n_negative += neg;
            // modified leading elements: S_ij += (x_j[0]' - x_j[0]) x_i[j - i] for i < j
            double cj[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) cj[j] = __shfl_sync(0xffffffffu, ca, j);
            __syncwarp();
            if (lane < 16) {
                const int i = lane;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (i < j) Sg[i * 16 + j] = fma(cj[j], X[i * LDX + c_lo + j], Sg[i * 16 + j]);
            }
            __syncwarp();
            if (lane < 16) X[lane * LDX + c_lo + lane] += ca;
            __syncwarp();
            if (lane < 16) {
                const int i = lane;
                double trow[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) trow[j] = (j == i) ? tau[i] : 0.0;
#pragma unroll
                for (int j = 1; j < 16; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int l = 0; l < 16; ++l)
                        if (l >= i && l < j) acc = fma(trow[l], Sg[l * 16 + j], acc);
                    if (i < j) trow[j] = -tau[j] * acc;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) Tg[i * 16 + j] = trow[j];
            }
        }
        __syncthreads();
        // ---- N[16g:, 16g:] <- N - ((N V) T^T) V^T, one 8-row tile per warp at a time.
        // B fragments of T^T blocks: lane (q,r) supplies T[row q'][col l] as B[k = l][n = q'].
        const double t11a = Tg[q * 16 + 2 * r], t11b = Tg[q * 16 + 2 * r + 1];
        const double t12a = Tg[q * 16 + 8 + 2 * r], t12b = Tg[q * 16 + 8 + 2 * r + 1];
        const double t22a = Tg[(8 + q) * 16 + 8 + 2 * r], t22b = Tg[(8 + q) * 16 + 8 + 2 * r + 1];
        const double *Xa = X, *Xb = X + 8 * LDX;
        for (int rt = 2 * g + w; rt < NT; rt += 8) {
            double2 *Nrow = reinterpret_cast<double2 *>(N + (size_t)(8 * rt + q) * n) + r;
            double ya0 = 0.0, ya1 = 0.0, ya2 = 0.0, ya3 = 0.0;  // Ya: two partial accumulators
            double yb0 = 0.0, yb1 = 0.0, yb2 = 0.0, yb3 = 0.0;  // Yb
            for (int ct = 2 * g; ct < NT; ++ct) {
                const double2 a = Nrow[4 * ct];
                const int c0 = 8 * ct + 2 * r;
                dmma8x8x4(ya0, ya1, a.x, Xa[q * LDX + c0]);
                dmma8x8x4(yb0, yb1, a.x, Xb[q * LDX + c0]);
                dmma8x8x4(ya2, ya3, a.y, Xa[q * LDX + c0 + 1]);
                dmma8x8x4(yb2, yb3, a.y, Xb[q * LDX + c0 + 1]);
            }
            const double Ya0 = ya0 + ya2, Ya1 = ya1 + ya3, Yb0 = yb0 + yb2, Yb1 = yb1 + yb3;
            // Z = Y T^T with T = [[T11, T12], [0, T22]]:  Za = Ya T11^T + Yb T12^T, Zb = Yb T22^T
            double za0 = 0.0, za1 = 0.0, zb0 = 0.0, zb1 = 0.0;
            dmma8x8x4(za0, za1, Ya0, t11a);
            dmma8x8x4(zb0, zb1, Yb0, t22a);
            dmma8x8x4(za0, za1, Ya1, t11b);
            dmma8x8x4(zb0, zb1, Yb1, t22b);
            dmma8x8x4(za0, za1, Yb0, t12a);
            dmma8x8x4(za0, za1, Yb1, t12b);
            for (int ct = 2 * g; ct < NT; ++ct) {
                double2 c2 = Nrow[4 * ct];
                const int cc = 8 * ct + q;
                dmma8x8x4(c2.x, c2.y, -za0, Xa[(2 * r) * LDX + cc]);
                dmma8x8x4(c2.x, c2.y, -za1, Xa[(2 * r + 1) * LDX + cc]);
                dmma8x8x4(c2.x, c2.y, -zb0, Xb[(2 * r) * LDX + cc]);
                dmma8x8x4(c2.x, c2.y, -zb1, Xb[(2 * r + 1) * LDX + cc]);
                Nrow[4 * ct] = c2;
            }
        }
        __syncthreads();
    }
    // ---- R[i][c] = D[i] N[c][i]; D[n-1] = (-1)^(n-1) prod D (functions.py:59-60)
    if (tid == 0) Dv[n - 1] = ((n_negative + n - 1) & 1) ? -1.0 : 1.0;
    __threadfence_block();
    __syncthreads();
    for (int e = tid; e < n * (n >> 1); e += 256) {
        const int row = e / (n >> 1), c2 = (e % (n >> 1)) * 2;
        double2 *p2 = reinterpret_cast<double2 *>(N + (size_t)row * n) + (c2 >> 1);
        double2 v = *p2;
        v.x *= Dv[c2];
        v.y *= Dv[c2 + 1];
        *p2 = v;
    }
}
static inline bool wy_big_supported(int n) { return n > 128 && n <= 512 && (n % 8) == 0; }
static inline int launch_basis_wy_big(cudaStream_t st, int n, const double *normals,
                                      double *store, int64_t tasks) {
    const size_t smem = (size_t)(16 * (n + 1) + n + 256 + 256 + 16) * sizeof(double);
    if (cudaFuncSetAttribute(k_basis_wy_big, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return -1;
    k_basis_wy_big<<<(unsigned)tasks, 256, smem, st>>>(n, normals, (int)normals_per_task(n),
                                                       store, tasks);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static int g_basis_wy = 1;  // 1: Householder sweep on the tensor pipe (k_basis_wy)

template <int NG>
static int launch_basis_fast_t(cudaStream_t st, uint32_t k0, uint32_t k1, uint64_t chain_id0,
                               int block, int n, const int64_t *vis, int vis_stride,
                               uint32_t e0_fixed, int cnt, double *store, int64_t tasks,
                               int64_t task0, int64_t store_task0,
                               const double *normals = nullptr, int nn_pad = 0) {
    constexpr int NP = NG * 8;
    constexpr int NW = (NG <= 8) ? 4 : NG / 2;
    if (g_basis_wy || NG > 8) {
        if (NG > 8 && !normals) return -3;  // large blocks take their normals from k_normals
        const int gram_t = 2 * NG * 64 > CB2_LOGTAB_DOUBLES ? 2 * NG * 64 : CB2_LOGTAB_DOUBLES;
        const size_t smem_wy = (size_t)(NP * (NP + 1) + 2 * NP + gram_t) * sizeof(double);
        cudaError_t e2 = cudaFuncSetAttribute(k_basis_wy<NG, NW>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_wy);
        if (e2 != cudaSuccess) return -1;
        k_basis_wy<NG, NW><<<(unsigned)tasks, NW * 32, smem_wy, st>>>(
            k0, k1, chain_id0, block, n, vis, vis_stride, e0_fixed, cnt, store, task0,
            store_task0, normals, nn_pad);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if constexpr (NG <= 8) {
        const int threads = 2 * NP < 32 ? 32 : 2 * NP;
        const size_t smem = (size_t)(NP * NP + 2 * NP) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(k_basis_fast<NG>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem);
        if (e != cudaSuccess) return -1;
        k_basis_fast<NG><<<(unsigned)tasks, threads, smem, st>>>(
            k0, k1, chain_id0, block, n, vis, vis_stride, e0_fixed, cnt, store, task0,
            store_task0);
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    return -1;
}

static int launch_basis_fast_any(cudaStream_t st, uint32_t k0, uint32_t k1, uint64_t chain_id0,
                                 int block, int n, const int64_t *vis, int vis_stride,
                                 uint32_t e0_fixed, int cnt, double *store, int64_t tasks,
                                 int64_t task0, int64_t store_task0,
                                 const double *normals = nullptr, int nn_pad = 0) {
    int NG = (n + 7) / 8;
    if (NG > 8) NG = (NG + 1) & ~1;  // large blocks: even group counts only (2 tiles per warp)
#define CB2_BF(G)                                                                           \
    case G:                                                                                 \
        return launch_basis_fast_t<G>(st, k0, k1, chain_id0, block, n, vis, vis_stride,     \
                                      e0_fixed, cnt, store, tasks, task0, store_task0,      \
                                      normals, nn_pad);
    switch (NG) {
        CB2_BF(1) CB2_BF(2) CB2_BF(3) CB2_BF(4) CB2_BF(5) CB2_BF(6) CB2_BF(7) CB2_BF(8)
        CB2_BF(10) CB2_BF(12) CB2_BF(14) CB2_BF(16)
    }
#undef CB2_BF
    return -1;
}

static inline int launch_basis_fast(cudaStream_t st, uint32_t k0, uint32_t k1,
                                    uint64_t chain_id0, int block, int n, const int64_t *vis,
                                    int vis_stride, int cnt, double *store, int64_t n_chains,
                                    const double *normals = nullptr) {
    return launch_basis_fast_any(st, k0, k1, chain_id0, block, n, vis, vis_stride, 0u, cnt,
                                 store, n_chains * cnt, 0, 0, normals,
                                 normals ? (int)normals_per_task(n) : 0);
}

static inline int launch_basis_fast_one(cudaStream_t st, uint32_t k0, uint32_t k1, uint64_t gid,
                                        int block, int n, uint32_t epoch, double *out) {
    // chain_id0 = gid, task 0 -> chain 0
    return launch_basis_fast_any(st, k0, k1, gid, block, n, nullptr, 0, epoch, 1, out, 1, 0, 0);
}

// =====================================================================================
// DMMA step kernel
// =====================================================================================
// Packed constant block staged into shared memory (all offsets in doubles, 16B aligned):
struct FastPackDesc {
    int NT;        // DP/8
    int n_modes;
    int tri_like;  // likelihood matrix lower-triangular in sorted coordinates
    int off_T;     // fragment-ordered T      : tri blocks, 64 doubles each
    int off_A;     // fragment-ordered L^-1 P : per mode, tri or dense blocks
    int blocks_A;  // blocks per mode
    int off_mu;    // [n_modes][DP]  means in sorted coordinates
    int off_c0;    // [n_modes]      d log 2pi + logdet
    int off_w;     // [n_modes]      weights
    int off_lower, off_upper, off_loc, off_mls, off_isc;  // [DP] each (mls = log-normalisation, isc = s)
    int off_pa, off_pb;  // [DP] shape parameters of the generic 1-D priors
    int off_flags; // [DP] as doubles: bit0 non-uniform prior, bit1 periodic, bits 8.. prior kind
    int off_iofj;  // [DP] as doubles: sampler index of sorted j (or -1 for padding)
    int off_klo, off_kup;  // [DP] order-preserving int64 keys of lower / upper
    int iofj_identity;  // D == DP, i_of_j[j] == j and the row stride keeps 16-byte alignment
    int vec_ok;         // every block with n_b >= 2 has even start and even size
    int total;     // doubles
};

// the one likelihood is a Gaussian mixture the register-resident kernels can evaluate
static inline bool fast_step_supported_like(const ModelDev &M) {
    const LikeDev &L = M.likes[0];
    if (L.kind != 0 || L.dim != M.D || L.derived) return false;
    if (L.n_modes > 4) return false;
    return true;
}
static inline bool fast_step_supported(const ModelDev &M, size_t n_likes) {
    if (M.drag || M.D > 64 || n_likes != 1) return false;
    return fast_step_supported_like(M);
}


__device__ __forceinline__ double quad_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// out[nt][h] += sum_k Mat[8nt+q .. ][k] * a[k]  for the 8 chains of the warp.
// frag: blocks of 64 doubles, block(nt, m) holds for lane l=(q,r): {Mat[8nt+q][8m+2r],
// Mat[8nt+q][8m+2r+1]}.  TRI: only m <= nt present, block index nt(nt+1)/2 + m.
template <int NT, bool TRI>
__device__ __forceinline__ void warp_matvec8(const double *__restrict__ frag, int lane,
                                             const double (&a)[NT][2], double (&out)[NT][2]) {
    const double2 *f2 = reinterpret_cast<const double2 *>(frag) + lane;
#pragma unroll
    for (int m = 0; m < NT; ++m) {
        // all B fragments of this k-block first, then the MMAs ordered so that consecutive
        // instructions update DIFFERENT accumulators (no back-to-back dependent DMMAs)
        double2 b[NT];
#pragma unroll
        for (int nt = TRI ? m : 0; nt < NT; ++nt) {
            const int blk = TRI ? (nt * (nt + 1)) / 2 + m : nt * NT + m;
            b[nt] = f2[blk * 32];
        }
#pragma unroll
        for (int nt = TRI ? m : 0; nt < NT; ++nt) dmma8x8x4(out[nt][0], out[nt][1], a[m][0], b[nt].x);
#pragma unroll
        for (int nt = TRI ? m : 0; nt < NT; ++nt) dmma8x8x4(out[nt][0], out[nt][1], a[m][1], b[nt].y);
    }
}

// loads that must be ISSUED where they are written (prefetch): asm volatile keeps their
// program order relative to the (asm volatile) MMAs, the scoreboard wait happens at first use
__device__ __forceinline__ double ldg_f64_early(const double *p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ldg_f64x2_early(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// direction of a planned step in the lane's fragment layout (zero outside the block;
// +1 at the block's single coordinate for a 1-parameter block).  vec: every block starts at
// an even sorted index and has even size -> 16-byte loads.
template <int NT>
__device__ __forceinline__ void fetch_direction(const ModelDev &M, const WindowDev &W,
                                                int2 pl, int r, bool vec, double (&u)[NT][2]) {
    const int b = pl.y, nb = M.bsize[b], j0 = M.jstart[b];
    if (nb >= 2) {
        const double *Rk = W.basis[b] + (size_t)pl.x * (size_t)nb;
        if (vec) {
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int j = 8 * n + 2 * r - j0;
                if (j >= 0 && j < nb) {
                    const double2 v2 = ldg_f64x2_early(reinterpret_cast<const double2 *>(Rk + j));
                    u[n][0] = v2.x;
                    u[n][1] = v2.y;
                } else {
                    u[n][0] = 0.0;
                    u[n][1] = 0.0;
                }
            }
        } else {
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j = 8 * n + 2 * r + h - j0;
                    u[n][h] = (j >= 0 && j < nb) ? ldg_f64_early(Rk + j) : 0.0;
                }
        }
    } else {
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int h = 0; h < 2; ++h) u[n][h] = (8 * n + 2 * r + h == j0) ? 1.0 : 0.0;
    }
}
__device__ __forceinline__ int2 ldg_int2_early(const int2 *p) {
    int2 v;
    asm volatile("ld.global.nc.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// Proposal-side randomness of a window, generated ahead of the serial accept chain:
// draws[(chain*len + s)] = { r (signed for 1-parameter blocks), Exp(1) of the accept test }
// (proposal.py:71-93, mcmc.py:683).  None of it depends on the chain state.
__global__ void k_draws(ModelDev M, const uint8_t *__restrict__ tape, int tape_len,
                        int64_t tape_base, uint8_t const_b, int64_t n_chains, uint64_t t0,
                        int n_steps, double2 *__restrict__ draws) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= n_chains * n_steps) return;
    const int64_t chain = e / n_steps;
    const int s = (int)(e % n_steps);
    const uint64_t t = t0 + (uint64_t)s;
    const uint64_t gid = M.chain_id0 + (uint64_t)chain;
    const int b = tape ? tape[chain * tape_len + (int64_t)(t - tape_base)] : const_b;
    double r, sign;
    draw_radial(M, gid, t, 0, M.bsize[b], r, sign);
    double2 o;
    o.x = (M.bsize[b] >= 2) ? r : sign * r;
    o.y = draw_accept_exp(M, gid, t, 0);
    draws[e] = o;
}

// Where every step of the window finds its direction: CyclicIndexRandomizer.next +
// RandDirectionProposer's loop_index / basis epoch (proposal.py:46-55,63-68) advanced for all
// steps of the window, one thread per chain.  plan[chain*n_steps + s] = { row index of the
// direction inside the block's basis store (units of n_b doubles), block }.  Also advances the
// persistent visit counters (none of this depends on accept/reject).
__global__ void k_plan(ModelDev M, ChainState S, WindowDev W, int64_t n_chains, uint64_t t0,
                       int n_steps, int2 *__restrict__ plan) {
    const int64_t chain = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (chain >= n_chains) return;
    const int NB = M.n_blocks, NV = NB + 1;
    int kcur[CB2_MAX_BLOCKS], slot[CB2_MAX_BLOCKS], cntv[CB2_MAX_BLOCKS];
    for (int b = 0; b < NB; ++b) {
        kcur[b] = (int)(S.vis[chain * NV + b] % M.bsize[b]);
        slot[b] = 0;
        cntv[b] = 0;
    }
    bool bad = false;
    for (int s = 0; s < n_steps; ++s) {
        const uint64_t t = t0 + (uint64_t)s;
        const int b = W.tape_main ? W.tape_main[chain * W.len_main + (int64_t)(t - W.base_main)]
                                  : W.const_main;
        const int nb = M.bsize[b];
        int sl = slot[b];
        if (nb >= 2 && sl >= W.cnt[b]) { bad = true; sl = 0; }
        int2 o;
        o.x = (nb >= 2) ? (int)((chain * W.cnt[b] + sl) * nb + kcur[b]) : 0;
        o.y = b;
        plan[chain * n_steps + s] = o;
        const bool wrap = (kcur[b] + 1 == nb);
        kcur[b] = wrap ? 0 : kcur[b] + 1;
        slot[b] += wrap ? 1 : 0;
        cntv[b] += 1;
    }
    for (int b = 0; b < NB; ++b) S.vis[chain * NV + b] += cntv[b];
    if (bad) atomicOr(&S.flags[chain], CB2_FLAG_INTERNAL);
}

template <int NT>
__global__ void __launch_bounds__(256, 1)
k_step_fast(ModelDev M, ChainState S, WindowDev W, const double *__restrict__ gpack,
            FastPackDesc P, const double2 *__restrict__ draws,
            const int2 *__restrict__ plan, int64_t n_chains, uint64_t t0, int n_steps) {
    constexpr int DP = NT * 8;
    extern __shared__ __align__(16) double fsm[];
    __shared__ __align__(8) unsigned long long mbar;
    double *pack = fsm;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
    // ---- stage the constant block with a TMA bulk copy (global -> shared, mbarrier)
    const uint32_t bytes = (uint32_t)P.total * 8u;
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(&mbar);
    const uint32_t dst_a = (uint32_t)__cvta_generic_to_shared(pack);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"(dst_a), "l"(gpack), "r"(bytes), "r"(mbar_a)
            : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(mbar_a)
                : "memory");
        }
    }

    const int q = lane >> 2, r = lane & 3;
    const int64_t tile = blockIdx.x * (int64_t)nwarps + wid;
    const int64_t chain_raw = tile * 8 + q;
    const bool active = chain_raw < n_chains;
    const int64_t chain = active ? chain_raw : (n_chains - 1);
    const int D = M.D;
    const double *Tf = pack + P.off_T, *Af = pack + P.off_A;
    const double *lower = pack + P.off_lower, *upper = pack + P.off_upper;
    const int *iofj = reinterpret_cast<const int *>(pack + P.off_iofj);
    const int *pflag = reinterpret_cast<const int *>(pack + P.off_flags);

    // ---- load chain state (sorted coordinates): element j = 8n + 2r + h
    double xs[NT][2];
    uint32_t m_norm = 0, m_per = 0;  // bit (2n+h): normal prior / periodic
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 8 * n + 2 * r + h;
            const int i = iofj[j];
            xs[n][h] = (i >= 0) ? S.x[chain * D + i] : 0.0;
            const int fl = pflag[j];
            m_norm |= (uint32_t)(fl & 1) << (2 * n + h);
            m_per |= (uint32_t)((fl >> 1) & 1) << (2 * n + h);
        }
    double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike = S.ll[chain];
    long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
              burn_left = S.burn_left[chain], added_w = S.added_w[chain],
              n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
    uint32_t flags = S.flags[chain];
    const bool any_special = M.any_periodic || M.any_normal;
    const double2 *my_draws = draws + chain * (int64_t)n_steps;
    const int2 *my_plan = plan + chain * (int64_t)n_steps;
    const double inv_T = M.temperature;
    const bool vec = P.vec_ok != 0;

    // software pipeline: the direction and draws of step s+1 are loaded during step s
    double un[NT][2];
    double2 dn;
    fetch_direction<NT>(M, W, my_plan[0], r, vec, un);
    dn = my_draws[0];
    for (int s = 0; s < n_steps; ++s) {
        // ---- v = R[:,k] * r * scale (proposal.py:69) / +-r*scale (:86-93)
        double v[NT][2];
        const double rs = dn.x, e_acc = dn.y;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            v[n][0] = un[n][0] * rs * M.proposal_scale;
            v[n][1] = un[n][1] * rs * M.proposal_scale;
        }
        if (s + 1 < n_steps) {  // prefetch (independent of the accept chain)
            fetch_direction<NT>(M, W, ldg_int2_early(my_plan + s + 1), r, vec, un);
            dn = ldg_f64x2_early(my_draws + s + 1);
        }
        // ---- trial = x + T v  (proposal.py:224) on the FP64 tensor pipe
        double xt[NT][2];
#pragma unroll
        for (int n = 0; n < NT; ++n) { xt[n][0] = 0.0; xt[n][1] = 0.0; }
        warp_matvec8<NT, true>(Tf, lane, v, xt);
#pragma unroll
        for (int n = 0; n < NT; ++n) { xt[n][0] += xs[n][0]; xt[n][1] += xs[n][1]; }
        // ---- reduce_periodic (prior.py:658-676), bounds + prior (prior.py:733-763)
        bool bad = false;
        double ps = 0.0;
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const double2 lo2 = *reinterpret_cast<const double2 *>(lower + 8 * n + 2 * r);
            const double2 up2 = *reinterpret_cast<const double2 *>(upper + 8 * n + 2 * r);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double xv = xt[n][h];
                const double lo = h ? lo2.y : lo2.x, up = h ? up2.y : up2.x;
                if (any_special) {
                    const int j = 8 * n + 2 * r + h;
                    if ((m_per >> (2 * n + h)) & 1u) {
                        double qq = (xv - lo) / (up - lo);
                        qq = qq - floor(qq);
                        xv = qq * (up - lo) + lo;
                        xt[n][h] = xv;
                    }
                    if ((m_norm >> (2 * n + h)) & 1u) {
                        const double zz = (xv - pack[P.off_loc + j]) / pack[P.off_isc + j];
                        if (M.any_generic)
                            ps += pack[P.off_mls + j] +
                                  prior1d_shape(pflag[j] >> 8, zz, pack[P.off_pa + j],
                                                pack[P.off_pb + j]);
                        else
                            ps += pack[P.off_mls + j] - zz * zz / 2;
                    }
                }
                if (!(xv <= up) || !(xv >= lo) || !isfinite(xv)) bad = true;
            }
        }
        bad = __shfl_xor_sync(0xffffffffu, (int)bad, 1) | (int)bad;
        bad = __shfl_xor_sync(0xffffffffu, (int)bad, 2) | (int)bad;
        double t_prior, t_like = 0.0, t_post;
        if (M.any_normal) ps = quad_sum(ps);
        t_prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
        // ---- GaussianMixture.logp (gaussian_mixture.py:138-163); computed for every chain
        // of the warp (the MMA is warp-wide), discarded where the prior is -inf
        {
            double lp0 = 0.0, lp1 = 0.0, lp2 = 0.0, lp3 = 0.0;
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
        }
        t_post = bad ? -CUDART_INF : (t_prior + t_like);
        // ---- metropolis_accept (mcmc.py:670-683)
        bool acc;
        if (t_post == -CUDART_INF) acc = false;
        else if (t_post > logpost) acc = true;
        else acc = e_acc > (logpost - t_post) / inv_T;
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
            for (int n = 0; n < NT; ++n) { xs[n][0] = xt[n][0]; xs[n][1] = xt[n][1]; }
            logpost = t_post; logprior = t_prior; loglike = t_like;
            weight = 1; prior_rej = 0; n_acc += 1;
        } else {
            weight += 1;
            if (t_prior == -CUDART_INF) prior_rej += 1;
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
        }
    }
}

template <int NT>
static int launch_step_fast_t(cudaStream_t st, const ModelDev &M, const ChainState &S,
                              const WindowDev &W, const double *gpack, const FastPackDesc &P,
                              const double2 *draws, const int2 *plan, int64_t n_chains,
                              uint64_t t0, int n_steps, int sm_count) {
    const int64_t tiles = (n_chains + 7) / 8;
    int wpc = (int)((tiles + sm_count - 1) / sm_count);
    if (wpc < 1) wpc = 1;
    if (wpc > 8) wpc = 8;
    const int grid = (int)((tiles + wpc - 1) / wpc);
    const size_t smem = (size_t)P.total * 8;
    if (smem > 200 * 1024) return -2;
    if (cudaFuncSetAttribute(k_step_fast<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return -1;
    k_step_fast<NT><<<grid, wpc * 32, smem, st>>>(M, S, W, gpack, P, draws, plan, n_chains, t0,
                                                  n_steps);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static inline int launch_draws(cudaStream_t st, const ModelDev &M, const ChainState &S,
                               const WindowDev &W, int64_t n_chains, uint64_t t0, int n_steps,
                               double2 *draws, int2 *plan) {
    const int64_t tot = n_chains * n_steps;
    const int bs = 256;
    k_draws<<<(unsigned)((tot + bs - 1) / bs), bs, 0, st>>>(M, W.tape_main, W.len_main,
                                                            W.base_main, W.const_main, n_chains,
                                                            t0, n_steps, draws);
    k_plan<<<(unsigned)((n_chains + 127) / 128), 128, 0, st>>>(M, S, W, n_chains, t0, n_steps,
                                                               plan);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

static inline int launch_step_fast(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                   const WindowDev &W, const double *gpack,
                                   const FastPackDesc &P, const double2 *draws,
                                   const int2 *plan, int64_t n_chains, uint64_t t0, int n_steps,
                                   int sm_count) {
#define CB2_SF(N)                                                                       \
    case N:                                                                             \
        return launch_step_fast_t<N>(st, M, S, W, gpack, P, draws, plan, n_chains, t0,        \
                                     n_steps, sm_count);
    switch (P.NT) {
        CB2_SF(1) CB2_SF(2) CB2_SF(3) CB2_SF(4) CB2_SF(5) CB2_SF(6) CB2_SF(7) CB2_SF(8)
    }
#undef CB2_SF
    return -1;
}

// =====================================================================================
// Producer/consumer step kernel (single Gaussian mode, no periodic parameters)
// =====================================================================================
// Every quantity on the proposal side of a Metropolis step is independent of the chain
// state: with delta_s = T v_s (proposal.py:224) and, by linearity of the Gaussian's
// whitening map, L^-1 (x + delta_s - mu) = y + L^-1 delta_s, both matrix products of a
// proposal can be formed ahead of the serial accept chain.  Per SM one CTA of 2*wpc warps:
//   * producer warp p : for its 8 chains, streams steps s = 0,1,...: direction + radius ->
//     delta_s = T v_s and w_s = L^-1 P delta_s on the FP64 tensor pipe (m8n8k4 DMMA), written
//     to a shared-memory ring in fragment order, mbarrier "full";
//   * consumer warp p : keeps x and y = L^-1 P (x - mu) in registers; per step reads
//     (delta_s, w_s): bounds/prior of x + delta, |y + w|^2, Metropolis test
//     (mcmc.py:670-683), bookkeeping and row store (mcmc.py:685-748), mbarrier "empty".
// The tensor pipe therefore never waits for an accept decision.  y is recomputed from x at
// the start of every window, which bounds the drift of the incremental update.
#define CB2_PC_RING 3

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait_backoff(uint32_t a, uint32_t parity) {
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(64);  // leave the issue slots to the producer warps
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}

static inline bool pc_step_supported(const ModelDev &M, const FastPackDesc &P) {
    return P.n_modes == 1 && !M.any_periodic && !M.any_generic;
}

template <int NT, bool HAS_NORMAL>
__global__ void __launch_bounds__(448, 1)
k_step_pc(ModelDev M, ChainState S, WindowDev W, const double *__restrict__ gpack,
          FastPackDesc P, const double2 *__restrict__ draws, const int2 *__restrict__ plan,
          int64_t n_chains, uint64_t t0, int n_steps, int wpc) {
    constexpr int DP = NT * 8;
    constexpr int SLOT = 2 * NT * 32 * 2 + 8;  // ring slot: delta + w (fragment order) + 8 accept draws
    extern __shared__ __align__(16) double fsm[];
    __shared__ __align__(8) unsigned long long mbar_pack;
    __shared__ __align__(8) unsigned long long mbar_full[8 * CB2_PC_RING];
    __shared__ __align__(8) unsigned long long mbar_empty[8 * CB2_PC_RING];
    double *pack = fsm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp >= wpc;
    const int pair = is_producer ? warp - wpc : warp;
    const uint32_t bytes = (uint32_t)P.total * 8u;
    const uint32_t mb_pack = (uint32_t)__cvta_generic_to_shared(&mbar_pack);
    if (tid == 0) {
        mbar_init(mb_pack, 1);
        for (int i = 0; i < wpc * CB2_PC_RING; ++i) {
            mbar_init((uint32_t)__cvta_generic_to_shared(&mbar_full[i]), 1);
            mbar_init((uint32_t)__cvta_generic_to_shared(&mbar_empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_pack),
                     "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
            ::"r"((uint32_t)__cvta_generic_to_shared(pack)), "l"(gpack), "r"(bytes), "r"(mb_pack)
            : "memory");
    }
    mbar_wait(mb_pack, 0);

    double *ring = fsm + P.total + (size_t)pair * CB2_PC_RING * SLOT;
    const int q = lane >> 2, r = lane & 3;
    const int64_t tile = blockIdx.x * (int64_t)wpc + pair;
    const int64_t chain_raw = tile * 8 + q;
    const bool active = chain_raw < n_chains;
    const int64_t chain = active ? chain_raw : (n_chains - 1);
    const int D = M.D;
    const double2 *my_draws = draws + chain * (int64_t)n_steps;
    const int2 *my_plan = plan + chain * (int64_t)n_steps;
    const bool vec = P.vec_ok != 0;
    const int *iofj = reinterpret_cast<const int *>(pack + P.off_iofj);

    if (is_producer) {
        // ================================ producer ====================================
        const double *Tf = pack + P.off_T, *Af = pack + P.off_A;
        // `un` holds the direction of the next step; it becomes v in place, and is refilled
        // (prefetch) as soon as T v has been issued
        double un[NT][2];
        double2 dr_next;  // {radius, Exp(1) of the accept test} of the next step
        fetch_direction<NT>(M, W, my_plan[0], r, vec, un);
        dr_next = ldg_f64x2_early(my_draws);
        for (int s = 0; s < n_steps; ++s) {
            const int slot = s % CB2_PC_RING;
            const uint32_t use = (uint32_t)(s / CB2_PC_RING);
            const double rs_cur = dr_next.x, e_cur = dr_next.y;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                un[n][0] = un[n][0] * rs_cur * M.proposal_scale;
                un[n][1] = un[n][1] * rs_cur * M.proposal_scale;
            }
            double dl[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) { dl[n][0] = 0.0; dl[n][1] = 0.0; }
            warp_matvec8<NT, true>(Tf, lane, un, dl);
            if (s + 1 < n_steps) {
                fetch_direction<NT>(M, W, ldg_int2_early(my_plan + s + 1), r, vec, un);
                dr_next = ldg_f64x2_early(my_draws + s + 1);
            }
            // wait until the consumer released this slot (first pass: free)
            if (use > 0)
                mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar_empty[pair * CB2_PC_RING + slot]),
                          (use - 1) & 1u);
            double2 *sl = reinterpret_cast<double2 *>(ring + (size_t)slot * SLOT);
#pragma unroll
            for (int n = 0; n < NT; ++n) sl[n * 32 + lane] = make_double2(dl[n][0], dl[n][1]);
            double wv[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) { wv[n][0] = 0.0; wv[n][1] = 0.0; }
            if (P.tri_like) warp_matvec8<NT, true>(Af, lane, dl, wv);
            else warp_matvec8<NT, false>(Af, lane, dl, wv);
#pragma unroll
            for (int n = 0; n < NT; ++n)
                sl[(NT + n) * 32 + lane] = make_double2(wv[n][0], wv[n][1]);
            if (r == 0) ring[(size_t)slot * SLOT + 2 * NT * 64 + q] = e_cur;
            __syncwarp();
            if (lane == 0)
                mbar_arrive((uint32_t)__cvta_generic_to_shared(&mbar_full[pair * CB2_PC_RING + slot]));
        }
    } else {
        // ================================ consumer ====================================
        const double *lower = pack + P.off_lower, *upper = pack + P.off_upper;
        const int *pflag = reinterpret_cast<const int *>(pack + P.off_flags);
        double xs[NT][2], ys[NT][2];
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
        {   // y = L^-1 P (x - mu), refreshed at every window start
            const double *mu = pack + P.off_mu;
            double z[NT][2];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const double2 m2 = *reinterpret_cast<const double2 *>(mu + 8 * n + 2 * r);
                z[n][0] = xs[n][0] - m2.x;
                z[n][1] = xs[n][1] - m2.y;
                ys[n][0] = 0.0;
                ys[n][1] = 0.0;
            }
            if (P.tri_like) warp_matvec8<NT, true>(pack + P.off_A, lane, z, ys);
            else warp_matvec8<NT, false>(pack + P.off_A, lane, z, ys);
        }
        double logpost = S.logpost[chain], logprior = S.logprior[chain], loglike = S.ll[chain];
        long long weight = S.weight[chain], prior_rej = S.prior_rej[chain],
                  burn_left = S.burn_left[chain], added_w = S.added_w[chain],
                  n_rows = S.n_rows[chain], n_acc = S.n_acc[chain];
        uint32_t flags = S.flags[chain];
        const double c0 = pack[P.off_c0];
        for (int s = 0; s < n_steps; ++s) {
            const int slot = s % CB2_PC_RING;
            const uint32_t use = (uint32_t)(s / CB2_PC_RING);
            mbar_wait((uint32_t)__cvta_generic_to_shared(&mbar_full[pair * CB2_PC_RING + slot]),
                      use & 1u);
            const double e_acc = ring[(size_t)slot * SLOT + 2 * NT * 64 + q];
            const double2 *sl = reinterpret_cast<const double2 *>(ring + (size_t)slot * SLOT);
            bool bad = false;
            double ps = 0.0, qsum = 0.0;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const double2 d2 = sl[n * 32 + lane];
                const double2 w2 = sl[(NT + n) * 32 + lane];
                const double2 lo2 = *reinterpret_cast<const double2 *>(lower + 8 * n + 2 * r);
                const double2 up2 = *reinterpret_cast<const double2 *>(upper + 8 * n + 2 * r);
                const double x0 = xs[n][0] + d2.x, x1 = xs[n][1] + d2.y;
                const double y0 = ys[n][0] + w2.x, y1 = ys[n][1] + w2.y;
                qsum += y0 * y0 + y1 * y1;
                if (!(x0 <= up2.x) || !(x0 >= lo2.x) || !isfinite(x0)) bad = true;
                if (!(x1 <= up2.y) || !(x1 >= lo2.y) || !isfinite(x1)) bad = true;
                if (HAS_NORMAL) {
                    if ((m_norm >> (2 * n)) & 1u) {
                        const int j = 8 * n + 2 * r;
                        const double zz = (x0 - pack[P.off_loc + j]) / pack[P.off_isc + j];
                        ps += pack[P.off_mls + j] - zz * zz / 2;
                    }
                    if ((m_norm >> (2 * n + 1)) & 1u) {
                        const int j = 8 * n + 2 * r + 1;
                        const double zz = (x1 - pack[P.off_loc + j]) / pack[P.off_isc + j];
                        ps += pack[P.off_mls + j] - zz * zz / 2;
                    }
                }
            }
            bad = __shfl_xor_sync(0xffffffffu, (int)bad, 1) | (int)bad;
            bad = __shfl_xor_sync(0xffffffffu, (int)bad, 2) | (int)bad;
            qsum = quad_sum(qsum);
            if (HAS_NORMAL) ps = quad_sum(ps);
            const double t_prior = bad ? -CUDART_INF : (M.uniform_logp + ps);
            const double t_like = -0.5 * (c0 + qsum);
            const double t_post = bad ? -CUDART_INF : (t_prior + t_like);
            bool acc;
            if (t_post == -CUDART_INF) acc = false;
            else if (t_post > logpost) acc = true;
            else acc = e_acc > (logpost - t_post) / M.temperature;
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
                                if (P.iofj_identity) {  // D == DP, sorted == sampler order
#pragma unroll
                                    for (int n = 0; n < NT; ++n)
                                        *reinterpret_cast<double2 *>(row + 2 + 8 * n + 2 * r) =
                                            make_double2(xs[n][0], xs[n][1]);
                                } else {
#pragma unroll
                                    for (int n = 0; n < NT; ++n)
#pragma unroll
                                        for (int h = 0; h < 2; ++h) {
                                            const int i = iofj[8 * n + 2 * r + h];
                                            if (i >= 0) row[2 + i] = xs[n][h];
                                        }
                                }
                            }
                            n_rows += 1;
                        }
                    }
                } else burn_left -= 1;
#pragma unroll
                for (int n = 0; n < NT; ++n) {  // same sums as in the evaluation above
                    const double2 d2 = sl[n * 32 + lane];
                    const double2 w2 = sl[(NT + n) * 32 + lane];
                    xs[n][0] += d2.x; xs[n][1] += d2.y;
                    ys[n][0] += w2.x; ys[n][1] += w2.y;
                }
                logpost = t_post; logprior = t_prior; loglike = t_like;
                weight = 1; prior_rej = 0; n_acc += 1;
            } else {
                weight += 1;
                if (t_prior == -CUDART_INF) prior_rej += 1;
                const long long sgn = (burn_left > 0) - (burn_left < 0);
                if (weight - prior_rej > M.max_tries * (1 + 9 * sgn)) flags |= CB2_FLAG_STUCK;
            }
            __syncwarp();
            if (lane == 0)  // hand the slot back to the producer
                mbar_arrive((uint32_t)__cvta_generic_to_shared(&mbar_empty[pair * CB2_PC_RING + slot]));
        }
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
                S.n_rows[chain] = n_rows; S.n_acc[chain] = n_acc;
                if (flags) atomicOr(&S.flags[chain], flags);
            }
        }
    }
}

template <int NT>
static int launch_step_pc_t(cudaStream_t st, const ModelDev &M, const ChainState &S,
                            const WindowDev &W, const double *gpack, const FastPackDesc &P,
                            const double2 *draws, const int2 *plan, int64_t n_chains,
                            uint64_t t0, int n_steps, int sm_count) {
    const int64_t tiles = (n_chains + 7) / 8;
    int wpc = (int)((tiles + sm_count - 1) / sm_count);
    if (wpc < 1) wpc = 1;
    if (wpc > 7) wpc = 7;
    const int grid = (int)((tiles + wpc - 1) / wpc);
    const size_t slot = (size_t)2 * NT * 32 * 2 + 8;
    const size_t smem = ((size_t)P.total + (size_t)wpc * CB2_PC_RING * slot) * 8;
    if (smem > 220 * 1024) return -2;
    cudaError_t e;
    if (M.any_normal) {
        e = cudaFuncSetAttribute(k_step_pc<NT, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1000 - (int)e;
        k_step_pc<NT, true><<<grid, 2 * wpc * 32, smem, st>>>(M, S, W, gpack, P, draws, plan,
                                                               n_chains, t0, n_steps, wpc);
    } else {
        e = cudaFuncSetAttribute(k_step_pc<NT, false>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return -1000 - (int)e;
        k_step_pc<NT, false><<<grid, 2 * wpc * 32, smem, st>>>(M, S, W, gpack, P, draws, plan,
                                                                n_chains, t0, n_steps, wpc);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -2000 - (int)e;
}

static inline int launch_step_pc(cudaStream_t st, const ModelDev &M, const ChainState &S,
                                 const WindowDev &W, const double *gpack, const FastPackDesc &P,
                                 const double2 *draws, const int2 *plan, int64_t n_chains,
                                 uint64_t t0, int n_steps, int sm_count) {
#define CB2_PC(N)                                                                         \
    case N:                                                                               \
        return launch_step_pc_t<N>(st, M, S, W, gpack, P, draws, plan, n_chains, t0, n_steps, \
                                   sm_count);
    switch (P.NT) {
        CB2_PC(1) CB2_PC(2) CB2_PC(3) CB2_PC(4) CB2_PC(5) CB2_PC(6) CB2_PC(7) CB2_PC(8)
    }
#undef CB2_PC
    return -1;
}
