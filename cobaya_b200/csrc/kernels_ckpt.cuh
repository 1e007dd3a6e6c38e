// kernels_ckpt.cuh -- the D x D part of check_convergence_and_learn_proposal on the device
// for D <= 64 (sm_100a): from the all-reduced sufficient statistics of the chains
// (cb2_moments, mcmc.py:785-793) one CTA forms
//   W = sum N_c C_c / sum N_c, B = cov of the chain means (mcmc.py:856-860),
//   R-1 = max eigenvalue of L^-1 (B / d d^T) L^-T with L = chol(W / d d^T), d = sqrt(diag B)
//         (mcmc.py:864-889: numpy cholesky, scipy dtrtri, numpy eigvalsh on the host),
//   the new proposal transform T = S L' of BlockedProposer.set_covariance applied to W
//         (proposal.py:226-260, tools.py:761-788), in block-sorted coordinates,
// and k_ckpt_pack writes T and G = (L^-1 P) T into the buffers the step kernels read
// (fragment-ordered T and G, column-major T) -- no host LAPACK, no re-upload.
// Everything lives in shared memory (five 64 x 64 matrices); the eigenvalues come from a
// two-sided Jacobi iteration with the round-robin parallel ordering (32 disjoint rotations
// per step, applied as 32 x 32 independent 2 x 2 blocks, one per thread).  The symmetric matrix is positive semi-definite, so max |eig| = max eig.
#pragma once
#include "kernels_fast.cuh"

#define CB2_CK_MAXD 64
#define CB2_CK_LD 65
#define CB2_CK_THREADS 1024

// index at position i of the round-robin order after `step` rotations (n2 even)
// (step already reduced modulo n2 - 1: no integer division in the sweep)
__device__ __forceinline__ int rr_index(int i, int step, int n2) {
    if (i == 0) return 0;
    int v = i - 1 - step;
    if (v < 0) v += n2 - 1;
    return v + 1;
}

struct CkptOut {          // device -> host, 8 + D doubles
    double M, N, acceptance, Rminus1, proposal_ok, chol_ok, sweeps, pad;
};

__device__ __forceinline__ int ck_cholesky(double *A, int n, int tid, int nt, int *fail) {
    // in-place lower Cholesky of the symmetric matrix A[n][LD] (upper part ignored)
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        if (tid == 0) {
            const double p = A[k * CB2_CK_LD + k];
            if (!(p > 0.0) || !isfinite(p)) *fail = 1;
            A[k * CB2_CK_LD + k] = sqrt(p > 0.0 ? p : 1.0);
        }
        __syncthreads();
        const double dk = A[k * CB2_CK_LD + k];
        for (int i = k + 1 + tid; i < n; i += nt) A[i * CB2_CK_LD + k] /= dk;
        __syncthreads();
        // trailing update of the lower triangle
        const int m = n - k - 1;
        for (int e = tid; e < m * m; e += nt) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) A[i * CB2_CK_LD + j] -= A[i * CB2_CK_LD + k] * A[j * CB2_CK_LD + k];
        }
    }
    __syncthreads();
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        if (j > i) A[i * CB2_CK_LD + j] = 0.0;
    }
    __syncthreads();
    return 0;
}

// sums: {M, S1 = sum N_c, S2 = sum N_c a_c, Sm[D], Smm[D*D], SC[D*D]} (cb2_moments)
// shift: the vector the means were shifted by; i_of_j: sorted index -> sampler index.
// Tnew[D*D]: row-major T in sorted coordinates; Wout[D*D]: W in sampler order.
__global__ void __launch_bounds__(CB2_CK_THREADS, 1)
k_ckpt_device(const double *__restrict__ sums, const double *__restrict__ shift,
              const int32_t *__restrict__ i_of_j, int D, double *__restrict__ Tnew,
              double *__restrict__ Wout, double *__restrict__ out) {
    extern __shared__ double cs[];
    double *W = cs;                              // [D][LD]
    double *B = W + CB2_CK_MAXD * CB2_CK_LD;     // corr of means, later the Jacobi matrix
    double *L = B + CB2_CK_MAXD * CB2_CK_LD;     // chol(norm W), later chol(corr)
    double *Li = L + CB2_CK_MAXD * CB2_CK_LD;    // L^-1, later scratch
    double *X = Li + CB2_CK_MAXD * CB2_CK_LD;    // products
    __shared__ double dvec[CB2_CK_MAXD], cvec[32], svec[32];
    __shared__ int fail, fail2;
    __shared__ double offnorm, diagnorm;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int DD = D * D;
    const double M = sums[0], S1 = sums[1], S2 = sums[2];
    const double *Sm = sums + 3, *Smm = sums + 3 + D, *SC = sums + 3 + D + DD;
    if (tid == 0) { fail = 0; fail2 = 0; }
    for (int e = tid; e < DD; e += nt) {
        const int i = e / D, j = e % D;
        const double w = 0.5 * (SC[i * D + j] + SC[j * D + i]) / S1;
        W[i * CB2_CK_LD + j] = w;
        Wout[e] = w;
        const double mi = Sm[i] / M, mj = Sm[j] / M;
        const double b = 0.5 * ((Smm[i * D + j] - M * mi * mj) + (Smm[j * D + i] - M * mj * mi)) /
                         (M - 1.0);
        B[i * CB2_CK_LD + j] = b;
    }
    __syncthreads();
    for (int i = tid; i < D; i += nt) {
        dvec[i] = sqrt(B[i * CB2_CK_LD + i]);
        out[8 + i] = Sm[i] / M + shift[i];
    }
    __syncthreads();
    for (int e = tid; e < DD; e += nt) {           // mcmc.py:864-866
        const int i = e / D, j = e % D;
        B[i * CB2_CK_LD + j] = B[i * CB2_CK_LD + j] / dvec[i] / dvec[j];
        L[i * CB2_CK_LD + j] = W[i * CB2_CK_LD + j] / dvec[i] / dvec[j];
    }
    __syncthreads();
    ck_cholesky(L, D, tid, nt, &fail);             // mcmc.py:871
    // L^-1 (dtrtri): column j by forward substitution, one thread per column
    for (int e = tid; e < CB2_CK_MAXD * CB2_CK_LD; e += nt) Li[e] = 0.0;
    __syncthreads();
    for (int j = tid; j < D; j += nt) {
        for (int i = j; i < D; ++i) {
            double s = (i == j) ? 1.0 : 0.0;
            for (int k = j; k < i; ++k) s -= L[i * CB2_CK_LD + k] * Li[k * CB2_CK_LD + j];
            Li[i * CB2_CK_LD + j] = s / L[i * CB2_CK_LD + i];
        }
    }
    __syncthreads();
    // X = Li B ; A = X Li^T  (symmetric) -> stored in B
    for (int e = tid; e < DD; e += nt) {
        const int i = e / D, j = e % D;
        double s = 0.0;
        for (int k = 0; k <= i; ++k) s += Li[i * CB2_CK_LD + k] * B[k * CB2_CK_LD + j];
        X[i * CB2_CK_LD + j] = s;
    }
    __syncthreads();
    for (int e = tid; e < DD; e += nt) {
        const int i = e / D, j = e % D;
        double s = 0.0;
        for (int k = 0; k <= j; ++k) s += X[i * CB2_CK_LD + k] * Li[j * CB2_CK_LD + k];
        L[i * CB2_CK_LD + j] = s;  // L is free now
    }
    __syncthreads();
    for (int e = tid; e < DD; e += nt) {
        const int i = e / D, j = e % D;
        B[i * CB2_CK_LD + j] = 0.5 * (L[i * CB2_CK_LD + j] + L[j * CB2_CK_LD + i]);
    }
    // ---- eigenvalues of B (n = D, padded to even): two-sided Jacobi, round-robin ordering
    const int n2 = (D + 1) & ~1, half = n2 / 2;
    if ((D & 1) && tid < n2) {  // padding row/column of zeros
        B[(n2 - 1) * CB2_CK_LD + tid] = 0.0;
        B[tid * CB2_CK_LD + n2 - 1] = 0.0;
    }
    __syncthreads();
    int sweeps = 0, rr = 0;
    const int blk_k = tid / half, blk_l = tid % half;
    for (int sw = 0; sw < 30; ++sw) {
        // convergence: off-diagonal norm against the diagonal
        if (tid == 0) { offnorm = 0.0; diagnorm = 0.0; }
        __syncthreads();
        {
            double so = 0.0, sd = 0.0;
            for (int e = tid; e < n2 * n2; e += nt) {
                const int i = e / n2, j = e % n2;
                const double v = B[i * CB2_CK_LD + j];
                if (i == j) sd += v * v; else so += v * v;
            }
            so = warp_sum(so); sd = warp_sum(sd);
            if ((tid & 31) == 0) { atomicAdd(&offnorm, so); atomicAdd(&diagnorm, sd); }
        }
        __syncthreads();
        if (offnorm <= 1e-26 * diagnorm || !(diagnorm > 0.0)) break;
        sweeps = sw + 1;
        for (int step = 0; step < n2 - 1; ++step, rr = (rr + 1 == n2 - 1) ? 0 : rr + 1) {
            // pair k of this step: positions k and n2-1-k of the round-robin order
            // order[0] = 0, order[i] = ((i - 1 - step) mod (n2 - 1)) + 1
            if (tid < half) {
                int p = rr_index(tid, rr, n2), q = rr_index(n2 - 1 - tid, rr, n2);
                if (p > q) { const int t_ = p; p = q; q = t_; }
                const double apq = B[p * CB2_CK_LD + q];
                double c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double tau = (B[q * CB2_CK_LD + q] - B[p * CB2_CK_LD + p]) / (2.0 * apq);
                    const double t_ = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t_ * t_);
                    s = t_ * c;
                }
                cvec[tid] = c; svec[tid] = s;
            }
            __syncthreads();
            // A <- J^T A J one 2 x 2 block (pair k rows, pair l columns) per thread: every
            // block is read and written by its own thread only
            if (tid < half * half) {  // half <= 32: one block per thread (k, l fixed)
                const int k = blk_k, l = blk_l;
                int pk = rr_index(k, rr, n2), qk = rr_index(n2 - 1 - k, rr, n2);
                if (pk > qk) { const int t_ = pk; pk = qk; qk = t_; }
                int pl = rr_index(l, rr, n2), ql = rr_index(n2 - 1 - l, rr, n2);
                if (pl > ql) { const int t_ = pl; pl = ql; ql = t_; }
                const double ck = cvec[k], sk = svec[k], cl = cvec[l], sl = svec[l];
                const double a00 = B[pk * CB2_CK_LD + pl], a01 = B[pk * CB2_CK_LD + ql];
                const double a10 = B[qk * CB2_CK_LD + pl], a11 = B[qk * CB2_CK_LD + ql];
                // columns (pair l), then rows (pair k)
                const double b00 = cl * a00 - sl * a01, b01 = sl * a00 + cl * a01;
                const double b10 = cl * a10 - sl * a11, b11 = sl * a10 + cl * a11;
                B[pk * CB2_CK_LD + pl] = ck * b00 - sk * b10;
                B[pk * CB2_CK_LD + ql] = ck * b01 - sk * b11;
                B[qk * CB2_CK_LD + pl] = sk * b00 + ck * b10;
                B[qk * CB2_CK_LD + ql] = sk * b01 + ck * b11;
            }
            __syncthreads();
        }
    }
    double lam = 0.0;
    for (int i = tid; i < D; i += nt) lam = fmax(lam, fabs(B[i * CB2_CK_LD + i]));
    for (int o = 16; o > 0; o >>= 1) lam = fmax(lam, __shfl_xor_sync(0xffffffffu, lam, o));
    if (tid == 0) { offnorm = 0.0; }
    __syncthreads();
    if ((tid & 31) == 0) {
        // max over warps through an integer-ordered atomic (values are >= 0)
        atomicMax(reinterpret_cast<unsigned long long *>(&offnorm),
                  (unsigned long long)__double_as_longlong(lam));
    }
    __syncthreads();
    const double rminus1 = offnorm;
    // ---- the new proposal transform from W (proposal.py:226-260): sorted coordinates,
    // T = S L', L' = chol(corr) with the diagonal of corr set to 1 (tools.py:779-788)
    for (int e = tid; e < DD; e += nt) {
        const int j = e / D, k = e % D;
        X[j * CB2_CK_LD + k] = W[i_of_j[j] * CB2_CK_LD + i_of_j[k]];
    }
    __syncthreads();
    for (int j = tid; j < D; j += nt) dvec[j] = sqrt(X[j * CB2_CK_LD + j]);
    __syncthreads();
    for (int e = tid; e < DD; e += nt) {
        const int j = e / D, k = e % D;
        const double inv_j = 1.0 / dvec[j], inv_k = 1.0 / dvec[k];
        L[j * CB2_CK_LD + k] = (j == k) ? 1.0 : inv_j * X[j * CB2_CK_LD + k] * inv_k;
    }
    __syncthreads();
    ck_cholesky(L, D, tid, nt, &fail2);
    for (int e = tid; e < DD; e += nt) {
        const int j = e / D, k = e % D;
        Tnew[e] = (k <= j) ? dvec[j] * L[j * CB2_CK_LD + k] : 0.0;
    }
    if (tid == 0) {
        out[0] = M; out[1] = S1; out[2] = S2 / S1;
        out[3] = fail ? CUDART_NAN : rminus1;
        out[4] = fail2 ? 0.0 : 1.0;
        out[5] = fail ? 0.0 : 1.0;
        out[6] = (double)sweeps;
        out[7] = 0.0;
    }
}

// element (row a, col j) of a lower-triangular matrix stored in B-fragment order (pack_frag)
__device__ __forceinline__ double frag_tri_read(const double *frag, int a, int j) {
    if (j > a) return 0.0;
    const int nt = a >> 3, m = j >> 3;
    const int blk = (nt * (nt + 1)) / 2 + m;
    const int lane = ((a & 7) << 2) | ((j & 7) >> 1);
    return frag[blk * 64 + lane * 2 + (j & 1)];
}

// write T (row-major, sorted coordinates) where the step kernels read it: fragment-ordered T
// inside the fast pack, fragment-ordered G = (L^-1 P) T, column-major TT.  One CTA.
__global__ void __launch_bounds__(256, 1)
k_ckpt_pack(const double *__restrict__ Tnew, int D, int NT, double *__restrict__ fastpack,
            int off_T, int off_A, int have_G, double *__restrict__ fastG,
            double *__restrict__ TT) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int DP = 8 * NT, blocks = NT * (NT + 1) / 2;
    // TT[k*D + j] = T[j][k]
    for (int e = tid; e < D * D; e += nt) {
        const int k = e / D, j = e % D;
        TT[e] = (k <= j) ? Tnew[j * D + k] : 0.0;
    }
    for (int e = tid; e < blocks * 64; e += nt) {
        const int blk = e / 64, w = e % 64, lane = w >> 1, comp = w & 1;
        // blk -> (nt_, m): nt_(nt_+1)/2 + m
        int nt_ = 0;
        while ((nt_ + 1) * (nt_ + 2) / 2 <= blk) ++nt_;
        const int m = blk - nt_ * (nt_ + 1) / 2;
        const int row = 8 * nt_ + (lane >> 2), col = 8 * m + 2 * (lane & 3) + comp;
        double t = 0.0;
        if (row < D && col < D && col <= row) t = Tnew[row * D + col];
        fastpack[off_T + e] = t;
        if (have_G) {
            // G[row][col] = sum_{j = col..row} A[row][j] T[j][col]
            double g = 0.0;
            if (row < D && col < D && col <= row)
                for (int j = col; j <= row; ++j)
                    g = fma(frag_tri_read(fastpack + off_A, row, j), Tnew[j * D + col], g);
            fastG[e] = g;
        }
    }
    (void)DP;
}
