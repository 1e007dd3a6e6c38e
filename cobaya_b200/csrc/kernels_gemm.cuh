// kernels_gemm.cuh -- k_rowgemm: the D x D x (directions) products of the streamed path for
// 128 < D <= 512 on the FP64 tensor pipe (sm_100a), hand-written (replaces the cublasDgemm
// calls of round 1).
//
//   out[d][j] = sum_k in[d][k] * W[j][k],   d < rows (millions), j < N, k < K  (N, K <= 512)
//
// used three ways per window (kernels_stream.cuh, engine.cu):
//   delta^[d][:] = T[:, j0:j0+n] u_d        (proposal.py:224; u_d = a direction of the Haar basis)
//   w^[d][:]     = (L^-1 P) delta^[d][:]    (gaussian_mixture.py:148, by linearity on the proposal)
//   y[c][:]      = (L^-1 P)(x_c - mu)       (whitening at window start)
// T and L^-1 P are LOWER TRIANGULAR in block-sorted coordinates (W[j][k] = 0 for
// k + tri_off > j): k-chunks above the diagonal of an output block are skipped, which halves
// the work of the square products -- cuBLAS computed the full ones.
//
// Tiling: CTA = 128 directions x 64 outputs, 8 warps of 32 x 32 (4 x 4 m8n8k4 tiles, the
// directions as the M dimension), k-chunks of 32 staged with cp.async into a double-buffered
// shared-memory tile (row stride 36 doubles: conflict-free 64-bit fragment loads).  Per
// k-step of 4 a warp loads 4 A and 4 B fragments for 16 DMMAs.  Bound: FP64 tensor pipe
// (2 rows N K flop, halved by the triangle) above ~64 flop/B; the operands stream once.
#pragma once
#include "kernels_fast.cuh"

#define CB2_GEMM_BM 128
#define CB2_GEMM_BN 64
#define CB2_GEMM_BK 32
#define CB2_GEMM_LD 36

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory");
}

// stage rows [r0, r0+NR) x cols [k0, k0+32) of a row-major matrix into tile[NR][LD];
// out-of-range elements are zero.  `vec`: 16-byte copies are legal (ld even, base aligned).
template <int NR>
__device__ __forceinline__ void gemm_stage(double *tile, const double *__restrict__ src,
                                           int64_t ld, int64_t r0, int64_t n_rows, int k0, int K,
                                           bool vec, int tid) {
    constexpr int CH = CB2_GEMM_BK / 2;  // 16-byte chunks per row
    for (int e = tid; e < NR * CH; e += 256) {
        const int row = e / CH, c = (e % CH) * 2;
        double *dst = tile + row * CB2_GEMM_LD + c;
        const int64_t gr = r0 + row;
        const int k = k0 + c;
        if (gr < n_rows && k + 1 < K && vec) {
            cp_async16((uint32_t)__cvta_generic_to_shared(dst), src + gr * ld + k);
        } else {
            dst[0] = (gr < n_rows && k < K) ? src[gr * ld + k] : 0.0;
            dst[1] = (gr < n_rows && k + 1 < K) ? src[gr * ld + k + 1] : 0.0;
        }
    }
}

__global__ void __launch_bounds__(256, 2)
k_rowgemm(const double *__restrict__ in, int64_t ldin, const double *__restrict__ W, int64_t ldw,
          double *__restrict__ out, int64_t ldout, int64_t rows, int N, int K, int tri,
          int tri_off) {
    extern __shared__ __align__(16) double gsm[];
    double *As = gsm;                                            // [2][BM][LD]
    double *Ws = gsm + 2 * CB2_GEMM_BM * CB2_GEMM_LD;            // [2][BN][LD]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = lane >> 2, r = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;  // warp tile: rows 32 wm, cols 32 wn
    // consecutive CTAs take the output blocks of the SAME 128 directions: the direction tile is
    // read from HBM once and from L2 by the others
    const int n_nblk = (N + CB2_GEMM_BN - 1) / CB2_GEMM_BN;
    const int64_t r0 = (int64_t)(blockIdx.x / n_nblk) * CB2_GEMM_BM;
    const int j0 = (int)(blockIdx.x % n_nblk) * CB2_GEMM_BN;
    // k range that can touch this output block: k + tri_off <= j0 + BN - 1
    int k_end = K;
    if (tri) k_end = min(K, j0 + CB2_GEMM_BN - tri_off);
    const int n_chunks = k_end > 0 ? (k_end + CB2_GEMM_BK - 1) / CB2_GEMM_BK : 0;
    const bool vec_a = ((ldin & 1) == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0);
    const bool vec_w = ((ldw & 1) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    if (n_chunks > 0) {
        gemm_stage<CB2_GEMM_BM>(As, in, ldin, r0, rows, 0, K, vec_a, tid);
        gemm_stage<CB2_GEMM_BN>(Ws, W, ldw, j0, N, 0, K, vec_w, tid);
        cp_async_commit();
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int buf = c & 1;
        if (c + 1 < n_chunks) {  // next chunk into the other buffer
            gemm_stage<CB2_GEMM_BM>(As + (buf ^ 1) * CB2_GEMM_BM * CB2_GEMM_LD, in, ldin, r0, rows,
                                    (c + 1) * CB2_GEMM_BK, K, vec_a, tid);
            gemm_stage<CB2_GEMM_BN>(Ws + (buf ^ 1) * CB2_GEMM_BN * CB2_GEMM_LD, W, ldw, j0, N,
                                    (c + 1) * CB2_GEMM_BK, K, vec_w, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const double *At = As + buf * CB2_GEMM_BM * CB2_GEMM_LD + (32 * wm + q) * CB2_GEMM_LD + r;
        const double *Wt = Ws + buf * CB2_GEMM_BN * CB2_GEMM_LD + (32 * wn + q) * CB2_GEMM_LD + r;
#pragma unroll
        for (int ks = 0; ks < CB2_GEMM_BK; ks += 4) {
            double a[4], b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[t] = At[8 * t * CB2_GEMM_LD + ks];
                b[t] = Wt[8 * t * CB2_GEMM_LD + ks];
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma8x8x4(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
        __syncthreads();
    }
    // C fragment: lane (q, r) holds out[row q][cols 2r, 2r+1] of every 8 x 8 tile
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
        const int64_t row = r0 + 32 * wm + 8 * mt + q;
        if (row >= rows) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int col = j0 + 32 * wn + 8 * nt + 2 * r;
            double *o = out + row * ldout + col;
            if (col + 1 < N && (ldout & 1) == 0)
                *reinterpret_cast<double2 *>(o) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            else {
                if (col < N) o[0] = acc[mt][nt][0];
                if (col + 1 < N) o[1] = acc[mt][nt][1];
            }
        }
    }
}

static inline int launch_rowgemm(cudaStream_t st, const double *in, int64_t ldin, const double *W,
                                 int64_t ldw, double *out, int64_t ldout, int64_t rows, int N,
                                 int K, bool tri, int tri_off) {
    if (rows <= 0) return 0;
    const size_t smem = (size_t)2 * (CB2_GEMM_BM + CB2_GEMM_BN) * CB2_GEMM_LD * sizeof(double);
    if (cudaFuncSetAttribute(k_rowgemm, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return -1;
    const int64_t gx = ((rows + CB2_GEMM_BM - 1) / CB2_GEMM_BM) * ((N + CB2_GEMM_BN - 1) / CB2_GEMM_BN);
    if (gx > 0x7fffffff) return -1;
    k_rowgemm<<<(unsigned)gx, 256, smem, st>>>(in, ldin, W, ldw, out, ldout, rows, N, K, tri ? 1 : 0,
                                       tri_off);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
