// kernels_rows.cuh -- moving stored sample rows in bulk (sm_100a).
//
// The reference appends every accepted point to a host-side SampleCollection
// (cobaya/collection.py:402-427,519-571).  Here rows are written by the step kernels into
// rows[chain][cap][W] in HBM; these kernels hand them to the host without one copy per
// chain: the rows [first_c, n_c) of every chain are compacted chain-major into one
// contiguous staging buffer (a pure HBM copy: 2 x bytes moved), which then leaves the
// device with ONE asynchronous D2H on a copy stream.  k_rows_regrow re-lays the store out
// for a larger per-chain capacity.
#pragma once
#include "common.cuh"

// counts[c] = max(0, n_rows[c] - first[c]) (first == NULL: 0), offsets = exclusive prefix
// sum, offsets[C] = total.  One CTA of 1024 threads; serial chunk per thread + block scan.
// If `advance_cursor` the cursor first[c] is moved to n_rows[c] afterwards (drain).
__global__ void __launch_bounds__(1024, 1)
k_rows_offsets(const int64_t *__restrict__ n_rows, int64_t *__restrict__ first,
               int64_t n_chains, int64_t *__restrict__ counts, int64_t *__restrict__ offsets,
               int64_t *__restrict__ begin, int advance_cursor) {
    __shared__ long long part[1024];
    const int tid = threadIdx.x;
    const int64_t per = (n_chains + 1023) / 1024;
    const int64_t c0 = tid * per, c1 = min(c0 + per, n_chains);
    long long s = 0;
    for (int64_t c = c0; c < c1; ++c) {
        const long long f = first ? first[c] : 0;
        const long long k = n_rows[c] - f;
        s += k > 0 ? k : 0;
    }
    part[tid] = s;
    __syncthreads();
    // inclusive Hillis-Steele scan over the 1024 partial sums
    for (int o = 1; o < 1024; o <<= 1) {
        long long v = (tid >= o) ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - s;
    for (int64_t c = c0; c < c1; ++c) {
        const long long f = first ? first[c] : 0;
        long long k = n_rows[c] - f;
        k = k > 0 ? k : 0;
        counts[c] = k;
        offsets[c] = run;
        begin[c] = f;
        run += k;
        if (advance_cursor && first) first[c] = f + k;
    }
    if (tid == 1023) offsets[n_chains] = part[1023];
}

// dst[offsets[c] + i][:] = rows[c][begin[c] + i][:], i < counts[c].  grid = (chains, slices):
// the rows of a chain are one contiguous run of counts[c]*W doubles on both sides.
__global__ void __launch_bounds__(256)
k_rows_gather(const double *__restrict__ rows, int64_t cap, int W,
              const int64_t *__restrict__ begin, const int64_t *__restrict__ counts,
              const int64_t *__restrict__ offsets, int64_t max_rows,
              double *__restrict__ dst) {
    const int64_t c = blockIdx.x;
    const int64_t k = counts[c], off = offsets[c];
    if (k <= 0 || off + k > max_rows) return;  // caller checks the total before trusting dst
    const double *src = rows + ((size_t)c * cap + begin[c]) * W;
    double *out = dst + (size_t)off * W;
    const int64_t n = k * W;
    const int64_t per = (((n + gridDim.y - 1) / gridDim.y) + 1) & ~(int64_t)1;  // even slices
    const int64_t e0 = min(blockIdx.y * per, n), e1 = min(e0 + per, n);
    // 16-byte path when both runs start on an even double
    const bool al = ((((size_t)c * cap + begin[c]) * W) % 2 == 0) && (((size_t)off * W) % 2 == 0);
    if (al) {
        const double2 *s2 = reinterpret_cast<const double2 *>(src + e0);
        double2 *d2 = reinterpret_cast<double2 *>(out + e0);
        const int64_t n2 = (e1 - e0) / 2;
        for (int64_t e = threadIdx.x; e < n2; e += blockDim.x) d2[e] = __ldcs(s2 + e);
        if (((e1 - e0) & 1) && threadIdx.x == 0) out[e1 - 1] = src[e1 - 1];
    } else {
        for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) out[e] = src[e];
    }
}

// new_rows[c][i][:] = old_rows[c][i][:], i < n_rows[c]: same chains, larger capacity
__global__ void __launch_bounds__(256)
k_rows_regrow(const double *__restrict__ old_rows, int64_t old_cap, double *__restrict__ new_rows,
              int64_t new_cap, int W, const int64_t *__restrict__ n_rows) {
    const int64_t c = blockIdx.x;
    const int64_t n = min(n_rows[c], old_cap) * (int64_t)W;
    const double *src = old_rows + (size_t)c * old_cap * W;
    double *dst = new_rows + (size_t)c * new_cap * W;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t e0 = blockIdx.y * per, e1 = min(e0 + per, n);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) dst[e] = src[e];
}

// inverse of k_rows_gather for restoring a snapshot: rows[c][i][:] = src[offsets[c] + i][:]
__global__ void __launch_bounds__(256)
k_rows_scatter(double *__restrict__ rows, int64_t cap, int W, const int64_t *__restrict__ counts,
               const int64_t *__restrict__ offsets, const double *__restrict__ src) {
    const int64_t c = blockIdx.x;
    const int64_t n = counts[c] * (int64_t)W;
    const double *in = src + (size_t)offsets[c] * W;
    double *out = rows + (size_t)c * cap * W;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t e0 = blockIdx.y * per, e1 = min(e0 + per, n);
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) out[e] = in[e];
}
