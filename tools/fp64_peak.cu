// Microbenchmark: what the FP64 datapath of this GPU actually delivers, as the denominator of
// the FP64 roofline in bench.py.  Measures, with CUDA events,
//   * mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64 with ILP independent accumulators,
//   * DFMA with ILP independent accumulators,
//   * a mix (even warps DMMA, odd warps DFMA),
// for several resident-warp counts per SM, and the dependent-issue latency of one m8n8k4.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&d)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_mma884(double *out, int iters, double a0, double b0) {
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    const double a = a0 + threadIdx.x * 1e-12, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_mma1684(double *out, int iters, double a0, double b0) {
    double acc[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-9 + j;
    double a[2] = {a0 + threadIdx.x * 1e-12, a0};
    const double b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma1684(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_mma1688(double *out, int iters, double a0, double b0) {
    double acc[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-9 + j;
    double a[4] = {a0 + threadIdx.x * 1e-12, a0, a0 * 0.5, a0 * 0.25};
    double b[2] = {b0, b0 * 0.5};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma1688(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_mma16816(double *out, int iters, double a0, double b0) {
    double acc[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = threadIdx.x * 1e-9 + j;
    double a[8], b[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = a0 + threadIdx.x * 1e-12 + j * 1e-3;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = b0 + j * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) mma16816(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a0, double b0) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    const double a = a0, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// even warps: m8n8k4 DMMA, odd warps: DFMA (same instruction count per warp)
template <int ILP>
__global__ void k_mix(double *out, int iters, double a0, double b0) {
    const int warp = threadIdx.x >> 5;
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    const double a = a0 + threadIdx.x * 1e-12, b = b0;
    if (warp & 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) acc[i][0] = fma(acc[i][0], a, b);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) mma884(acc[i][0], acc[i][1], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

// dependent-issue latency: one warp per SM sub-partition, one accumulator
__global__ void k_lat884(double *out, long long *cyc, int iters, double a0, double b0) {
    double d0 = threadIdx.x * 1e-9, d1 = 0.5;
    const double a = a0, b = b0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) mma884(d0, d1, a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (d0 + d1 == 123.456) out[0] = d0;
}
// C -> A dependency (the step kernel's pattern: an MMA's C fragment is the next one's A)
__global__ void k_lat884_ca(double *out, long long *cyc, int iters, double b0) {
    double d0 = threadIdx.x * 1e-9, d1 = 0.5;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double e0 = 0.0, e1 = 0.0;
        mma884(e0, e1, d0, b0);
        d0 = e0; d1 = e1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    if (d0 + d1 == 123.456) out[0] = d0;
}

template <typename F>
static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();  // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        launch();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double *out; cudaMalloc(&out, 64);
    long long *cyc; cudaMalloc(&cyc, 64);
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, clk_khz);
    const int wps_list[4] = {4, 8, 16, 32};
#define RUN(NAME, KERNEL, ILP, FMA_PER_INST)                                              \
    for (int wi = 0; wi < 4; ++wi) {                                                      \
        const int wps = wps_list[wi];                                                     \
        float ms = time_ms([&] { KERNEL<ILP><<<sms, wps * 32>>>(out, iters, 1.0000001, 1e-9); }); \
        double fl = 2.0 * FMA_PER_INST * (double)ILP * iters * wps * sms;                 \
        printf(" \"%s_ilp%d_w%d\": {\"ms\": %.4f, \"tflops\": %.2f},\n", NAME, ILP, wps, ms, \
               fl / ms * 1e-9);                                                           \
    }
    RUN("mma_m8n8k4", k_mma884, 8, 256.0)
    RUN("mma_m8n8k4", k_mma884, 2, 256.0)
    RUN("mma_m16n8k4", k_mma1684, 4, 512.0)
    RUN("mma_m16n8k8", k_mma1688, 4, 1024.0)
    RUN("mma_m16n8k16", k_mma16816, 4, 2048.0)
    RUN("dfma", k_dfma, 8, 32.0)
    RUN("mix_dmma_dfma(flops counted as if all DMMA)", k_mix, 8, 256.0)
    {
        k_lat884<<<1, 32>>>(out, cyc, 4096, 1.0000001, 1e-9);
        long long c = 0; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" \"lat_m8n8k4_C_dep_cycles\": %.2f,\n", (double)c / 4096);
        k_lat884_ca<<<1, 32>>>(out, cyc, 4096, 1e-9);
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf(" \"lat_m8n8k4_C_to_A_cycles\": %.2f,\n", (double)c / 4096);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf(" \"status\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
