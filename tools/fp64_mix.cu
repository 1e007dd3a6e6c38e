// Microbenchmark: what a vector FP64 instruction costs the FP64 tensor pipe (DMMA m8n8k4) of
// its SM sub-partition when both are in flight -- the question behind the consumer design of
// k_step_pc2 (cobaya_b200/csrc/kernels_pc2.cuh).
//
// One CTA per SM; per sub-partition two warps issue DMMAs back to back (fixed count) and one
// "consumer" warp issues `bursts` bursts of `blen` independent DFMAs separated by `gap`
// dependent integer instructions.  Reported: kernel time, and the extra pipe cycles per DFMA
// warp-instruction relative to the DMMA-only run.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o fp64_mix fp64_mix.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int BLEN>
__global__ void __launch_bounds__(384, 1)
k_mix(double *out, int mma_iters, int bursts, int gap, double a0, double b0) {
    const int warp = threadIdx.x >> 5;
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    const double a = a0 + threadIdx.x * 1e-12, b = b0;
    double s = 0;
    if (warp < 8) {  // two DMMA warps per sub-partition
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) mma884(acc[i][0], acc[i][1], a, b);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s += acc[i][0] + acc[i][1];
    } else {         // one vector-FP64 warp per sub-partition
        double f[BLEN];
#pragma unroll
        for (int i = 0; i < BLEN; ++i) f[i] = threadIdx.x * 1e-9 + i;
        unsigned x = threadIdx.x;
        for (int bu = 0; bu < bursts; ++bu) {
#pragma unroll
            for (int i = 0; i < BLEN; ++i) f[i] = fma(f[i], a, b);
            for (int g = 0; g < gap; ++g) x = x * 1664525u + 1013904223u;  // dependent integer work
        }
#pragma unroll
        for (int i = 0; i < BLEN; ++i) s += f[i];
        s += (double)x;
    }
    if (s == 123.456) out[0] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double *out; cudaMalloc(&out, 64);
    const int mma_iters = 16384;  // x4 DMMAs per warp, 2 warps per sub-partition
    const double mma_cycles = 2.0 * 4 * mma_iters * 16.0;
    const double ghz = 1.965;
    printf("{\"gpu\": \"%s\", \"dmma_only_ideal_ms\": %.4f,\n", p.name, mma_cycles / ghz * 1e-6);
    float base = time_ms([&] { k_mix<16><<<sms, 384>>>(out, mma_iters, 0, 0, 1.0000001, 1e-9); });
    printf(" \"dmma_only_ms\": %.4f,\n", base);
#define RUN(BLEN, BURSTS, GAP)                                                              \
    {                                                                                       \
        float ms = time_ms([&] { k_mix<BLEN><<<sms, 384>>>(out, mma_iters, BURSTS, GAP, 1.0000001, 1e-9); }); \
        double extra = (ms - base) * 1e6 * ghz / ((double)BLEN * BURSTS);                    \
        printf(" \"burst%d_x%d_gap%d\": {\"ms\": %.4f, \"extra_pipe_cycles_per_dfma\": %.2f},\n", \
               BLEN, BURSTS, GAP, ms, extra);                                               \
    }
    RUN(1, 16384, 64)
    RUN(1, 16384, 16)
    RUN(4, 8192, 64)
    RUN(16, 4096, 64)
    RUN(16, 4096, 256)
    RUN(48, 1024, 256)
    RUN(48, 2048, 64)
    RUN(48, 512, 1024)
    cudaError_t e = cudaDeviceSynchronize();
    printf(" \"status\": \"%s\"}\n", cudaGetErrorString(e));
    return 0;
}
