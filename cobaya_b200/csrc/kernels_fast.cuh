// kernels_fast.cuh -- register-resident fast paths (filled in below the general path).
#pragma once
#include "common.cuh"
#include "kernels_general.cuh"

static inline bool fast_step_supported(const ModelDev &, size_t) { return false; }
static inline int launch_step_fast(cudaStream_t, const ModelDev &, const ChainState &,
                                   const WindowDev &, const double *, int64_t, uint64_t, int,
                                   int) { return -1; }
static inline bool fast_basis_supported(int) { return false; }
static inline int launch_basis_fast(cudaStream_t, uint32_t, uint32_t, uint64_t, int, int,
                                    const int64_t *, int, int, double *, int64_t) { return -1; }
static inline int launch_basis_fast_one(cudaStream_t, uint32_t, uint32_t, uint64_t, int, int,
                                        uint32_t, double *) { return -1; }
