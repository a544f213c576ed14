// Shared helpers for the sm_100a kernels behind include/d2b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/d2b200.h"

#define D2B_API extern "C" __attribute__((visibility("default")))

#define D2B_CHECK_LAUNCH()                               \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

#define D2B_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t e__ = (expr);                            \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

__host__ __device__ static inline int d2b_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;  // B200
