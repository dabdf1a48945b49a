// common.cuh -- shared helpers for libpbsed_b200 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pbsed_b200.h"

#ifndef __CUDA_ARCH__
#define PBSED_HOST_ONLY 1
#endif

extern long long g_pbsed_launches;   // defined in api.cu
extern const char* g_pbsed_last_kernel;   // name of the main kernel the last entry point dispatched (bench.py)
static inline void pbsed_note_kernel(const char* name) { g_pbsed_last_kernel = name; }

static inline int pbsed_after_launch() {
  ++g_pbsed_launches;
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
