// common.cuh -- shared helpers for libpbsed_b200 (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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

// ---- activation storage type (precision 'bf16': conv-stack maps and their gradients live in HBM as bf16) ----
// dtype codes of the C ABI: 0 = fp32, 1 = bf16.  `idx` counts ELEMENTS from the base pointer and is a
// multiple of 4; a bf16 quad is one 8-byte access.
#define PBSED_F32 0
#define PBSED_BF16 1
__device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 float4_to_bf16x4(float4 v) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
  return u;
}
__device__ __forceinline__ float4 ld_act4(const void* base, long long idx, int bf16) {
  if (bf16) return bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx)));
  return __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
}
__device__ __forceinline__ void st_act4(void* base, long long idx, float4 v, int bf16) {
  if (bf16) *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = float4_to_bf16x4(v);
  else *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = v;
}
__device__ __forceinline__ float ld_act1(const void* base, long long idx, int bf16) {
  if (bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  return __ldg(reinterpret_cast<const float*>(base) + idx);
}
