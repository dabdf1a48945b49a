// wgrad_mma.cu -- weight gradient of the narrow 3x3 conv layers (16 / 32 input channels) on the warp-level
// tensor-core path (mma.sync m16n8k8 TF32, 3-pass hi/lo split = fp32-equivalent).  FALLBACK of the row-stacked
// tcgen05 tiles (tapgemm_tc.cu, tapgemm_wgrad_stack_dispatch): it runs when those do not apply -- strided maps -- or
// are switched off (PBSED_WG_STACK=0).
//
// Reference: backward of padertorch CNN2d layers 1-4 (pb_sed/experiments/weak_label_crnn/training.py:
// 158-169: channels 16,16,32,32,64).  dW[tap][n][c] = sum_t dout[t][n] * a[t + dt][c] is a GEMM with
// M = Cout (16..64), N = Cin (16/32), K = frames.  Each warp owns one 16 (n) x 16 (c) plane for all nine taps
// (18 accumulator tiles = 72 registers) and, when the layer has fewer planes than warps, a share of the K steps;
// partial tiles meet in shared memory before ONE coalesced set of global atomics per CTA.  Measured on B200: the
// warp-level TF32 MMA runs at ~300 TFLOP/s, so three passes bound this kernel near 100 algorithmic TFLOP/s.
#include "common.cuh"
#include <cstdlib>

namespace {

struct WmParams {
  int B, F, T, relu, mask_out, single;
  int in_stride, out_stride;
  int in_bf16, out_bf16;      // storage type of `in` / `dout` (bf16 activation maps)
  long long w_tap_stride, w_sn, w_sc;
};

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // both pieces rounded to nearest TF32 (the tensor core would truncate the low 13 bits): |x - hi - lo| <= 2^-22 |x|
  hi = (__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x00001000u) & 0xFFFFE000u;
}

// ---------------------------------------------------------------------------------------------------
// Frequency walking: a persistent CTA owns a contiguous run of (clip, frame tile,
// frequency row) steps with the row index fastest, so that walking down the frequency axis every input
// strip and every dout row is fetched from global memory ONCE (the kernel above fetches each strip for
// three rows and waits for it synchronously).  The fetches are cp.async copies issued one step ahead
// (input strips in a ring of four, dout rows double buffered; out-of-range frames / rows are zero-filled
// by the copy itself), the owner thread of each 16-byte quad applies norm + ReLU in place once the copy
// has landed AND splits it into its (hi, lo) TF32 pair -- once per element instead of once per fragment
// use, so the inner loop is 64-bit shared loads and MMAs only -- and there is one __syncthreads per step.
// bf16 maps land in a small raw staging area and are widened by the same owner pass.  The MMAs of one
// fragment group are issued pass-major (three independent accumulators between two dependent instructions).
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst, const void* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = ok ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void mma_tf32_nv(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int COUT, int CIN, int TW>
__global__ void __launch_bounds__(256, 2)
wgrad_walk_kernel(WmParams p, const float* __restrict__ in, const float* __restrict__ scale,
                  const float* __restrict__ shift, const int* __restrict__ seq_len,
                  const float* __restrict__ dout, float* __restrict__ dW, float* __restrict__ dbias) {
  // rows of (hi, lo) TF32 pairs; a pitch == 8 (mod 32) words keeps the 64-bit fragment loads conflict free
  constexpr int PZ = 2 * (COUT + 4), PA = 2 * (CIN + 4);
  constexpr int TS = (COUT / 16) * (CIN / 16);
  constexpr int KS = 8 / TS;
  static_assert(TS >= 1 && TS <= 8 && KS * TS == 8 && TW % 8 == 0, "warp mapping");
  constexpr int ZQ = COUT / 4, AQ = CIN / 4;
  constexpr int ZN = TW * ZQ, AN = (TW + 2) * AQ;      // quads of one dout row tile / one input strip
  static_assert(32 % ZQ == 0 && 32 % AQ == 0, "a row's quads belong to one warp (in-place widening)");
  extern __shared__ __align__(16) float sm[];
  float* zs = sm;                                    // [2][TW][PZ]
  float* as = sm + 2 * TW * PZ;                      // [4][TW + 2][PA]
  uint2* zr = reinterpret_cast<uint2*>(as + 4 * (TW + 2) * PA);   // bf16 maps: raw quads [ZN], [3][AN]
  uint2* ar = zr + ZN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int ts = warp % TS, ksel = warp / TS;
  const int n0 = (ts / (CIN / 16)) * 16, c0 = (ts % (CIN / 16)) * 16;
  const bool xform = scale != nullptr || p.relu;
  float4 sc4 = make_float4(1.f, 1.f, 1.f, 1.f), sh4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (scale) {                                       // tid % AQ is the same for every quad this thread owns
    sc4 = __ldg(reinterpret_cast<const float4*>(scale + (tid % AQ) * 4));
    sh4 = __ldg(reinterpret_cast<const float4*>(shift + (tid % AQ) * 4));
  }

  float acc[9][2][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[k][j][e] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);

  const int t_tiles = (p.T + TW - 1) / TW;
  const long long total = (long long)p.B * t_tiles * p.F;
  const int s0 = (int)(total * blockIdx.x / gridDim.x), s1 = (int)(total * (blockIdx.x + 1) / gridDim.x);
  int b = 0, t0 = 0, len_b = 0, len_out = 0;

  // fp32 maps: the raw quad q of a row lands in words [4q, 4q + 4) of that row's (wider) image
  auto issue_x = [&](int fs, int stage) {            // strip fs -> ring slot fs & 3 (raw stage `stage` for bf16)
    const bool f_ok = fs >= 0 && fs < p.F;
    const long long a0 = ((long long)b * p.F + (f_ok ? fs : 0)) * p.T * p.in_stride;
    float* dst = as + ((fs + 4) & 3) * (TW + 2) * PA;
    for (int i = tid; i < AN; i += 256) {
      const int r = i / AQ, q = i % AQ;
      const int t = t0 + r - 1;
      const bool ok = f_ok && t >= 0 && t < len_b;
      const long long off = ok ? a0 + (long long)t * p.in_stride + q * 4 : 0;
      if (p.in_bf16) cp_async_8(ar + stage * AN + i, reinterpret_cast<const __nv_bfloat16*>(in) + off, ok);
      else cp_async_16(dst + r * PA + q * 4, in + off, ok);
    }
  };
  auto issue_z = [&](int f) {
    const long long z0 = ((long long)b * p.F + f) * p.T * p.out_stride;
    float* dst = zs + (f & 1) * TW * PZ;
    for (int i = tid; i < ZN; i += 256) {
      const int r = i / ZQ, q = i % ZQ;
      const int t = t0 + r;
      const bool ok = t < len_out;
      const long long off = ok ? z0 + (long long)t * p.out_stride + q * 4 : 0;
      if (p.out_bf16) cp_async_8(zr + i, reinterpret_cast<const __nv_bfloat16*>(dout) + off, ok);
      else cp_async_16(dst + r * PZ + q * 4, dout + off, ok);
    }
  };
  auto widen = [&](float* cell8, float4 v) {         // four values -> four (hi, lo) pairs
    uint4 w0, w1;
    split_tf32(v.x, w0.x, w0.y); split_tf32(v.y, w0.z, w0.w);
    split_tf32(v.z, w1.x, w1.y); split_tf32(v.w, w1.z, w1.w);
    *reinterpret_cast<uint4*>(cell8) = w0;
    *reinterpret_cast<uint4*>(cell8 + 4) = w1;
  };
  // owner pass: the thread that copied a quad widens it in place once its copy has landed.  A row's raw
  // quads and its image overlap, but they all belong to lanes of ONE warp: read, __syncwarp, write.
  auto own_x = [&](int fs, int stage) {
    const bool f_ok = fs >= 0 && fs < p.F;
    float* dst = as + ((fs + 4) & 3) * (TW + 2) * PA;
    for (int i0 = 0; i0 < AN; i0 += 256) {
      const int i = i0 + tid;
      const bool mine = i < AN;
      const int r = i / AQ, q = i % AQ;
      const int t = t0 + r - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mine) v = p.in_bf16 ? bf16x4_to_float4(ar[stage * AN + i]) : *reinterpret_cast<const float4*>(dst + r * PA + q * 4);
      if (xform && f_ok && t >= 0 && t < len_b) {
        v.x = fmaf(v.x, sc4.x, sh4.x); v.y = fmaf(v.y, sc4.y, sh4.y);
        v.z = fmaf(v.z, sc4.z, sh4.z); v.w = fmaf(v.w, sc4.w, sh4.w);
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      __syncwarp();
      if (mine) widen(dst + r * PA + q * 8, v);
    }
  };
  auto own_z = [&](int f) {
    float* dst = zs + (f & 1) * TW * PZ;
    for (int i0 = 0; i0 < ZN; i0 += 256) {
      const int i = i0 + tid;
      const bool mine = i < ZN;
      const int r = i / ZQ, q = i % ZQ;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mine) v = p.out_bf16 ? bf16x4_to_float4(zr[i]) : *reinterpret_cast<const float4*>(dst + r * PZ + q * 4);
      bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
      __syncwarp();
      if (mine) widen(dst + r * PZ + q * 8, v);
    }
  };

  bool fresh = true;
  for (int s = s0; s < s1; ++s) {
    const int f = s % p.F;
    if (f == 0) fresh = true;
    if (fresh) {
      const int col = s / p.F;
      b = col / t_tiles; t0 = (col % t_tiles) * TW;
      len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
      len_out = p.mask_out ? len_b : p.T;
      __syncthreads();                               // the previous step's readers are done with the rings
      issue_x(f - 1, 0); issue_x(f, 1); issue_x(f + 1, 2); issue_z(f);
      cp_async_commit();
    }
    cp_async_wait_all();
    if (fresh) { own_x(f - 1, 0); own_x(f, 1); }
    own_x(f + 1, 2);
    own_z(f);
    __syncthreads();
    if (s + 1 < s1 && f + 1 < p.F) { issue_x(f + 2, 2); issue_z(f + 1); cp_async_commit(); }
    fresh = false;
    const float* zt = zs + (f & 1) * TW * PZ + 2 * (n0 + g);
#pragma unroll 1
    for (int ks = ksel; ks < TW / 8; ks += KS) {
      const int r0 = ks * 8;
      uint32_t ah[4], al[4];
      {
        const uint2 z0 = *reinterpret_cast<const uint2*>(zt + (r0 + tig) * PZ);
        const uint2 z1 = *reinterpret_cast<const uint2*>(zt + (r0 + tig) * PZ + 16);
        const uint2 z2 = *reinterpret_cast<const uint2*>(zt + (r0 + tig + 4) * PZ);
        const uint2 z3 = *reinterpret_cast<const uint2*>(zt + (r0 + tig + 4) * PZ + 16);
        ah[0] = z0.x; al[0] = z0.y; ah[1] = z1.x; al[1] = z1.y;
        ah[2] = z2.x; al[2] = z2.y; ah[3] = z3.x; al[3] = z3.y;
      }
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float* strip = as + ((f + d + 3) & 3) * (TW + 2) * PA + (r0 + tig) * PA + 2 * (c0 + g);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float* col = strip + 16 * j;
          uint2 x[6];
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            x[e] = *reinterpret_cast<const uint2*>(col + e * PA);
            x[3 + e] = *reinterpret_cast<const uint2*>(col + (e + 4) * PA);
          }
#pragma unroll
          for (int e = 0; e < 3; ++e) mma_tf32_nv(acc[d * 3 + e][j], ah, x[e].x, x[3 + e].x);
          if (!p.single) {
#pragma unroll
            for (int e = 0; e < 3; ++e) mma_tf32_nv(acc[d * 3 + e][j], al, x[e].x, x[3 + e].x);
#pragma unroll
            for (int e = 0; e < 3; ++e) mma_tf32_nv(acc[d * 3 + e][j], ah, x[e].y, x[3 + e].y);
          }
        }
      }
    }
  }
  // ---- reduce the KS partial planes in shared memory, then one coalesced set of global atomics per tap
  float* red = sm;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    __syncthreads();
    for (int i = tid; i < COUT * CIN; i += 256) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = c0 + 8 * j + 2 * tig;
      if (KS == 1) {
        red[(n0 + g) * CIN + c] = acc[k][j][0]; red[(n0 + g) * CIN + c + 1] = acc[k][j][1];
        red[(n0 + g + 8) * CIN + c] = acc[k][j][2]; red[(n0 + g + 8) * CIN + c + 1] = acc[k][j][3];
      } else {
        atomicAdd(&red[(n0 + g) * CIN + c], acc[k][j][0]); atomicAdd(&red[(n0 + g) * CIN + c + 1], acc[k][j][1]);
        atomicAdd(&red[(n0 + g + 8) * CIN + c], acc[k][j][2]); atomicAdd(&red[(n0 + g + 8) * CIN + c + 1], acc[k][j][3]);
      }
    }
    __syncthreads();
    for (int i = tid; i < COUT * CIN; i += 256) {
      const float v = red[i];
      const int n = i / CIN, c = i % CIN;
      if (v != 0.f) atomicAdd(dW + (long long)k * p.w_tap_stride + (long long)n * p.w_sn + (long long)c * p.w_sc, v);
    }
  }
  if (dbias) {
    float* db = dbias + (tid % ZQ) * 4;
    if (bsum.x != 0.f) atomicAdd(db + 0, bsum.x);
    if (bsum.y != 0.f) atomicAdd(db + 1, bsum.y);
    if (bsum.z != 0.f) atomicAdd(db + 2, bsum.z);
    if (bsum.w != 0.f) atomicAdd(db + 3, bsum.w);
  }
}

template <int COUT, int CIN, int TW>
int launch_walk(const WmParams& p, const float* in, const float* scale, const float* shift, const int* seq_len,
                const float* dout, float* dW, float* dbias, cudaStream_t st) {
  size_t smem = (2 * (size_t)TW * 2 * (COUT + 4) + 4 * (size_t)(TW + 2) * 2 * (CIN + 4)) * sizeof(float);
  if (p.in_bf16 || p.out_bf16) smem += ((size_t)TW * (COUT / 4) + 3 * (size_t)(TW + 2) * (CIN / 4)) * sizeof(uint2);
  cudaError_t e = cudaFuncSetAttribute(wgrad_walk_kernel<COUT, CIN, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const long long total = (long long)p.B * p.F * cdiv(p.T, TW);
  if (total > 0x7fffffffLL) return PBSED_EINVAL;
  int grid = 148 * 2;
  if (grid > total) grid = (int)total;
  pbsed_note_kernel("wgrad_walk_kernel");
  wgrad_walk_kernel<COUT, CIN, TW><<<grid, 256, smem, st>>>(p, in, scale, shift, seq_len, dout, dW, dbias);
  return pbsed_after_launch();
}

}  // namespace

int wgrad_mma_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale, const float* shift,
                       const int* seq_len, const float* dout, int mask_out, float* dW, float* dbias,
                       cudaStream_t st, int* handled) {
  *handled = 0;
  if (d->precision == 0) return 0;                     // exact-fp32 mode keeps the FFMA kernel
  if (d->ntaps != 9 || d->F_in != d->F_out || d->per_f) return 0;
  for (int i = 0; i < 9; ++i)
    if (d->df[i] != i / 3 - 1 || d->dt[i] != i % 3 - 1) return 0;
  WmParams p;
  p.B = d->B; p.F = d->F_in; p.T = d->T; p.relu = d->relu; p.mask_out = mask_out;
  p.single = d->precision == 3;
  p.in_stride = d->in_stride > 0 ? d->in_stride : d->Cin;
  p.in_bf16 = d->in_dtype == PBSED_BF16; p.out_bf16 = d->out_dtype == PBSED_BF16;
  p.out_stride = d->out_stride > 0 ? d->out_stride : d->Cout;
  p.w_tap_stride = d->w_tap_stride; p.w_sn = d->w_sn; p.w_sc = d->w_sc;
  if (p.in_stride % 4 || p.out_stride % 4) return 0;
  if ((((uintptr_t)in | (uintptr_t)dout | (uintptr_t)scale | (uintptr_t)shift) & 15) != 0) return 0;
  int rc;
  if (d->Cout == 16 && d->Cin == 16)      rc = launch_walk<16, 16, 64>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 16) rc = launch_walk<32, 16, 64>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 32) rc = launch_walk<32, 32, 48>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 64 && d->Cin == 32) rc = launch_walk<64, 32, 40>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else return 0;
  *handled = 1;
  return rc;
}
