// wgrad_mma.cu -- weight gradient of the narrow 3x3 conv layers (16 / 32 input channels) on the
// warp-level tensor-core path (mma.sync m16n8k8 TF32, 3-pass hi/lo split = fp32-equivalent).
//
// Reference: backward of padertorch CNN2d layers 1-4 (pb_sed/experiments/weak_label_crnn/training.py:
// 158-169: channels 16,16,32,32,64).  dW[tap][n][c] = sum_t dout[t][n] * a[t + dt][c] is a GEMM with
// M = Cout (16..64), N = Cin (16/32), K = frames: far too narrow for a 128-lane tcgen05 tile (the
// tcgen05 weight-gradient kernel loses to plain FFMA here), but a natural fit for
// m16n8k8 fragments.  Staging is the one of wgrad_small.cu (dout tile + three input strips with
// norm + ReLU + mask applied, ONE staging feeds all nine taps); the pitches are == 8 (mod 32) words so
// that every fragment load is bank-conflict free.  Each warp owns one 16 (n) x 16 (c) plane for all nine
// taps (18 accumulator tiles = 72 registers) and, when the layer has fewer planes than warps, a share
// of the K steps; partial tiles meet in shared memory before ONE coalesced set of global atomics per CTA.
#include "common.cuh"
#include <cstdlib>

namespace {

struct WmParams {
  int B, F, T, relu, mask_out, single;
  int in_stride, out_stride;
  int in_bf16, out_bf16;      // storage type of `in` / `dout` (bf16 activation maps)
  long long w_tap_stride, w_sn, w_sc;
};

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // both pieces rounded to nearest TF32 (the tensor core would truncate the low 13 bits): |x - hi - lo| <= 2^-22 |x|
  hi = (__float_as_uint(x) + 0x00001000u) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x00001000u) & 0xFFFFE000u;
}

constexpr int TT = 64;                       // frames per work unit

template <int COUT, int CIN>
__global__ void __launch_bounds__(256, 2)
wgrad_mma_kernel(WmParams p, const float* __restrict__ in, const float* __restrict__ scale,
                 const float* __restrict__ shift, const int* __restrict__ seq_len,
                 const float* __restrict__ dout, float* __restrict__ dW, float* __restrict__ dbias) {
  constexpr int LDZ = COUT + 8, LDA = CIN + 8;
  constexpr int TS = (COUT / 16) * (CIN / 16);       // 16 x 16 planes
  constexpr int KS = 8 / TS;                         // warps sharing one plane split the K steps
  static_assert(TS >= 1 && TS <= 8 && KS * TS == 8, "warp mapping");
  extern __shared__ __align__(16) float sm[];
  float* zs = sm;                                    // [TT][LDZ]
  float* as = sm + TT * LDZ;                         // [3][TT + 2][LDA]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int ts = warp % TS, ksel = warp / TS;
  const int n0 = (ts / (CIN / 16)) * 16, c0 = (ts % (CIN / 16)) * 16;

  float acc[9][2][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[k][j][e] = 0.f;
  float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);     // this thread's channel quad of the bias gradient
  constexpr int ZQ = COUT / 4;
  static_assert(256 % ZQ == 0, "bias mapping");

  const int t_tiles = (p.T + TT - 1) / TT;
  const int units = p.B * p.F * t_tiles;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int tt = u % t_tiles, gq = u / t_tiles;
    const int f = gq % p.F, b = gq / p.F;
    const int t0 = tt * TT;
    const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
    const int len_out = p.mask_out ? len_b : p.T;
    __syncthreads();                                   // previous unit's readers are done
    {   // dout tile (+ bias partial sums: tid % ZQ is loop invariant)
      const long long z0 = ((long long)b * p.F + f) * p.T * p.out_stride;
      for (int i = tid; i < TT * ZQ; i += 256) {
        const int r = i / ZQ, q = i % ZQ;
        const int t = t0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < len_out) v = ld_act4(dout, z0 + (long long)t * p.out_stride + q * 4, p.out_bf16);
        *reinterpret_cast<float4*>(zs + r * LDZ + q * 4) = v;
        bsum.x += v.x; bsum.y += v.y; bsum.z += v.z; bsum.w += v.w;
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {   // input strips f-1, f, f+1, frames t0-1 .. t0+TT
      const int fs = f + d - 1;
      const bool f_ok = fs >= 0 && fs < p.F;
      const long long a0 = ((long long)b * p.F + (f_ok ? fs : 0)) * p.T * p.in_stride;
      for (int i = tid; i < (TT + 2) * (CIN / 4); i += 256) {
        const int r = i / (CIN / 4), q = i % (CIN / 4);
        const int t = t0 + r - 1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f_ok && t >= 0 && t < len_b) {
          v = ld_act4(in, a0 + (long long)t * p.in_stride + q * 4, p.in_bf16);
          if (scale) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + q * 4));
            const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + q * 4));
            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
            v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
          }
          if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        *reinterpret_cast<float4*>(as + (d * (TT + 2) + r) * LDA + q * 4) = v;
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int ks = ksel; ks < TT / 8; ks += KS) {
      const int r0 = ks * 8;
      // A fragment = dout^T: (row n, col t)
      uint32_t ah[4], al[4];
      split_tf32(zs[(r0 + tig) * LDZ + n0 + g], ah[0], al[0]);
      split_tf32(zs[(r0 + tig) * LDZ + n0 + g + 8], ah[1], al[1]);
      split_tf32(zs[(r0 + tig + 4) * LDZ + n0 + g], ah[2], al[2]);
      split_tf32(zs[(r0 + tig + 4) * LDZ + n0 + g + 8], ah[3], al[3]);
#pragma unroll
      for (int d = 0; d < 3; ++d) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          // strip rows r0 + tig + {0,1,2} and r0 + tig + 4 + {0,1,2} (row 0 of the strip = frame t0 - 1)
          const float* col = as + (d * (TT + 2) + r0 + tig) * LDA + c0 + 8 * j + g;
          uint32_t bh[6], bl[6];
#pragma unroll
          for (int e = 0; e < 3; ++e) {
            split_tf32(col[e * LDA], bh[e], bl[e]);
            split_tf32(col[(e + 4) * LDA], bh[3 + e], bl[3 + e]);
          }
#pragma unroll
          for (int e = 0; e < 3; ++e) {                // dt = e - 1
            const uint32_t b_hi[2] = {bh[e], bh[3 + e]}, b_lo[2] = {bl[e], bl[3 + e]};
            mma_tf32_16x8x8(acc[d * 3 + e][j], ah, b_hi);
            if (!p.single) {
              mma_tf32_16x8x8(acc[d * 3 + e][j], al, b_hi);
              mma_tf32_16x8x8(acc[d * 3 + e][j], ah, b_lo);
            }
          }
        }
      }
    }
  }
  // ---- reduce the KS partial planes in shared memory, then one coalesced set of global atomics per tap
  float* red = sm;                                     // [COUT][CIN] (<= 8 KB)
#pragma unroll                                          // (full unroll keeps acc[][] in registers)
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

template <int COUT, int CIN>
int launch_wm(const WmParams& p, const float* in, const float* scale, const float* shift, const int* seq_len,
              const float* dout, float* dW, float* dbias, cudaStream_t st) {
  const size_t smem = ((size_t)TT * (COUT + 8) + 3 * (size_t)(TT + 2) * (CIN + 8)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(wgrad_mma_kernel<COUT, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int units = p.B * p.F * cdiv(p.T, TT);
  int grid = 148 * 2;
  if (grid > units) grid = units;
  pbsed_note_kernel("wgrad_mma_kernel");
  wgrad_mma_kernel<COUT, CIN><<<grid, 256, smem, st>>>(p, in, scale, shift, seq_len, dout, dW, dbias);
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
  // measured at B = 32 (profiles/r01_s6_*): the 16-input-channel layers are bound by the synchronous
  // staging of their 16 k-frame units, not by arithmetic, and the FFMA kernel's lighter inner loop wins
  // there (0.41 vs 0.53 ms, 0.34 vs 0.35 ms); with 32 input channels the tensor-core version is ahead
  // (0.54 vs 0.69 ms, 0.41 vs 0.55 ms).  PBSED_WGRAD_MMA16=1 forces it for the narrow layers too.
  static const bool narrow_too = getenv("PBSED_WGRAD_MMA16") != nullptr;
  int rc;
  if (d->Cin == 16 && !narrow_too) return 0;
  if (d->Cout == 16 && d->Cin == 16)      rc = launch_wm<16, 16>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 16) rc = launch_wm<32, 16>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 32) rc = launch_wm<32, 32>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 64 && d->Cin == 32) rc = launch_wm<64, 32>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else return 0;
  *handled = 1;
  return rc;
}
