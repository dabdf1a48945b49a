// conv_cin1.cu -- the first CNN2d layer: ONE input channel (the log-mel map), <= 32 output channels.
//
// Reference: layer 0 of padertorch CNN2d as configured at
// pb_sed/experiments/weak_label_crnn/training.py:218-230 (input_layer: no norm / activation; 3x3,
// 1 -> 16 channels).  K = 9 MACs per output: a GEMM tile would be all padding, the layer is pure
// HBM streaming (read 1 float, write Cout floats per frame), so it gets a direct kernel:
// one thread per output frame, weights in shared memory, 64-byte contiguous stores per thread.
#include "common.cuh"

struct Cin1Params {
  int B, F_in, F_out, T, Cout, ntaps;
  int df[PBSED_MAX_TAPS], dt[PBSED_MAX_TAPS];
  int relu;
  int out_stride;
  int out_bf16;          // `out` (forward) / `dout` (weight gradient) is stored as bf16
};

__device__ __forceinline__ float cin1_load(const float* __restrict__ in, const Cin1Params& p, int b, int f,
                                           int t, int len_b, float sc, float sh, bool affine) {
  if (f < 0 || f >= p.F_in || t < 0 || t >= len_b) return 0.f;
  float v = __ldg(in + ((long long)b * p.F_in + f) * p.T + t);
  if (affine) v = fmaf(v, sc, sh);
  if (p.relu) v = fmaxf(v, 0.f);
  return v;
}

template <int COUT>
__global__ void __launch_bounds__(256)
conv_cin1_fwd_kernel(Cin1Params p, const float* __restrict__ in, const float* __restrict__ scale,
                     const float* __restrict__ shift, const int* __restrict__ seq_len,
                     const float* __restrict__ W, const float* __restrict__ bias,
                     float* __restrict__ out) {
  __shared__ float ws[PBSED_MAX_TAPS][COUT];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < p.ntaps * COUT; i += 256) ws[i / COUT][i % COUT] = __ldg(W + i);   // (tap, n, c=0)
  for (int i = threadIdx.x; i < COUT; i += 256) bs[i] = bias ? __ldg(bias + i) : 0.f;
  __syncthreads();
  const int fo = blockIdx.y, b = blockIdx.z;
  const int t = blockIdx.x * 256 + threadIdx.x;           // frames >= T compute on (masked) padding and store nothing
  const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
  const bool affine = scale != nullptr;
  const float sc = affine ? __ldg(scale) : 1.f, sh = affine ? __ldg(shift) : 0.f;
  float acc[COUT];
#pragma unroll
  for (int n = 0; n < COUT; ++n) acc[n] = bs[n];
  for (int tap = 0; tap < p.ntaps; ++tap) {
    const float a = cin1_load(in, p, b, fo + p.df[tap], t + p.dt[tap], len_b, sc, sh, affine);
#pragma unroll
    for (int n = 0; n < COUT; ++n) acc[n] = fmaf(a, ws[tap][n], acc[n]);
  }
  // stores: a thread owns one frame's COUT channels (64 / 128 bytes, lanes 64 / 128 bytes apart); the warp's 32 frames are
  // contiguous in memory, so the tile goes through shared memory and leaves as fully coalesced 16-byte-per-lane rows
  __shared__ __align__(16) float tile[8][32][COUT + 4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < COUT; n += 4)
    *reinterpret_cast<float4*>(&tile[warp][lane][n]) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
  __syncwarp();
  const int t_w = blockIdx.x * 256 + warp * 32;                      // first frame of this warp
  const long long o = (((long long)b * p.F_out + fo) * p.T + t_w) * p.out_stride;
  constexpr int Q = COUT / 4;
  if (p.out_stride == COUT) {
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const int i = lane + 32 * k, fr = i / Q, q = i % Q;
      if (t_w + fr < p.T) st_act4(out, o + (long long)fr * COUT + 4 * q, *reinterpret_cast<const float4*>(&tile[warp][fr][4 * q]), p.out_bf16);
    }
  } else {
#pragma unroll
    for (int n = 0; n < COUT; n += 4)
      st_act4(out, o + (long long)lane * p.out_stride + n, make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]), p.out_bf16);
  }
}

// dW[tap][n] += sum_rows dout[row][n] * a[row + off(tap)];  dbias[n] += sum_rows dout[row][n]
// CTA = one (b, fo) row group at a time: the <= 3 source rows it touches are staged (transformed) in
// shared memory once, then thread (frame lane, channel n) streams dout and reads the taps from smem.
constexpr int CIN1_TMAX = 2048;
template <int COUT>
__global__ void __launch_bounds__(256)
conv_cin1_wgrad_kernel(Cin1Params p, const float* __restrict__ in, const float* __restrict__ scale,
                       const float* __restrict__ shift, const int* __restrict__ seq_len,
                       const float* __restrict__ dout, int mask_out, float* __restrict__ dW,
                       float* __restrict__ dbias, int groups_per_cta, int df_min, int n_rows, int dt_min, int halo) {
  // thread = (frame lane rl, channel QUAD q): one 16-byte (8-byte bf16) dout load feeds 4 x ntaps FMAs, the strip value of
  // a tap is read once per quad instead of once per channel
  constexpr int Q = COUT / 4, RL = 256 / Q, RW = 32 / Q;     // frame lanes per CTA / per warp
  extern __shared__ float strip[];                         // [n_rows][T + halo]
  __shared__ float4 red[8][PBSED_MAX_TAPS + 1][Q];
  const int q = threadIdx.x % Q, rl = threadIdx.x / Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool affine = scale != nullptr;
  const float sc = affine ? __ldg(scale) : 1.f, sh = affine ? __ldg(shift) : 0.f;
  const int LD = p.T + halo;
  float4 acc[PBSED_MAX_TAPS + 1];                           // [ntaps] = bias sum
#pragma unroll
  for (int i = 0; i <= PBSED_MAX_TAPS; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int total = p.B * p.F_out;
  const int g0 = blockIdx.x * groups_per_cta, g1 = min(g0 + groups_per_cta, total);
  for (int g = g0; g < g1; ++g) {
    const int b = g / p.F_out, fo = g % p.F_out;
    const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
    const int len_out = mask_out ? len_b : p.T;
    __syncthreads();
    for (int i = threadIdx.x; i < n_rows * LD; i += 256) {
      const int r = i / LD, tt = i % LD;
      strip[i] = cin1_load(in, p, b, fo + df_min + r, tt + dt_min, len_b, sc, sh, affine);
    }
    __syncthreads();
    const long long z0 = ((long long)b * p.F_out + fo) * p.T * p.out_stride + 4 * q;
    for (int t = rl; t < len_out; t += RL) {
      const float4 dz = ld_act4(dout, z0 + (long long)t * p.out_stride, p.out_bf16);
      acc[PBSED_MAX_TAPS].x += dz.x; acc[PBSED_MAX_TAPS].y += dz.y; acc[PBSED_MAX_TAPS].z += dz.z; acc[PBSED_MAX_TAPS].w += dz.w;
#pragma unroll
      for (int tap = 0; tap < PBSED_MAX_TAPS; ++tap) {
        if (tap < p.ntaps) {
          const float a = strip[(p.df[tap] - df_min) * LD + t + p.dt[tap] - dt_min];
          acc[tap].x = fmaf(dz.x, a, acc[tap].x); acc[tap].y = fmaf(dz.y, a, acc[tap].y);
          acc[tap].z = fmaf(dz.z, a, acc[tap].z); acc[tap].w = fmaf(dz.w, a, acc[tap].w);
        }
      }
    }
  }
  // reduce over the frame lanes: shuffles inside the warp (lanes q + Q * k), then the 8 warps through shared memory
#pragma unroll
  for (int tap = 0; tap <= PBSED_MAX_TAPS; ++tap) {
    if (tap < p.ntaps || tap == PBSED_MAX_TAPS) {
      float4 v = acc[tap];
#pragma unroll
      for (int o = Q; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o); v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, o); v.w += __shfl_xor_sync(0xffffffffu, v.w, o);
      }
      if (lane < Q) red[warp][tap][lane] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (PBSED_MAX_TAPS + 1) * Q; i += 256) {
    const int tap = i / Q, qq = i % Q;
    if (tap >= p.ntaps && tap != PBSED_MAX_TAPS) continue;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < 8; ++w) { const float4 u = red[w][tap][qq]; s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w; }
    float* dst = tap < p.ntaps ? dW + tap * COUT + 4 * qq : (dbias ? dbias + 4 * qq : nullptr);
    if (dst) {
      if (s.x != 0.f) atomicAdd(dst + 0, s.x);
      if (s.y != 0.f) atomicAdd(dst + 1, s.y);
      if (s.z != 0.f) atomicAdd(dst + 2, s.z);
      if (s.w != 0.f) atomicAdd(dst + 3, s.w);
    }
  }
}

static bool cin1_ok(const pbsed_tapgemm_desc* d) {
  const int in_stride = d->in_stride > 0 ? d->in_stride : d->Cin;
  const int out_stride = d->out_stride > 0 ? d->out_stride : d->Cout;
  return d->Cin == 1 && in_stride == 1 && (d->Cout == 16 || d->Cout == 32) && out_stride % 4 == 0 &&
         d->w_sn == 1 && d->w_sc == 1 && d->w_tap_stride == d->Cout && d->per_f == 0 &&
         d->F_out <= 65535 && d->B <= 65535 && d->in_dtype == PBSED_F32;
}

static void cin1_fill(const pbsed_tapgemm_desc* d, Cin1Params& p) {
  p.B = d->B; p.F_in = d->F_in; p.F_out = d->F_out; p.T = d->T; p.Cout = d->Cout; p.ntaps = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) { p.df[i] = d->df[i]; p.dt[i] = d->dt[i]; }
  p.relu = d->relu;
  p.out_stride = d->out_stride > 0 ? d->out_stride : d->Cout;
  p.out_bf16 = d->out_dtype == PBSED_BF16;
}

int conv_cin1_fwd_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                           const float* shift, const int* seq_len, const float* W, const float* bias,
                           float* out, const float* ep_src, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!cin1_ok(d) || ep_src || (((uintptr_t)out) & 15)) return 0;
  Cin1Params p;
  cin1_fill(d, p);
  dim3 grid(cdiv(p.T, 256), p.F_out, p.B);
  pbsed_note_kernel("conv_cin1_fwd_kernel");
  if (p.Cout == 16) conv_cin1_fwd_kernel<16><<<grid, 256, 0, st>>>(p, in, scale, shift, seq_len, W, bias, out);
  else              conv_cin1_fwd_kernel<32><<<grid, 256, 0, st>>>(p, in, scale, shift, seq_len, W, bias, out);
  *handled = 1;
  return pbsed_after_launch();
}

int conv_cin1_wgrad_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                             const float* shift, const int* seq_len, const float* dout, int mask_out,
                             float* dW, float* dbias, cudaStream_t st, int* handled) {
  *handled = 0;
  if (!cin1_ok(d)) return 0;
  Cin1Params p;
  cin1_fill(d, p);
  if (p.T > CIN1_TMAX) return 0;
  int df_min = 0, df_max = 0, dt_min = 0, dt_max = 0;
  for (int i = 0; i < p.ntaps; ++i) {
    df_min = p.df[i] < df_min ? p.df[i] : df_min; df_max = p.df[i] > df_max ? p.df[i] : df_max;
    dt_min = p.dt[i] < dt_min ? p.dt[i] : dt_min; dt_max = p.dt[i] > dt_max ? p.dt[i] : dt_max;
  }
  const int n_rows = df_max - df_min + 1, halo = dt_max - dt_min;
  const size_t smem = (size_t)n_rows * (p.T + halo) * sizeof(float);
  if (smem > 160 * 1024) return 0;
  const int total = p.B * p.F_out;
  int gpc = cdiv(total, 148 * 8);
  if (gpc < 1) gpc = 1;
  dim3 grid(cdiv(total, gpc));
  cudaError_t e;
  pbsed_note_kernel("conv_cin1_wgrad_kernel");
  if (p.Cout == 16) {
    e = cudaFuncSetAttribute(conv_cin1_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    conv_cin1_wgrad_kernel<16><<<grid, 256, smem, st>>>(p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, gpc, df_min, n_rows, dt_min, halo);
  } else {
    e = cudaFuncSetAttribute(conv_cin1_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    conv_cin1_wgrad_kernel<32><<<grid, 256, smem, st>>>(p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, gpc, df_min, n_rows, dt_min, halo);
  }
  *handled = 1;
  return pbsed_after_launch();
}
