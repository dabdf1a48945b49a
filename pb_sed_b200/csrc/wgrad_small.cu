// wgrad_small.cu -- weight gradient of the narrow 3x3 conv layers (16 / 32 channels on the input side) in EXACT fp32
// (precision 0: the reference the tensor-core kernels are tested against, and the 'fp32' mode of the library).
//
// Reference: backward of padertorch CNN2d layers 1-4 (pb_sed/experiments/weak_label_crnn/training.py:
// 158-169: channels 16,16,32,32,64).  These layers are HBM streams (each map is 131 MB at B = 32) with
// only 2.3k-18k MACs per frame, and the generic FFMA kernel re-reads both maps once per tap.  Here ONE CTA forms all
// nine taps from a single staged tile: dout tile (64 frames x Cout) + three input strips (f-1, f, f+1; 66 frames x
// Cin; norm + ReLU + mask applied while staging), each thread keeps NPT x CPT x 9 accumulators in registers across
// all its work units and slides a 3-frame window over the input strip, so shared memory is read ~once per 6 FMAs.
// (In the tensor-core modes these layers run as row-stacked tcgen05 tiles, tapgemm_tc.cu.)
#include "common.cuh"

struct WsParams {
  int B, F, T, relu, mask_out;
  int in_stride, out_stride;
  int in_bf16, out_bf16;      // storage type of `in` / `dout` (bf16 activation maps)
  long long w_tap_stride, w_sn, w_sc;
};

template <int COUT, int CIN, int NPT, int CPT>
__global__ void __launch_bounds__(256)
wgrad_small_kernel(WsParams p, const float* __restrict__ in, const float* __restrict__ scale,
                   const float* __restrict__ shift, const int* __restrict__ seq_len,
                   const float* __restrict__ dout, float* __restrict__ dW, float* __restrict__ dbias) {
  constexpr int TT = 64;
  constexpr int P = (COUT / NPT) * (CIN / CPT);      // threads covering one (n, c) plane
  constexpr int TH = 256 / P;                        // frame sub-ranges processed side by side
  static_assert(P * TH == 256 && TT % TH == 0, "thread mapping");
  constexpr int LDA = CIN + 4;                       // row pitch of the input strips (bank spread)
  __shared__ __align__(16) float zs[TT][COUT];
  __shared__ __align__(16) float as[3][TT + 2][LDA];

  const int tid = threadIdx.x;
  const int th = tid / P, pid = tid % P;
  const int cb = pid % (CIN / CPT), nb = pid / (CIN / CPT);
  const int c_base = cb * CPT, n_base = nb * NPT;

  float acc[9][NPT][CPT];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) acc[k][i][j] = 0.f;
  float bsum[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) bsum[i] = 0.f;

  const int t_tiles = (p.T + TT - 1) / TT;
  const int units = p.B * p.F * t_tiles;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    const int tt = u % t_tiles, gq = u / t_tiles;
    const int f = gq % p.F, b = gq / p.F;
    const int t0 = tt * TT;
    const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
    const int len_out = p.mask_out ? len_b : p.T;
    __syncthreads();                                   // previous unit's readers are done
    {   // dout tile
      const long long z0 = ((long long)b * p.F + f) * p.T * p.out_stride;
      for (int i = tid; i < TT * (COUT / 4); i += 256) {
        const int r = i / (COUT / 4), q = i % (COUT / 4);
        const int t = t0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < len_out) v = ld_act4(dout, z0 + (long long)t * p.out_stride + q * 4, p.out_bf16);
        *reinterpret_cast<float4*>(&zs[r][q * 4]) = v;
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
        *reinterpret_cast<float4*>(&as[d][r][q * 4]) = v;
      }
    }
    __syncthreads();
    // frames [th*TT/TH, (th+1)*TT/TH) of the tile; window w[d][0..2] = strip rows r-1, r, r+1 (+1 halo shift)
    const int r_begin = th * (TT / TH), r_end = r_begin + TT / TH;
    float w0[3][CPT], w1[3][CPT], w2[3][CPT];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int j = 0; j < CPT; ++j) { w0[d][j] = as[d][r_begin][c_base + j]; w1[d][j] = as[d][r_begin + 1][c_base + j]; }
#pragma unroll 4
    for (int r = r_begin; r < r_end; ++r) {
      float z[NPT];
#pragma unroll
      for (int i = 0; i < NPT; ++i) z[i] = zs[r][n_base + i];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int j = 0; j < CPT; ++j) w2[d][j] = as[d][r + 2][c_base + j];
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        bsum[i] += z[i];
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            acc[d * 3 + 0][i][j] = fmaf(z[i], w0[d][j], acc[d * 3 + 0][i][j]);
            acc[d * 3 + 1][i][j] = fmaf(z[i], w1[d][j], acc[d * 3 + 1][i][j]);
            acc[d * 3 + 2][i][j] = fmaf(z[i], w2[d][j], acc[d * 3 + 2][i][j]);
          }
      }
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int j = 0; j < CPT; ++j) { w0[d][j] = w1[d][j]; w1[d][j] = w2[d][j]; }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const float v = acc[k][i][j];
        if (v != 0.f)
          atomicAdd(dW + (long long)k * p.w_tap_stride + (long long)(n_base + i) * p.w_sn + (long long)(c_base + j) * p.w_sc, v);
      }
  if (dbias && cb == 0) {
#pragma unroll
    for (int i = 0; i < NPT; ++i)
      if (bsum[i] != 0.f) atomicAdd(dbias + n_base + i, bsum[i]);
  }
}

template <int COUT, int CIN, int NPT, int CPT>
static int launch_ws(const WsParams& p, const float* in, const float* scale, const float* shift,
                     const int* seq_len, const float* dout, float* dW, float* dbias, cudaStream_t st) {
  const int units = p.B * p.F * cdiv(p.T, 64);
  int grid = 148 * 4;
  if (grid > units) grid = units;
  pbsed_note_kernel("wgrad_small_kernel");
  wgrad_small_kernel<COUT, CIN, NPT, CPT><<<grid, 256, 0, st>>>(p, in, scale, shift, seq_len, dout, dW, dbias);
  return pbsed_after_launch();
}

int wgrad_small_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                         const float* shift, const int* seq_len, const float* dout, int mask_out,
                         float* dW, float* dbias, cudaStream_t st, int* handled) {
  *handled = 0;
  if (d->ntaps != 9 || d->F_in != d->F_out || d->per_f) return 0;
  for (int i = 0; i < 9; ++i)
    if (d->df[i] != i / 3 - 1 || d->dt[i] != i % 3 - 1) return 0;
  WsParams p;
  p.B = d->B; p.F = d->F_in; p.T = d->T; p.relu = d->relu; p.mask_out = mask_out;
  p.in_stride = d->in_stride > 0 ? d->in_stride : d->Cin;
  p.in_bf16 = d->in_dtype == PBSED_BF16; p.out_bf16 = d->out_dtype == PBSED_BF16;
  p.out_stride = d->out_stride > 0 ? d->out_stride : d->Cout;
  p.w_tap_stride = d->w_tap_stride; p.w_sn = d->w_sn; p.w_sc = d->w_sc;
  if (p.in_stride % 4 || p.out_stride % 4) return 0;
  if ((((uintptr_t)in | (uintptr_t)dout | (uintptr_t)scale | (uintptr_t)shift) & 15) != 0) return 0;
  int rc;
  if (d->Cout == 16 && d->Cin == 16)      rc = launch_ws<16, 16, 2, 1>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 16) rc = launch_ws<32, 16, 2, 1>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 32 && d->Cin == 32) rc = launch_ws<32, 32, 2, 2>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else if (d->Cout == 64 && d->Cin == 32) rc = launch_ws<64, 32, 4, 2>(p, in, scale, shift, seq_len, dout, dW, dbias, st);
  else return 0;
  *handled = 1;
  return rc;
}
