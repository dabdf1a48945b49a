// tapgemm.cu -- exact-fp32 (FFMA) tap-GEMM: forward / data-gradient / weight-gradient of the
// contraction behind CNN2d, CNN1d, the GRU projections and the output_net 1x1 convs.
//
//   out[(b,fo,t), n] = bias[n] + sum_tap sum_c a[(b, fo+df, t+dt), c] * W[tap][n][c]
//
// Reference: padertorch.contrib.je.modules.conv.{CNN2d,CNN1d} layer bodies as configured at
// pb_sed/experiments/weak_label_crnn/training.py:218-242: pre-activation
// norm -> ReLU -> zero 'same' pad -> conv(+bias).  The norm-apply + ReLU + sequence mask are
// fused into the operand load (scale/shift per channel), so the normalised tensor never
// round-trips through HBM.
//
// Tiling: one CTA owns BM consecutive frames t of one (b, fo) row group and BN output channels;
// a tap whose source row fo+df falls outside the map is skipped for the whole CTA (the 'same'
// zero padding along f costs nothing), padding along t is predicated per row.
#include "common.cuh"

struct TapParams {
  int B, F_in, F_out, T, Cin, Cout, ntaps;
  int df[PBSED_MAX_TAPS];
  int dt[PBSED_MAX_TAPS];
  int relu, per_f;
  long long w_tap_stride, w_sn, w_sc;
  int in_stride, out_stride;    // row strides in floats (>= Cin / Cout)
};

static int fill_params(const pbsed_tapgemm_desc* d, TapParams& p) {
  if (!d || d->ntaps < 1 || d->ntaps > PBSED_MAX_TAPS) return PBSED_EINVAL;
  if (d->B < 1 || d->F_in < 1 || d->F_out < 1 || d->T < 1 || d->Cin < 1 || d->Cout < 1) return PBSED_EINVAL;
  p.B = d->B; p.F_in = d->F_in; p.F_out = d->F_out; p.T = d->T; p.Cin = d->Cin; p.Cout = d->Cout;
  p.ntaps = d->ntaps;
  for (int i = 0; i < d->ntaps; ++i) { p.df[i] = d->df[i]; p.dt[i] = d->dt[i]; }
  p.relu = d->relu; p.per_f = d->per_f;
  p.w_tap_stride = d->w_tap_stride; p.w_sn = d->w_sn; p.w_sc = d->w_sc;
  p.in_stride = d->in_stride > 0 ? d->in_stride : d->Cin;
  p.out_stride = d->out_stride > 0 ? d->out_stride : d->Cout;
  if (p.in_stride < p.Cin || p.out_stride < p.Cout) return PBSED_EINVAL;
  return 0;
}

// ---------------------------------------------------------------- operand load (shared by all)
// loads a [ROWS x BK] tile of the transformed input a[(b,f_src,t_base+r), c0+k] into smem
// TRANSPOSED (dst[k][r], leading dim LD) or ROW-major (dst[r][k]) depending on TRANSPOSE.
template <int ROWS, int BK, int LD, int NT, bool TRANSPOSE>
__device__ __forceinline__ void load_a_tile(float* __restrict__ dst, const TapParams& p,
                                            const float* __restrict__ in,
                                            const float* __restrict__ scale,
                                            const float* __restrict__ shift,
                                            int b, int f_src, int t_base, int c0, int len_b, int tid) {
  const long long row0 = ((long long)b * p.F_in + f_src) * p.T;
  const int aff_base = p.per_f ? f_src * p.Cin : 0;
  const bool vec = ((p.Cin & 3) == 0) && ((p.in_stride & 3) == 0) && (c0 + BK <= p.Cin);
  if (vec) {
    constexpr int KV = BK / 4;
    for (int i = tid; i < ROWS * KV; i += NT) {
      const int r = i / KV, kv = i % KV;
      const int t = t_base + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t >= 0 && t < len_b) {
        const int c = c0 + kv * 4;
        v = __ldg(reinterpret_cast<const float4*>(in + (row0 + t) * p.in_stride + c));
        if (scale) {
          const float4 sc = __ldg(reinterpret_cast<const float4*>(scale + aff_base + c));
          const float4 sh = __ldg(reinterpret_cast<const float4*>(shift + aff_base + c));
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
          v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      }
      if (TRANSPOSE) {
        dst[(kv * 4 + 0) * LD + r] = v.x; dst[(kv * 4 + 1) * LD + r] = v.y;
        dst[(kv * 4 + 2) * LD + r] = v.z; dst[(kv * 4 + 3) * LD + r] = v.w;
      } else {
        *reinterpret_cast<float4*>(dst + r * LD + kv * 4) = v;
      }
    }
  } else {
    for (int i = tid; i < ROWS * BK; i += NT) {
      const int r = i / BK, k = i % BK;
      const int t = t_base + r, c = c0 + k;
      float v = 0.f;
      if (t >= 0 && t < len_b && c < p.Cin) {
        v = __ldg(in + (row0 + t) * p.in_stride + c);
        if (scale) v = fmaf(v, __ldg(scale + aff_base + c), __ldg(shift + aff_base + c));
        if (p.relu) v = fmaxf(v, 0.f);
      }
      if (TRANSPOSE) dst[k * LD + r] = v; else dst[r * LD + k] = v;
    }
  }
}

// ---------------------------------------------------------------- forward / dgrad kernel
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
tapgemm_ffma_kernel(TapParams p, const float* __restrict__ in, const float* __restrict__ scale,
                    const float* __restrict__ shift, const int* __restrict__ seq_len,
                    const float* __restrict__ W, const float* __restrict__ bias,
                    float* __restrict__ out, const float* __restrict__ ep_src,
                    const float* __restrict__ ep_scale, const float* __restrict__ ep_shift,
                    int n_tiles) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDA = BM + 4, LDB = BN + 4;
  __shared__ __align__(16) float As[BK * LDA];
  __shared__ __align__(16) float Bs[BK * LDB];

  const int tid = threadIdx.x;
  const int tile_n = blockIdx.x % n_tiles, tile_t = blockIdx.x / n_tiles;
  const int fo = blockIdx.y, b = blockIdx.z;
  const int t0 = tile_t * BM, n0 = tile_n * BN;
  const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < p.ntaps; ++tap) {
    const int f_src = fo + p.df[tap];
    if (f_src < 0 || f_src >= p.F_in) continue;           // whole-CTA uniform
    const int t_base = t0 + p.dt[tap];
    if (t_base >= len_b || t_base + BM <= 0) continue;     // tile entirely in the zero padding
    const float* Wt = W + (long long)tap * p.w_tap_stride;
    for (int c0 = 0; c0 < p.Cin; c0 += BK) {
      load_a_tile<BM, BK, LDA, NT, true>(As, p, in, scale, shift, b, f_src, t_base, c0, len_b, tid);
      // weights: Bs[k][n] = W[tap][(n0+n)][(c0+k)]
      if (p.w_sc == 1) {
        for (int i = tid; i < BK * BN; i += NT) {
          const int n = i / BK, k = i % BK;
          float v = 0.f;
          if (n0 + n < p.Cout && c0 + k < p.Cin) v = __ldg(Wt + (long long)(n0 + n) * p.w_sn + (c0 + k));
          Bs[k * LDB + n] = v;
        }
      } else {
        for (int i = tid; i < BK * BN; i += NT) {
          const int k = i / BN, n = i % BN;
          float v = 0.f;
          if (n0 + n < p.Cout && c0 + k < p.Cin)
            v = __ldg(Wt + (long long)(n0 + n) * p.w_sn + (long long)(c0 + k) * p.w_sc);
          Bs[k * LDB + n] = v;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], bb[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          const float4 v = *reinterpret_cast<const float4*>(&As[k * LDA + ty * TM + i]);
          a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        }
        if (TN % 4 == 0) {
#pragma unroll
          for (int j = 0; j < TN; j += 4) {
            const float4 v = *reinterpret_cast<const float4*>(&Bs[k * LDB + tx * TN + j]);
            bb[j] = v.x; bb[(j + 1) % TN] = v.y; bb[(j + 2) % TN] = v.z; bb[(j + 3) % TN] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < TN; ++j) bb[j] = Bs[k * LDB + tx * TN + j];
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // epilogue
  const long long orow0 = ((long long)b * p.F_out + fo) * p.T;
  const int ep_base = p.per_f ? fo * p.Cout : 0;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int t = t0 + ty * TM + i;
    if (t >= p.T) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= p.Cout) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + n);
      if (ep_src) {
        float keep = 0.f;
        if (t < len_b) {
          float s = __ldg(ep_src + (orow0 + t) * p.out_stride + n);
          if (ep_scale) s = fmaf(s, __ldg(ep_scale + ep_base + n), __ldg(ep_shift + ep_base + n));
          keep = s > 0.f ? 1.f : 0.f;
        }
        v *= keep;
      }
      out[(orow0 + t) * p.out_stride + n] = v;
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_fwd(const TapParams& p, const float* in, const float* scale, const float* shift,
                      const int* seq_len, const float* W, const float* bias, float* out,
                      const float* ep_src, const float* ep_scale, const float* ep_shift,
                      cudaStream_t st) {
  const int n_tiles = cdiv(p.Cout, BN);
  dim3 grid(cdiv(p.T, BM) * n_tiles, p.F_out, p.B);
  if (grid.y > 65535 || grid.z > 65535) return PBSED_EINVAL;
  pbsed_note_kernel("tapgemm_ffma_kernel");
  tapgemm_ffma_kernel<BM, BN, BK, TM, TN><<<grid, (BM / TM) * (BN / TN), 0, st>>>(
      p, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift, n_tiles);
  return pbsed_after_launch();
}

int tapgemm_tc_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                        const float* shift, const int* seq_len, const float* W, const float* bias,
                        float* out, const float* ep_src, const float* ep_scale,
                        const float* ep_shift, double* out_stats, const float* ep_mean,
                        const float* ep_rstd, double* ep_sums, void* workspace, long long ws_bytes,
                        cudaStream_t st, int* handled);

int tapgemm_fw_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale, const float* shift,
                        const int* seq_len, const float* W, const float* bias, float* out, const float* ep_src,
                        const float* ep_scale, const float* ep_shift, double* out_stats, const float* ep_mean,
                        const float* ep_rstd, double* ep_sums, void* workspace, long long ws_bytes,
                        cudaStream_t st, int* handled);
int conv_cin1_fwd_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                           const float* shift, const int* seq_len, const float* W, const float* bias,
                           float* out, const float* ep_src, cudaStream_t st, int* handled);
int conv_cin1_wgrad_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                             const float* shift, const int* seq_len, const float* dout, int mask_out,
                             float* dW, float* dbias, cudaStream_t st, int* handled);

int wgrad_small_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                         const float* shift, const int* seq_len, const float* dout, int mask_out,
                         float* dW, float* dbias, cudaStream_t st, int* handled);

static int tapgemm_plain(const pbsed_tapgemm_desc* d, const TapParams& p, const float* in,
                         const float* scale, const float* shift, const int* seq_len,
                         const int* ep_seq_len, const float* W, const float* bias, float* out,
                         const float* ep_src, const float* ep_scale, const float* ep_shift,
                         cudaStream_t st) {
  int rc;
  {
    int handled = 0;
    rc = conv_cin1_fwd_dispatch(d, in, scale, shift, seq_len, W, bias, out, ep_src, st, &handled);
    if (handled || rc) return rc;
  }
  if (p.Cout <= 16)
    return launch_fwd<128, 16, 16, 8, 2>(p, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift, st);
  if (p.Cout <= 32)
    return launch_fwd<128, 32, 16, 8, 4>(p, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift, st);
  return launch_fwd<128, 64, 16, 8, 8>(p, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift, st);
}

extern "C" int pbsed_tapgemm(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                             const float* shift, const int* seq_len, const float* W,
                             const float* bias, float* out, const float* ep_src,
                             const float* ep_scale, const float* ep_shift, double* out_stats,
                             const float* ep_mean, const float* ep_rstd, double* ep_sums,
                             void* workspace, long long workspace_bytes, void* stream) {
  TapParams p;
  int rc = fill_params(d, p);
  if (rc) return rc;
  if (!in || !W || !out) return PBSED_EINVAL;
  if ((scale == nullptr) != (shift == nullptr)) return PBSED_EINVAL;
  if (ep_sums && (!ep_src || !ep_mean || !ep_rstd)) return PBSED_EINVAL;
  if ((out_stats || ep_sums) && p.out_stride != p.Cout) return PBSED_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  const int* load_seq = d->no_input_mask ? nullptr : seq_len;
  int fw_done = 0;
  {                                 // first layer (one input channel): direct kernel, may write a bf16 map
    int handled = 0;
    rc = conv_cin1_fwd_dispatch(d, in, scale, shift, load_seq, W, bias, out, ep_src, st, &handled);
    if (rc) return rc;
    if (handled) goto sums;
  }
  if (d->precision != 0) {
    // narrow 3x3 layers: the frequency-walking persistent kernel
    rc = tapgemm_fw_dispatch(d, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift,
                             out_stats, ep_mean, ep_rstd, ep_sums, workspace, workspace_bytes, st, &fw_done);
    if (fw_done || rc) return rc;
    {
      int handled = 0;
      rc = tapgemm_tc_dispatch(d, in, scale, shift, seq_len, W, bias, out, ep_src, ep_scale, ep_shift,
                               out_stats, ep_mean, ep_rstd, ep_sums, workspace, workspace_bytes, st, &handled);
      if (handled || rc) return rc;
    }
  }
  if (d->in_dtype != PBSED_F32 || d->out_dtype != PBSED_F32) return PBSED_EINVAL;   // bf16 maps: tensor-core kernels only
  rc = tapgemm_plain(d, p, in, scale, shift, load_seq, seq_len, W, bias, out, ep_src, ep_scale, ep_shift, st);
  if (rc) return rc;
sums:
  // kernels without fused reductions: run them as separate passes over the finished map
  if (out_stats) {
    rc = pbsed_channel_stats(out, p.B, p.F_out, p.T, p.Cout, p.per_f, seq_len, out_stats, d->out_dtype, stream);
    if (rc) return rc;
  }
  if (ep_sums)
    rc = pbsed_norm_bwd_reduce(out, ep_src, p.B, p.F_out, p.T, p.Cout, p.per_f, seq_len, ep_mean, ep_rstd, ep_sums, PBSED_F32, stream);
  return rc;
}

// ---------------------------------------------------------------- weight gradient
//   dW[tap][n][c] += sum_rows dout[row][n] * a[src(row,tap)][c]
// CTA = (group of G (b,fo) row groups) x tap x (BN x BC) output tile; reduction over frames in
// steps of BKR rows; one fp32 atomicAdd per output element per CTA at the end.
template <int BN, int BC, int TN, int TC, int BKR>
__global__ void __launch_bounds__((BN / TN) * (BC / TC))
tapgemm_wgrad_kernel(TapParams p, const float* __restrict__ in, const float* __restrict__ scale,
                     const float* __restrict__ shift, const int* __restrict__ seq_len,
                     const float* __restrict__ dout, int mask_out, float* __restrict__ dW,
                     float* __restrict__ dbias, int groups_per_cta, int c_tiles) {
  constexpr int NT = (BN / TN) * (BC / TC);
  constexpr int LDZ = BN + 4, LDA = BC + 4;
  __shared__ __align__(16) float Zs[BKR * LDZ];
  __shared__ __align__(16) float As[BKR * LDA];

  const int tid = threadIdx.x;
  const int tap = blockIdx.y;
  const int tile_c = blockIdx.z % c_tiles, tile_n = blockIdx.z / c_tiles;
  const int n0 = tile_n * BN, c0 = tile_c * BC;
  const int tn = tid / (BC / TC), tc = tid % (BC / TC);
  const bool do_bias = (dbias != nullptr) && tap == 0 && tile_c == 0;
  const bool zvec = ((p.Cout & 3) == 0) && ((p.out_stride & 3) == 0) && (n0 + BN <= p.Cout);

  float acc[TN][TC];
#pragma unroll
  for (int i = 0; i < TN; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;

  const int total_groups = p.B * p.F_out;
  const int g_begin = blockIdx.x * groups_per_cta;
  const int g_end = min(g_begin + groups_per_cta, total_groups);
  for (int g = g_begin; g < g_end; ++g) {
    const int b = g / p.F_out, fo = g % p.F_out;
    const int f_src = fo + p.df[tap];
    const bool f_ok = (f_src >= 0 && f_src < p.F_in);
    if (!f_ok && !do_bias) continue;
    const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
    const int len_out = mask_out ? len_b : p.T;
    const long long zrow0 = ((long long)b * p.F_out + fo) * p.T;
    for (int t0 = 0; t0 < len_out; t0 += BKR) {
      // dout tile Zs[kr][n]
      if (zvec) {
        constexpr int NV = BN / 4;
        for (int i = tid; i < BKR * NV; i += NT) {
          const int r = i / NV, nv = i % NV;
          const int t = t0 + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (t < len_out) v = __ldg(reinterpret_cast<const float4*>(dout + (zrow0 + t) * p.out_stride + n0 + nv * 4));
          *reinterpret_cast<float4*>(Zs + r * LDZ + nv * 4) = v;
        }
      } else {
        for (int i = tid; i < BKR * BN; i += NT) {
          const int r = i / BN, n = i % BN;
          const int t = t0 + r;
          float v = 0.f;
          if (t < len_out && n0 + n < p.Cout) v = __ldg(dout + (zrow0 + t) * p.out_stride + n0 + n);
          Zs[r * LDZ + n] = v;
        }
      }
      if (f_ok)
        load_a_tile<BKR, BC, LDA, NT, false>(As, p, in, scale, shift, b, f_src, t0 + p.dt[tap], c0, len_b, tid);
      __syncthreads();
      if (f_ok) {
#pragma unroll 8
        for (int kr = 0; kr < BKR; ++kr) {
          float z[TN], a[TC];
#pragma unroll
          for (int i = 0; i < TN; ++i) z[i] = Zs[kr * LDZ + tn * TN + i];
#pragma unroll
          for (int j = 0; j < TC; ++j) a[j] = As[kr * LDA + tc * TC + j];
#pragma unroll
          for (int i = 0; i < TN; ++i)
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(z[i], a[j], acc[i][j]);
        }
      }
      if (do_bias && tid < BN) {
#pragma unroll 8
        for (int kr = 0; kr < BKR; ++kr) bsum += Zs[kr * LDZ + tid];
      }
      __syncthreads();
    }
  }

  float* dWt = dW + (long long)tap * p.w_tap_stride;
#pragma unroll
  for (int i = 0; i < TN; ++i) {
    const int n = n0 + tn * TN + i;
    if (n >= p.Cout) continue;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const int c = c0 + tc * TC + j;
      if (c >= p.Cin) continue;
      const float v = acc[i][j];
      if (v != 0.f) atomicAdd(dWt + (long long)n * p.w_sn + (long long)c * p.w_sc, v);
    }
  }
  if (do_bias && tid < BN && n0 + tid < p.Cout && bsum != 0.f) atomicAdd(dbias + n0 + tid, bsum);
}

template <int BN, int BC, int TN, int TC, int BKR>
static int launch_wgrad(const TapParams& p, const float* in, const float* scale, const float* shift,
                        const int* seq_len, const float* dout, int mask_out, float* dW, float* dbias,
                        cudaStream_t st) {
  const int n_tiles = cdiv(p.Cout, BN), c_tiles = cdiv(p.Cin, BC);
  const int total_groups = p.B * p.F_out;
  // aim for ~8 waves of 148 SMs worth of CTAs, but never fewer than 1 group per CTA
  long long ctas_per_group = (long long)p.ntaps * n_tiles * c_tiles;
  int gpc = (int)((total_groups * ctas_per_group + 148LL * 16 - 1) / (148LL * 16));
  if (gpc < 1) gpc = 1;
  if (gpc > 64) gpc = 64;
  dim3 grid(cdiv(total_groups, gpc), p.ntaps, n_tiles * c_tiles);
  if (grid.z > 65535) return PBSED_EINVAL;
  pbsed_note_kernel("tapgemm_wgrad_kernel");
  tapgemm_wgrad_kernel<BN, BC, TN, TC, BKR><<<grid, (BN / TN) * (BC / TC), 0, st>>>(
      p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, gpc, c_tiles);
  return pbsed_after_launch();
}

int tapgemm_wgrad_tc_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                              const float* shift, const int* seq_len, const float* dout,
                              int mask_out, float* dW, float* dbias, cudaStream_t st, int* handled);
int tapgemm_wgrad_stack_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                                 const float* shift, const int* seq_len, const float* dout,
                                 int mask_out, float* dW, float* dbias, cudaStream_t st, int* handled);
int wgrad_mma_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale, const float* shift,
                       const int* seq_len, const float* dout, int mask_out, float* dW, float* dbias,
                       cudaStream_t st, int* handled);

extern "C" int pbsed_tapgemm_wgrad(const pbsed_tapgemm_desc* d, const float* in, const float* scale,
                                   const float* shift, const int* seq_len, const float* dout,
                                   int mask_out, float* dW, float* dbias, void* stream) {
  TapParams p;
  int rc = fill_params(d, p);
  if (rc) return rc;
  if (!in || !dout || !dW) return PBSED_EINVAL;
  if ((scale == nullptr) != (shift == nullptr)) return PBSED_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  {
    int handled = 0;
    rc = conv_cin1_wgrad_dispatch(d, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st, &handled);
    if (handled || rc) return rc;
    rc = tapgemm_wgrad_stack_dispatch(d, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st, &handled);
    if (handled || rc) return rc;
    rc = wgrad_mma_dispatch(d, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st, &handled);
    if (handled || rc) return rc;
    rc = wgrad_small_dispatch(d, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st, &handled);
    if (handled || rc) return rc;
  }
  if (d->precision != 0) {
    int handled = 0;
    rc = tapgemm_wgrad_tc_dispatch(d, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st, &handled);
    if (handled || rc) return rc;
  }
  if (d->in_dtype != PBSED_F32 || d->out_dtype != PBSED_F32) return PBSED_EINVAL;   // bf16 maps: tensor-core / narrow kernels only
  if (p.Cout <= 16 && p.Cin <= 16)
    return launch_wgrad<16, 16, 2, 2, 32>(p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st);
  if (p.Cout <= 32 || p.Cin <= 32)
    return launch_wgrad<32, 32, 4, 4, 32>(p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st);
  return launch_wgrad<64, 64, 4, 4, 32>(p, in, scale, shift, seq_len, dout, mask_out, dW, dbias, st);
}
