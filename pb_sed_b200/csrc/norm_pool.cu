// norm_pool.cu -- masked batch statistics, running-statistics update, frequency max-pool and
// their backward passes.
//
// Reference semantics (padertorch.contrib.je.modules.norm.Normalization as used
// with norm='batch', norm_kwargs={'eps': 1e-3} at pb_sed/experiments/weak_label_crnn/training.py:
// 224-225,236-237,255-256): statistics over (b, f, t) per channel [2-D] / (b, t) per feature [1-D],
// masked by seq_len along t, biased variance, momentum 0.95 running mean / running power,
// learnable scale/shift; padded frames are zeroed.  F.max_pool2d((2,1)) after the conv.
#include "common.cuh"

// ------------------------------------------------------------------ channel statistics
// generic two-quantity column reduction over valid rows:
//   MODE 0: (x, x^2)            MODE 1: (g, g * (x - mean) * rstd)
template <int MODE>
__global__ void __launch_bounds__(256)
colreduce_kernel(const float* __restrict__ a, const float* __restrict__ x, int F, int T, int C,
                 int per_f, const int* __restrict__ seq_len, const float* __restrict__ mean,
                 const float* __restrict__ rstd, double* __restrict__ out, int t_chunk) {
  __shared__ float red[2][256];
  const int tid = threadIdx.x;
  const int g = blockIdx.x;                     // (b, f) row group
  const int b = g / F, f = g % F;
  const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
  const int t0 = blockIdx.y * t_chunk;
  const int t1 = min(t0 + t_chunk, len_b);
  if (t0 >= t1) return;                          // uniform per CTA
  const long long row0 = (long long)g * T;
  const int Cw = C < 256 ? C : 256;
  const int rl = 256 / Cw;                       // row lanes
  const int lane_r = tid / Cw, c_local = tid % Cw;
  const bool active = lane_r < rl;
  const int idx_base = per_f ? f * C : 0;
  for (int c = c_local; c < C; c += Cw) {        // >1 trip only when C > 256 (then rl == 1)
    float s0 = 0.f, s1 = 0.f;
    if (active) {
      float mu = 0.f, rs = 1.f;
      if (MODE == 1) { mu = __ldg(mean + idx_base + c); rs = __ldg(rstd + idx_base + c); }
      for (int t = t0 + lane_r; t < t1; t += rl) {
        const long long o = (row0 + t) * C + c;
        if (MODE == 0) {
          const float v = __ldg(a + o);
          s0 += v; s1 = fmaf(v, v, s1);
        } else {
          const float gv = __ldg(a + o);
          const float xh = (__ldg(x + o) - mu) * rs;
          s0 += gv; s1 = fmaf(gv, xh, s1);
        }
      }
    }
    if (rl == 1) {
      atomicAdd(out + 2 * (idx_base + c), (double)s0);
      atomicAdd(out + 2 * (idx_base + c) + 1, (double)s1);
    } else {
      red[0][tid] = s0; red[1][tid] = s1;
      __syncthreads();
      if (lane_r == 0) {
        double d0 = 0., d1 = 0.;
        for (int r = 0; r < rl; ++r) { d0 += (double)red[0][r * Cw + c_local]; d1 += (double)red[1][r * Cw + c_local]; }
        atomicAdd(out + 2 * (idx_base + c), d0);
        atomicAdd(out + 2 * (idx_base + c) + 1, d1);
      }
      __syncthreads();
    }
  }
}

// float4 variant (C % 4 == 0, C <= 1024): thread = (row lane, channel quad); 4 independent rows in flight
template <int MODE>
__global__ void __launch_bounds__(256)
colreduce4_kernel(const void* __restrict__ a, const void* __restrict__ x, int F, int T, int C,
                  int per_f, const int* __restrict__ seq_len, const float* __restrict__ mean,
                  const float* __restrict__ rstd, double* __restrict__ out, int t_chunk, int bf) {
  __shared__ float4 red[2][256];
  const int tid = threadIdx.x;
  const int g = blockIdx.x;
  const int b = g / F, f = g % F;
  const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
  const int t0 = blockIdx.y * t_chunk;
  const int t1 = min(t0 + t_chunk, len_b);
  if (t0 >= t1) return;
  const int C4 = C >> 2;
  const int rl = 256 / C4;
  const int lane_r = tid / C4, q = tid % C4;
  const int idx_base = (per_f ? f * C : 0) + q * 4;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (lane_r < rl) {
    float4 mu = s0, rs = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == 1) {
      mu = __ldg(reinterpret_cast<const float4*>(mean + idx_base));
      rs = __ldg(reinterpret_cast<const float4*>(rstd + idx_base));
    }
    const long long e0 = (long long)g * T * C + 4 * q;        // element index of (row group, frame 0, this quad)
    for (int t = t0 + lane_r; t < t1; t += 4 * rl) {
      float4 v[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int tt = t + k * rl;
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f); w[k] = mu;
        if (tt < t1) {
          v[k] = ld_act4(a, e0 + (long long)tt * C, bf);
          if (MODE == 1) w[k] = ld_act4(x, e0 + (long long)tt * C, bf);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == 0) {
          s0.x += v[k].x; s0.y += v[k].y; s0.z += v[k].z; s0.w += v[k].w;
          s1.x = fmaf(v[k].x, v[k].x, s1.x); s1.y = fmaf(v[k].y, v[k].y, s1.y);
          s1.z = fmaf(v[k].z, v[k].z, s1.z); s1.w = fmaf(v[k].w, v[k].w, s1.w);
        } else {
          s0.x += v[k].x; s0.y += v[k].y; s0.z += v[k].z; s0.w += v[k].w;
          s1.x = fmaf(v[k].x, (w[k].x - mu.x) * rs.x, s1.x); s1.y = fmaf(v[k].y, (w[k].y - mu.y) * rs.y, s1.y);
          s1.z = fmaf(v[k].z, (w[k].z - mu.z) * rs.z, s1.z); s1.w = fmaf(v[k].w, (w[k].w - mu.w) * rs.w, s1.w);
        }
      }
    }
  }
  red[0][tid] = s0; red[1][tid] = s1;
  __syncthreads();
  if (lane_r == 0) {
    double d0[4] = {0., 0., 0., 0.}, d1[4] = {0., 0., 0., 0.};
    for (int r = 0; r < rl; ++r) {
      const float4 u0 = red[0][r * C4 + q], u1 = red[1][r * C4 + q];
      d0[0] += u0.x; d0[1] += u0.y; d0[2] += u0.z; d0[3] += u0.w;
      d1[0] += u1.x; d1[1] += u1.y; d1[2] += u1.z; d1[3] += u1.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      atomicAdd(out + 2 * (idx_base + k), d0[k]);
      atomicAdd(out + 2 * (idx_base + k) + 1, d1[k]);
    }
  }
}

static int stats_t_chunk(int B, int F, int T) {
  // enough CTAs to fill 148 SMs a few times over, chunks of >= 32 frames
  long long groups = (long long)B * F;
  int chunks = (int)((148LL * 8 + groups - 1) / groups);
  if (chunks < 1) chunks = 1;
  int tc = (T + chunks - 1) / chunks;
  if (tc < 32) tc = 32;
  return tc;
}

extern "C" int pbsed_channel_stats(const float* x, int B, int F, int T, int C, int per_f,
                                   const int* seq_len, double* stats, int act_dtype, void* stream) {
  if (!x || !stats || B < 1 || F < 1 || T < 1 || C < 1) return PBSED_EINVAL;
  const int tc = stats_t_chunk(B, F, T);
  dim3 grid(B * F, cdiv(T, tc));
  if ((C & 3) == 0 && C <= 1024 && (((uintptr_t)x) & 15) == 0)
    colreduce4_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, nullptr, F, T, C, per_f, seq_len,
                                                                 nullptr, nullptr, stats, tc, act_dtype);
  else if (act_dtype != PBSED_F32)
    return PBSED_EINVAL;                      // bf16 maps: channel counts that are multiples of 4 only
  else
    colreduce_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(x, nullptr, F, T, C, per_f, seq_len,
                                                                nullptr, nullptr, stats, tc);
  return pbsed_after_launch();
}

extern "C" int pbsed_norm_bwd_reduce(const float* g, const float* x, int B, int F, int T, int C,
                                     int per_f, const int* seq_len, const float* save_mean,
                                     const float* save_rstd, double* sums, int act_dtype, void* stream) {
  if (!g || !x || !sums || !save_mean || !save_rstd || B < 1 || F < 1 || T < 1 || C < 1) return PBSED_EINVAL;
  const int tc = stats_t_chunk(B, F, T);
  dim3 grid(B * F, cdiv(T, tc));
  if ((C & 3) == 0 && C <= 1024 && ((((uintptr_t)g) | ((uintptr_t)x) | ((uintptr_t)save_mean) | ((uintptr_t)save_rstd)) & 15) == 0)
    colreduce4_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g, x, F, T, C, per_f, seq_len,
                                                                 save_mean, save_rstd, sums, tc, act_dtype);
  else if (act_dtype != PBSED_F32)
    return PBSED_EINVAL;
  else
    colreduce_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(g, x, F, T, C, per_f, seq_len,
                                                                save_mean, save_rstd, sums, tc);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ finalize
__global__ void norm_finalize_kernel(const double* __restrict__ stats, double count, int nch,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps, float momentum, int training,
                                     float* __restrict__ running_mean, float* __restrict__ running_power,
                                     float* __restrict__ num_tracked, float* __restrict__ scale,
                                     float* __restrict__ shift, float* __restrict__ save_mean,
                                     float* __restrict__ save_rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nch) return;
  if (training && count <= 0.) count = stats[2 * nch];     // device-side count (data-parallel exact statistics)
  const float ga = gamma ? gamma[i] : 1.f;
  const float be = beta ? beta[i] : 0.f;
  float mu, var;
  if (training) {
    const double m = stats[2 * i] / count;
    double v = stats[2 * i + 1] / count - m * m;
    if (v < 0.) v = 0.;
    mu = (float)m; var = (float)v;
    const float power = (float)(v + m * m);
    if (save_mean) { save_mean[i] = mu; save_rstd[i] = rsqrtf(var + eps); }
    if (running_mean) {
      if (momentum >= 0.f) {
        running_mean[i] = momentum * running_mean[i] + (1.f - momentum) * mu;
        running_power[i] = momentum * running_power[i] + (1.f - momentum) * power;
        num_tracked[i] += (float)count;
      } else {
        const float n = (float)count;
        const float tot = num_tracked[i] + n;
        const float rm = running_mean[i] + (mu - running_mean[i]) * n / tot;
        const float rp = running_power[i] + (power - running_power[i]) * n / tot;
        running_mean[i] = rm; running_power[i] = rp; num_tracked[i] = tot;
        // interpolation_factor = 1: normalise with the updated cumulative statistics
        mu = rm;
        var = (rp - rm * rm) * tot / fmaxf(tot - 1.f, 1.f);
      }
    }
  } else {
    mu = running_mean[i];
    var = running_power[i] - mu * mu;
    if (momentum < 0.f) {
      const float n = num_tracked[i];
      var = var * n / fmaxf(n - 1.f, 1.f);
    }
  }
  const float sc = ga / sqrtf(var + eps);
  scale[i] = sc;
  shift[i] = be - mu * sc;
}

extern "C" int pbsed_norm_finalize(const double* stats, double count, int nch, const float* gamma,
                                   const float* beta, float eps, float momentum, int training,
                                   float* running_mean, float* running_power, float* num_tracked,
                                   float* scale, float* shift, float* save_mean, float* save_rstd,
                                   void* stream) {
  if (nch < 1 || !scale || !shift) return PBSED_EINVAL;
  if (training && !stats) return PBSED_EINVAL;
  if (!training && (!running_mean || !running_power)) return PBSED_EINVAL;
  if (running_mean && (!running_power || !num_tracked)) return PBSED_EINVAL;
  norm_finalize_kernel<<<cdiv(nch, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, count, nch, gamma, beta, eps, momentum, training, running_mean, running_power,
      num_tracked, scale, shift, save_mean, save_rstd);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ norm backward apply
__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ x, int F, int T, int C,
                      int per_f, const int* __restrict__ seq_len, const float* __restrict__ mean,
                      const float* __restrict__ rstd, const float* __restrict__ gamma,
                      const double* __restrict__ sums, float inv_n, float* __restrict__ dx,
                      float* __restrict__ dgamma, float* __restrict__ dbeta, long long total,
                      int nch) {
  if (inv_n <= 0.f) inv_n = (float)(1.0 / sums[2 * nch]);   // device-side count
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const long long row = i / C;
    const int t = (int)(row % T);
    const long long gq = row / T;
    const int f = (int)(gq % F), b = (int)(gq / F);
    const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
    float r = 0.f;
    if (t < len_b) {
      const int idx = (per_f ? f * C : 0) + c;
      const float rs = __ldg(rstd + idx);
      const float xh = (__ldg(x + i) - __ldg(mean + idx)) * rs;
      const float s0 = (float)sums[2 * idx] * inv_n, s1 = (float)sums[2 * idx + 1] * inv_n;
      const float ga = gamma ? __ldg(gamma + idx) : 1.f;
      r = ga * rs * (__ldg(g + i) - s0 - xh * s1);
    }
    dx[i] = r;
  }
  if (blockIdx.x == 0 && dgamma) {
    for (int j = threadIdx.x; j < nch; j += blockDim.x) {
      dbeta[j] += (float)sums[2 * j];
      dgamma[j] += (float)sums[2 * j + 1];
    }
  }
}

// dx = A[c] * g + Bx[c] * x + C0[c] on valid rows (0 behind the clip end), with
//   A = gamma*rstd,  Bx = -A*rstd*s1/n,  C0 = -A*s0/n - Bx*mean      (s0 = sum g, s1 = sum g*xhat)
// The grid stride is a multiple of C/4, so a thread keeps its channel quad: the coefficients live in registers
// (recomputed only when the frequency row changes in the per-(f,c) case) and the loop body is 2 loads + 1 store
// per 16 output bytes, four rows in flight per thread.
struct NbaCoef { float4 a, bx, c0; };
__device__ __forceinline__ NbaCoef nba_coef(const float4* __restrict__ mean, const float4* __restrict__ rstd,
                                            const float4* __restrict__ gamma, const double* __restrict__ sums,
                                            float inv_n, int i4) {
  const float4 rs = __ldg(rstd + i4), mu = __ldg(mean + i4);
  const float4 ga = gamma ? __ldg(gamma + i4) : make_float4(1.f, 1.f, 1.f, 1.f);
  const double* sp = sums + 8 * (long long)i4;
  NbaCoef k;
  k.a = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
  k.bx = make_float4(-k.a.x * rs.x * ((float)sp[1] * inv_n), -k.a.y * rs.y * ((float)sp[3] * inv_n),
                     -k.a.z * rs.z * ((float)sp[5] * inv_n), -k.a.w * rs.w * ((float)sp[7] * inv_n));
  k.c0 = make_float4(-k.a.x * ((float)sp[0] * inv_n) - k.bx.x * mu.x, -k.a.y * ((float)sp[2] * inv_n) - k.bx.y * mu.y,
                     -k.a.z * ((float)sp[4] * inv_n) - k.bx.z * mu.z, -k.a.w * ((float)sp[6] * inv_n) - k.bx.w * mu.w);
  return k;
}

// BF = 0: fp32 maps, one channel quad (16 bytes) per thread and row; BF = 1: bf16 maps, TWO quads (8 channels = 16 bytes)
// per thread and row, so both variants move 16 bytes per access and share the row arithmetic (32-bit: rows < 2^31).
template <int BF>
__global__ void __launch_bounds__(256)
norm_bwd_apply4_kernel(const void* __restrict__ g, const void* __restrict__ x, int F, int T, int C4,
                       int per_f, const int* __restrict__ seq_len, const float4* __restrict__ mean,
                       const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                       const double* __restrict__ sums, float inv_n, void* __restrict__ dx,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, long long total4, int nch) {
  constexpr int NQ = BF ? 2 : 1, U = 4;
  if (inv_n <= 0.f) inv_n = (float)(1.0 / sums[2 * nch]);   // device-side count
  const unsigned CV = (unsigned)C4 / NQ;                      // channel vectors per row
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned stride = gridDim.x * blockDim.x;             // multiple of CV (256 % CV == 0)
  const unsigned v = gtid % CV, q0 = v * NQ;
  const unsigned rstep = stride / CV, rows = (unsigned)(total4 / C4);
  int f_cur = -1;
  NbaCoef k[NQ];
  if (!per_f) {
#pragma unroll
    for (int j = 0; j < NQ; ++j) k[j] = nba_coef(mean, rstd, gamma, sums, inv_n, q0 + j);
  }
  for (unsigned row0 = gtid / CV; row0 < rows; row0 += U * rstep) {
    float4 gv[U][NQ], xv[U][NQ];
    bool ok[U];
    int fr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned row = row0 + u * rstep;
      ok[u] = false;
      fr[u] = 0;
      if (row < rows) {
        const unsigned t = row % (unsigned)T, gq = row / (unsigned)T;
        fr[u] = (int)(gq % (unsigned)F);
        const unsigned b = gq / (unsigned)F;
        const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
        ok[u] = (int)t < len_b;
        if (ok[u]) {
          const long long e = 4LL * ((long long)row * C4 + q0);
          if (BF) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(g) + e));
            const uint4 c = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x) + e));
            gv[u][0] = bf16x4_to_float4(make_uint2(a.x, a.y)); gv[u][NQ - 1] = bf16x4_to_float4(make_uint2(a.z, a.w));
            xv[u][0] = bf16x4_to_float4(make_uint2(c.x, c.y)); xv[u][NQ - 1] = bf16x4_to_float4(make_uint2(c.z, c.w));
          } else {
            gv[u][0] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(g) + e));
            xv[u][0] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + e));
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned row = row0 + u * rstep;
      if (row >= rows) break;
      float4 r[NQ];
#pragma unroll
      for (int j = 0; j < NQ; ++j) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok[u]) {
        if (per_f && fr[u] != f_cur) {
          f_cur = fr[u];
#pragma unroll
          for (int j = 0; j < NQ; ++j) k[j] = nba_coef(mean, rstd, gamma, sums, inv_n, f_cur * C4 + q0 + j);
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
          r[j].x = fmaf(k[j].a.x, gv[u][j].x, fmaf(k[j].bx.x, xv[u][j].x, k[j].c0.x));
          r[j].y = fmaf(k[j].a.y, gv[u][j].y, fmaf(k[j].bx.y, xv[u][j].y, k[j].c0.y));
          r[j].z = fmaf(k[j].a.z, gv[u][j].z, fmaf(k[j].bx.z, xv[u][j].z, k[j].c0.z));
          r[j].w = fmaf(k[j].a.w, gv[u][j].w, fmaf(k[j].bx.w, xv[u][j].w, k[j].c0.w));
        }
      }
      const long long e = 4LL * ((long long)row * C4 + q0);
      if (BF) {
        const uint2 lo = float4_to_bf16x4(r[0]), hi = float4_to_bf16x4(r[NQ - 1]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dx) + e) = make_uint4(lo.x, lo.y, hi.x, hi.y);
      } else {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dx) + e) = r[0];
      }
    }
  }
  if (blockIdx.x == 0 && dgamma) {
    for (int j = threadIdx.x; j < nch; j += blockDim.x) {
      dbeta[j] += (float)sums[2 * j];
      dgamma[j] += (float)sums[2 * j + 1];
    }
  }
}

extern "C" int pbsed_norm_bwd_apply(const float* g, const float* x, int B, int F, int T, int C,
                                    int per_f, const int* seq_len, const float* save_mean,
                                    const float* save_rstd, const float* gamma, const double* sums,
                                    double count, float* dx, float* dgamma, float* dbeta,
                                    int act_dtype, void* stream) {
  if (!g || !x || !sums || !dx || !save_mean || !save_rstd) return PBSED_EINVAL;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return PBSED_EINVAL;
  const float inv_count = count > 0. ? (float)(1.0 / count) : 0.f;   // <= 0: count = sums[2*nch] on the device
  const long long total = (long long)B * F * T * C;
  const int nch = per_f ? F * C : C;
  if ((C & (act_dtype == PBSED_BF16 ? 7 : 3)) == 0 && 256 % (C / 4) == 0 && total / C < (1LL << 31) &&
      ((((uintptr_t)g) | ((uintptr_t)x) | ((uintptr_t)dx) | ((uintptr_t)save_mean) |
                        ((uintptr_t)save_rstd) | ((uintptr_t)gamma)) & 15) == 0) {
    const long long total4 = total / 4;
    int blocks4 = (int)((total4 + 255) / 256);
    if (blocks4 > 148 * 8) blocks4 = 148 * 8;
    if (act_dtype == PBSED_BF16)
      norm_bwd_apply4_kernel<1><<<(blocks4 + 1) / 2, 256, 0, (cudaStream_t)stream>>>(
          g, x, F, T, C / 4, per_f, seq_len,
          reinterpret_cast<const float4*>(save_mean), reinterpret_cast<const float4*>(save_rstd),
          reinterpret_cast<const float4*>(gamma), sums, inv_count, dx,
          dgamma, dbeta, total4, nch);
    else
      norm_bwd_apply4_kernel<0><<<blocks4, 256, 0, (cudaStream_t)stream>>>(
          g, x, F, T, C / 4, per_f, seq_len,
          reinterpret_cast<const float4*>(save_mean), reinterpret_cast<const float4*>(save_rstd),
          reinterpret_cast<const float4*>(gamma), sums, inv_count, dx,
          dgamma, dbeta, total4, nch);
    return pbsed_after_launch();
  }
  if (act_dtype != PBSED_F32) return PBSED_EINVAL;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  norm_bwd_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
      g, x, F, T, C, per_f, seq_len, save_mean, save_rstd, gamma, sums, inv_count, dx,
      dgamma, dbeta, total, nch);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ frequency max-pool
__global__ void __launch_bounds__(256)
maxpool_f_kernel(const float* __restrict__ x, int F, long long TC, int pool, float* __restrict__ y,
                 uint8_t* __restrict__ idx, long long total_out) {
  const int Fo = F / pool;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_out; i += stride) {
    const long long tc = i % TC;
    const long long gq = i / TC;
    const int fo = (int)(gq % Fo);
    const long long b = gq / Fo;
    const float* src = x + ((b * F + (long long)fo * pool) * TC + tc);
    float best = __ldg(src);
    int bi = 0;
    for (int k = 1; k < pool; ++k) {
      const float v = __ldg(src + (long long)k * TC);
      if (v > best || (v != v)) { best = v; bi = k; }     // first max wins, NaN propagates (ATen)
    }
    y[i] = best;
    if (idx) idx[i] = (uint8_t)bi;
  }
}

__global__ void __launch_bounds__(256)
maxpool_f_bwd_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ idx, int F,
                     long long TC, int pool, float* __restrict__ dx, long long total_in) {
  const int Fo = F / pool;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_in; i += stride) {
    const long long tc = i % TC;
    const long long gq = i / TC;
    const int f = (int)(gq % F);
    const long long b = gq / F;
    const int fo = f / pool;
    float v = 0.f;
    if (fo < Fo) {
      const long long o = (b * Fo + fo) * TC + tc;
      if ((int)idx[o] == f - fo * pool) v = __ldg(dy + o);
    }
    dx[i] = v;
  }
}

// pool == 2, C % 4 == 0: float4 in, float4 out, uchar4 argmax.  STATS: also accumulate the per-channel
// sum / sum of squares of the POOLED map over valid frames (the next layer's batch statistics), which
// saves that layer a full pass over the map.  The grid stride is a multiple of C/4, so a thread keeps
// the same channel quad for its whole loop and the sums live in registers until the end.
template <bool STATS>
__global__ void __launch_bounds__(256)
maxpool2_f4_kernel(const void* __restrict__ x, int Fo, int TC4, void* __restrict__ y,
                   uchar4* __restrict__ idx, int rows, int C4, const int* __restrict__ seq_len,
                   double* __restrict__ stats, int bf_in, int bf_out) {
  // grid (gx, gy): blockIdx.y strides over the pooled rows gq = b * Fo + fo, blockIdx.x * 256 + tid over the row's
  // T * C / 4 quads -- 32-bit index arithmetic, no division in the inner loop, two row pairs in flight per thread
  // (r02: the flat 64-bit-indexed version ran at 2.3 TB/s, instruction bound)
  __shared__ float red[2][1024];
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), ss = s;
  const int tc0 = blockIdx.x * 256 + threadIdx.x, tstep = gridDim.x * 256;
  auto pool1 = [&](const float4 a, const float4 c, long long o, bool valid) {
    float4 r; uchar4 k;
    k.x = (c.x > a.x || c.x != c.x) ? 1 : 0; r.x = k.x ? c.x : a.x;
    k.y = (c.y > a.y || c.y != c.y) ? 1 : 0; r.y = k.y ? c.y : a.y;
    k.z = (c.z > a.z || c.z != c.z) ? 1 : 0; r.z = k.z ? c.z : a.z;
    k.w = (c.w > a.w || c.w != c.w) ? 1 : 0; r.w = k.w ? c.w : a.w;
    st_act4(y, 4 * o, r, bf_out);
    if (idx) idx[o] = k;
    if (STATS && valid) {
      s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w;
      ss.x = fmaf(r.x, r.x, ss.x); ss.y = fmaf(r.y, r.y, ss.y); ss.z = fmaf(r.z, r.z, ss.z); ss.w = fmaf(r.w, r.w, ss.w);
    }
  };
  for (int gq = blockIdx.y; gq < rows; gq += gridDim.y) {
    const long long in0 = (long long)(2 * gq) * TC4, in1 = in0 + TC4, out0 = (long long)gq * TC4;
    const int tc_valid = (STATS && seq_len) ? min(__ldg(seq_len + gq / Fo), TC4 / C4) * C4 : TC4;   // quads of valid frames
    int tc = tc0;
    for (; tc + tstep < TC4; tc += 2 * tstep) {
      const float4 a0 = ld_act4(x, 4 * (in0 + tc), bf_in), c0 = ld_act4(x, 4 * (in1 + tc), bf_in);
      const float4 a1 = ld_act4(x, 4 * (in0 + tc + tstep), bf_in), c1 = ld_act4(x, 4 * (in1 + tc + tstep), bf_in);
      pool1(a0, c0, out0 + tc, tc < tc_valid);
      pool1(a1, c1, out0 + tc + tstep, tc + tstep < tc_valid);
    }
    if (tc < TC4) pool1(ld_act4(x, 4 * (in0 + tc), bf_in), ld_act4(x, 4 * (in1 + tc), bf_in), out0 + tc, tc < tc_valid);
  }
  if (STATS) {
    const int C = 4 * C4;
    for (int j = threadIdx.x; j < C; j += 256) { red[0][j] = 0.f; red[1][j] = 0.f; }
    __syncthreads();
    const int q = tc0 % C4;                                 // 256 % C4 == 0: the quad is the same for every element of the thread
    atomicAdd(&red[0][4 * q + 0], s.x); atomicAdd(&red[0][4 * q + 1], s.y);
    atomicAdd(&red[0][4 * q + 2], s.z); atomicAdd(&red[0][4 * q + 3], s.w);
    atomicAdd(&red[1][4 * q + 0], ss.x); atomicAdd(&red[1][4 * q + 1], ss.y);
    atomicAdd(&red[1][4 * q + 2], ss.z); atomicAdd(&red[1][4 * q + 3], ss.w);
    __syncthreads();
    for (int j = threadIdx.x; j < C; j += 256) {
      atomicAdd(stats + 2 * j, (double)red[0][j]);
      atomicAdd(stats + 2 * j + 1, (double)red[1][j]);
    }
  }
}
__global__ void __launch_bounds__(256)
maxpool2_f4_bwd_kernel(const void* __restrict__ dy, const uchar4* __restrict__ idx, long long TC4,
                       void* __restrict__ dx, long long total4_out, int bf_in, int bf_out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4_out; i += stride) {
    const long long tc = i % TC4, gq = i / TC4;
    const float4 g = ld_act4(dy, 4 * i, bf_in);
    const uchar4 k = idx[i];
    float4 lo = z, hi = z;
    if (k.x) hi.x = g.x; else lo.x = g.x;
    if (k.y) hi.y = g.y; else lo.y = g.y;
    if (k.z) hi.z = g.z; else lo.z = g.z;
    if (k.w) hi.w = g.w; else lo.w = g.w;
    st_act4(dx, 4 * ((2 * gq) * TC4 + tc), lo, bf_out);
    st_act4(dx, 4 * ((2 * gq + 1) * TC4 + tc), hi, bf_out);
  }
}

static int ew_blocks(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int pbsed_maxpool_f(const float* x, int B, int F, int T, int C, int pool, float* y,
                               uint8_t* idx, const int* seq_len, double* out_stats, int in_dtype, int out_dtype,
                               void* stream) {
  if (!x || !y || pool < 1 || pool > 255 || F / pool < 1) return PBSED_EINVAL;
  const long long total = (long long)B * (F / pool) * T * C;
  if (pool == 2 && (F & 1) == 0 && (C & 3) == 0 && ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)idx)) & 15) == 0) {
    const bool fused = out_stats && C <= 1024 && (256 % (C / 4)) == 0;
    const long long tc4 = (long long)T * C / 4, rows = (long long)B * (F / 2);
    if (tc4 > 0x7fffffffLL / 8 || rows > 0x7fffffffLL) return PBSED_EINVAL;
    int gx = (int)((tc4 + 511) / 512);                     // two quads per thread and pass
    if (gx > 8) gx = 8;
    long long gy = (148 * 8 + gx - 1) / gx;
    if (gy > rows) gy = rows;
    const dim3 grid(gx, (unsigned)gy);
    if (fused)
      maxpool2_f4_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(
          x, F / 2, (int)tc4, y, reinterpret_cast<uchar4*>(idx), (int)rows, C / 4, seq_len, out_stats, in_dtype, out_dtype);
    else
      maxpool2_f4_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(
          x, F / 2, (int)tc4, y, reinterpret_cast<uchar4*>(idx), (int)rows, C / 4, nullptr, nullptr, in_dtype, out_dtype);
    int rc = pbsed_after_launch();
    if (rc || fused || !out_stats) return rc;
    return pbsed_channel_stats(y, B, F / pool, T, C, 0, seq_len, out_stats, out_dtype, stream);
  }
  if (in_dtype != PBSED_F32 || out_dtype != PBSED_F32) return PBSED_EINVAL;
  maxpool_f_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, F, (long long)T * C, pool, y, idx, total);
  int rc = pbsed_after_launch();
  if (rc || !out_stats) return rc;
  return pbsed_channel_stats(y, B, F / pool, T, C, 0, seq_len, out_stats, PBSED_F32, stream);
}

extern "C" int pbsed_maxpool_f_bwd(const float* dy, const uint8_t* idx, int B, int F, int T, int C,
                                   int pool, float* dx, int in_dtype, int out_dtype, void* stream) {
  if (!dy || !idx || !dx || pool < 1 || F / pool < 1) return PBSED_EINVAL;
  const long long total = (long long)B * F * T * C;
  if (pool == 2 && (F & 1) == 0 && (C & 3) == 0 && ((((uintptr_t)dx) | ((uintptr_t)dy) | ((uintptr_t)idx)) & 15) == 0) {
    const long long total4_out = total / 8;
    maxpool2_f4_bwd_kernel<<<ew_blocks(total4_out), 256, 0, (cudaStream_t)stream>>>(
        dy, reinterpret_cast<const uchar4*>(idx), (long long)T * C / 4, dx, total4_out, in_dtype, out_dtype);
    return pbsed_after_launch();
  }
  if (in_dtype != PBSED_F32 || out_dtype != PBSED_F32) return PBSED_EINVAL;
  maxpool_f_bwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(dy, idx, F, (long long)T * C, pool, dx, total);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ tag-condition concat
__global__ void __launch_bounds__(256)
concat_cond_kernel(const float* __restrict__ x, const float* __restrict__ cond, long long FT, int C0,
                   int K, float* __restrict__ out, long long total) {
  const int C = C0 + K;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C);
    const long long row = i / C;
    const long long b = row / FT;
    out[i] = c < C0 ? __ldg(x + row * C0 + c) : __ldg(cond + b * K + (c - C0));
  }
}
__global__ void __launch_bounds__(256)
split_cond_bwd_kernel(const float* __restrict__ dout, int C0, int K, float* __restrict__ dx,
                      long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % C0);
    const long long row = i / C0;
    dx[i] = __ldg(dout + row * (C0 + K) + c);
  }
}

extern "C" int pbsed_concat_cond(const float* x, const float* cond, int B, int F, int T, int C0,
                                 int K, float* out, void* stream) {
  if (!x || !cond || !out || C0 < 1 || K < 1) return PBSED_EINVAL;
  const long long total = (long long)B * F * T * (C0 + K);
  concat_cond_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(x, cond, (long long)F * T, C0, K, out, total);
  return pbsed_after_launch();
}
extern "C" int pbsed_split_cond_bwd(const float* dout, int B, int F, int T, int C0, int K, float* dx,
                                    void* stream) {
  if (!dout || !dx || C0 < 1 || K < 1) return PBSED_EINVAL;
  const long long total = (long long)B * F * T * C0;
  split_cond_bwd_kernel<<<ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(dout, C0, K, dx, total);
  return pbsed_after_launch();
}
