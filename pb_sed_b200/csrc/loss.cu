// loss.cu -- K4: bounded sigmoid, FBCRNN weak/strong forward-backward BCE (value + gradient in one
// pass), BiCRNN frame BCE.
//
// Reference (pb_sed's own code, restated -- not third party):
//   CRNN.sigmoid                       pb_sed/models/weak_label/crnn.py:58-59
//   CRNN.review loss assembly          pb_sed/models/weak_label/crnn.py:117-153
//   compute_weak_fwd_bwd_loss          pb_sed/models/weak_label/crnn.py:180-192
//   compute_strong_fwd_bwd_loss        pb_sed/models/weak_label/crnn.py:194-206
//   strong-label review                pb_sed/models/strong_label/crnn.py:107-112
// BCE follows torch.nn.BCELoss: log clamped at -100, backward (y - t) / max(y (1-y), 1e-12).
#include "common.cuh"

__device__ __forceinline__ float bce_(float y, float t) {
  const float l1 = fmaxf(logf(y), -100.f), l0 = fmaxf(logf(1.f - y), -100.f);
  return -(t * l1 + (1.f - t) * l0);
}
__device__ __forceinline__ float dbce_(float y, float t) {
  return (y - t) / fmaxf((1.f - y) * y, 1e-12f);
}

__device__ __forceinline__ float block_sum(float v, float* sh) {   // 256 threads
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += sh[i];
  return r;
}

// ------------------------------------------------------------------ sigmoid (+ layout change)
__global__ void __launch_bounds__(256)
sigmoid_btk_to_bkt_kernel(const float* __restrict__ z, int T, int K, float min_score,
                          float* __restrict__ y, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int t = (int)(i % T);
    const long long g = i / T;
    const int k = (int)(g % K);
    const long long b = g / K;
    const float s = 1.f / (1.f + expf(-__ldg(z + (b * T + t) * K + k)));
    y[i] = min_score + (1.f - 2.f * min_score) * s;
  }
}
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, int T, int K,
                   float min_score, float* __restrict__ dz, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % K);
    const long long g = i / K;
    const int t = (int)(g % T);
    const long long b = g / T;
    const float s = 1.f / (1.f + expf(-__ldg(z + i)));
    dz[i] = __ldg(dy + (b * K + k) * T + t) * (1.f - 2.f * min_score) * s * (1.f - s);
  }
}

extern "C" int pbsed_sigmoid_btk_to_bkt(const float* z, int B, int T, int K, float min_score,
                                        float* y, void* stream) {
  if (!z || !y || B < 1 || T < 1 || K < 1) return PBSED_EINVAL;
  const long long total = (long long)B * T * K;
  sigmoid_btk_to_bkt_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, T, K, min_score, y, total);
  return pbsed_after_launch();
}
extern "C" int pbsed_sigmoid_bwd(const float* dy, const float* z, int B, int T, int K,
                                 float min_score, float* dz, void* stream) {
  if (!dy || !z || !dz || B < 1 || T < 1 || K < 1) return PBSED_EINVAL;
  const long long total = (long long)B * T * K;
  sigmoid_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, z, T, K, min_score, dz, total);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ FBCRNN loss
// workspace: ws[0] = sum of weights
__global__ void __launch_bounds__(256)
fbcrnn_weights_kernel(const float* __restrict__ weak, const float* __restrict__ cw, int BK, int K,
                      float* __restrict__ ws, float* __restrict__ loss_out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < BK; i += 256) {
    const float wt = weak[i];
    const float mw = (wt < .01f || wt > .99f) ? 1.f : 0.f;
    s += cw ? mw * cw[i % K] : mw;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) { ws[0] = s; loss_out[0] = 0.f; loss_out[1] = s; }
}

constexpr int LOSS_MAXSEG = 16;   // frames per thread; T <= 4096

__global__ void __launch_bounds__(256)
fbcrnn_loss_kernel(const float* __restrict__ y_fwd, const float* __restrict__ y_bwd,
                   const float* __restrict__ weak, const float* __restrict__ boundary,
                   const float* __restrict__ cw, const int* __restrict__ seq_len, int K, int T,
                   float strong_weight, float smooth, const float* __restrict__ ws,
                   float* __restrict__ loss_out, float* __restrict__ dy_fwd,
                   float* __restrict__ dy_bwd) {
  __shared__ float sh[8];
  __shared__ float pmax[256], smax[256];
  const int tid = threadIdx.x;
  const int bk = blockIdx.x, b = bk / K, k = bk % K;
  const int len = seq_len ? min(__ldg(seq_len + b), T) : T;
  const long long base = (long long)bk * T;
  const int L = (T + 255) / 256;
  const int ts = tid * L, te = min(ts + L, T);

  float wt = weak[bk];
  const float mw = (wt < .01f || wt > .99f) ? 1.f : 0.f;
  wt *= mw;
  float tw = wt;
  if (smooth > 0.f) tw = fminf(fmaxf(tw, smooth), 1.f - smooth);
  const bool strong = strong_weight > 0.f && boundary != nullptr;

  // boundary mask statistics + per-segment running maxima
  float bt[LOSS_MAXSEG];
  float cnt = 0.f, lmax = -INFINITY;
  if (strong) {
    for (int i = 0; i < L; ++i) {
      const int t = ts + i;
      float v = 0.f;
      if (t < T) {
        v = __ldg(boundary + base + t);
        cnt += (v > .99f || v < .01f) ? 1.f : 0.f;
        float vs = v;
        if (smooth > 0.f) vs = fminf(fmaxf(vs, smooth), 1.f - smooth);
        lmax = fmaxf(lmax, vs);
      }
      bt[i] = v;
    }
  }
  float excl_fwd = -INFINITY, excl_bwd = -INFINITY;   // max over segments before / after mine
  bool full_mask = false;
  if (strong) {
    cnt = block_sum(cnt, sh);
    full_mask = (cnt / (float)T > .999f) && (wt > .99f);
    pmax[tid] = lmax; smax[tid] = lmax;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {          // Hillis-Steele inclusive prefix / suffix max
      const float a = tid >= off ? pmax[tid - off] : -INFINITY;
      const float c = tid + off < 256 ? smax[tid + off] : -INFINITY;
      __syncthreads();
      pmax[tid] = fmaxf(pmax[tid], a); smax[tid] = fmaxf(smax[tid], c);
      __syncthreads();
    }
    excl_fwd = tid > 0 ? pmax[tid - 1] : -INFINITY;
    excl_bwd = tid < 255 ? smax[tid + 1] : -INFINITY;
  }

  const float wsum = ws[0];
  const float wbk = cw ? mw * cw[k] : mw;
  const float denom_t = seq_len ? ((float)len + 1e-6f) : (float)T;   // padertorch Mean(axis=-1)
  const float gf = wbk / (wsum * denom_t);

  // suffix maxima inside my segment (for the backward targets)
  float sufmax[LOSS_MAXSEG];
  if (strong) {
    float r = excl_bwd;
    for (int i = L - 1; i >= 0; --i) {
      const int t = ts + i;
      if (t < T) {
        float vs = bt[i];
        if (smooth > 0.f) vs = fminf(fmaxf(vs, smooth), 1.f - smooth);
        r = fmaxf(r, vs);
      }
      sufmax[i] = r;
    }
  }
  const float y_last = (!y_bwd && len > 0) ? __ldg(y_fwd + base + len - 1) : 0.f;

  float lsum = 0.f, coef_last = 0.f, run = excl_fwd;
  for (int i = 0; i < L; ++i) {
    const int t = ts + i;
    if (t >= T) break;
    float sw = 0.f, t_f = 0.f, t_b = 0.f;
    if (strong) {
      float vs = bt[i];
      if (smooth > 0.f) vs = fminf(fmaxf(vs, smooth), 1.f - smooth);
      run = fmaxf(run, vs);
      t_f = run; t_b = sufmax[i];
      const float mb = (bt[i] > .99f || bt[i] < .01f) ? 1.f : 0.f;
      sw = full_mask ? mb * strong_weight : 0.f;
    }
    float gfw = 0.f, gbw = 0.f;
    if (t < len) {
      const float yf = __ldg(y_fwd + base + t);
      const float yb = y_bwd ? __ldg(y_bwd + base + t) : 0.f;
      float l = 0.f;
      // weak part
      const float cwk = (1.f - sw) * mw;
      if (y_bwd) {
        const float ym = fmaxf(yf, yb);
        l += cwk * bce_(ym, tw);
        const float d = cwk * dbce_(ym, tw);
        if (yf > yb) gfw += d; else if (yb > yf) gbw += d; else { gfw += .5f * d; gbw += .5f * d; }
      } else {
        l += cwk * bce_(y_last, tw);
        coef_last += cwk;
      }
      if (sw > 0.f) {
        if (y_bwd) {
          l += sw * .5f * (bce_(yf, t_f) + bce_(yb, t_b));
          gfw += sw * .5f * dbce_(yf, t_f);
          gbw += sw * .5f * dbce_(yb, t_b);
        } else {
          l += sw * bce_(yf, t_f);
          gfw += sw * dbce_(yf, t_f);
        }
      }
      lsum += l;
    }
    if (dy_fwd) dy_fwd[base + t] = gfw * gf;
    if (dy_bwd) dy_bwd[base + t] = gbw * gf;
  }
  lsum = block_sum(lsum, sh);
  if (!y_bwd) {
    coef_last = block_sum(coef_last, sh);
    if (tid == 0 && dy_fwd && len > 0) dy_fwd[base + len - 1] += coef_last * dbce_(y_last, tw) * gf;
  }
  if (tid == 0 && wbk != 0.f) atomicAdd(loss_out, lsum / denom_t * wbk / wsum);
}

extern "C" int pbsed_fbcrnn_loss(const float* y_fwd, const float* y_bwd, const float* weak,
                                 const float* boundary, const float* class_weights,
                                 const int* seq_len, int B, int K, int T, float strong_weight,
                                 float label_smoothing, float* loss_out, float* dy_fwd,
                                 float* dy_bwd, float* workspace, void* stream) {
  if (!y_fwd || !weak || !loss_out || !workspace || B < 1 || K < 1 || T < 1) return PBSED_EINVAL;
  if (T > 256 * LOSS_MAXSEG) return PBSED_EINVAL;
  if (strong_weight > 0.f && !boundary) return PBSED_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  fbcrnn_weights_kernel<<<1, 256, 0, st>>>(weak, class_weights, B * K, K, workspace, loss_out);
  int rc = pbsed_after_launch();
  if (rc) return rc;
  fbcrnn_loss_kernel<<<B * K, 256, 0, st>>>(y_fwd, y_bwd, weak, boundary, class_weights, seq_len, K,
                                            T, strong_weight, label_smoothing, workspace, loss_out,
                                            dy_fwd, dy_bwd);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ BiCRNN loss
__global__ void __launch_bounds__(256)
bicrnn_mask_count_kernel(const float* __restrict__ strong, long long total, float* __restrict__ ws,
                         float* __restrict__ loss_out) {
  __shared__ float sh[8];
  float s = 0.f;
  for (long long i = threadIdx.x; i < total; i += 256) {
    const float v = __ldg(strong + i);
    s += (v > .99f || v < .01f) ? 1.f : 0.f;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0) { ws[0] = s; loss_out[0] = 0.f; loss_out[1] = s; }
}
__global__ void __launch_bounds__(256)
bicrnn_loss_kernel(const float* __restrict__ y, const float* __restrict__ strong,
                   const int* __restrict__ seq_len, int K, int T, const float* __restrict__ ws,
                   float* __restrict__ loss_out, float* __restrict__ dy) {
  __shared__ float sh[8];
  const int bk = blockIdx.x, b = bk / K;
  const int len = seq_len ? min(__ldg(seq_len + b), T) : T;
  const long long base = (long long)bk * T;
  const float inv = 1.f / ws[0];
  float s = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) {
    float g = 0.f;
    if (t < len) {
      const float st = __ldg(strong + base + t), yy = __ldg(y + base + t);
      if (st > .99f || st < .01f) { s += bce_(yy, st); g = dbce_(yy, st) * inv; }
    }
    if (dy) dy[base + t] = g;
  }
  s = block_sum(s, sh);
  if (threadIdx.x == 0 && s != 0.f) atomicAdd(loss_out, s * inv);
}

extern "C" int pbsed_bicrnn_loss(const float* y, const float* strong, const int* seq_len, int B,
                                 int K, int T, float* loss_out, float* dy, float* workspace,
                                 void* stream) {
  if (!y || !strong || !loss_out || !workspace || B < 1 || K < 1 || T < 1) return PBSED_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  bicrnn_mask_count_kernel<<<1, 256, 0, st>>>(strong, (long long)B * K * T, workspace, loss_out);
  int rc = pbsed_after_launch();
  if (rc) return rc;
  bicrnn_loss_kernel<<<B * K, 256, 0, st>>>(y, strong, seq_len, K, T, workspace, loss_out, dy);
  return pbsed_after_launch();
}
