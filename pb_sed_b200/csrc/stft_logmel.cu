// stft_logmel.cu -- K1: framing + window + real FFT + |.|^2 + mel filterbank + log, fused.
//
// Reference: paderbox.transform.module_stft.stft as configured at
// pb_sed/data_preparation/provider.py:315-323 (shift 320, window_length 960, size 1024,
// fading='half', pad=True; periodic Blackman window) and invoked on the CPU at
// pb_sed/data_preparation/transform.py:53; followed by the MelTransform part of
// NormalizedLogMelExtractor (call site pb_sed/models/weak_label/crnn.py:86-90):
// power -> unit-sum HTK-mel triangles -> log(. + 1e-18) -> (B, n_mels, T).
//
// One CTA (256 threads) produces FR = 32 consecutive frames of one clip.  Frames are
// transformed two at a time as ONE complex FFT (frame A in the real part, frame B in the
// imaginary part; the two real spectra are separated with the conjugate-symmetry identity),
// radix-2 Stockham autosort in shared memory.  The triangular filters are applied from a
// sparse (lo, hi, weights) table, and the 32 x n_mels result tile is staged in shared memory
// so that the (B, n_mels, T) store is 128-byte coalesced along t.  The per-band sum / sum of
// squares needed by the cumulative running normalisation is accumulated in the same pass.
#include "common.cuh"

constexpr int FR = 32;   // frames per CTA

template <bool FROM_AUDIO>
__global__ void __launch_bounds__(256)
logmel_kernel(const float* __restrict__ src, int S, int shift, int window_length, int N,
              int pad_front, int T, int n_bins, const float* __restrict__ window,
              const int* __restrict__ fb_lo, const int* __restrict__ fb_hi,
              const float* __restrict__ fb_w, int fb_stride, int n_mels, int fb_per_clip,
              const int* __restrict__ frame_start,
              const int* __restrict__ seq_len, float* __restrict__ logmel,
              double* __restrict__ stats) {
  extern __shared__ __align__(16) float smem[];
  // layout: bufA[N] float2 | bufB[N] float2 | tw[N/2] float2 | P[2][n_bins] | out[n_mels][FR+1] | win[window_length]
  float2* bufA = reinterpret_cast<float2*>(smem);
  float2* bufB = bufA + (FROM_AUDIO ? N : 0);
  float2* tw = bufB + (FROM_AUDIO ? N : 0);
  float* P = reinterpret_cast<float*>(tw + (FROM_AUDIO ? N / 2 : 0));
  float* outb = P + 2 * n_bins;
  float* win = outb + n_mels * (FR + 1);

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
  if (fb_per_clip) {                       // per-example warped filterbank (train-time MelWarping)
    fb_lo += (long long)b * n_mels; fb_hi += (long long)b * n_mels;
    fb_w += (long long)b * n_mels * fb_stride;
  }

  if (FROM_AUDIO) {
    for (int k = tid; k < N / 2; k += 256) {
      float sn, cs;
      sincospif(-2.f * (float)k / (float)N, &sn, &cs);
      tw[k] = make_float2(cs, sn);
    }
    for (int n = tid; n < window_length; n += 256) win[n] = __ldg(window + n);
  }
  __syncthreads();

  for (int pr = 0; pr < FR / 2; ++pr) {
    const int ta = t0 + 2 * pr, tb = ta + 1;
    if (ta >= T) break;                                    // uniform
    if (FROM_AUDIO) {
      const float* a = src + (long long)b * S;
      for (int n = tid; n < N; n += 256) {
        float xa = 0.f, xb = 0.f;
        if (n < window_length) {
          const float w = win[n];
          // frame onsets: uniform hop, or the piecewise-linear grid of TimeWarpedSTFT (transform.py:36-45)
          const int ia = (frame_start ? __ldg(frame_start + (long long)b * T + ta) : ta * shift) + n - pad_front;
          const int ib = (frame_start ? (tb < T ? __ldg(frame_start + (long long)b * T + tb) : 0) : tb * shift) + n - pad_front;
          if (ia >= 0 && ia < S) xa = __ldg(a + ia) * w;
          if (tb < T && ib >= 0 && ib < S) xb = __ldg(a + ib) * w;
        }
        bufA[n] = make_float2(xa, xb);
      }
      __syncthreads();
      float2* x = bufA;
      float2* y = bufB;
      for (int n = N, s = 1; n > 1; n >>= 1, s <<= 1) {
        const int m = n >> 1;
        for (int i = tid; i < N / 2; i += 256) {
          const int p = i / s, q = i - p * s;
          const float2 w = tw[p * s];
          const float2 u = x[q + s * p], v = x[q + s * (p + m)];
          const float2 d = make_float2(u.x - v.x, u.y - v.y);
          y[q + s * (2 * p)] = make_float2(u.x + v.x, u.y + v.y);
          y[q + s * (2 * p + 1)] = make_float2(d.x * w.x - d.y * w.y, d.x * w.y + d.y * w.x);
        }
        __syncthreads();
        float2* tmp = x; x = y; y = tmp;
      }
      // separate the two real spectra:  Xa = (Z[k] + conj Z[N-k]) / 2,  Xb = (Z[k] - conj Z[N-k]) / (2i)
      for (int k = tid; k < n_bins; k += 256) {
        const float2 zk = x[k], zn = x[(N - k) & (N - 1)];
        const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
        const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
        P[k] = ar * ar + ai * ai;
        P[n_bins + k] = br * br + bi * bi;
      }
    } else {
      // src = stft (B, T, n_bins, 2)
      const float2* sa = reinterpret_cast<const float2*>(src) + ((long long)b * T + ta) * n_bins;
      for (int k = tid; k < n_bins; k += 256) {
        const float2 za = __ldg(sa + k);
        P[k] = za.x * za.x + za.y * za.y;
        float pb = 0.f;
        if (tb < T) { const float2 zb = __ldg(sa + n_bins + k); pb = zb.x * zb.x + zb.y * zb.y; }
        P[n_bins + k] = pb;
      }
    }
    __syncthreads();
    for (int i = tid; i < 2 * n_mels; i += 256) {
      const int which = i / n_mels, m = i - which * n_mels;
      const int lo = __ldg(fb_lo + m), hi = __ldg(fb_hi + m);
      const float* w = fb_w + (long long)m * fb_stride;
      const float* pp = P + which * n_bins;
      float acc = 0.f;
      for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(w + (k - lo)), pp[k], acc);
      outb[m * (FR + 1) + 2 * pr + which] = logf(acc + 1e-18f);
    }
    __syncthreads();
  }
  __syncthreads();
  const int nfr = min(FR, T - t0);
  for (int i = tid; i < n_mels * FR; i += 256) {
    const int m = i / FR, f = i - m * FR;
    if (f < nfr) logmel[((long long)b * n_mels + m) * T + t0 + f] = outb[m * (FR + 1) + f];
  }
  if (stats) {
    const int nvalid = min(nfr, len_b - t0);
    for (int m = tid; m < n_mels; m += 256) {
      float s = 0.f, ss = 0.f;
      for (int f = 0; f < nvalid; ++f) { const float v = outb[m * (FR + 1) + f]; s += v; ss = fmaf(v, v, ss); }
      if (nvalid > 0) { atomicAdd(stats + 2 * m, (double)s); atomicAdd(stats + 2 * m + 1, (double)ss); }
    }
  }
}

// ------------------------------------------------------------------ N = 1024 fast path (the reference's STFT size)
// One WARP transforms one frame pair: 1024 = 32 x 32 Cooley-Tukey with both 32-point stages fully
// unrolled in registers (64 data registers per thread), one padded shared-memory transpose between
// them, exact twiddles from a shared table.  No block-wide barrier inside the transform; 8 warps per
// CTA cover 32 consecutive frames so the (B, n_mels, T) stores are 128-byte rows.
constexpr int FW_WARPS = 8;
constexpr int FW_TP = 33;                // transpose pitch in float2

__device__ __forceinline__ constexpr int brev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// in place, natural order in, bit-reversed order out (X[k] = v[brev5(k)]); radix-2 DIF, constants folded
__device__ __forceinline__ void fft32_regs(float2 (&v)[32]) {
  constexpr float C[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f,
                           0.f, -0.19509032201612825f, -0.38268343236508978f, -0.55557023301960218f,
                           -0.70710678118654757f, -0.83146961230254524f, -0.92387953251128674f, -0.98078528040323043f};
  constexpr float S[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                           0.70710678118654757f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
                           1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                           0.70710678118654757f, 0.55557023301960218f, 0.38268343236508978f, 0.19509032201612825f};
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int half = 16 >> s;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int grp = i / half, j = i % half;
      const int a = grp * 2 * half + j, b = a + half;
      const int tw = j * (16 / half);
      const float2 u = v[a], w = v[b];
      v[a] = make_float2(u.x + w.x, u.y + w.y);
      const float dx = u.x - w.x, dy = u.y - w.y;
      if (tw == 0) v[b] = make_float2(dx, dy);
      else if (tw == 8) v[b] = make_float2(dy, -dx);               // * (-i)
      else v[b] = make_float2(dx * C[tw] + dy * S[tw], dy * C[tw] - dx * S[tw]);   // * (C - iS)
    }
  }
}

__global__ void __launch_bounds__(FW_WARPS * 32)
logmel1024_kernel(const float* __restrict__ src, int S, int shift, int window_length, int pad_front, int T,
                  const float* __restrict__ window, const int* __restrict__ fb_lo,
                  const int* __restrict__ fb_hi, const float* __restrict__ fb_w, int fb_stride, int n_mels,
                  int fb_per_clip, const int* __restrict__ frame_start, const int* __restrict__ seq_len,
                  float* __restrict__ logmel, double* __restrict__ stats) {
  constexpr int N = 1024, NB = 513;
  extern __shared__ __align__(16) float smem[];
  float2* twid = reinterpret_cast<float2*>(smem);                       // [1024]  W_1024^i
  float* win = smem + 2 * N;                                            // [1024]
  float* outb = win + N;                                                // [n_mels][FR+1]
  float2* tbuf = reinterpret_cast<float2*>(outb + n_mels * (FR + 1) + ((n_mels * (FR + 1)) & 1));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float2* tb_w = tbuf + warp * (32 * FW_TP);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * FR;
  const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
  if (fb_per_clip) {
    fb_lo += (long long)b * n_mels; fb_hi += (long long)b * n_mels;
    fb_w += (long long)b * n_mels * fb_stride;
  }
  for (int i = tid; i < N; i += FW_WARPS * 32) {
    float sn, cs;
    sincospif(-2.f * (float)i / (float)N, &sn, &cs);
    twid[i] = make_float2(cs, sn);
    win[i] = i < window_length ? __ldg(window + i) : 0.f;
  }
  __syncthreads();
  const float* a = src + (long long)b * S;
#pragma unroll 1
  for (int pr = warp; pr < FR / 2; pr += FW_WARPS) {
    const int ta = t0 + 2 * pr, tb = ta + 1;
    if (ta >= T) break;                                                 // uniform within the warp
    const int sa = (frame_start ? __ldg(frame_start + (long long)b * T + ta) : ta * shift) - pad_front;
    const int sb = tb < T ? (frame_start ? __ldg(frame_start + (long long)b * T + tb) : tb * shift) - pad_front : -(1 << 30);
    float2 v[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) {
      const int n = 32 * n1 + lane;
      const float w = win[n];
      const int ia = sa + n, ib = sb + n;
      const float xa = (ia >= 0 && ia < S) ? __ldg(a + ia) : 0.f;
      const float xb = (ib >= 0 && ib < S) ? __ldg(a + ib) : 0.f;
      v[n1] = make_float2(xa * w, xb * w);
    }
    fft32_regs(v);
#pragma unroll
    for (int k1 = 0; k1 < 32; ++k1) {
      const float2 y = v[brev5(k1)], w = twid[(lane * k1) & (N - 1)];
      tb_w[k1 * FW_TP + lane] = make_float2(y.x * w.x - y.y * w.y, y.x * w.y + y.y * w.x);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) v[n2] = tb_w[lane * FW_TP + n2];
    __syncwarp();
    fft32_regs(v);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) tb_w[lane + 32 * k2] = v[brev5(k2)];   // X[k1 + 32 k2], k1 = lane
    __syncwarp();
    // separate the two real spectra and take |.|^2 (bins k = lane + 32 j <= 512)
    float pa[17], pb[17];
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int k = lane + 32 * j;
      pa[j] = pb[j] = 0.f;
      if (k < NB) {
        const float2 zk = tb_w[k], zn = tb_w[(N - k) & (N - 1)];
        const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
        const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
        pa[j] = ar * ar + ai * ai;
        pb[j] = br * br + bi * bi;
      }
    }
    __syncwarp();
    float* P = reinterpret_cast<float*>(tb_w);
#pragma unroll
    for (int j = 0; j < 17; ++j) {
      const int k = lane + 32 * j;
      if (k < NB) { P[k] = pa[j]; P[NB + k] = pb[j]; }
    }
    __syncwarp();
    for (int i = lane; i < 2 * n_mels; i += 32) {
      const int which = i / n_mels, m = i - which * n_mels;
      const int lo = __ldg(fb_lo + m), hi = __ldg(fb_hi + m);
      const float* w = fb_w + (long long)m * fb_stride;
      const float* pp = P + which * NB;
      float acc = 0.f;
      for (int k = lo; k < hi; ++k) acc = fmaf(__ldg(w + (k - lo)), pp[k], acc);
      outb[m * (FR + 1) + 2 * pr + which] = logf(acc + 1e-18f);
    }
    __syncwarp();
  }
  __syncthreads();
  const int nfr = min(FR, T - t0);
  for (int i = tid; i < n_mels * FR; i += FW_WARPS * 32) {
    const int m = i / FR, f = i - m * FR;
    if (f < nfr) logmel[((long long)b * n_mels + m) * T + t0 + f] = outb[m * (FR + 1) + f];
  }
  if (stats) {
    const int nvalid = min(nfr, len_b - t0);
    for (int m = tid; m < n_mels; m += FW_WARPS * 32) {
      float s = 0.f, ss = 0.f;
      for (int f = 0; f < nvalid; ++f) { const float v = outb[m * (FR + 1) + f]; s += v; ss = fmaf(v, v, ss); }
      if (nvalid > 0) { atomicAdd(stats + 2 * m, (double)s); atomicAdd(stats + 2 * m + 1, (double)ss); }
    }
  }
}

static size_t logmel1024_smem(int n_mels) {
  size_t fl = 2 * 1024 + 1024 + (size_t)n_mels * (FR + 1);
  fl += fl & 1;
  fl += (size_t)FW_WARPS * 32 * FW_TP * 2;
  return fl * sizeof(float);
}

static size_t logmel_smem(bool from_audio, int N, int n_bins, int n_mels, int window_length) {
  size_t fl = 0;
  if (from_audio) fl += 2 * (size_t)N * 2 + (size_t)N;   // bufA, bufB (float2), tw (N/2 float2)
  fl += 2 * (size_t)n_bins + (size_t)n_mels * (FR + 1);
  if (from_audio) fl += window_length;
  return fl * sizeof(float);
}

extern "C" int pbsed_stft_logmel(const float* audio, int B, int S, int shift, int window_length,
                                 int fft_size, int pad_front, int T, const float* window,
                                 const int* fbank_lo, const int* fbank_hi, const float* fbank_w,
                                 int fbank_stride, int n_mels, int fbank_per_clip, const int* frame_start,
                                 const int* seq_len, float* logmel, double* stats, void* stream) {
  if (!audio || !window || !fbank_lo || !fbank_hi || !fbank_w || !logmel) return PBSED_EINVAL;
  if (B < 1 || S < 1 || T < 1 || shift < 1 || n_mels < 1 || B > 65535) return PBSED_EINVAL;
  if (fft_size < 8 || fft_size > 4096 || (fft_size & (fft_size - 1))) return PBSED_EINVAL;
  if (window_length < 1 || window_length > fft_size) return PBSED_EINVAL;
  const int n_bins = fft_size / 2 + 1;
  if (fft_size == 1024 && logmel1024_smem(n_mels) <= 200 * 1024) {
    const size_t sm = logmel1024_smem(n_mels);
    cudaError_t e = cudaFuncSetAttribute(logmel1024_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(cdiv(T, FR), B);
    logmel1024_kernel<<<grid, FW_WARPS * 32, sm, (cudaStream_t)stream>>>(
        audio, S, shift, window_length, pad_front, T, window, fbank_lo, fbank_hi, fbank_w, fbank_stride,
        n_mels, fbank_per_clip, frame_start, seq_len, logmel, stats);
    return pbsed_after_launch();
  }
  const size_t smem = logmel_smem(true, fft_size, n_bins, n_mels, window_length);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  cudaError_t e = cudaFuncSetAttribute(logmel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(cdiv(T, FR), B);
  logmel_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(
      audio, S, shift, window_length, fft_size, pad_front, T, n_bins, window, fbank_lo, fbank_hi,
      fbank_w, fbank_stride, n_mels, fbank_per_clip, frame_start, seq_len, logmel, stats);
  return pbsed_after_launch();
}

extern "C" int pbsed_spec_logmel(const float* stft, int B, int T, int n_bins, const int* fbank_lo,
                                 const int* fbank_hi, const float* fbank_w, int fbank_stride,
                                 int n_mels, int fbank_per_clip, const int* seq_len, float* logmel,
                                 double* stats, void* stream) {
  if (!stft || !fbank_lo || !fbank_hi || !fbank_w || !logmel) return PBSED_EINVAL;
  if (B < 1 || T < 1 || n_bins < 1 || n_mels < 1 || B > 65535) return PBSED_EINVAL;
  const size_t smem = logmel_smem(false, 0, n_bins, n_mels, 0);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  cudaError_t e = cudaFuncSetAttribute(logmel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(cdiv(T, FR), B);
  logmel_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(
      stft, 0, 0, 0, 0, 0, T, n_bins, nullptr, fbank_lo, fbank_hi, fbank_w, fbank_stride, n_mels,
      fbank_per_clip, nullptr, seq_len, logmel, stats);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ normalise + clamp + mask (+ train-time augmentation)
// order (SURVEY App. A [R]): running-stat normalisation -> clamp -> time masks -> frequency masks ->
// additive Gaussian noise; frames behind seq_len stay exactly 0.
__global__ void __launch_bounds__(256)
logmel_normalize_kernel(float* __restrict__ x, int F, int T, const float* __restrict__ scale,
                        const float* __restrict__ shift, float clampv, const int* __restrict__ seq_len,
                        const int* __restrict__ tmask, int n_tmask, const int* __restrict__ fmask, int n_fmask,
                        const float* __restrict__ noise, const float* __restrict__ noise_scale,
                        long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int t = (int)(i % T);
    const long long g = i / T;
    const int f = (int)(g % F), b = (int)(g / F);
    const int len_b = seq_len ? min(__ldg(seq_len + b), T) : T;
    float v = 0.f;
    if (t < len_b) {
      v = fmaf(x[i], __ldg(scale + f), __ldg(shift + f));
      if (clampv > 0.f) v = fminf(fmaxf(v, -clampv), clampv);
      for (int j = 0; j < n_tmask; ++j) {
        const int on = __ldg(tmask + ((long long)b * n_tmask + j) * 2), w = __ldg(tmask + ((long long)b * n_tmask + j) * 2 + 1);
        if (t >= on && t < on + w) v = 0.f;
      }
      for (int j = 0; j < n_fmask; ++j) {
        const int on = __ldg(fmask + ((long long)b * n_fmask + j) * 2), w = __ldg(fmask + ((long long)b * n_fmask + j) * 2 + 1);
        if (f >= on && f < on + w) v = 0.f;
      }
      if (noise) v = fmaf(__ldg(noise_scale + b), __ldg(noise + i), v);
    }
    x[i] = v;
  }
}

extern "C" int pbsed_logmel_normalize(float* x, int B, int F, int T, const float* scale,
                                      const float* shift, float clampv, const int* seq_len,
                                      const int* time_masks, int n_time_masks, const int* freq_masks,
                                      int n_freq_masks, const float* noise, const float* noise_scale,
                                      void* stream) {
  if (!x || !scale || !shift || B < 1 || F < 1 || T < 1) return PBSED_EINVAL;
  if ((n_time_masks > 0 && !time_masks) || (n_freq_masks > 0 && !freq_masks) || n_time_masks < 0 || n_freq_masks < 0)
    return PBSED_EINVAL;
  if (noise && !noise_scale) return PBSED_EINVAL;
  const long long total = (long long)B * F * T;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  logmel_normalize_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      x, F, T, scale, shift, clampv, seq_len, time_masks, n_time_masks, freq_masks, n_freq_masks, noise,
      noise_scale, total);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ per-example warped mel filterbanks
// paderbox MelWarping [R] (configured at pb_sed/experiments/weak_label_crnn/training.py:195-208): the
// n_mels+2 equally mel-spaced edge frequencies are warped per example, piecewise linearly in the mel
// domain (vocal-tract-length-perturbation shape: slope alpha up to the break point, then a straight
// line to (m_hi, m_hi)), and the unit-sum triangles are rebuilt on the warped edges.
//   m_b = m_hi / (1 + ratio);  knee = m_b * min(alpha, 1) / alpha
//   m' = alpha * m                                                    (m <= knee)
//   m' = m_hi - (m_hi - m_b * min(alpha,1)) / (m_hi - knee) * (m_hi - m)   (m > knee)
// One CTA per clip, one thread per filter; float64 like the host-side table.
__device__ __forceinline__ double warp_mel(double m, double alpha, double m_b, double m_hi) {
  const double mn = fmin(alpha, 1.0);
  const double knee = m_b * mn / alpha;
  if (m <= knee) return alpha * m;
  return m_hi - (m_hi - m_b * mn) / (m_hi - knee) * (m_hi - m);
}

__global__ void make_warped_fbank_kernel(const float* __restrict__ alpha, const float* __restrict__ ratio,
                                         int n_mels, int n_bins, double mel_lo, double mel_hi,
                                         double mel_warp_hi, double bins_per_hz, int* __restrict__ lo_out,
                                         int* __restrict__ hi_out, float* __restrict__ w_out, int stride) {
  const int b = blockIdx.x;
  const double a = (double)alpha[b];
  const double m_b = mel_warp_hi / (1.0 + (double)ratio[b]);
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    double e[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double mel = mel_lo + (mel_hi - mel_lo) * (double)(m + j) / (double)(n_mels + 1);
      const double wm = warp_mel(mel, a, m_b, mel_warp_hi);
      e[j] = 700.0 * (pow(10.0, wm / 2595.0) - 1.0) * bins_per_hz;
    }
    int lo = (int)floor(e[0]) + 1, hi = (int)ceil(e[2]);
    lo = max(lo, 0); hi = min(hi, n_bins);
    if (hi - lo > stride) hi = lo + stride;
    if (hi < lo) hi = lo;
    float* w = w_out + ((long long)b * n_mels + m) * stride;
    double sum = 0.0;
    for (int k = lo; k < hi; ++k) {
      const double tri = fmax(fmin(((double)k - e[0]) / (e[1] - e[0]), (e[2] - (double)k) / (e[2] - e[1])), 0.0);
      sum += tri;
    }
    const double inv = 1.0 / (sum + 1e-6);
    for (int k = lo; k < hi; ++k) {
      const double tri = fmax(fmin(((double)k - e[0]) / (e[1] - e[0]), (e[2] - (double)k) / (e[2] - e[1])), 0.0);
      w[k - lo] = (float)(tri * inv);
    }
    for (int k = hi - lo; k < stride; ++k) w[k] = 0.f;
    lo_out[(long long)b * n_mels + m] = lo;
    hi_out[(long long)b * n_mels + m] = hi;
  }
}

extern "C" int pbsed_make_warped_fbank(const float* alpha, const float* ratio, int B, int n_mels, int n_bins,
                                       double mel_lo, double mel_hi, double mel_warp_hi, double bins_per_hz,
                                       int* fbank_lo, int* fbank_hi, float* fbank_w, int fbank_stride,
                                       void* stream) {
  if (!alpha || !ratio || !fbank_lo || !fbank_hi || !fbank_w) return PBSED_EINVAL;
  if (B < 1 || n_mels < 1 || n_bins < 2 || fbank_stride < 1 || !(mel_hi > mel_lo) || !(mel_warp_hi > 0.)) return PBSED_EINVAL;
  make_warped_fbank_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(alpha, ratio, n_mels, n_bins, mel_lo, mel_hi,
                                                               mel_warp_hi, bins_per_hz, fbank_lo, fbank_hi,
                                                               fbank_w, fbank_stride);
  return pbsed_after_launch();
}
