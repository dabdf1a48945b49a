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
              const float* __restrict__ fb_w, int fb_stride, int n_mels,
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
          const int ia = ta * shift + n - pad_front, ib = ia + shift;
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
                                 int fbank_stride, int n_mels, const int* seq_len, float* logmel,
                                 double* stats, void* stream) {
  if (!audio || !window || !fbank_lo || !fbank_hi || !fbank_w || !logmel) return PBSED_EINVAL;
  if (B < 1 || S < 1 || T < 1 || shift < 1 || n_mels < 1 || B > 65535) return PBSED_EINVAL;
  if (fft_size < 8 || fft_size > 4096 || (fft_size & (fft_size - 1))) return PBSED_EINVAL;
  if (window_length < 1 || window_length > fft_size) return PBSED_EINVAL;
  const int n_bins = fft_size / 2 + 1;
  const size_t smem = logmel_smem(true, fft_size, n_bins, n_mels, window_length);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  cudaError_t e = cudaFuncSetAttribute(logmel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(cdiv(T, FR), B);
  logmel_kernel<true><<<grid, 256, smem, (cudaStream_t)stream>>>(
      audio, S, shift, window_length, fft_size, pad_front, T, n_bins, window, fbank_lo, fbank_hi,
      fbank_w, fbank_stride, n_mels, seq_len, logmel, stats);
  return pbsed_after_launch();
}

extern "C" int pbsed_spec_logmel(const float* stft, int B, int T, int n_bins, const int* fbank_lo,
                                 const int* fbank_hi, const float* fbank_w, int fbank_stride,
                                 int n_mels, const int* seq_len, float* logmel, double* stats,
                                 void* stream) {
  if (!stft || !fbank_lo || !fbank_hi || !fbank_w || !logmel) return PBSED_EINVAL;
  if (B < 1 || T < 1 || n_bins < 1 || n_mels < 1 || B > 65535) return PBSED_EINVAL;
  const size_t smem = logmel_smem(false, 0, n_bins, n_mels, 0);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  cudaError_t e = cudaFuncSetAttribute(logmel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(cdiv(T, FR), B);
  logmel_kernel<false><<<grid, 256, smem, (cudaStream_t)stream>>>(
      stft, 0, 0, 0, 0, 0, T, n_bins, nullptr, fbank_lo, fbank_hi, fbank_w, fbank_stride, n_mels,
      seq_len, logmel, stats);
  return pbsed_after_launch();
}

// ------------------------------------------------------------------ normalise + clamp + mask
__global__ void __launch_bounds__(256)
logmel_normalize_kernel(float* __restrict__ x, int F, int T, const float* __restrict__ scale,
                        const float* __restrict__ shift, float clampv, const int* __restrict__ seq_len,
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
    }
    x[i] = v;
  }
}

extern "C" int pbsed_logmel_normalize(float* x, int B, int F, int T, const float* scale,
                                      const float* shift, float clampv, const int* seq_len,
                                      void* stream) {
  if (!x || !scale || !shift || B < 1 || F < 1 || T < 1) return PBSED_EINVAL;
  const long long total = (long long)B * F * T;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  logmel_normalize_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, F, T, scale, shift, clampv, seq_len, total);
  return pbsed_after_launch();
}
