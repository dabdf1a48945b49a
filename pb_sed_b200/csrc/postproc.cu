// postproc.cu -- score post-processing of the inference path on the GPU.
//
// Reference: pb_sed/filters.py:56-83 (medfilt = scipy.signal.medfilt per row, zero padded),
// pb_sed/filters.py:113-135 (stepfilt), pb_sed/models/base/inference.py:143-151 (sequence mask, then
// median filter, then boundary filter), :225-266 (filtering: scalar / per-class / per-(n, class) filter
// lengths), :269-289 (boundariesfilt = min(cummax(stepfilt(x)), flip(cummax(stepfilt(flip(x)))))),
// :170-183 (tag masking of SED scores).  The reference does all of this in numpy on the host after a
// D2H copy of the scores, row by row through np.apply_along_axis.
//
// Rows: the score tensor (B, [N,] K, T) flattened to (R, T); the filter length of row r is
// filt_len[r % filt_mod] (a device array: 1 entry = scalar, K = per class, N*K = per (n, class));
// clip b = r / rows_per_clip owns seq_len[b] valid frames, everything behind is treated as 0
// (inference.py:143-147).  One CTA per row; the row lives in shared memory.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned f2key(float f) {          // order-preserving float -> uint map
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------ median filter
// y[t] = (n-1)/2-th smallest of { x[t-h .. t+h] } with zeros outside [0, T): selection, so the result
// is bit-identical to scipy's.  The k-th smallest key is the largest v with #(key < v) <= k; v is
// built bit by bit (32 counting passes over the window, all in shared memory).
__global__ void __launch_bounds__(256)
medfilt_kernel(const float* __restrict__ x, int T, const int* __restrict__ filt_len, int filt_mod,
               const int* __restrict__ seq_len, int rows_per_clip, float* __restrict__ y) {
  extern __shared__ unsigned keys[];
  const int r = blockIdx.x;
  const float* xr = x + (long long)r * T;
  float* yr = y + (long long)r * T;
  const int len = seq_len ? min(__ldg(seq_len + r / rows_per_clip), T) : T;
  const int n = __ldg(filt_len + (r % filt_mod));
  for (int t = threadIdx.x; t < T; t += blockDim.x) keys[t] = f2key(t < len ? xr[t] : 0.f);
  __syncthreads();
  if (n <= 1) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) yr[t] = key2f(keys[t]);
    return;
  }
  const int h = (n - 1) >> 1;
  const unsigned key0 = 0x80000000u;                           // f2key(+0.0f)
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int lo = max(t - h, 0), hi = min(t + h, T - 1);
    const int nz = n - (hi - lo + 1);                          // zero padding inside the window
    unsigned ans = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const unsigned cand = ans | (1u << bit);
      int cnt = key0 < cand ? nz : 0;
      for (int i = lo; i <= hi; ++i) cnt += keys[i] < cand;
      if (cnt <= h) ans = cand;
    }
    yr[t] = key2f(ans);
  }
}

// ------------------------------------------------------------------ median filter, T <= 512 (a 10 s clip has 500 frames)
// The row is sorted ONCE (bitonic, 512 (key, index) pairs in shared memory); every window query then works
// on ranks: a table C[b][t] = #{ i < t : rank(i) < 32 (b + 1) } (16 rank buckets x 513 prefix counts) gives
// the number of window elements below any bucket boundary with two loads, and the answer is found by
// scanning the <= 32 ranks of ONE bucket.  ~60 shared-memory reads per output instead of 32 x n, and still
// a selection (bit-exact).  Zero padding is handled analytically: the window's elements below zero come
// first, then the nz padded zeros, then the rest.
constexpr int MS_P = 512, MS_NB = MS_P / 32;

__global__ void __launch_bounds__(256)
medfilt_sorted_kernel(const float* __restrict__ x, int T, const int* __restrict__ filt_len, int filt_mod,
                      const int* __restrict__ seq_len, int rows_per_clip, float* __restrict__ y) {
  __shared__ unsigned long long kv[MS_P];                 // (key << 32) | index, ascending after the sort
  __shared__ unsigned short rank_of[MS_P];
  __shared__ unsigned short C[MS_NB][MS_P + 2];
  __shared__ int warp_tot[8];
  __shared__ int r0_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r = blockIdx.x;
  const float* xr = x + (long long)r * T;
  float* yr = y + (long long)r * T;
  const int len = seq_len ? min(__ldg(seq_len + r / rows_per_clip), T) : T;
  const int n = __ldg(filt_len + (r % filt_mod));
  if (n <= 1) {                                           // uniform per CTA
    for (int t = tid; t < T; t += 256) yr[t] = t < len ? xr[t] : 0.f;
    return;
  }
  for (int i = tid; i < MS_P; i += 256) {
    const unsigned key = i < T ? f2key(i < len ? xr[i] : 0.f) : 0xFFFFFFFFu;
    kv[i] = ((unsigned long long)key << 32) | (unsigned)i;
  }
  __syncthreads();
  for (int k = 2; k <= MS_P; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      // one compare-exchange per thread: pair (i, i ^ j) with i the member whose bit j is clear
      const int i = ((tid & ~(j - 1)) << 1) | (tid & (j - 1));
      const int p = i | j;
      const unsigned long long a = kv[i], b = kv[p];
      const bool up = (i & k) == 0;
      if ((a > b) == up) { kv[i] = b; kv[p] = a; }
      __syncthreads();
    }
  for (int pos = tid; pos < MS_P; pos += 256) {
    const unsigned idx = (unsigned)(kv[pos] & 0xFFFFFFFFu);
    if (idx < (unsigned)T) rank_of[idx] = (unsigned short)pos;
  }
  if (tid == 0) {                                         // r0 = number of row elements below +0.0 (lower bound)
    int lo = 0, hi = T;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if ((unsigned)(kv[mid] >> 32) < 0x80000000u) lo = mid + 1; else hi = mid;
    }
    r0_s = lo;
  }
  __syncthreads();
  {   // C[b][t]: 16 block-wide exclusive scans, two consecutive t per thread
    const int i0 = 2 * tid;
    const int ra = i0 < T ? (int)rank_of[i0] : 1 << 20, rb = i0 + 1 < T ? (int)rank_of[i0 + 1] : 1 << 20;
    for (int b = 0; b < MS_NB; ++b) {
      const int thr = 32 * (b + 1);
      const int a = ra < thr, c = rb < thr;
      int v = a + c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
      if (lane == 31) warp_tot[warp] = v;
      __syncthreads();
      int off = 0;
      for (int w = 0; w < warp; ++w) off += warp_tot[w];
      const int excl = off + v - (a + c);
      C[b][i0] = (unsigned short)excl;
      C[b][i0 + 1] = (unsigned short)(excl + a);
      if (tid == 255) C[b][MS_P] = (unsigned short)(off + v);
      __syncthreads();
    }
  }
  const int h = (n - 1) >> 1;
  const int r0 = r0_s;
  for (int t = tid; t < T; t += 256) {
    const int lo = max(t - h, 0), hi = min(t + h, T - 1);
    const int nz = n - (hi - lo + 1);
    // m = number of window elements with rank < r0 (the negative ones)
    int m;
    {
      const int b = r0 >> 5;
      m = b > 0 ? (int)C[b - 1][hi + 1] - (int)C[b - 1][lo] : 0;
      for (int q = 32 * b; q < r0; ++q) {
        const int idx = (int)(kv[q] & 0xFFFFFFFFu);
        m += (idx >= lo && idx <= hi);
      }
    }
    int target;
    if (h < m) target = h;
    else if (h < m + nz) { yr[t] = 0.f; continue; }
    else target = h - nz;
    int b = 0, prev = 0;
    for (; b < MS_NB; ++b) {
      const int cle = (int)C[b][hi + 1] - (int)C[b][lo];
      if (cle > target) break;
      prev = cle;
    }
    unsigned ans = 0x80000000u;
    for (int q = 32 * b; q < 32 * b + 32; ++q) {
      const unsigned long long e = kv[q];
      const int idx = (int)(e & 0xFFFFFFFFu);
      if (idx >= lo && idx <= hi) {
        if (prev == target) { ans = (unsigned)(e >> 32); break; }
        ++prev;
      }
    }
    yr[t] = key2f(ans);
  }
}

// ------------------------------------------------------------------ boundaries filter
// block-wide inclusive scans over a shared-memory row (each thread owns a contiguous chunk)
template <bool IS_MAX, bool REVERSE>
__device__ void block_scan(double* v, int T, double* part) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int chunk = (T + nt - 1) / nt;
  const int a = min(tid * chunk, T), b = min(a + chunk, T);
  double acc = IS_MAX ? -1.0 / 0.0 : 0.0;
  if (!REVERSE) { for (int i = a; i < b; ++i) { acc = IS_MAX ? fmax(acc, v[i]) : acc + v[i]; v[i] = acc; } }
  else { for (int i = T - 1 - a; i > T - 1 - b; --i) { acc = IS_MAX ? fmax(acc, v[i]) : acc + v[i]; v[i] = acc; } }
  part[tid] = acc;
  __syncthreads();
  if (tid == 0) {
    double run = IS_MAX ? -1.0 / 0.0 : 0.0;
    for (int j = 0; j < nt; ++j) { const double p = part[j]; part[j] = run; run = IS_MAX ? fmax(run, p) : run + p; }
  }
  __syncthreads();
  const double off = part[tid];
  if (!REVERSE) { for (int i = a; i < b; ++i) v[i] = IS_MAX ? fmax(v[i], off) : v[i] + off; }
  else { for (int i = T - 1 - a; i > T - 1 - b; --i) v[i] = IS_MAX ? fmax(v[i], off) : v[i] + off; }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
boundariesfilt_kernel(const float* __restrict__ x, int T, const int* __restrict__ filt_len, int filt_mod,
                      const int* __restrict__ seq_len, int rows_per_clip, double* __restrict__ y) {
  extern __shared__ double sm[];
  double* P = sm;                 // [T+1] exclusive prefix sums, P[i] = sum_{u<i} x[u]
  double* Fw = P + (T + 1);       // [T]
  double* Bw = Fw + T;            // [T]
  double* part = Bw + T;          // [blockDim]
  const int r = blockIdx.x;
  const float* xr = x + (long long)r * T;
  const int len = seq_len ? min(__ldg(seq_len + r / rows_per_clip), T) : T;
  const int n = __ldg(filt_len + (r % filt_mod));
  const int h = n >> 1;
  for (int t = threadIdx.x; t < T; t += blockDim.x) P[t + 1] = t < len ? (double)xr[t] : 0.0;
  if (threadIdx.x == 0) P[0] = 0.0;
  __syncthreads();
  if (h > 0) {
    // keep the raw samples implicit: x[t] = P[t+1] - P[t] is not needed once the prefix exists
    block_scan<false, false>(P + 1, T, part);
    const double inv = 1.0 / (double)h;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const double pt = P[t], pt1 = P[t + 1];
      Fw[t] = ((P[min(t + h, T)] - pt) - (pt - P[max(t - h, 0)])) * inv;
      Bw[t] = ((pt1 - P[max(t + 1 - h, 0)]) - (P[min(t + 1 + h, T)] - pt1)) * inv;
    }
  } else {
    for (int t = threadIdx.x; t < T; t += blockDim.x) { Fw[t] = P[t + 1]; Bw[t] = P[t + 1]; }
  }
  __syncthreads();
  block_scan<true, false>(Fw, T, part);      // cummax over t
  block_scan<true, true>(Bw, T, part);       // cummax over the flipped axis, flipped back
  double* yr = y + (long long)r * T;
  for (int t = threadIdx.x; t < T; t += blockDim.x) yr[t] = fmin(Fw[t], Bw[t]);
}

// tag masking of detection scores (inference.py:170-183): scores (B, N, K, T) *= max(tags[b, k], 1 - apply[n, k])
__global__ void __launch_bounds__(256)
tag_mask_kernel(float* __restrict__ s, const float* __restrict__ tags, const float* __restrict__ apply,
                int NK, int K, int T, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long row = i / T;                 // (b, n, k)
    const int nk = (int)(row % NK);
    const long long b = row / NK;
    s[i] *= fmaxf(__ldg(tags + b * K + nk % K), 1.f - __ldg(apply + nk));
  }
}

}  // namespace

extern "C" int pbsed_medfilt(const float* x, int R, int T, const int* filt_len, int filt_mod,
                             const int* seq_len, int rows_per_clip, float* y, void* stream) {
  if (!x || !y || !filt_len || R < 1 || T < 1 || filt_mod < 1 || rows_per_clip < 1) return PBSED_EINVAL;
  if (T <= MS_P) {
    medfilt_sorted_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(x, T, filt_len, filt_mod, seq_len, rows_per_clip, y);
    return pbsed_after_launch();
  }
  const size_t smem = (size_t)T * sizeof(unsigned);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(medfilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  medfilt_kernel<<<R, 256, smem, (cudaStream_t)stream>>>(x, T, filt_len, filt_mod, seq_len, rows_per_clip, y);
  return pbsed_after_launch();
}

extern "C" int pbsed_boundariesfilt(const float* x, int R, int T, const int* filt_len, int filt_mod,
                                    const int* seq_len, int rows_per_clip, double* y, void* stream) {
  if (!x || !y || !filt_len || R < 1 || T < 1 || filt_mod < 1 || rows_per_clip < 1) return PBSED_EINVAL;
  const size_t smem = ((size_t)3 * T + 1 + 256) * sizeof(double);
  if (smem > 200 * 1024) return PBSED_EINVAL;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(boundariesfilt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  boundariesfilt_kernel<<<R, 256, smem, (cudaStream_t)stream>>>(x, T, filt_len, filt_mod, seq_len, rows_per_clip, y);
  return pbsed_after_launch();
}

extern "C" int pbsed_tag_mask(float* scores, const float* mask, const float* apply, int B, int N, int K,
                              int T, void* stream) {
  if (!scores || !mask || !apply || B < 1 || N < 1 || K < 1 || T < 1) return PBSED_EINVAL;
  const long long total = (long long)B * N * K * T;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tag_mask_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(scores, mask, apply, N * K, K, T, total);
  return pbsed_after_launch();
}
