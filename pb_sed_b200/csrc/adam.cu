// adam.cu -- K5: global gradient L2 norm, clip, Adam update over ONE flat parameter arena.
//
// Reference: padertorch.train.optimizer.Adam (= torch.nn.utils.clip_grad_norm_ followed by
// torch.optim.Adam with default betas/eps) as configured at
// pb_sed/experiments/weak_label_crnn/training.py:264-269 (lr 5e-4, gradient_clipping 1e10 for
// DESED :102,110; lr 1e-4, clipping 0.1 for AudioSet :139,150).
//   clip_coef = min(1, max_norm / (norm + 1e-6));   g <- g * clip_coef
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//   p -= lr / (1-b1^t) * m / ( sqrt(v) / sqrt(1-b2^t) + eps )
// hyper (device float[8]): lr, beta1, beta2, eps, max_norm, step, grad_scale, unused -- device
// resident so that the LR schedule (training.py:377-396) can change it under a captured CUDA graph.
#include "common.cuh"

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(const float* __restrict__ g, long long n, const float* __restrict__ hyper,
                  double* __restrict__ sumsq) {
  __shared__ double sh[8];
  const float gs = hyper ? hyper[6] : 1.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = __ldg(g + i) * gs;
    acc = fmaf(v, v, acc);
  }
  double d = warp_sum_d((double)acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(sumsq, t);
  }
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, const float* __restrict__ hyper, const double* __restrict__ sumsq,
            float* __restrict__ grad_norm_out, int zero_grad) {
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], max_norm = hyper[4];
  const float step = hyper[5] + 1.f, gs = hyper[6];
  const float norm = (float)sqrt(*sumsq);
  const float coef = fminf(max_norm / (norm + 1e-6f), 1.f) * gs;
  const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
  const float step_size = lr / bc1, rsbc2 = rsqrtf(bc2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    p[i] -= step_size * mi / (sqrtf(vi) * rsbc2 + eps);
    if (zero_grad) g[i] = 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && grad_norm_out) *grad_norm_out = norm;
}

__global__ void adam_advance_kernel(float* hyper, double* sumsq) { hyper[5] += 1.f; *sumsq = 0.; }

static int flat_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148LL * 8) b = 148LL * 8;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int pbsed_grad_sumsq(const float* g, long long n, const float* hyper, double* sumsq, void* stream) {
  if (!g || !sumsq || n < 1) return PBSED_EINVAL;
  grad_sumsq_kernel<<<flat_blocks(n), 256, 0, (cudaStream_t)stream>>>(g, n, hyper, sumsq);
  return pbsed_after_launch();
}

extern "C" int pbsed_adam_step(float* p, float* g, float* m, float* v, long long n, float* hyper,
                               double* sumsq, float* grad_norm_out, int zero_grad, void* stream) {
  if (!p || !g || !m || !v || !hyper || !sumsq || n < 1) return PBSED_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  adam_kernel<<<flat_blocks(n), 256, 0, st>>>(p, g, m, v, n, hyper, sumsq, grad_norm_out, zero_grad);
  int rc = pbsed_after_launch();
  if (rc) return rc;
  adam_advance_kernel<<<1, 1, 0, st>>>(hyper, sumsq);
  return pbsed_after_launch();
}
