// gru.cu -- persistent, register-resident GRU recurrence (forward and BPTT) for sm_100a.
//
// Reference: torch.nn.GRU inside padertorch.contrib.je.modules.rnn.GRU, called at
// pb_sed/models/weak_label/crnn.py:62,66 (rnn_fwd / rnn_bwd = same config + reverse=True,
// :338-340) and pb_sed/models/strong_label/crnn.py:92 (bidirectional=True); gate order r,z,n:
//     r = s(gi_r + W_hr h + b_hr)   z = s(gi_z + W_hz h + b_hz)
//     n = tanh(gi_n + r * (W_hn h + b_hn))     h' = (1-z) n + z h
// packed-sequence semantics: clip b only advances for its seq_len[b] valid frames; the
// reverse direction walks t = len-1 .. 0; outputs at padded frames are zero.
//
// Mapping: one thread-block CLUSTER of NC = H/32 CTAs per (direction, 8-clip batch slice).
// CTA `rank` owns hidden units [32*rank, 32*rank+32).  Thread (warp q, lane j) keeps the
// W_hh entries {gate g} x {unit u = 32*rank + j} x {k in [q*H/8, (q+1)*H/8)} in REGISTERS
// for the whole sequence (3*H/8 = 96 registers at H = 256), so the 500 dependent steps never
// re-read the weights.  Per step: h_{t-1} (8 x H, replicated in every CTA's shared memory)
// is multiplied against the register tile, the 8 k-slices are reduced through shared memory,
// the gate math runs on thread (q = clip, j = unit), and the new h is pushed to every CTA of
// the cluster through distributed shared memory, followed by one cluster barrier.
#include "common.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
namespace cg = cooperative_groups;

struct GruDirs {             // per-direction operands: directions may belong to different modules
  int reverse[4];
  const float* gi[4];        // (B,T,3H)
  const float* w_hh[4];      // (3H,H)
  const float* b_hh[4];      // (3H)
  float* h_out[4];           // rows (B,T,h_stride), H channels written
  float* save[4];            // (B,T,4H) or null
  const float* dh_out[4];    // backward: gradient w.r.t. h_out
  float* dgi[4];             // backward outputs (B,T,3H)
  float* dgh[4];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// packed 2 x fp32 FMA (Blackwell FFMA2): halves the instruction count of the recurrent mat-vec,
// which ncu showed to be issue-bound (57 % issue-active, 6144 of 8580 warp instructions per step)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a);
  unsigned long long rb = *reinterpret_cast<unsigned long long*>(&b);
  unsigned long long rc = *reinterpret_cast<unsigned long long*>(&c);
  unsigned long long rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}

// ---- cluster exchange without a cluster-scope fence in the time loop ------------------------
// The per-step state (h, or the gate gradients) is pushed into every CTA of the cluster with
// st.async: a remote shared-memory store that completes bytes on the DESTINATION CTA's mbarrier.
// Consumers wait on their local mbarrier, so the loop needs no barrier.cluster / release fence --
// which would also have to drain the step's global stores (h_out, saved gates): ncu showed that
// membar stall as the top stall of the first version (profiles/r01_gru_ncu.txt).
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
               ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void xbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void xbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "XW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra XD_%=;\n"
      "bra XW_%=;\n"
      "XD_%=:\n"
      "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

template <int H, int BC>
__global__ void __launch_bounds__(256, 1)
gru_fwd_kernel(const int* __restrict__ seq_len, int B, int T, GruDirs dirs, int h_stride) {
  constexpr int NC = H / 32, KS = H / 8;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, q = tid >> 5, j = tid & 31;
  const int u = rank * 32 + j;
  const int d = blockIdx.z;
  const bool reverse = dirs.reverse[d] != 0;
  const bool ew = q < BC;                       // warps 0..BC-1 run the gate math for clip q
  const int bq = ew ? blockIdx.y * BC + q : B;

  __shared__ __align__(16) float hbuf[2][BC][H];
  __shared__ float red[8][3][BC][32];
  __shared__ __align__(8) uint64_t hbar[2];            // "buffer i holds the complete h of a step"

  const float* w = dirs.w_hh[d];
  const float* bh = dirs.b_hh[d];
  float2 W[3][KS / 2];                         // (k, k+1) pairs: FFMA2 accumulates even / odd k separately
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int kk = 0; kk < KS / 2; ++kk)
      W[g][kk] = make_float2(__ldg(w + (long long)(g * H + u) * H + q * KS + 2 * kk),
                             __ldg(w + (long long)(g * H + u) * H + q * KS + 2 * kk + 1));
  const float bhr = __ldg(bh + u), bhz = __ldg(bh + H + u), bhn = __ldg(bh + 2 * H + u);

  for (int i = tid; i < 2 * BC * H; i += 256) (&hbuf[0][0][0])[i] = 0.f;
  if (tid == 0) {
    xbar_init(&hbar[0], 1); xbar_init(&hbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const int len = (bq < B) ? (seq_len ? min(__ldg(seq_len + bq), T) : T) : 0;
  const float* gi_b = dirs.gi[d] + (long long)bq * T * 3 * H;
  float* ho_b = dirs.h_out[d] + (long long)bq * T * h_stride;
  float* sv_b = dirs.save[d] ? dirs.save[d] + (long long)bq * T * 4 * H : nullptr;
  float hprev = 0.f;
  // the input projections of step s+1 are fetched while step s computes (they are the only global
  // loads on the recurrence's critical path)
  float nxt_r = 0.f, nxt_z = 0.f, nxt_n = 0.f;
  if (0 < len) {
    const float* p = gi_b + (long long)(reverse ? len - 1 : 0) * 3 * H + u;
    nxt_r = __ldg(p); nxt_z = __ldg(p + H); nxt_n = __ldg(p + 2 * H);
  }
  cluster.sync();

  for (int s = 0; s < T; ++s) {
    const int cur = s & 1;
    const bool active = s < len;
    const int t = reverse ? (len - 1 - s) : s;
    const float gir = nxt_r, giz = nxt_z, gin = nxt_n;
    // arm the barrier of the buffer this step fills, then wait for the buffer this step reads
    if (tid == 0) xbar_expect(&hbar[cur ^ 1], (uint32_t)(NC * BC * 32 * sizeof(float)));
    if (s > 0) xbar_wait(&hbar[cur], (uint32_t)(((s - 1) >> 1) & 1));
    if (s + 1 < len) {
      const float* p = gi_b + (long long)(reverse ? len - 2 - s : s + 1) * 3 * H + u;
      nxt_r = __ldg(p); nxt_z = __ldg(p + H); nxt_n = __ldg(p + 2 * H);
    }
    float2 acc[3][BC];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int b = 0; b < BC; ++b) acc[g][b] = make_float2(0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < KS; kk += 4) {
#pragma unroll
      for (int b = 0; b < BC; ++b) {
        const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][b][q * KS + kk]);
        const float2 h01 = make_float2(hv.x, hv.y), h23 = make_float2(hv.z, hv.w);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          acc[g][b] = ffma2(W[g][kk / 2], h01, acc[g][b]);
          acc[g][b] = ffma2(W[g][kk / 2 + 1], h23, acc[g][b]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int b = 0; b < BC; ++b) red[q][g][b][j] = acc[g][b].x + acc[g][b].y;
    __syncthreads();
    float ghr = bhr, ghz = bhz, ghn = bhn;
    if (ew) {
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) {
        ghr += red[qq][0][q][j]; ghz += red[qq][1][q][j]; ghn += red[qq][2][q][j];
      }
    }
    float hnew = hprev;
    if (active) {
      const float r = sigmoidf_(gir + ghr);
      const float z = sigmoidf_(giz + ghz);
      const float n = tanhf(fmaf(r, ghn, gin));
      hnew = (1.f - z) * n + z * hprev;
      ho_b[(long long)t * h_stride + u] = hnew;
      if (sv_b) {
        float* sp = sv_b + (long long)t * 4 * H + u;
        sp[0] = r; sp[H] = z; sp[2 * H] = n; sp[3 * H] = ghn;
      }
    } else if (bq < B) {
      ho_b[(long long)s * h_stride + u] = 0.f;       // padded frame t = s >= len
    }
    hprev = hnew;
    if (ew) {
      const uint32_t mine = smem_addr(&hbuf[cur ^ 1][q][u]), bar = smem_addr(&hbar[cur ^ 1]);
#pragma unroll
      for (int c = 0; c < NC; ++c) st_async_f32(mapa_rank(mine, c), hnew, mapa_rank(bar, c));
    }
    // red[] is rewritten by the next step's matvec only after every warp passed this step's reads:
    // the wait above (all threads) sits between them, but a CTA-local barrier keeps it explicit
    __syncthreads();
  }
  cluster.sync();      // no CTA may exit while peers can still push into its shared memory
}

// ------------------------------------------------------------------ BPTT
template <int H, int BC>
__global__ void __launch_bounds__(256, 1)
gru_bwd_kernel(const int* __restrict__ seq_len, int B, int T, GruDirs dirs, int h_stride) {
  constexpr int NC = H / 32, KS = H / 8;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, q = tid >> 5, j = tid & 31;
  const int u = rank * 32 + j;
  const int d = blockIdx.z;
  const bool reverse = dirs.reverse[d] != 0;
  const bool ew = q < BC;
  const int bq = ew ? blockIdx.y * BC + q : B;

  extern __shared__ __align__(16) float smem[];
  float* dg = smem;                            // [2][3][BC][H]
  float* red = smem + 2 * 3 * BC * H;          // [8][BC][32]
  __shared__ __align__(8) uint64_t gbar[2];

  const float* w = dirs.w_hh[d];
  float2 Wt[3][KS / 2];
#pragma unroll
  for (int g = 0; g < 3; ++g)
#pragma unroll
    for (int kk = 0; kk < KS / 2; ++kk)
      Wt[g][kk] = make_float2(__ldg(w + (long long)(g * H + q * KS + 2 * kk) * H + u),
                              __ldg(w + (long long)(g * H + q * KS + 2 * kk + 1) * H + u));

  const int len = (bq < B) ? (seq_len ? min(__ldg(seq_len + bq), T) : T) : 0;
  const float* dho_b = dirs.dh_out[d] + (long long)bq * T * h_stride;
  const float* ho_b = dirs.h_out[d] + (long long)bq * T * h_stride;
  const float* sv_b = dirs.save[d] + (long long)bq * T * 4 * H;
  float* dgi_b = dirs.dgi[d] + (long long)bq * T * 3 * H;
  float* dgh_b = dirs.dgh[d] + (long long)bq * T * 3 * H;
  float dh_carry = 0.f;
  // operands of the NEXT processed step (s-1) are fetched while step s computes
  float n_dho = 0.f, n_r = 0.f, n_z = 0.f, n_n = 0.f, n_ghn = 0.f, n_hp = 0.f;
  auto fetch = [&](int s) {
    if (s >= 0 && s < len) {
      const int t = reverse ? (len - 1 - s) : s;
      n_dho = __ldg(dho_b + (long long)t * h_stride + u);
      const float* sp = sv_b + (long long)t * 4 * H + u;
      n_r = __ldg(sp); n_z = __ldg(sp + H); n_n = __ldg(sp + 2 * H); n_ghn = __ldg(sp + 3 * H);
      n_hp = s > 0 ? __ldg(ho_b + (long long)(reverse ? t + 1 : t - 1) * h_stride + u) : 0.f;
    }
  };
  fetch(T - 1);
  if (tid == 0) {
    xbar_init(&gbar[0], 1); xbar_init(&gbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();
  int fills0 = 0, fills1 = 0;                  // completed fills per buffer -> wait parity

  for (int s = T - 1; s >= 0; --s) {
    const int cur = s & 1;
    const bool active = s < len;
    const int t = reverse ? (len - 1 - s) : s;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, direct = dh_carry;
    if (tid == 0) xbar_expect(&gbar[cur], (uint32_t)(3 * NC * BC * 32 * sizeof(float)));
    const float c_dho = n_dho, r = n_r, z = n_z, n = n_n, ghn = n_ghn, hp = n_hp;
    fetch(s - 1);
    if (active) {
      const float dh = dh_carry + c_dho;
      const float dn = dh * (1.f - z) * (1.f - n * n);
      const float dz = dh * (hp - n) * z * (1.f - z);
      const float dr = dn * ghn * r * (1.f - r);
      v0 = dr; v1 = dz; v2 = dn * r;
      float* gp = dgi_b + (long long)t * 3 * H + u;
      gp[0] = dr; gp[H] = dz; gp[2 * H] = dn;
      float* hp2 = dgh_b + (long long)t * 3 * H + u;
      hp2[0] = dr; hp2[H] = dz; hp2[2 * H] = v2;
      direct = dh * z;
    } else if (bq < B) {
      float* gp = dgi_b + (long long)s * 3 * H + u;
      gp[0] = 0.f; gp[H] = 0.f; gp[2 * H] = 0.f;
      float* hp2 = dgh_b + (long long)s * 3 * H + u;
      hp2[0] = 0.f; hp2[H] = 0.f; hp2[2 * H] = 0.f;
    }
    if (ew) {
      const uint32_t m0 = smem_addr(dg + ((cur * 3 + 0) * BC + q) * H + u);
      const uint32_t m1 = smem_addr(dg + ((cur * 3 + 1) * BC + q) * H + u);
      const uint32_t m2 = smem_addr(dg + ((cur * 3 + 2) * BC + q) * H + u);
      const uint32_t bar = smem_addr(&gbar[cur]);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const uint32_t rb = mapa_rank(bar, c);
        st_async_f32(mapa_rank(m0, c), v0, rb);
        st_async_f32(mapa_rank(m1, c), v1, rb);
        st_async_f32(mapa_rank(m2, c), v2, rb);
      }
    }
    {
      int& fills = cur ? fills1 : fills0;
      xbar_wait(&gbar[cur], (uint32_t)(fills & 1));
      ++fills;
    }
    float2 acc[BC];
#pragma unroll
    for (int b = 0; b < BC; ++b) acc[b] = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int kk = 0; kk < KS; kk += 4)
#pragma unroll
        for (int b = 0; b < BC; ++b) {
          const float4 dv = *reinterpret_cast<const float4*>(dg + ((cur * 3 + g) * BC + b) * H + q * KS + kk);
          acc[b] = ffma2(Wt[g][kk / 2], make_float2(dv.x, dv.y), acc[b]);
          acc[b] = ffma2(Wt[g][kk / 2 + 1], make_float2(dv.z, dv.w), acc[b]);
        }
#pragma unroll
    for (int b = 0; b < BC; ++b) red[(q * BC + b) * 32 + j] = acc[b].x + acc[b].y;
    __syncthreads();
    float sum = direct;
    if (ew) {
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) sum += red[(qq * BC + q) * 32 + j];
    }
    dh_carry = sum;
    __syncthreads();     // red[] is free for the next step
  }
  cluster.sync();        // no CTA may exit while peers can still push into its shared memory
}

// ------------------------------------------------------------------ launchers
template <int H, bool BWD, int BC>
static int launch_gru_bc(const int* seq_len, int B, int T, int ndir, const GruDirs& dirs, int h_stride,
                         cudaStream_t st) {
  size_t smem = 0;
  cudaError_t e;
  if (BWD) {
    smem = (size_t)(2 * 3 * BC * H + 8 * BC * 32) * sizeof(float);
    e = cudaFuncSetAttribute(gru_bwd_kernel<H, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H / 32, cdiv(B, BC), ndir);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / 32; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (BWD) e = cudaLaunchKernelEx(&cfg, gru_bwd_kernel<H, BC>, seq_len, B, T, dirs, h_stride);
  else     e = cudaLaunchKernelEx(&cfg, gru_fwd_kernel<H, BC>, seq_len, B, T, dirs, h_stride);
  ++g_pbsed_launches;
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

// how many clusters of H/32 CTAs are co-resident on THIS device (an 8-CTA cluster needs 8 free SMs of one GPC: 14 on
// the B200s measured); asked from the runtime once per kernel variant, sized for the smallest clips-per-cluster variant
template <int H, bool BWD>
static int gru_max_clusters() {
  static int cached = 0;
  if (cached) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H / 32, 64, 1);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = BWD ? (size_t)(2 * 3 * 4 * H + 8 * 4 * 32) * sizeof(float) : 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / 32; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  cudaError_t e;
  if (BWD) {
    cudaFuncSetAttribute(gru_bwd_kernel<H, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    e = cudaOccupancyMaxActiveClusters(&n, gru_bwd_kernel<H, 4>, &cfg);
  } else {
    e = cudaOccupancyMaxActiveClusters(&n, gru_fwd_kernel<H, 4>, &cfg);
  }
  if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = (H / 32 == 8) ? 14 : 148 / (H / 32); }
  cached = n;
  return n;
}

// clips per cluster: 8 (measured faster than 4 at B = 32 even though 4 halves the FMAs per step;
// PBSED_GRU_BC=4 forces the other variant for experiments)
template <int H, bool BWD>
static int launch_gru(const int* seq_len, int B, int T, int ndir, const GruDirs& dirs, int h_stride,
                      cudaStream_t st) {
  static const int force_bc = getenv("PBSED_GRU_BC") ? atoi(getenv("PBSED_GRU_BC")) : 0;
  int bc = force_bc;
  if (bc == 0) {
    // fewest clips per cluster whose clusters are all co-resident: an 8-CTA cluster needs 8 free SMs of
    // one GPC and ~14 of them fit a B200 at once (16 clusters of the 4-clip variant at B = 32 ran in two
    // rounds = 2x the time); fewer clips = fewer FMAs on the 500-step critical path
    const int max_clusters = gru_max_clusters<H, BWD>();
    bc = 8;
    for (int c : {4, 5, 6}) if (cdiv(B, c) * ndir <= max_clusters) { bc = c; break; }
  }
  switch (bc) {
    case 4: return launch_gru_bc<H, BWD, 4>(seq_len, B, T, ndir, dirs, h_stride, st);
    case 5: return launch_gru_bc<H, BWD, 5>(seq_len, B, T, ndir, dirs, h_stride, st);
    case 6: return launch_gru_bc<H, BWD, 6>(seq_len, B, T, ndir, dirs, h_stride, st);
    default: return launch_gru_bc<H, BWD, 8>(seq_len, B, T, ndir, dirs, h_stride, st);
  }
}

template <bool BWD>
static int dispatch_gru(const int* seq_len, int B, int T, int H, int ndir, const GruDirs& dirs,
                        int h_stride, cudaStream_t st) {
  switch (H) {
    case 32:  return launch_gru<32, BWD>(seq_len, B, T, ndir, dirs, h_stride, st);
    case 64:  return launch_gru<64, BWD>(seq_len, B, T, ndir, dirs, h_stride, st);
    case 128: return launch_gru<128, BWD>(seq_len, B, T, ndir, dirs, h_stride, st);
    case 256: return launch_gru<256, BWD>(seq_len, B, T, ndir, dirs, h_stride, st);
    default:  return PBSED_EINVAL;
  }
}

extern "C" int pbsed_gru_fwd(const float* const* gi, const float* const* w_hh,
                             const float* const* b_hh, const int* seq_len, int B, int T, int H,
                             int ndir, const int* reverse_host, float* const* h_out, int h_stride,
                             float* const* save, void* stream) {
  if (ndir < 1 || ndir > 4 || !reverse_host || !gi || !w_hh || !b_hh || !h_out) return PBSED_EINVAL;
  if (B < 1 || T < 1 || h_stride < H) return PBSED_EINVAL;
  GruDirs dirs = {};
  for (int d = 0; d < ndir; ++d) {
    if (!gi[d] || !w_hh[d] || !b_hh[d] || !h_out[d]) return PBSED_EINVAL;
    dirs.reverse[d] = reverse_host[d];
    dirs.gi[d] = gi[d]; dirs.w_hh[d] = w_hh[d]; dirs.b_hh[d] = b_hh[d];
    dirs.h_out[d] = h_out[d]; dirs.save[d] = save ? save[d] : nullptr;
  }
  return dispatch_gru<false>(seq_len, B, T, H, ndir, dirs, h_stride, (cudaStream_t)stream);
}

extern "C" int pbsed_gru_bwd(const float* const* dh_out, const float* const* h_out,
                             const float* const* save, const float* const* w_hh,
                             const int* seq_len, int B, int T, int H, int ndir,
                             const int* reverse_host, float* const* dgi, float* const* dgh,
                             int h_stride, void* stream) {
  if (ndir < 1 || ndir > 4 || !reverse_host || !dh_out || !h_out || !save || !w_hh || !dgi || !dgh) return PBSED_EINVAL;
  if (B < 1 || T < 1 || h_stride < H) return PBSED_EINVAL;
  GruDirs dirs = {};
  for (int d = 0; d < ndir; ++d) {
    if (!dh_out[d] || !h_out[d] || !save[d] || !w_hh[d] || !dgi[d] || !dgh[d]) return PBSED_EINVAL;
    dirs.reverse[d] = reverse_host[d];
    dirs.dh_out[d] = dh_out[d]; dirs.h_out[d] = const_cast<float*>(h_out[d]); dirs.save[d] = const_cast<float*>(save[d]);
    dirs.w_hh[d] = w_hh[d]; dirs.dgi[d] = dgi[d]; dirs.dgh[d] = dgh[d];
  }
  return dispatch_gru<true>(seq_len, B, T, H, ndir, dirs, h_stride, (cudaStream_t)stream);
}
