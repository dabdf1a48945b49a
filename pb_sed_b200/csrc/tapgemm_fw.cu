// tapgemm_fw.cu -- "frequency-walking" tcgen05 tap-GEMM for the NARROW 3x3 conv layers (16 / 32 channels)
// of the CNN2d stack (padertorch CNN2d layer bodies, pb_sed/experiments/weak_label_crnn/training.py:158-169,
// 218-230: layers 16->16 @ F128, 16->32 @ F64, 32->32 @ F64) and their data gradients.
//
// Why a second kernel: with N = Cout = 16 the generic kernel (tapgemm_tc.cu) issues M128 x N16 x K8 MMAs whose
// A-operand fetch (4 KB of shared memory per instruction) bounds the tensor pipe at 1/4 of its rate, re-reads
// every input row for three output rows, and runs 16 k short-lived CTAs whose load -> convert -> MMA -> epilogue
// chain never overlaps itself (measured r02: 277 us for 262 MB of algorithmic traffic, DRAM at 15 %, tensor pipe
// at 14 %).  Here one PERSISTENT CTA per SM walks the frequency axis of a (clip, 128-frame) tile:
//   * every input strip (row f, 130 frames incl. halo, all Cin channels: one CONTIGUOUS block of the (B,F,T,C) map)
//     is fetched ONCE -- by a 1-D bulk copy (cp.async.bulk -> UBLKCP) into a 4/8-deep raw ring; a tensor-map
//     box with 64-byte rows measured 3x slower, the TMA unit pays per row -- converted once (norm + ReLU + mask +
//     hi/lo split) and
//     feeds the THREE output rows f-1, f, f+1 with one MMA per time tap: the weights of df = +1 / 0 / -1 are
//     stacked along N (N = 3 x Cout = 48 / 96), the three rows' accumulators are adjacent TMEM columns;
//   * a whole layer's weight image (<= 72 KB incl. hi/lo) stays resident in shared memory;
//   * accumulators for R = 256 / Cout output rows live in one half of TMEM, the other half belongs to the
//     neighbouring work item, so the epilogue (TMEM -> registers -> bias / ReLU-mask -> 64-byte row stores)
//     of finished rows overlaps the MMAs of the next rows and of the next item; drained columns are re-zeroed
//     by the epilogue (tcgen05.st), so every MMA accumulates.
// Same arithmetic as tapgemm_tc.cu (3xTF32 split or one TF32 pass, fp32 accumulation), same tap-GEMM contract
// (include/pbsed_b200.h); fused column sums are left to the caller's separate passes.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cstdlib>

namespace {

constexpr int FW_TM = 128;                 // frames per tile (= MMA M)
constexpr int FW_ROWS = FW_TM + 2;         // strip rows incl. one halo frame each side
constexpr int FW_KB = 16, FW_KCH = 4;      // channels / 16-byte chunks per stage
constexpr int FW_NA = 3;                   // converted operand stages
constexpr int FW_NRAW = 8;                 // raw (bulk-copy) stages at Cin = 16; 4 at Cin = 32 (same bytes)
constexpr int FW_MAXR = 16;                // output rows per TMEM half (Cout = 16)
constexpr uint32_t FW_A_LBO = FW_ROWS * 16;
constexpr uint32_t FW_A_PART = FW_KCH * FW_A_LBO;      // 8320 B: one hi (or lo) strip stage
constexpr uint32_t FW_A_STAGE = 2 * FW_A_PART;
constexpr int FW_THREADS = 512;            // warps 0-3 converters, 4-7 + 10-13 two epilogue groups, 8 / 14 / 15 MMA issuers, 9 strip loader

struct FwParams {
  int B, F, T, Cin, Cout;
  int relu, single;
  int nkb, R;                 // channel blocks of 16; output rows per item (= 256 / Cout)
  int t_tiles, f_chunks, items;
  unsigned w_bytes;
  int in_bf16, out_bf16;      // storage of `in` / of `out` and `ep_src` (bf16 activation maps)
  int bfm;                    // 1: kind::f16 MMAs (bf16 operands, 8 channels per 16-byte chunk, all Cin channels in ONE stage)
  int nraw;                   // raw ring depth
  unsigned raw_stage;         // bytes of one raw strip: 130 frames x Cin elements (fp32 or bf16)
};

struct __align__(16) FwCtl {
  uint64_t raw_full[FW_NRAW], raw_empty[FW_NRAW], a_full[FW_NA], a_empty[FW_NA];
  uint64_t acc_full[2][FW_MAXR], acc_empty[2];
  uint32_t tmem_base;
  alignas(16) float colacc[2][32];   // per-CTA column sums of the fused statistics
  alignas(16) float cvec[5][32];          // per-channel epilogue constants: bias, ep_scale, ep_shift, ep_mean, ep_rstd
};
constexpr int FW_STAT_REP = 32;   // global accumulators are replicated: CTAs hash onto them, a fold kernel sums

// weight image: [dt 3][kb][part hi|lo][k-chunk 4][n' = blk * Cout + n][4 floats], blk 0/1/2 <-> df = +1/0/-1,
// i.e. the output rows f-1, f, f+1 a source strip f contributes to
__global__ void __launch_bounds__(256)
fwprep_kernel(const float* __restrict__ W, long long w_tap_stride, long long w_sn, long long w_sc,
              int Cin, int Cout, int t00, int t01, int t02, int t10, int t11, int t12, int t20, int t21, int t22,
              float* __restrict__ img, int single, double* __restrict__ rep, int rep_count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rep_count; i += gridDim.x * blockDim.x) rep[i] = 0.;
  const int tapidx[3][3] = {{t00, t01, t02}, {t10, t11, t12}, {t20, t21, t22}};   // [df+1][dt+1]
  const int nkb = Cin / FW_KB, N3 = 3 * Cout;
  const int total = 3 * nkb * FW_KCH * N3 * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 3;
    int q = i >> 2;
    const int np = q % N3; q /= N3;
    const int kc = q % FW_KCH; q /= FW_KCH;
    const int kb = q % nkb;
    const int dt = q / nkb;
    const int blk = np / Cout, n = np - blk * Cout;
    const int tap = tapidx[2 - blk][dt];                  // df = 1 - blk
    const int cin = kb * FW_KB + kc * 4 + e;
    const float w = __ldg(W + (long long)tap * w_tap_stride + (long long)n * w_sn + (long long)cin * w_sc);
    const float hi = tf32_rn(w);
    const long long part = (long long)FW_KCH * N3 * 4;
    const long long base = (long long)(dt * nkb + kb) * 2 * part + ((long long)kc * N3 + np) * 4 + e;
    img[base] = hi;
    img[base + part] = tf32_rn(w - hi);
  }
}

// bf16 weight image (kind::f16): [dt 3][k-chunk Cin/8][n' = blk * Cout + n][8 bf16]
__global__ void __launch_bounds__(256)
fwprep_bf16_kernel(const float* __restrict__ W, long long w_tap_stride, long long w_sn, long long w_sc,
                   int Cin, int Cout, int t00, int t01, int t02, int t10, int t11, int t12, int t20, int t21, int t22,
                   __nv_bfloat16* __restrict__ img, double* __restrict__ rep, int rep_count) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rep_count; i += gridDim.x * blockDim.x) rep[i] = 0.;
  const int tapidx[3][3] = {{t00, t01, t02}, {t10, t11, t12}, {t20, t21, t22}};   // [df+1][dt+1]
  const int nch = Cin / 8, N3 = 3 * Cout;
  const int total = 3 * nch * N3 * 8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i & 7;
    int q = i >> 3;
    const int np = q % N3; q /= N3;
    const int kc = q % nch;
    const int dt = q / nch;
    const int blk = np / Cout, n = np - blk * Cout;
    const int tap = tapidx[2 - blk][dt];
    const int cin = kc * 8 + e;
    img[i] = __float2bfloat16_rn(__ldg(W + (long long)tap * w_tap_stride + (long long)n * w_sn + (long long)cin * w_sc));
  }
}

__global__ void __launch_bounds__(FW_THREADS, 1)
tapgemm_fw_kernel(FwParams p, const float* __restrict__ in, const float* __restrict__ scale,
                  const float* __restrict__ shift, const int* __restrict__ seq_len,
                  const int* __restrict__ load_seq_len, const float* __restrict__ wimg,
                  const float* __restrict__ bias, float* __restrict__ out, const float* __restrict__ ep_src,
                  const float* __restrict__ ep_scale, const float* __restrict__ ep_shift,
                  double* __restrict__ out_stats, const float* __restrict__ ep_mean,
                  const float* __restrict__ ep_rstd, double* __restrict__ ep_sums) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* w_smem = smem_raw;
  uint8_t* a_smem = w_smem + p.w_bytes;
  uint8_t* r_smem = a_smem + FW_NA * FW_A_STAGE;
  FwCtl* ctl = reinterpret_cast<FwCtl*>(r_smem + p.nraw * p.raw_stage);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Cout = p.Cout, R = p.R;

  if (tid == 0) {
    for (int i = 0; i < p.nraw; ++i) { mbar_init(&ctl->raw_full[i], 1); mbar_init(&ctl->raw_empty[i], 128); }
    const uint32_t n_issuers = (p.single || p.bfm) ? 1u : 3u;
    for (int i = 0; i < FW_NA; ++i) { mbar_init(&ctl->a_full[i], 128); mbar_init(&ctl->a_empty[i], n_issuers); }
    for (int h = 0; h < 2; ++h) {
      for (int i = 0; i < FW_MAXR; ++i) mbar_init(&ctl->acc_full[h][i], n_issuers);
      mbar_init(&ctl->acc_empty[h], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) tmem_alloc(&ctl->tmem_base, 512);
  if (tid < 64) ctl->colacc[tid >> 5][tid & 31] = 0.f;
  if (tid >= 64 && tid < 224) {               // stage the per-channel epilogue vectors once (broadcast LDS later)
    const int which = (tid - 64) >> 5, ch = tid & 31;
    const float* srcv = which == 0 ? bias : which == 1 ? ep_scale : which == 2 ? ep_shift : which == 3 ? ep_mean : ep_rstd;
    ctl->cvec[which][ch] = (srcv && ch < Cout) ? __ldg(srcv + ch) : (which == 1 || which == 4 ? 1.f : 0.f);
  }
  for (unsigned i = tid; i < p.w_bytes / 16; i += FW_THREADS)          // the layer's weights: resident for the whole kernel
    reinterpret_cast<float4*>(w_smem)[i] = __ldg(reinterpret_cast<const float4*>(wimg) + i);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (warp >= 4 && warp < 8) {                // all 512 accumulator columns start at zero: every MMA accumulates (group 0 does it)
    for (int col = 0; col < 512; col += 16) tmem_st16_zero(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + col);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 4) {
    // ============================== converters ==============================
    const int c = tid & 3, r0 = tid >> 2;
    int it = 0, sit = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
      const int fc = item % p.f_chunks, tt = (item / p.f_chunks) % p.t_tiles, b = item / (p.f_chunks * p.t_tiles);
      const int fo0 = fc * R, t0 = tt * FW_TM;
      const int len_in = load_seq_len ? min(__ldg(load_seq_len + b), p.T) : p.T;
      const int f_lo = max(fo0 - 1, 0), f_hi = min(fo0 + R, p.F - 1);
      for (int f = f_lo; f <= f_hi; ++f, ++sit) {
        const int rs = sit % p.nraw;
        const uint8_t* raw = r_smem + rs * p.raw_stage;
        mbar_wait(&ctl->raw_full[rs], (sit / p.nraw) & 1);             // the raw strip has landed
        if (p.bfm) {
          // bf16 operands: the whole strip (Cin = 16 / 32 channels = 2 / 4 chunks of 8) is ONE stage
          const int nch = p.Cin >> 3, cb = tid % nch, rb = tid / nch, rstep = 128 / nch;
          float4 s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0, h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
          if (scale) {
            s0 = __ldg(reinterpret_cast<const float4*>(scale + cb * 8)); s1 = __ldg(reinterpret_cast<const float4*>(scale + cb * 8 + 4));
            h0 = __ldg(reinterpret_cast<const float4*>(shift + cb * 8)); h1 = __ldg(reinterpret_cast<const float4*>(shift + cb * 8 + 4));
          }
          const int slot = it % FW_NA;
          uint8_t* img = a_smem + slot * FW_A_STAGE;
          mbar_wait(&ctl->a_empty[slot], ((it / FW_NA) & 1) ^ 1);
          for (int r = rb; r < FW_ROWS; r += rstep) {
            const int t = t0 + r - 1;
            uint4 o4 = make_uint4(0u, 0u, 0u, 0u);
            if (t >= 0 && t < len_in) {
              o4 = *reinterpret_cast<const uint4*>(raw + 2 * (uint32_t)(r * p.Cin + cb * 8));
              if (scale || p.relu) {
                float4 x0 = bf16x4_to_float4(make_uint2(o4.x, o4.y)), x1 = bf16x4_to_float4(make_uint2(o4.z, o4.w));
                if (scale) {
                  x0.x = fmaf(x0.x, s0.x, h0.x); x0.y = fmaf(x0.y, s0.y, h0.y); x0.z = fmaf(x0.z, s0.z, h0.z); x0.w = fmaf(x0.w, s0.w, h0.w);
                  x1.x = fmaf(x1.x, s1.x, h1.x); x1.y = fmaf(x1.y, s1.y, h1.y); x1.z = fmaf(x1.z, s1.z, h1.z); x1.w = fmaf(x1.w, s1.w, h1.w);
                }
                if (p.relu) {
                  x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
                  x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
                }
                const uint2 lo2 = float4_to_bf16x4(x0), hi2 = float4_to_bf16x4(x1);
                o4 = make_uint4(lo2.x, lo2.y, hi2.x, hi2.y);
              }
            }
            *reinterpret_cast<uint4*>(img + (uint32_t)(cb * FW_ROWS + r) * 16) = o4;
          }
          fence_async_smem();
          mbar_arrive(&ctl->a_full[slot]);
          ++it;
          mbar_arrive(&ctl->raw_empty[rs]);
          continue;
        }
        for (int kb = 0; kb < p.nkb; ++kb, ++it) {
          float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
          if (scale) {
            sc = __ldg(reinterpret_cast<const float4*>(scale + kb * FW_KB + c * 4));
            sh = __ldg(reinterpret_cast<const float4*>(shift + kb * FW_KB + c * 4));
          }
          const int slot = it % FW_NA;
          uint8_t* hi_base = a_smem + slot * FW_A_STAGE;
          uint8_t* lo_base = hi_base + FW_A_PART;
          mbar_wait(&ctl->a_empty[slot], ((it / FW_NA) & 1) ^ 1);
#pragma unroll
          for (int u = 0; u < (FW_ROWS + 31) / 32; ++u) {
            const int r = r0 + 32 * u, t = t0 + r - 1;
            if (r >= FW_ROWS) break;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < len_in) {
              const uint32_t e = (uint32_t)(r * p.Cin + kb * FW_KB + c * 4);
              x = p.in_bf16 ? bf16x4_to_float4(*reinterpret_cast<const uint2*>(raw + 2 * e))
                            : *reinterpret_cast<const float4*>(raw + 4 * e);
              if (scale) {
                x.x = fmaf(x.x, sc.x, sh.x); x.y = fmaf(x.y, sc.y, sh.y);
                x.z = fmaf(x.z, sc.z, sh.z); x.w = fmaf(x.w, sc.w, sh.w);
              }
              if (p.relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            }
            const uint32_t o = (uint32_t)(c * FW_ROWS + r) * 16;
            if (p.single) {
              *reinterpret_cast<float4*>(hi_base + o) = make_float4(tf32_rn(x.x), tf32_rn(x.y), tf32_rn(x.z), tf32_rn(x.w));
            } else {
              float4 h;
              h.x = tf32_rn(x.x);
              h.y = tf32_rn(x.y);
              h.z = tf32_rn(x.z);
              h.w = tf32_rn(x.w);
              *reinterpret_cast<float4*>(hi_base + o) = h;
              *reinterpret_cast<float4*>(lo_base + o) = make_float4(tf32_rn(x.x - h.x), tf32_rn(x.y - h.y), tf32_rn(x.z - h.z), tf32_rn(x.w - h.w));
            }
          }
          fence_async_smem();
          mbar_arrive(&ctl->a_full[slot]);
        }
        mbar_arrive(&ctl->raw_empty[rs]);              // every channel block of the strip is converted
      }
    }
  } else if (warp < 8 || (warp >= 10 && warp < 14)) {
    // ============================== epilogue ==============================
    // Work is flattened into (output row, 16-column block) steps, dealt alternately to TWO groups of 4 warps
    // (ncu r02: one group was busy 83 % of the time while converters and the tensor pipe waited for it): with
    // Cout = 32 a group owns one 16-column half of every row, with Cout = 16 every other row -- either way a
    // thread sees ONE fixed 16-channel block, so its column sums are 2 x 16 registers.  The ReLU-mask source of
    // the group's next step is in flight while the current one is processed; per-channel vectors come from
    // shared memory.
    const int ew = warp & 3;                   // TMEM lane quarter (hardware: warp id mod 4)
    const int grp = warp >= 10 ? 1 : 0;
    const int row = ew * 32 + lane;
    const bool want_sums = out_stats != nullptr || ep_sums != nullptr;
    const int CH = Cout >> 4;
    const int cc = CH == 2 ? grp * 16 : 0;     // this thread's channel block
    float s0[16], s1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { s0[i] = 0.f; s1[i] = 0.f; }
    int li = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++li) {
      const int fc = item % p.f_chunks, tt = (item / p.f_chunks) % p.t_tiles, b = item / (p.f_chunks * p.t_tiles);
      const int fo0 = fc * R, t = tt * FW_TM + row;
      const int h = li & 1;
      const int len_b = seq_len ? min(__ldg(seq_len + b), p.T) : p.T;
      const bool in_map = t < p.T, valid = t < len_b;
      const bool use_src = ep_src != nullptr && valid;
      const long long orow0 = ((long long)b * p.F + fo0) * p.T + t;
      const int nsteps = R * CH;
      float4 nxt[4];
      if (use_src && grp < nsteps) {
        const long long sp = (orow0 + (long long)(grp / CH) * p.T) * Cout + cc;
#pragma unroll
        for (int q = 0; q < 4; ++q) nxt[q] = ld_act4(ep_src, sp + 4 * q, p.out_bf16);
      }
      for (int step = grp; step < nsteps; step += 2) {
        const int r = step / CH;
        const long long obase = (orow0 + (long long)r * p.T) * Cout + cc;
        float4 cur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
        if (use_src && step + 2 < nsteps) {
          const long long sp = (orow0 + (long long)((step + 2) / CH) * p.T) * Cout + cc;
#pragma unroll
          for (int q = 0; q < 4; ++q) nxt[q] = ld_act4(ep_src, sp + 4 * q, p.out_bf16);
        }
        mbar_wait(&ctl->acc_full[h][r], (li >> 1) & 1);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(h * 256 + r * Cout + cc);
        float v[16];
        tmem_ld16(ta, v);
        tmem_st16_zero(ta);                            // the column block is free for the item after next
        if (in_map) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            const float4 bb = *reinterpret_cast<const float4*>(&ctl->cvec[0][cc + j]);
            o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            const float4 x = cur[j >> 2];
            if (ep_src) {                              // ReLU / sequence mask of a data-gradient pass
              float4 sv = make_float4(0.f, 0.f, 0.f, 0.f);
              if (valid) {
                const float4 es = *reinterpret_cast<const float4*>(&ctl->cvec[1][cc + j]);
                const float4 eh = *reinterpret_cast<const float4*>(&ctl->cvec[2][cc + j]);
                sv.x = fmaf(x.x, es.x, eh.x); sv.y = fmaf(x.y, es.y, eh.y);
                sv.z = fmaf(x.z, es.z, eh.z); sv.w = fmaf(x.w, es.w, eh.w);
              }
              o.x = sv.x > 0.f ? o.x : 0.f; o.y = sv.y > 0.f ? o.y : 0.f;
              o.z = sv.z > 0.f ? o.z : 0.f; o.w = sv.w > 0.f ? o.w : 0.f;
            }
            st_act4(out, obase + j, o, p.out_bf16);
            if (want_sums && valid) {
              float4 w2 = o;                           // out_stats: sum, sum of squares
              if (!out_stats) {                        // ep_sums: sum g, sum g * xhat
                const float4 mu = *reinterpret_cast<const float4*>(&ctl->cvec[3][cc + j]);
                const float4 rs = *reinterpret_cast<const float4*>(&ctl->cvec[4][cc + j]);
                w2 = make_float4((x.x - mu.x) * rs.x, (x.y - mu.y) * rs.y, (x.z - mu.z) * rs.z, (x.w - mu.w) * rs.w);
              }
              s0[j] += o.x; s0[j + 1] += o.y; s0[j + 2] += o.z; s0[j + 3] += o.w;
              s1[j] = fmaf(o.x, w2.x, s1[j]); s1[j + 1] = fmaf(o.y, w2.y, s1[j + 1]);
              s1[j + 2] = fmaf(o.z, w2.z, s1[j + 2]); s1[j + 3] = fmaf(o.w, w2.w, s1[j + 3]);
            }
          }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&ctl->acc_empty[h]);
    }
    if (want_sums) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a0 = warp_sum(s0[i]), a1 = warp_sum(s1[i]);
        if (lane == 0) { atomicAdd(&ctl->colacc[0][cc + i], a0); atomicAdd(&ctl->colacc[1][cc + i], a1); }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int et = tid - 128;                 // group 0's threads flush
      if (et >= 0 && et < Cout) {
        double* dst = (out_stats ? out_stats : ep_sums) + ((long long)(blockIdx.x % FW_STAT_REP) * Cout + et) * 2;
        atomicAdd(dst, (double)ctl->colacc[0][et]);
        atomicAdd(dst + 1, (double)ctl->colacc[1][et]);
      }
    }
  } else if (warp == 8 || warp >= 14) {
    // ============================== MMA issuers ==============================
    // One thread feeding tcgen05.mma spends ~100 cycles of dependent scalar work per instruction (ncu r02: the
    // issuing warp never idles while the tensor pipe is 17-30 % busy), far more than an M128 x N48 x K8 MMA runs.
    // The three passes of the split (hi*hi, lo*hi, hi*lo) are therefore issued by THREE threads in three warps,
    // each with fixed operand images; accumulation order is irrelevant, every barrier counts all issuers.
    const int pass = warp == 8 ? 0 : warp - 13;               // 0: hi*hi, 1: lo*hi, 2: hi*lo
    if (lane == 0 && (pass == 0 || !(p.single || p.bfm))) {
      uint32_t idesc[4];
      for (int n = 1; n <= 3; ++n) idesc[n] = make_idesc_tf32(FW_TM, n * Cout);
      const uint32_t W_LBO = (uint32_t)(3 * Cout) * 16, W_PART = FW_KCH * W_LBO;
      const uint32_t w_base = smem_u32(w_smem), a_base = smem_u32(a_smem);
      const uint32_t w_dt_stride = (uint32_t)(p.nkb * 2) * W_PART;
      const uint64_t a_d0 = make_desc(0, FW_A_LBO, 128), w_d0 = make_desc(0, W_LBO, 128);
      int it = 0, li = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++li) {
        const int fc = item % p.f_chunks;
        const int fo0 = fc * R, h = li & 1;
        const int f_lo = max(fo0 - 1, 0), f_hi = min(fo0 + R, p.F - 1);
        mbar_wait(&ctl->acc_empty[h], ((li >> 1) & 1) ^ 1);          // the epilogue drained + re-zeroed this half
        tc_fence_after();
        for (int f = f_lo; f <= f_hi; ++f) {
          const int rlo = max(f - 1, fo0), rhi = min(f + 1, fo0 + R - 1);
          const int blk0 = rlo - (f - 1), nblk = rhi - rlo + 1;
          const uint32_t d = tmem_base + (uint32_t)(h * 256 + (rlo - fo0) * Cout);
          for (int kb = 0; kb < p.nkb; ++kb, ++it) {
            const int slot = it % FW_NA;
            mbar_wait(&ctl->a_full[slot], (it / FW_NA) & 1);
            tc_fence_after();
            const uint32_t a_hi = a_base + slot * FW_A_STAGE, a_lo = a_hi + FW_A_PART;
            if (p.bfm) {                 // kind::f16: K = 16 = two 8-channel chunks per instruction, one pass
              const uint32_t nch = (uint32_t)p.Cin >> 3, idb = make_idesc_bf16(FW_TM, nblk * Cout);
              for (uint32_t dt = 0; dt < 3; ++dt)
                for (uint32_t ks = 0; ks < nch / 2; ++ks)
                  mma_bf16(d, desc_at(a_d0, a_hi + dt * 16 + (ks * 2) * FW_A_LBO),
                           desc_at(w_d0, w_base + (dt * nch + ks * 2) * W_LBO + (uint32_t)(blk0 * Cout) * 16), idb, 1u);
              mma_commit(&ctl->a_empty[slot]);
              continue;
            }
            const uint32_t w_kb = w_base + (uint32_t)(kb * 2) * W_PART + (uint32_t)(blk0 * Cout) * 16;
            const uint32_t id = idesc[nblk];
#pragma unroll
            for (int dt = 0; dt < 3; ++dt) {
              const uint32_t w_hi = w_kb + (uint32_t)dt * w_dt_stride, w_lo = w_hi + W_PART;
#pragma unroll
              for (int ks = 0; ks < FW_KB / 8; ++ks) {
                const uint32_t ao = (uint32_t)dt * 16 + (uint32_t)(ks * 2) * FW_A_LBO, wo = (uint32_t)(ks * 2) * W_LBO;
                mma_tf32(d, desc_at(a_d0, (pass == 1 ? a_lo : a_hi) + ao), desc_at(w_d0, (pass == 2 ? w_lo : w_hi) + wo), id, 1u);
              }
            }
            mma_commit(&ctl->a_empty[slot]);
          }
          // strip f was the last contribution to output row f-1 (and, at the bottom of the map, to row f)
          if (f - 1 >= fo0) mma_commit(&ctl->acc_full[h][f - 1 - fo0]);
          if (f == f_hi)
            for (int r = max(f, fo0); r <= fo0 + R - 1; ++r) mma_commit(&ctl->acc_full[h][r - fo0]);
        }
      }
    }
  } else {
    // ============================== strip loader ==============================
    // one 1-D bulk copy per strip: frames [max(t0-1, 0), min(t0+128, T-1)] of row group (b, f) are contiguous;
    // the copy never leaves the row group, the converters never read the rows it left untouched
    if (lane == 0) {
      int sit = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
        const int fc = item % p.f_chunks, tt = (item / p.f_chunks) % p.t_tiles, b = item / (p.f_chunks * p.t_tiles);
        const int fo0 = fc * R, t0 = tt * FW_TM;
        const int f_lo = max(fo0 - 1, 0), f_hi = min(fo0 + R, p.F - 1);
        const int ta = max(t0 - 1, 0), tb = min(t0 + FW_TM, p.T - 1);
        const uint32_t esz = p.in_bf16 ? 2u : 4u;
        const uint32_t bytes = (uint32_t)(tb - ta + 1) * p.Cin * esz;
        for (int f = f_lo; f <= f_hi; ++f, ++sit) {
          const int rs = sit % p.nraw;
          mbar_wait(&ctl->raw_empty[rs], ((sit / p.nraw) & 1) ^ 1);
          mbar_expect_tx(&ctl->raw_full[rs], bytes);
          bulk_g2s(r_smem + rs * p.raw_stage + (uint32_t)(ta - (t0 - 1)) * p.Cin * esz,
                   reinterpret_cast<const uint8_t*>(in) + (((long long)b * p.F + f) * p.T + ta) * p.Cin * esz, bytes,
                   &ctl->raw_full[rs]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void fw_stat_fold_kernel(const double* __restrict__ rep, int n, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.;
  for (int r = 0; r < FW_STAT_REP; ++r) s += rep[(long long)r * n + i];
  dst[i] += s;
}

}  // namespace

// *handled = 1: `out` holds the finished map (bias / ReLU-mask epilogue applied) and the fused column sums
// (out_stats / ep_sums, when asked for) have been added to the caller's accumulators.
int tapgemm_fw_dispatch(const pbsed_tapgemm_desc* d, const float* in, const float* scale, const float* shift,
                        const int* seq_len, const float* W, const float* bias, float* out, const float* ep_src,
                        const float* ep_scale, const float* ep_shift, double* out_stats, const float* ep_mean,
                        const float* ep_rstd, double* ep_sums, void* workspace, long long ws_bytes,
                        cudaStream_t st, int* handled) {
  *handled = 0;
  if (out_stats && ep_sums) return 0;
  if (ep_sums && (!ep_src || !ep_mean || !ep_rstd)) return 0;
  static const int enabled = getenv("PBSED_FW") ? atoi(getenv("PBSED_FW")) : 1;
  if (!enabled || (d->precision != 1 && d->precision != 3) || !workspace) return 0;
  if (d->ntaps != 9 || d->F_in != d->F_out || d->per_f) return 0;
  if ((d->Cin != 16 && d->Cin != 32) || (d->Cout != 16 && d->Cout != 32)) return 0;
  if ((d->in_stride > 0 && d->in_stride != d->Cin) || (d->out_stride > 0 && d->out_stride != d->Cout)) return 0;
  int tapidx[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) tapidx[i][j] = -1;
  for (int i = 0; i < 9; ++i) {
    if (d->df[i] < -1 || d->df[i] > 1 || d->dt[i] < -1 || d->dt[i] > 1) return 0;
    tapidx[d->df[i] + 1][d->dt[i] + 1] = i;
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) if (tapidx[i][j] < 0) return 0;
  FwParams p = {};
  p.B = d->B; p.F = d->F_in; p.T = d->T; p.Cin = d->Cin; p.Cout = d->Cout;
  p.relu = d->relu; p.single = d->precision == 3;
  p.nkb = d->Cin / FW_KB;
  p.R = 256 / d->Cout;
  if (p.F % p.R) return 0;
  p.t_tiles = cdiv(p.T, FW_TM);
  p.f_chunks = p.F / p.R;
  const long long items = (long long)p.B * p.t_tiles * p.f_chunks;
  const long long rows = (long long)p.B * p.F * p.T;
  if (items > (1 << 30) || rows >= (1LL << 31) - 256) return 0;
  p.items = (int)items;
  static const int use_bfm = getenv("PBSED_BF16_MMA") ? atoi(getenv("PBSED_BF16_MMA")) : 1;
  p.bfm = use_bfm && d->in_dtype == PBSED_BF16 && d->precision == 3;
  if (p.bfm) p.nkb = 1;
  p.w_bytes = p.bfm ? 3u * (p.Cin / 8) * (3u * p.Cout) * 16u : 3u * p.nkb * 2u * FW_KCH * (3u * p.Cout) * 16u;
  const bool want_sums = out_stats || ep_sums;
  const long long rep_bytes = (long long)FW_STAT_REP * p.Cout * 2 * sizeof(double);
  if ((long long)p.w_bytes + 256 + (want_sums ? rep_bytes : 0) > ws_bytes) return 0;
  double* rep = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + ((p.w_bytes + 255) / 256) * 256);
  if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)workspace | (uintptr_t)bias | (uintptr_t)scale | (uintptr_t)shift |
        (uintptr_t)ep_src | (uintptr_t)ep_scale | (uintptr_t)ep_shift | (uintptr_t)ep_mean | (uintptr_t)ep_rstd) & 15) != 0)
    return 0;
  p.in_bf16 = d->in_dtype == PBSED_BF16; p.out_bf16 = d->out_dtype == PBSED_BF16;
  p.raw_stage = (unsigned)FW_ROWS * p.Cin * 4;                    // sized for fp32; bf16 strips use half of it
  p.nraw = FW_NRAW * FW_KB / p.Cin;
  float* img = reinterpret_cast<float*>(workspace);
  if (p.bfm)
    fwprep_bf16_kernel<<<cdiv(p.w_bytes / 2, 256), 256, 0, st>>>(W, d->w_tap_stride, d->w_sn, d->w_sc, p.Cin, p.Cout,
        tapidx[0][0], tapidx[0][1], tapidx[0][2], tapidx[1][0], tapidx[1][1], tapidx[1][2],
        tapidx[2][0], tapidx[2][1], tapidx[2][2], reinterpret_cast<__nv_bfloat16*>(img), rep, want_sums ? FW_STAT_REP * p.Cout * 2 : 0);
  else
  fwprep_kernel<<<cdiv(p.w_bytes / 8, 256), 256, 0, st>>>(W, d->w_tap_stride, d->w_sn, d->w_sc, p.Cin, p.Cout,
      tapidx[0][0], tapidx[0][1], tapidx[0][2], tapidx[1][0], tapidx[1][1], tapidx[1][2],
      tapidx[2][0], tapidx[2][1], tapidx[2][2], img, p.single, rep, want_sums ? FW_STAT_REP * p.Cout * 2 : 0);
  int rc = pbsed_after_launch();
  if (rc) return rc;
  const size_t smem = (size_t)p.w_bytes + FW_NA * FW_A_STAGE + (size_t)p.nraw * p.raw_stage + sizeof(FwCtl) + 128;
  cudaError_t e = cudaFuncSetAttribute(tapgemm_fw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  const int grid = p.items < n_sm ? p.items : n_sm;
  pbsed_note_kernel("tapgemm_fw_kernel");
  tapgemm_fw_kernel<<<grid, FW_THREADS, smem, st>>>(p, in, scale, shift, seq_len,
                                                   d->no_input_mask ? nullptr : seq_len, img, bias, out, ep_src,
                                                   ep_scale, ep_shift, out_stats ? rep : nullptr, ep_mean, ep_rstd,
                                                   ep_sums ? rep : nullptr);
  *handled = 1;
  rc = pbsed_after_launch();
  if (rc || !want_sums) return rc;
  fw_stat_fold_kernel<<<cdiv(2 * p.Cout, 64), 64, 0, st>>>(rep, 2 * p.Cout, out_stats ? out_stats : ep_sums);
  return pbsed_after_launch();
}
