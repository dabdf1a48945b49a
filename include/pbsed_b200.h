/* pbsed_b200.h -- C ABI of libpbsed_b200.so (hand-written sm_100a kernels for the
 * pb_sed FBCRNN / BiCRNN hot path).
 *
 * Conventions:
 *   - every entry point returns 0 on success, a negative PBSED_E* code for a bad
 *     argument, or a positive cudaError_t for a launch failure; nothing throws;
 *   - nothing allocates, frees or synchronises: the caller owns every buffer and
 *     passes the cudaStream_t (as void*) the work is enqueued on;
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *   - activations are fp32 "rows x channels" matrices, rows = (b, f, t) in that
 *     order, channels contiguous:  2-D maps (B,F,T,C), 1-D maps (B,T,C) (F = 1).
 *     The reference's (B,C,F,T) log-mel with C = 1 is the same memory as (B,F,T,1).
 *   - seq_len (int32[B], nullable = all frames valid) carries the reference's
 *     sequence-length masking (padertorch Normalization / Mean / pack_padded).
 *
 * Each function names the reference interface it replaces; pb_sed paths are
 * relative to /root/reference, third-party ones are padertorch@b7ba24a /
 * paderbox@809b272 (pinned in README.md:40-41; source not vendored).
 */
#ifndef PBSED_B200_H
#define PBSED_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBSED_EINVAL   (-1)   /* bad argument / unsupported shape            */
#define PBSED_EWORKSPACE (-2) /* workspace too small                         */
#define PBSED_MAX_TAPS 16
#define PBSED_F32  0          /* activation storage types (the *_dtype arguments below) */
#define PBSED_BF16 1

/* library identification: returns the ABI version (bumped on signature change) */
int pbsed_abi_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long pbsed_launch_count(void);
/* name of the main kernel the most recent entry point dispatched (static string; bench.py's per-kernel
 * roofline pass groups its per-call device timings by it) */
const char* pbsed_last_kernel(void);

/* ---------------------------------------------------------------------------
 * K1  STFT -> |.|^2 -> mel -> log      (replaces paderbox stft as configured at
 *     pb_sed/data_preparation/provider.py:315-323 and called at
 *     pb_sed/data_preparation/transform.py:53, plus MelTransform inside
 *     NormalizedLogMelExtractor, call site pb_sed/models/weak_label/crnn.py:86-90)
 *
 * audio (B, S) fp32 -> logmel (B, n_mels, T) fp32 = log(mel_power + 1e-18).
 * fading='half' zero padding, pad=True tail, periodic Blackman window of
 * `window_length`, zero padded to `fft_size` (power of two, <= 2048).
 * window:  (window_length) fp32.   fbank_lo/fbank_hi: int32[n_mels] first/last+1
 * fft bin of each triangular filter, fbank_w: (n_mels, fbank_stride) fp32 weights
 * for bins lo..hi-1 (row-normalised HTK-mel triangles).
 * stats (nullable): double[n_mels][2], accumulates sum / sum-of-squares of logmel
 * over valid frames (t < seq_len[b]) for the cumulative running normalisation.
 * fbank_per_clip != 0: every clip has its own filterbank (train-time MelWarping,
 * training.py:195-208): fbank_lo/hi are (B, n_mels), fbank_w is (B, n_mels, fbank_stride).
 * frame_start (nullable, int32 (B, T)): first sample of every frame in the 'fading'-padded
 * signal instead of t*shift -- the non-uniform frame grid of TimeWarpedSTFT
 * (pb_sed/data_preparation/transform.py:36-45, samplers provider.py:329-338).
 */
int pbsed_stft_logmel(const float* audio, int B, int S,
                      int shift, int window_length, int fft_size, int pad_front, int T,
                      const float* window,
                      const int* fbank_lo, const int* fbank_hi, const float* fbank_w,
                      int fbank_stride, int n_mels, int fbank_per_clip, const int* frame_start,
                      const int* seq_len, float* logmel, double* stats, void* stream);

/* power spectrogram input variant (reference-compatible 5-D `stft` input,
 * (B, T, n_bins, 2) fp32 re/im):  same outputs as above. */
int pbsed_spec_logmel(const float* stft, int B, int T, int n_bins,
                      const int* fbank_lo, const int* fbank_hi, const float* fbank_w,
                      int fbank_stride, int n_mels, int fbank_per_clip,
                      const int* seq_len, float* logmel, double* stats, void* stream);

/* per-example warped HTK-mel filterbanks (paderbox MelWarping as configured at
 * pb_sed/experiments/weak_label_crnn/training.py:195-208; piecewise-linear warp of the filter edge
 * frequencies in the mel domain, see stft_logmel.cu).  alpha / ratio: device float[B] (warp factor,
 * boundary-frequency ratio).  mel_lo / mel_hi: mel of the lowest / highest filter edge, mel_warp_hi:
 * mel of MelWarping.highest_frequency, bins_per_hz = fft_size / sample_rate.  Writes the sparse
 * tables pbsed_stft_logmel reads with fbank_per_clip = 1. */
int pbsed_make_warped_fbank(const float* alpha, const float* ratio, int B, int n_mels, int n_bins,
                            double mel_lo, double mel_hi, double mel_warp_hi, double bins_per_hz,
                            int* fbank_lo, int* fbank_hi, float* fbank_w, int fbank_stride,
                            void* stream);

/* normalise + clamp + mask in place:  x = clamp(x * scale[f] + shift[f], +-clamp) * (t < seq_len[b])
 * with scale = rsqrt(var+eps), shift = -mean*scale from pbsed_norm_finalize
 * (padertorch Normalization('bcft', statistics_axis='bt', no affine) + clamp(+-6)) */
/* train-time augmentation fused into the same pass (all nullable / 0; kwargs at
 * training.py:210-216): time_masks int32 (B, n_time_masks, 2) = (onset frame, width),
 * freq_masks int32 (B, n_freq_masks, 2) = (onset band, width) -> zeroed after the clamp;
 * then x += noise_scale[b] * noise (noise (B,F,T) standard normal). */
int pbsed_logmel_normalize(float* x, int B, int F, int T, const float* scale, const float* shift,
                           float clamp, const int* seq_len,
                           const int* time_masks, int n_time_masks, const int* freq_masks, int n_freq_masks,
                           const float* noise, const float* noise_scale, void* stream);

/* ---------------------------------------------------------------------------
 * K2  tap-GEMM: the one contraction behind CNN2d / CNN1d / GRU projections / output_net
 *     (replaces padertorch.contrib.je.modules.conv.{CNN2d,CNN1d} layer bodies,
 *     configured at pb_sed/experiments/weak_label_crnn/training.py:218-242, and the
 *     nn.GRU input projection / output_net 1x1 convs, training.py:243-260)
 *
 *   out[(b,fo,t), n] = bias[n] + sum_tap sum_c  a[(b, fo+df[tap], t+dt[tap]), c] * W[tap][n][c]
 *   a[(b,f,t), c]    = act( in[(b,f,t), c] * scale[.] + shift[.] )  if 0<=f<F_in, 0<=t<T, t<seq_len[b]
 *                    = 0 otherwise   ('same' zero padding happens AFTER norm+ReLU: pre-activation)
 *   scale/shift index = c (per_f == 0) or f*Cin + c (per_f == 1); scale == NULL -> identity.
 *   relu != 0 -> act = max(.,0).
 *   W layout: [ntaps][w_tap_stride] with element (n, c) at n*w_sn + c*w_sc  (so forward uses
 *   w_sn = Cin, w_sc = 1 and the data-gradient pass reuses the same weights transposed with
 *   w_sn = 1, w_sc = Cout_fwd and negated taps).
 *   epilogue (nullable extras):
 *     out_mul_relu_of: if given (same shape as out) with ep_scale/ep_shift, multiplies the
 *       result by [ (v*ep_scale+ep_shift) > 0 ] * (t < seq_len[b])  -- the ReLU/mask backward of a
 *       pre-activation layer when this call is a data-gradient pass.
 */
typedef struct {
  int B, F_in, F_out, T;
  int Cin, Cout;
  int ntaps;
  int df[PBSED_MAX_TAPS];
  int dt[PBSED_MAX_TAPS];
  int relu;          /* ReLU on the loaded operand               */
  int per_f;         /* scale/shift (and ep_*) indexed by f*C + c */
  long long w_tap_stride, w_sn, w_sc;
  int in_stride;     /* row stride of `in` in floats  (0 -> Cin;  lets a GRU direction read its half of a (B,T,2H) map) */
  int out_stride;    /* row stride of `out` / `dout` / `ep_src` in floats (0 -> Cout) */
  int precision;     /* 0 = exact fp32 FFMA; 1 = 3xTF32 tcgen05 (fp32-equivalent split);
                        3 = one TF32 tcgen05 pass, fp32 accumulation (reduced precision, >= bf16 mantissa) */
  int no_input_mask; /* 1: seq_len does NOT zero the loaded operand (a bare conv reads its input unmasked,
                        the reference masks inside Normalization only); seq_len still masks the fused
                        statistics / epilogue */
  int in_dtype;      /* storage type of `in`:  PBSED_F32 (0) or PBSED_BF16 (1).  Activation maps of the conv
                        stacks (and their gradients) may live in HBM as bf16 (BASELINE "bf16" configurations);
                        arithmetic and accumulation stay as `precision` says, weights / statistics are fp32 */
  int out_dtype;     /* storage type of `out` and `ep_src` (pbsed_tapgemm) / of `dout` (pbsed_tapgemm_wgrad).
                        bf16 maps are handled by the tensor-core kernels only (channel counts they accept) */
} pbsed_tapgemm_desc;

int pbsed_tapgemm(const pbsed_tapgemm_desc* d_host,
                  const float* in, const float* scale, const float* shift, const int* seq_len,
                  const float* W, const float* bias, float* out,
                  const float* ep_src, const float* ep_scale, const float* ep_shift,
                  double* out_stats, const float* ep_mean, const float* ep_rstd, double* ep_sums,
                  void* workspace, long long workspace_bytes, void* stream);
/* fused column reductions (all nullable; they save a full pass over the output map):
 *   out_stats[idx][0..1] += sum / sum of squares of `out` over frames t < seq_len[b]
 *                            (idx = n, or fo*Cout + n when per_f) -- the next layer's batch statistics
 *   ep_sums[idx][0..1]   += sum out, sum out * (ep_src - ep_mean[idx]) * ep_rstd[idx]
 *                            -- pass 1 of the batch-norm backward when this call is a data-gradient pass */
/* bytes of caller-owned scratch the tensor-core path needs (pre-tiled hi/lo weight image);
 * 0 for precision 0.  A NULL / too small workspace with precision != 0 runs the exact-fp32 kernel. */
long long pbsed_tapgemm_workspace_bytes(const pbsed_tapgemm_desc* d_host);

/* weight gradient of the same contraction:
 *   dW[tap][n][c] += sum_{b,fo,t}  dout[(b,fo,t), n] * a[(b, fo+df, t+dt), c]       (a as above)
 *   dbias[n]      += sum_{b,fo,t}  dout[(b,fo,t), n]                                (nullable)
 * dW uses the forward layout (w_sn = Cin, w_sc = 1).  ACCUMULATES: caller zeroes the gradient arena.
 * rows with t >= seq_len[b] of dout are ignored when mask_out != 0. */
int pbsed_tapgemm_wgrad(const pbsed_tapgemm_desc* d_host,
                        const float* in, const float* scale, const float* shift, const int* seq_len,
                        const float* dout, int mask_out, float* dW, float* dbias, void* stream);

/* ---------------------------------------------------------------------------
 * Normalisation / pooling helpers (padertorch Normalization + max_pool2d)
 * *_dtype: storage type of the activation maps (x, y, g, dx ...; PBSED_F32 / PBSED_BF16); a `float*` that is
 * declared bf16 points at 2-byte elements.  Statistics, sums and per-channel vectors are always fp32 / fp64.
 */
/* stats[idx][0..1] += sum, sum of squares over valid rows; idx = c or f*C+c (per_f). double[.][2] */
int pbsed_channel_stats(const float* x, int B, int F, int T, int C, int per_f,
                        const int* seq_len, double* stats, int act_dtype, void* stream);
/* batch statistics -> affine used on load, and running-statistics update.
 *   n = count (valid rows per channel), mean = s/n, var = ss/n - mean^2 (biased)
 *   scale = gamma * rsqrt(var+eps); shift = beta - mean*scale; save_mean/save_rstd for backward.
 *   momentum >= 0: running = momentum*running + (1-momentum)*batch    (CNN layers, 0.95)
 *   momentum <  0: cumulative average over num_tracked (feature-extractor norm)
 *   training == 0: scale/shift from running statistics, nothing updated.
 *   cumulative (momentum<0) & training: scale/shift from the UPDATED running stats
 *   (interpolation_factor = 1), var unbiased n/(n-1) as padertorch does for momentum None.
 *   count <= 0 (training): the count is read on the device from stats[2*nch] -- the data-parallel
 *   "exact" mode all-reduces (sum, sum of squares, count) of every replica in one buffer
 *    and never brings the global count back to the host. */
int pbsed_norm_finalize(const double* stats, double count, int nch,
                        const float* gamma, const float* beta, float eps, float momentum, int training,
                        float* running_mean, float* running_power, float* num_tracked,
                        float* scale, float* shift, float* save_mean, float* save_rstd, void* stream);
/* frequency max-pool by `pool` (rows (B,F,T,C) -> (B,F/pool,T,C)); idx (uint8) = argmax offset.
 * out_stats (nullable, double[C][2]): += per-channel sum / sum of squares of the POOLED map over frames
 * t < seq_len[b] -- the next layer's batch statistics, fused into the pooling pass. */
int pbsed_maxpool_f(const float* x, int B, int F, int T, int C, int pool,
                    float* y, uint8_t* idx, const int* seq_len, double* out_stats,
                    int in_dtype, int out_dtype, void* stream);
int pbsed_maxpool_f_bwd(const float* dy, const uint8_t* idx, int B, int F, int T, int C, int pool,
                        float* dx, int in_dtype, int out_dtype, void* stream);
/* batch-norm backward, two passes.  g = gradient w.r.t. the normalised+affine output (already
 * multiplied by the ReLU mask), x = the layer input the statistics were taken on.
 *   pass 1: sums[idx][0] += sum g ; sums[idx][1] += sum g * xhat       (valid rows only)
 *   pass 2: dx = gamma*rstd * ( g - sums0/n - xhat * sums1/n ) ; dgamma += sums1 ; dbeta += sums0
 *           rows t >= seq_len[b] get dx = 0.  dx may alias g.
 *           count <= 0: n is read on the device from sums[2*nch] (see pbsed_norm_finalize);
 *           dgamma / dbeta nullable (data-parallel: the replica adds its LOCAL sums itself). */
int pbsed_norm_bwd_reduce(const float* g, const float* x, int B, int F, int T, int C, int per_f,
                          const int* seq_len, const float* save_mean, const float* save_rstd,
                          double* sums, int act_dtype, void* stream);
int pbsed_norm_bwd_apply(const float* g, const float* x, int B, int F, int T, int C, int per_f,
                         const int* seq_len, const float* save_mean, const float* save_rstd,
                         const float* gamma, const double* sums, double count,
                         float* dx, float* dgamma, float* dbeta, int act_dtype, void* stream);
/* out = [add +] x  with the tag condition broadcast (B,K) -> extra channels; see pbsed_concat_cond */
/* rows (B,F,T,C) <- concat( x (B,F,T,C0), cond (B,K) broadcast over f,t )  (strong_label/crnn.py:86-91) */
int pbsed_concat_cond(const float* x, const float* cond, int B, int F, int T, int C0, int K,
                      float* out, void* stream);
/* dx (B,F,T,C0) <- dout[..., :C0]  (cond carries no gradient) */
int pbsed_split_cond_bwd(const float* dout, int B, int F, int T, int C0, int K, float* dx, void* stream);

/* ---------------------------------------------------------------------------
 * K3  GRU recurrence   (replaces torch.nn.GRU inside padertorch.contrib.je.modules.rnn.GRU,
 *     call sites pb_sed/models/weak_label/crnn.py:62,66 and strong_label/crnn.py:92)
 *
 * One launch runs `ndir` (<= 4) independent recurrences of the same shape -- the two directions of
 * a bidirectional layer, or the same layer of the reference's separate rnn_fwd / rnn_bwd modules.
 * Every operand is a HOST array of `ndir` device pointers:
 *   gi[d]    : (B,T,3H) input projection x@W_ih^T + b_ih  (gate order r,z,n)
 *   w_hh[d]  : (3H,H),   b_hh[d] : (3H)
 *   reverse_host[d] != 0 : time runs seq_len[b]-1 .. 0 (the reference's reverse=True / the backward
 *                          direction of bidirectional=True)
 *   h_out[d] : rows (B,T,h_stride), H channels written (zeros at t >= seq_len[b]); a bidirectional
 *              layer passes h_out[1] = h_out[0] + H with h_stride = 2H
 *   save[d]  : (B,T,4H) = r, z, n, (W_hn h + b_hn)  for backward (save or save[d] nullable)
 * H must be a multiple of 32 and <= 256 (one thread-block cluster of H/32 CTAs per 8-clip batch
 * slice and direction, W_hh resident in registers); h0 = 0.
 */
int pbsed_gru_fwd(const float* const* gi, const float* const* w_hh, const float* const* b_hh,
                  const int* seq_len, int B, int T, int H, int ndir, const int* reverse_host,
                  float* const* h_out, int h_stride, float* const* save, void* stream);
/* backward through time.  dh_out[d]: gradient w.r.t. h_out[d] (same addressing).
 * writes dgi[d] (B,T,3H) = dL/d(gi) = [dr, dz, dn] and dgh[d] (B,T,3H) = dL/d(W_hh h + b_hh) =
 * [dr, dz, dn*r]; zeros at t >= seq_len[b].  The weight gradients are then plain tap-GEMM wgrads
 * (dW_ih: dgi x input, dW_hh: dgh x h_out shifted one step against the direction of time,
 * dt = -1 / +1 with seq_len masking). */
int pbsed_gru_bwd(const float* const* dh_out, const float* const* h_out, const float* const* save,
                  const float* const* w_hh, const int* seq_len, int B, int T, int H, int ndir,
                  const int* reverse_host, float* const* dgi, float* const* dgh, int h_stride,
                  void* stream);

/* ---------------------------------------------------------------------------
 * K4  scores + losses
 *     bounded sigmoid (pb_sed/models/weak_label/crnn.py:58-59; min_score = 0 -> nn.Sigmoid of
 *     strong_label/crnn.py:93) with the (B,T,K) -> (B,K,T) transpose the model API returns.
 */
int pbsed_sigmoid_btk_to_bkt(const float* z, int B, int T, int K, float min_score, float* y, void* stream);
/* dz (B,T,K) = dy (B,K,T) * (1-2*min) * s(1-s), s = sigmoid(z) */
int pbsed_sigmoid_bwd(const float* dy, const float* z, int B, int T, int K, float min_score,
                      float* dz, void* stream);
/* FBCRNN review loss, forward + backward in one pass
 *   (pb_sed/models/weak_label/crnn.py:107-153,180-206).
 * y_fwd, y_bwd (B,K,T) scores (y_bwd nullable), weak (B,K), boundary (B,K,T) (nullable when
 * strong_weight == 0), class_weights (K, nullable).
 * out_host-free: loss_out[0] = loss, loss_out[1] = sum of weights; dy_* (nullable) gradients.
 * workspace: float[2*B*K + 8]. */
int pbsed_fbcrnn_loss(const float* y_fwd, const float* y_bwd, const float* weak, const float* boundary,
                      const float* class_weights, const int* seq_len, int B, int K, int T,
                      float strong_weight, float label_smoothing,
                      float* loss_out, float* dy_fwd, float* dy_bwd, float* workspace, void* stream);
/* BiCRNN review loss (pb_sed/models/strong_label/crnn.py:107-112) */
int pbsed_bicrnn_loss(const float* y, const float* strong, const int* seq_len, int B, int K, int T,
                      float* loss_out, float* dy, float* workspace, void* stream);

/* ---------------------------------------------------------------------------
 * K5  gradient-norm clip + Adam over one flat arena
 *     (replaces padertorch.train.optimizer.Adam = clip_grad_norm_ + torch.optim.Adam,
 *     configured at pb_sed/experiments/weak_label_crnn/training.py:264-269)
 * hyper (device float[8]): lr, beta1, beta2, eps, max_norm, step (incremented here), grad_scale, -
 * sumsq (device double[1]) is zeroed by step 1 of the NEXT call (the caller zeroes it once).
 * grad_norm_out (device float[1]) receives the pre-clip global L2 norm.
 */
int pbsed_grad_sumsq(const float* g, long long n, const float* hyper, double* sumsq, void* stream);
int pbsed_adam_step(float* p, float* g, float* m, float* v, long long n, float* hyper,
                    double* sumsq, float* grad_norm_out, int zero_grad, void* stream);

/* ---------------------------------------------------------------------------
 * K6  score post-processing of the inference path (the reference does this in
 *     numpy after a D2H copy: pb_sed/models/base/inference.py:143-183,225-289, pb_sed/filters.py)
 *
 * Scores (B, [N,] K, T) are addressed as R rows of T frames.  filt_len (device int32[filt_mod]):
 * the filter length of row r is filt_len[r % filt_mod] (1 = scalar, K = per class, N*K = per
 * (n, class): the three cases of inference.py:225-266).  seq_len (nullable, device int32[B]):
 * clip b = r / rows_per_clip; frames t >= seq_len[b] are read as 0 (inference.py:143-147).
 */
/* y[r,t] = median of x[r, t-h .. t+h], zeros outside [0,T), h = (n-1)/2, n odd (n <= 1: copy);
 * = scipy.signal.medfilt per row (pb_sed/filters.py:56-83).  Selection: bit-exact. */
int pbsed_medfilt(const float* x, int R, int T, const int* filt_len, int filt_mod,
                  const int* seq_len, int rows_per_clip, float* y, void* stream);
/* boundariesfilt (inference.py:269-289): s = stepfilt(x, n) (filters.py:113-135; n even, n = 0: s = x),
 * y = min( cummax_t(s_fwd), flip(cummax(s of the flipped row)) ), computed and returned in float64
 * like the reference (np.correlate with a float64 kernel). */
int pbsed_boundariesfilt(const float* x, int R, int T, const int* filt_len, int filt_mod,
                         const int* seq_len, int rows_per_clip, double* y, void* stream);
/* scores (B,N,K,T) *= max(tags[b,k], 1 - apply[n,k])   (inference.py:170-183; tags (B,K), apply (N,K)) */
int pbsed_tag_mask(float* scores, const float* tags, const float* apply, int B, int N, int K, int T,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PBSED_B200_H */
