"""CPU restatement of the padertorch / paderbox modules pb_sed composes.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): parity unpinned at this
third-party boundary.  Each class cites the pb_sed call site that fixes its
signature ([CS]) and SURVEY.md Appendix A for the recollected semantics ([R]).

Third-party pins (reference ``README.md:40-41``):
    padertorch @ b7ba24a42a05745d127a74a519af08a876319a95
    paderbox   @ 809b27251c478f1997d2720b89fe455aac23234e
"""
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


# --------------------------------------------------------------------------
# paderbox.transform.module_stft.stft  (called via pb_sed/data_preparation/transform.py:53,
# parameters pb_sed/data_preparation/provider.py:315-323)
# --------------------------------------------------------------------------
def blackman_periodic(window_length):
    """[R] scipy.signal.windows.blackman(window_length + 1)[:-1] (symmetric_window=False)."""
    n = np.arange(window_length, dtype=np.float64)
    m = window_length  # periodic: denominator is N, not N-1
    return (0.42 - 0.5 * np.cos(2 * np.pi * n / m)
            + 0.08 * np.cos(4 * np.pi * n / m))


def stft_frames(num_samples, shift=320, window_length=960, fading='half', pad=True):
    """number of frames the reference STFT yields for ``num_samples``."""
    if fading == 'half':
        num_samples = num_samples + (window_length - shift) // 2 \
            + int(math.ceil((window_length - shift) / 2))
    elif fading in (True, 'full'):
        num_samples = num_samples + 2 * (window_length - shift)
    if pad:
        return max(int(math.ceil((num_samples - window_length) / shift)) + 1, 1)
    return (num_samples - window_length) // shift + 1


def stft(audio, shift=320, window_length=960, size=1024, fading='half', pad=True, frame_start=None):
    """audio (..., S) -> complex (..., T, size//2+1), float64 math.

    frame_start (B, T) int: TimeWarpedSTFT [R] -- first sample of every frame in the faded signal
    instead of t*shift (audio must then be (B, S) or (B, 1, S)).

    [R] fading='half' zero-pads (window_length-shift)//2 in front and
    ceil((window_length-shift)/2) at the end; pad=True keeps the last partial
    frame zero padded (segment_axis(end='pad')); periodic Blackman window on
    the window_length samples; rfft(n=size); no normalisation factor.
    """
    audio = np.asarray(audio, dtype=np.float64)
    if fading == 'half':
        front = (window_length - shift) // 2
        back = int(math.ceil((window_length - shift) / 2))
    elif fading in (True, 'full'):
        front = back = window_length - shift
    else:
        front = back = 0
    pad_width = [(0, 0)] * (audio.ndim - 1) + [(front, back)]
    audio = np.pad(audio, pad_width, mode='constant')
    n = audio.shape[-1]
    t = stft_frames(n, shift, window_length, fading=None, pad=pad)
    if pad:
        need = (t - 1) * shift + window_length
        if need > n:
            audio = np.pad(audio, [(0, 0)] * (audio.ndim - 1) + [(0, need - n)])
    if frame_start is not None:
        fs = np.asarray(frame_start)
        assert fs.shape[-1] == t, (fs.shape, t)
        idx = fs[..., None] + np.arange(window_length)                  # (B, T, W)
        flat = audio.reshape(fs.shape[0], -1)
        flat = np.pad(flat, [(0, 0), (0, max(int(idx.max()) + 1 - flat.shape[-1], 0))])
        frames = np.take_along_axis(flat[:, None, :], idx.reshape(fs.shape[0], 1, -1), axis=-1)
        frames = frames.reshape(fs.shape[0], t, window_length) * blackman_periodic(window_length)
        return np.fft.rfft(frames, n=size, axis=-1).reshape(audio.shape[:-1] + (t, size // 2 + 1))
    idx = np.arange(t)[:, None] * shift + np.arange(window_length)[None, :]
    frames = audio[..., idx] * blackman_periodic(window_length)
    return np.fft.rfft(frames, n=size, axis=-1)


def time_warp_grid(anchor, anchor_shift, num_frames, shift):
    """TimeWarpedSTFT [R] (transform.py:36-45; samplers provider.py:329-338: anchor ~ U(.4,.6),
    anchor_shift ~ U(-.1,.1), both relative to the clip length): output frame t reads the source
    position src(t) (in frames), piecewise linear with src(0)=0, src((anchor+anchor_shift) T) = anchor T,
    src(T) = T.  Returns (src (B,T) float64, frame_start (B,T) int = round(src * shift))."""
    a_in = np.asarray(anchor, dtype=np.float64)[:, None] * num_frames
    a_out = np.clip((np.asarray(anchor, dtype=np.float64) + np.asarray(anchor_shift, dtype=np.float64))[:, None]
                    * num_frames, 1., num_frames - 1.)
    t = np.arange(num_frames, dtype=np.float64)[None]
    src = np.where(t <= a_out, t * a_in / a_out,
                   a_in + (t - a_out) * (num_frames - a_in) / (num_frames - a_out))
    return src, np.floor(src * shift + .5).astype(np.int64)


class STFT:
    """padertorch.contrib.je.data.transforms.STFT [CS] provider.py:315-323, transform.py:53-54.

    example['audio_data'] (C, S) -> example['stft'] (C, T, F, 2) float32.
    """

    def __init__(self, shift, size, window_length=None, window='blackman',
                 symmetric_window=False, pad=True, fading='full',
                 alignment_keys=None):
        assert window == 'blackman' and not symmetric_window
        self.shift = shift
        self.size = size
        self.window_length = size if window_length is None else window_length
        self.pad = pad
        self.fading = fading
        self.alignment_keys = alignment_keys

    def __call__(self, example):
        x = stft(example['audio_data'], self.shift, self.window_length,
                 self.size, self.fading, self.pad)
        example['stft'] = np.stack([x.real, x.imag], axis=-1).astype(np.float32)
        return example


# --------------------------------------------------------------------------
# paderbox.transform.module_fbank.get_fbanks  [R]
# --------------------------------------------------------------------------
def hz2mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel2hz(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def get_fbanks(sample_rate, stft_size, number_of_filters,
               lowest_frequency=50., highest_frequency=None):
    """(number_of_filters, stft_size//2+1) float64 HTK-mel triangles, unit row sum.

    [R] number_of_filters+2 equally mel-spaced edges between lowest_frequency
    and highest_frequency (default sample_rate/2), evaluated at bin centres
    k*sample_rate/stft_size, rows normalised by (sum + 1e-6).
    """
    if highest_frequency is None:
        highest_frequency = sample_rate / 2
    edges = mel2hz(np.linspace(hz2mel(lowest_frequency), hz2mel(highest_frequency),
                               number_of_filters + 2))
    edges = edges * stft_size / sample_rate            # in (fractional) bins
    k = np.arange(stft_size // 2 + 1, dtype=np.float64)[None, :]
    lo, ce, hi = edges[:-2, None], edges[1:-1, None], edges[2:, None]
    fb = np.maximum(np.minimum((k - lo) / (ce - lo), (hi - k) / (hi - ce)), 0.)
    return fb / (fb.sum(-1, keepdims=True) + 1e-6)


def warp_mel(m, alpha, ratio, m_hi):
    """paderbox MelWarping [R] (kwargs training.py:195-208): VTLP-shaped piecewise-linear warp applied in
    the mel domain.  m_b = m_hi / (1 + ratio); slope alpha below the knee m_b*min(alpha,1)/alpha, then a
    straight line to (m_hi, m_hi)."""
    m = np.asarray(m, dtype=np.float64)
    mn = np.minimum(alpha, 1.)
    m_b = m_hi / (1. + ratio)
    knee = m_b * mn / alpha
    return np.where(m <= knee, alpha * m, m_hi - (m_hi - m_b * mn) / (m_hi - knee) * (m_hi - m))


def get_warped_fbanks(alpha, ratio, sample_rate, stft_size, number_of_filters, lowest_frequency=50.,
                      highest_frequency=None, warp_highest_frequency=None):
    """(B, number_of_filters, F) float64: ``get_fbanks`` on per-example warped edge frequencies."""
    if highest_frequency is None:
        highest_frequency = sample_rate / 2
    if warp_highest_frequency is None:
        warp_highest_frequency = sample_rate / 2
    alpha = np.asarray(alpha, dtype=np.float64)[:, None]
    ratio = np.asarray(ratio, dtype=np.float64)[:, None]
    mel = np.linspace(hz2mel(lowest_frequency), hz2mel(highest_frequency), number_of_filters + 2)[None]
    edges = mel2hz(warp_mel(mel, alpha, ratio, hz2mel(warp_highest_frequency))) * stft_size / sample_rate
    k = np.arange(stft_size // 2 + 1, dtype=np.float64)[None, None, :]
    lo, ce, hi = edges[:, :-2, None], edges[:, 1:-1, None], edges[:, 2:, None]
    fb = np.maximum(np.minimum((k - lo) / (ce - lo), (hi - k) / (hi - ce)), 0.)
    return fb / (fb.sum(-1, keepdims=True) + 1e-6)


# --------------------------------------------------------------------------
# padertorch.ops.sequence.mask.compute_mask  [CS] weak_label/crnn.py:238
# --------------------------------------------------------------------------
def compute_mask(x, sequence_lengths, batch_axis=0, sequence_axis=1):
    """0/1 mask broadcastable to x; all-ones when sequence_lengths is None."""
    if sequence_lengths is None:
        return torch.ones_like(x)
    if batch_axis < 0:
        batch_axis += x.dim()
    if sequence_axis < 0:
        sequence_axis += x.dim()
    seq = torch.as_tensor(np.asarray(sequence_lengths), device=x.device).long()
    shape = [1] * x.dim()
    shape[batch_axis] = x.shape[batch_axis]
    seq = seq.reshape(shape)
    shape = [1] * x.dim()
    shape[sequence_axis] = x.shape[sequence_axis]
    idx = torch.arange(x.shape[sequence_axis], device=x.device).reshape(shape)
    return (idx < seq).to(x.dtype).expand(x.shape)


# --------------------------------------------------------------------------
# padertorch.contrib.je.modules.reduce  [CS] weak_label/crnn.py:147,158; strong_label/crnn.py:112,202
# --------------------------------------------------------------------------
class _Reduce(nn.Module):
    def __init__(self, axis=-1, keepdims=False):
        super().__init__()
        self.axis = axis
        self.keepdims = keepdims


class Sum(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is not None:
            x = x * compute_mask(x, seq_len, 0, self.axis)
        return x.sum(self.axis, keepdim=self.keepdims)


class Mean(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is None:
            return x.mean(self.axis, keepdim=self.keepdims)
        mask = compute_mask(x, seq_len, 0, self.axis)
        return (x * mask).sum(self.axis, keepdim=self.keepdims) / (
            mask.sum(self.axis, keepdim=self.keepdims) + 1e-6)


class Max(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is not None:
            mask = compute_mask(x, seq_len, 0, self.axis)
            x = x * mask + torch.finfo(x.dtype).min * (1 - mask)
        return x.max(self.axis, keepdim=self.keepdims)


class TakeLast(_Reduce):
    def forward(self, x, seq_len=None):
        axis = self.axis % x.dim()
        if seq_len is None:
            out = x.narrow(axis, x.shape[axis] - 1, 1)
        else:
            idx = torch.as_tensor(np.asarray(seq_len), device=x.device).long() - 1
            shape = [1] * x.dim()
            shape[0] = x.shape[0]
            idx = idx.reshape(shape)
            exp = list(x.shape)
            exp[axis] = 1
            out = x.gather(axis, idx.expand(exp))
        return out if self.keepdims else out.squeeze(axis)


# --------------------------------------------------------------------------
# padertorch.contrib.je.modules.norm.Normalization  [R]
# --------------------------------------------------------------------------
class Normalization(nn.Module):
    """Masked normalisation over ``statistics_axis`` with running statistics.

    data_format e.g. 'bcft' / 'bct'.  Learnable ``scale``/``shift`` over
    ``independent_axis`` (None -> no affine).  Buffers ``running_mean``,
    ``running_power`` (second raw moment), ``num_tracked_values``.
    momentum=None -> cumulative average.  Training: biased batch statistics,
    ``y = (x-mean)/sqrt(var+eps)*scale+shift``, padded frames zeroed.
    ``interpolation_factor=1`` -> value of the running-statistics
    normalisation (after the update) is returned in training.
    Eval: ``running_var = n/(n-1)*(running_power-running_mean**2)`` when
    momentum is None (unbiased cumulative estimate) else plain.
    """

    def __init__(self, data_format, shape, statistics_axis='b', independent_axis='c',
                 batch_axis='b', sequence_axis='t', shift=True, scale=True,
                 eps=1e-3, momentum=0.95, interpolation_factor=0.,
                 track_running_stats=True):
        super().__init__()
        self.data_format = data_format.lower()
        self.batch_axis = self.data_format.index(batch_axis)
        self.sequence_axis = self.data_format.index(sequence_axis)
        self.statistics_axis = tuple(self.data_format.index(a) for a in statistics_axis.lower())
        self.eps = eps
        self.momentum = momentum
        self.interpolation_factor = interpolation_factor
        self.track_running_stats = track_running_stats
        self.frozen_stats = False      # [R] Normalization.freeze(freeze_stats=True): running statistics in train mode
        self.shift_on = shift
        self.scale_on = scale
        reduced = [1 if i in self.statistics_axis else s for i, s in enumerate(shape)]
        assert all(s is not None for s in reduced), (data_format, shape, statistics_axis)
        if track_running_stats:
            self.register_buffer('num_tracked_values', torch.zeros(reduced))
            self.register_buffer('running_mean', torch.zeros(reduced))
            self.register_buffer('running_power', torch.ones(reduced))
        if independent_axis is not None:
            ind = tuple(self.data_format.index(a) for a in independent_axis.lower())
            pshape = [s if i in ind else 1 for i, s in enumerate(shape)]
            self.scale = nn.Parameter(torch.ones(pshape))
            self.shift = nn.Parameter(torch.zeros(pshape))
        else:
            self.scale = self.shift = None

    def _stats(self, x, seq_len):
        mask = compute_mask(x, seq_len, self.batch_axis, self.sequence_axis)
        n = mask.sum(self.statistics_axis, keepdim=True)
        mean = (x * mask).sum(self.statistics_axis, keepdim=True) / n
        var = (((x - mean) * mask) ** 2).sum(self.statistics_axis, keepdim=True) / n
        return mask, n, mean, var

    def running_var(self):
        var = self.running_power - self.running_mean ** 2
        if self.momentum is None:
            n = self.num_tracked_values
            var = var * n / torch.clamp(n - 1, min=1.)
        return var

    def forward(self, x, sequence_lengths=None):
        mask = compute_mask(x, sequence_lengths, self.batch_axis, self.sequence_axis)
        if (self.training and not self.frozen_stats) or not self.track_running_stats:
            _, n, mean, var = self._stats(x, sequence_lengths)
            y = x
            if self.shift_on:
                y = y - mean
            if self.scale_on:
                y = y / torch.sqrt(var + self.eps)
            if self.track_running_stats:
                with torch.no_grad():
                    power = var + mean ** 2
                    if self.momentum is None:
                        tot = self.num_tracked_values + n
                        self.running_mean += (mean - self.running_mean) * n / tot
                        self.running_power += (power - self.running_power) * n / tot
                    else:
                        m = self.momentum
                        self.running_mean.mul_(m).add_((1 - m) * mean)
                        self.running_power.mul_(m).add_((1 - m) * power)
                    self.num_tracked_values += n
                if self.interpolation_factor > 0.:
                    y_run = self._running_norm(x)
                    y = y + self.interpolation_factor * (y_run - y).detach()
        else:
            y = self._running_norm(x)
        if self.scale is not None:
            y = y * self.scale + self.shift
        return y * mask

    def _running_norm(self, x):
        y = x
        if self.shift_on:
            y = y - self.running_mean
        if self.scale_on:
            y = y / torch.sqrt(self.running_var() + self.eps)
        return y


# --------------------------------------------------------------------------
# padertorch.contrib.je.modules.features.NormalizedLogMelExtractor
# [CS] weak_label/crnn.py:86-90, experiments/weak_label_crnn/training.py:190-217
# --------------------------------------------------------------------------
class NormalizedLogMelExtractor(nn.Module):
    """stft (B,C,T,F,2) -> normalised, clamped log-mel (B,C,n_mels,T).

    [R] power -> unit-sum HTK-mel filterbank -> log(.+1e-18) -> 'bcft' ->
    cumulative running mean/var normalisation over 'bt' (eps 1e-5, no affine,
    interpolation_factor 1) -> clamp(+-6); all under no_grad.  The train-only
    random augmentations (mel warping, time/frequency masks, noise) are
    SURVEY section 8f row 2 ("next") and are rejected here if requested with a
    non-zero strength in training mode.
    """

    def __init__(self, sample_rate, stft_size, number_of_filters, num_channels=1,
                 lowest_frequency=50., highest_frequency=None,
                 add_deltas=False, add_delta_deltas=False, norm_eps=1e-5, clamp=6.,
                 frequency_warping_fn=None, blur_sigma=0.,
                 n_time_masks=0, max_masked_time_steps=70, max_masked_time_rate=.2,
                 n_frequency_masks=0, max_masked_frequency_bands=20,
                 max_masked_frequency_rate=.2, max_noise_scale=0.,
                 augment=False):
        super().__init__()
        assert not add_deltas and not add_delta_deltas
        self.sample_rate, self.stft_size = sample_rate, stft_size
        self.number_of_filters = number_of_filters
        self.clamp = clamp
        self.augment = augment  # oracle: deterministic path only
        fb = get_fbanks(sample_rate, stft_size, number_of_filters,
                        lowest_frequency, highest_frequency)
        self.register_buffer('fbanks', torch.from_numpy(fb.T.copy()).float(), persistent=False)  # (F, n_mels)
        self.norm = Normalization(
            'bcft', (None, num_channels, number_of_filters, None), statistics_axis='bt',
            independent_axis=None, eps=norm_eps, momentum=None, interpolation_factor=1.)

    def forward(self, x, seq_len=None, targets=None, augmentation=None):
        """augmentation (training only; the random draws are made by the caller so that both sides of a
        parity test see the same ones): dict with any of alpha, ratio (B,) mel warping; time_masks,
        freq_masks (B, n, 2) int (onset, width); noise (B, C, F, T) standard normal, noise_scale (B,).
        Order [R]: warped filterbank -> log -> running norm -> clamp -> time masks -> frequency masks ->
        noise; frames behind seq_len stay 0."""
        assert not (self.training and self.augment and augmentation is None), 'pass the draws explicitly'
        aug = augmentation if (self.training and augmentation) else {}
        with torch.no_grad():
            power = (x ** 2).sum(-1)                              # b c t f
            if 'alpha' in aug:
                fb = get_warped_fbanks(aug['alpha'], aug['ratio'], self.sample_rate, self.stft_size,
                                       self.number_of_filters)
                fb = torch.from_numpy(fb).float().transpose(1, 2)[:, None]        # b 1 f m
                mel = torch.log(power @ fb + 1e-18)
            else:
                mel = torch.log(power @ self.fbanks + 1e-18)      # b c t m
            x = mel.transpose(-2, -1)                             # b c m t
            x = self.norm(x, seq_len)
            if self.clamp is not None:
                x = torch.clamp(x, -self.clamp, self.clamp)
            B, _, F, T = x.shape
            for key, axis in (('time_masks', 3), ('freq_masks', 2)):
                if key in aug:
                    for b in range(B):
                        for on, w in np.asarray(aug[key])[b]:
                            if axis == 3:
                                x[b, :, :, on:on + w] = 0.
                            else:
                                x[b, :, on:on + w, :] = 0.
            if 'noise' in aug:
                x = x + torch.as_tensor(aug['noise_scale']).float().reshape(B, 1, 1, 1) * torch.as_tensor(aug['noise']).float()
            if aug and seq_len is not None:
                x = x * compute_mask(x, seq_len, 0, -1)
        if targets is None:
            return x, seq_len
        return x, seq_len, targets


# --------------------------------------------------------------------------
# padertorch.contrib.je.modules.conv  [CS] experiments/weak_label_crnn/training.py:218-242
# --------------------------------------------------------------------------
class Pad(nn.Module):
    """[CS] weak_label/crnn.py:289-290: Pad(side)(x, size) zero-pads the last axis."""

    def __init__(self, side='both', mode='constant'):
        super().__init__()
        self.side = side

    def forward(self, x, size):
        size = int(size)
        if self.side == 'front':
            p = (size, 0)
        elif self.side == 'end':
            p = (0, size)
        elif self.side == 'both':
            p = (size // 2, int(math.ceil(size / 2)))
        else:
            raise ValueError(self.side)
        return F.pad(x, p)


class _ConvBlock(nn.Module):
    """[R] padertorch Conv1d / Conv2d module: ``conv`` (torch conv) + ``norm`` -> state-dict keys
    ``convs.<i>.conv.{weight,bias}``, ``convs.<i>.norm.*`` ([CS] weak_label_crnn/training.py:331-340 reads the
    layer index from the second dotted component)."""

    def __init__(self, conv, norm):
        super().__init__()
        self.conv = conv
        self.norm = norm


class _CNN(nn.Module):
    """Shared stack logic of CNN1d / CNN2d.

    [R] pre_activation: layer i = norm -> relu -> zero 'same' pad -> conv(+bias)
    -> max-pool; layer 0 with input_layer=True has no norm/activation; with
    pre_activation and output_layer=False nothing follows the last conv.
    Post-activation (pre_activation=False; used by rnn output_net):
    conv -> norm -> relu for all but the last layer when output_layer=True.
    """
    ndim = None

    def __init__(self, in_channels, out_channels, kernel_size, pool_size=1,
                 residual_connections=None, norm='batch', norm_kwargs=None,
                 activation_fn='relu', pre_activation=False, dropout=0.,
                 input_layer=True, output_layer=True):
        super().__init__()
        assert activation_fn == 'relu' and dropout == 0.
        assert residual_connections is None or all(r is None for r in residual_connections), \
            'residual (deep) config is SURVEY 8f row 4'
        n = len(out_channels)
        self.in_channels = in_channels
        self.out_channels = list(out_channels)
        self.kernel_sizes = kernel_size if isinstance(kernel_size, (list, tuple)) else [kernel_size] * n
        self.pool_sizes = pool_size if isinstance(pool_size, (list, tuple)) and len(pool_size) == n \
            else [pool_size] * n
        self.pre_activation = pre_activation
        self.input_layer, self.output_layer = input_layer, output_layer
        norm_kwargs = dict(norm_kwargs or {})
        Conv = nn.Conv2d if self.ndim == 2 else nn.Conv1d
        self.convs = nn.ModuleList()
        norms = []
        convs = []
        c = in_channels
        for i, co in enumerate(self.out_channels):
            conv = Conv(c, co, self.kernel_sizes[i])
            nn.init.xavier_uniform_(conv.weight)
            nn.init.zeros_(conv.bias)
            convs.append(conv)
            if pre_activation:
                has_norm = not (i == 0 and input_layer)
                nc = c
            else:
                has_norm = not (i == n - 1 and output_layer)
                nc = co
            if norm == 'batch' and has_norm:
                fmt = 'bcft' if self.ndim == 2 else 'bct'
                shape = (None, nc, None, None) if self.ndim == 2 else (None, nc, None)
                norms.append(Normalization(
                    fmt, shape, statistics_axis='bft' if self.ndim == 2 else 'bt',
                    independent_axis='c', momentum=0.95, **norm_kwargs))
            else:
                norms.append(None)
            c = co
        for conv, norm in zip(convs, norms):
            self.convs.append(_ConvBlock(conv, norm))

    @property
    def norms(self):
        return [blk.norm for blk in self.convs]

    def _pad(self, x, k):
        if self.ndim == 2:
            kf = kt = k if isinstance(k, int) else None
            if kf is None:
                kf, kt = k
            return F.pad(x, ((kt - 1) // 2, int(math.ceil((kt - 1) / 2)),
                             (kf - 1) // 2, int(math.ceil((kf - 1) / 2))))
        return F.pad(x, ((k - 1) // 2, int(math.ceil((k - 1) / 2))))

    def _pool(self, x, p):
        if p == 1 or p == (1, 1) or p is None:
            return x
        if self.ndim == 2:
            p = (p, p) if isinstance(p, int) else tuple(p)
            assert p[1] == 1, 'time pooling not on the hot path'
            return F.max_pool2d(x, p)
        raise AssertionError('1-D pooling not on the hot path')

    def forward(self, x, seq_len=None):
        n = len(self.convs)
        for i in range(n):
            if self.pre_activation:
                if self.norms[i] is not None:
                    x = torch.relu(self.norms[i](x, seq_len))
                elif not (i == 0 and self.input_layer):
                    x = torch.relu(x)
                x = self.convs[i].conv(self._pad(x, self.kernel_sizes[i]))
            else:
                x = self.convs[i].conv(self._pad(x, self.kernel_sizes[i]))
                if not (i == n - 1 and self.output_layer):
                    if self.norms[i] is not None:
                        x = self.norms[i](x, seq_len)
                    x = torch.relu(x)
            x = self._pool(x, self.pool_sizes[i])
        return x, seq_len

    def freeze(self, num_layers=None, freeze_norm_stats=True):
        """[CS] experiments/weak_label_crnn/training.py:343-350.  [R] frozen norms with ``freeze_norm_stats``
        use their running statistics in train mode and stop updating them."""
        n = len(self.convs) if num_layers is None else num_layers
        for i in range(n):
            for p in self.convs[i].conv.parameters():
                p.requires_grad = False
            if self.norms[i] is not None:
                for p in self.norms[i].parameters():
                    p.requires_grad = False
                self.norms[i].frozen_stats = bool(freeze_norm_stats)


class CNN2d(_CNN):
    ndim = 2


class CNN1d(_CNN):
    ndim = 1


class CNN(nn.Module):
    """padertorch.contrib.je.modules.hybrid.CNN  [CS] weak_label/crnn.py:93,319,326-330;
    strong_label/crnn.py:86,180-184.  (B,C,F,T) -> cnn_2d -> 'b c f t -> b (c f) t' -> cnn_1d."""

    def __init__(self, cnn_2d, cnn_1d, input_height=None, positional_encoding=False,
                 conditional_dims=0):
        super().__init__()
        assert not positional_encoding
        self.cnn_2d, self.cnn_1d = cnn_2d, cnn_1d
        self.input_height = input_height
        self.conditional_dims = conditional_dims

    def forward(self, x, seq_len=None, condition=None):
        if self.conditional_dims:
            b, _, f, t = x.shape
            cond = condition.to(x.dtype)
            if cond.dim() == 3:
                cond = cond.unsqueeze(2)                 # (B,K,1,1)
            x = torch.cat([x, cond.expand(b, cond.shape[1], f, t)], dim=1)
        x, seq_len = self.cnn_2d(x, seq_len)
        b, c, f, t = x.shape
        x = x.reshape(b, c * f, t)
        return self.cnn_1d(x, seq_len)


# --------------------------------------------------------------------------
# padertorch.contrib.je.modules.rnn.GRU  [CS] experiments/weak_label_crnn/training.py:243-260;
# weak_label/crnn.py:62,66,340; strong_label/crnn.py:92,189-195
# --------------------------------------------------------------------------
def reverse_sequence(x, seq_len):
    """flip the valid part of each sequence along the last axis (x: B,F,T)."""
    if seq_len is None:
        return x.flip(-1)
    t = x.shape[-1]
    sl = torch.as_tensor(np.asarray(seq_len), device=x.device).long()[:, None]
    idx = torch.arange(t, device=x.device)[None, :]
    src = torch.where(idx < sl, sl - 1 - idx, idx)
    return x.gather(-1, src[:, None, :].expand(x.shape))


class GRU(nn.Module):
    def __init__(self, rnn, output_net, reverse=False):
        super().__init__()
        self.rnn, self.output_net, self.reverse = rnn, output_net, reverse

    def forward(self, x, seq_len=None):
        if self.reverse:
            x = reverse_sequence(x, seq_len)
        if self.rnn is not None:
            h = x.transpose(1, 2)                                 # b t f
            if seq_len is not None:
                packed = nn.utils.rnn.pack_padded_sequence(
                    h, torch.as_tensor(np.asarray(seq_len)).long().cpu(),
                    batch_first=True, enforce_sorted=False)
                out, _ = self.rnn(packed)
                h, _ = nn.utils.rnn.pad_packed_sequence(
                    out, batch_first=True, total_length=x.shape[-1])
            else:
                h, _ = self.rnn(h)
            x = h.transpose(1, 2)
        y, seq_len = self.output_net(x, seq_len)
        if self.reverse:
            y = reverse_sequence(y, seq_len)
        return y, seq_len


# --------------------------------------------------------------------------
# minimal padertorch.Model  (base class of pb_sed.models.base.SoundEventModel)
# --------------------------------------------------------------------------
class Model(nn.Module):
    def example_to_device(self, example, device=None):
        out = {}
        for k, v in example.items():
            if isinstance(v, np.ndarray) and v.dtype.kind in 'fiub':
                v = torch.from_numpy(v)
            if torch.is_tensor(v):
                v = v.to(device)
            out[k] = v
        return out

    def modify_summary(self, summary):
        return summary
