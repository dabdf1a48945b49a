"""Host-side mirror of the padertorch modules pb_sed composes, backed by the sm_100a kernels.

Same class names, constructor kwargs and call signatures as the ``factory`` targets pb_sed
wires up (``pb_sed/models/weak_label/crnn.py:318-321``, ``pb_sed/models/strong_label/crnn.py:
169-171``, kwargs from ``pb_sed/experiments/weak_label_crnn/training.py:190-260``):

    NormalizedLogMelExtractor(x, seq_len=, targets=) -> (x (B,1,F,T), seq_len[, targets])
    CNN(x, seq_len[, condition])                     -> (h (B,D,T), seq_len)
    GRU(h, seq_len)                                  -> (y (B,K,T), seq_len)      [.rnn, .output_net]
    CNN2d / CNN1d / Pad / TakeLast / Mean / Sum / Max / compute_mask

Parameters live in kernel-native layouts (conv weight ``(taps, Cout, Cin)``), but
``state_dict()`` / ``load_state_dict()`` speak the reference layouts (torch ``Conv2d`` /
``Conv1d`` / ``nn.GRU`` shapes, Normalization buffers ``(1,C,1,1)``), so checkpoints
(``training.py:327-342``) interchange.  All tensors returned in reference layout are zero-copy
permuted views of the native ``(B,F,T,C)`` / ``(B,T,C)`` maps.
"""
import math

import numpy as np
import torch
from torch import nn

from . import ops
from .ops import SeqLen, to_native, from_native


# =============================================================== feature tables (host, float64)
def hz2mel(f):
    return 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)


def mel2hz(m):
    return 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)


def mel_filterbank(sample_rate, stft_size, number_of_filters, lowest_frequency=50.,
                   highest_frequency=None):
    """(n_mels, n_bins) HTK-mel triangles, rows normalised to unit sum."""
    if highest_frequency is None:
        highest_frequency = sample_rate / 2
    pts = mel2hz(np.linspace(hz2mel(lowest_frequency), hz2mel(highest_frequency),
                             number_of_filters + 2)) * stft_size / sample_rate
    k = np.arange(stft_size // 2 + 1, dtype=np.float64)[None]
    left, centre, right = pts[:-2, None], pts[1:-1, None], pts[2:, None]
    tri = np.minimum((k - left) / (centre - left), (right - k) / (right - centre))
    tri = np.maximum(tri, 0.)
    return tri / (tri.sum(-1, keepdims=True) + 1e-6)


def sparse_filterbank(fb):
    """dense (n_mels, n_bins) -> (lo, hi, weights (n_mels, stride)) for the kernel."""
    n_mels = fb.shape[0]
    lo = np.zeros(n_mels, np.int32)
    hi = np.zeros(n_mels, np.int32)
    for m in range(n_mels):
        nz = np.nonzero(fb[m])[0]
        if len(nz):
            lo[m], hi[m] = nz[0], nz[-1] + 1
    stride = max(int((hi - lo).max()), 1)
    w = np.zeros((n_mels, stride), np.float32)
    for m in range(n_mels):
        w[m, :hi[m] - lo[m]] = fb[m, lo[m]:hi[m]]
    return lo, hi, w, stride


def blackman_window(window_length):
    """periodic Blackman (scipy.signal.windows.blackman(N+1)[:-1])."""
    n = np.arange(window_length, dtype=np.float64)
    return 0.42 - 0.5 * np.cos(2 * np.pi * n / window_length) \
        + 0.08 * np.cos(4 * np.pi * n / window_length)


def stft_num_frames(num_samples, shift, window_length, fading='half', pad=True):
    if fading == 'half':
        num_samples += (window_length - shift) // 2 + int(math.ceil((window_length - shift) / 2))
    elif fading in (True, 'full'):
        num_samples += 2 * (window_length - shift)
    if pad:
        return max(int(math.ceil((num_samples - window_length) / shift)) + 1, 1)
    return (num_samples - window_length) // shift + 1


# =============================================================== mask / reduce (tiny host-side glue)
def compute_mask(x, sequence_lengths, batch_axis=0, sequence_axis=1):
    """padertorch.ops.sequence.mask.compute_mask (call site weak_label/crnn.py:238)."""
    if sequence_lengths is None:
        return torch.ones_like(x)
    batch_axis %= x.dim()
    sequence_axis %= x.dim()
    seq = torch.as_tensor(np.asarray(sequence_lengths), device=x.device).long()
    shp = [1] * x.dim()
    shp[batch_axis] = x.shape[batch_axis]
    seq = seq.reshape(shp)
    shp = [1] * x.dim()
    shp[sequence_axis] = x.shape[sequence_axis]
    idx = torch.arange(x.shape[sequence_axis], device=x.device).reshape(shp)
    return (idx < seq).to(x.dtype).expand(x.shape)


class _Reduce(nn.Module):
    def __init__(self, axis=-1, keepdims=False):
        super().__init__()
        self.axis, self.keepdims = axis, keepdims


class Sum(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is not None:
            x = x * compute_mask(x, seq_len, 0, self.axis)
        return x.sum(self.axis, keepdim=self.keepdims)


class Mean(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is None:
            return x.mean(self.axis, keepdim=self.keepdims)
        m = compute_mask(x, seq_len, 0, self.axis)
        return (x * m).sum(self.axis, keepdim=self.keepdims) / (m.sum(self.axis, keepdim=self.keepdims) + 1e-6)


class Max(_Reduce):
    def forward(self, x, seq_len=None):
        if seq_len is not None:
            m = compute_mask(x, seq_len, 0, self.axis)
            x = x * m + torch.finfo(x.dtype).min * (1 - m)
        return x.max(self.axis, keepdim=self.keepdims)


class TakeLast(_Reduce):
    def forward(self, x, seq_len=None):
        axis = self.axis % x.dim()
        if seq_len is None:
            out = x.narrow(axis, x.shape[axis] - 1, 1)
        else:
            idx = torch.as_tensor(np.asarray(seq_len), device=x.device).long() - 1
            shp = [1] * x.dim()
            shp[0] = x.shape[0]
            tgt = list(x.shape)
            tgt[axis] = 1
            out = x.gather(axis, idx.reshape(shp).expand(tgt))
        return out if self.keepdims else out.squeeze(axis)


class Pad(nn.Module):
    """padertorch.contrib.je.modules.conv.Pad (call sites weak_label/crnn.py:289-290)."""

    def __init__(self, side='both', mode='constant'):
        super().__init__()
        self.side = side

    def forward(self, x, size):
        size = int(size)
        p = {'front': (size, 0), 'end': (0, size),
             'both': (size // 2, int(math.ceil(size / 2)))}[self.side]
        return torch.nn.functional.pad(x, p)


# =============================================================== Normalization (parameter holder)
class Normalization(nn.Module):
    """parameters / running statistics of one masked batch-norm; the arithmetic is fused into the
    consuming tap-GEMM's operand load (ops.ConvLayerFn).  ``flatten`` = (C, Fh) when the
    reference sees ``C*Fh`` flattened features (index c*Fh+f) while the native map is (f, c)."""

    def __init__(self, num_channels, ndim, eps=1e-3, momentum=0.95, affine=True, flatten=None):
        super().__init__()
        self.num_channels, self.ndim = num_channels, ndim
        self.eps, self.momentum = eps, momentum
        self.flatten = flatten
        self.frozen_stats = False      # freeze(freeze_norm_stats=True): running statistics even in train mode
        if affine:
            self.scale = nn.Parameter(torch.ones(num_channels))
            self.shift = nn.Parameter(torch.zeros(num_channels))
        else:
            self.scale = self.shift = None
        self.register_buffer('num_tracked_values', torch.zeros(num_channels))
        self.register_buffer('running_mean', torch.zeros(num_channels))
        self.register_buffer('running_power', torch.ones(num_channels))

    def _ref_shape(self):
        return (1, self.num_channels) + (1,) * self.ndim

    def _to_ref(self, v):
        if self.flatten is not None:
            C, Fh = self.flatten
            v = v.reshape(Fh, C).t()
        return v.reshape(self._ref_shape())

    def _from_ref(self, v):
        v = v.reshape(-1)
        if self.flatten is not None:
            C, Fh = self.flatten
            v = v.reshape(C, Fh).t().reshape(-1)
        return v

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for name in ('scale', 'shift', 'num_tracked_values', 'running_mean', 'running_power'):
            v = getattr(self, name)
            if v is not None:
                destination[prefix + name] = self._to_ref(v.detach()).clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        for name in ('scale', 'shift', 'num_tracked_values', 'running_mean', 'running_power'):
            v = getattr(self, name)
            if v is None:
                continue
            if prefix + name in state_dict:
                with torch.no_grad():
                    v.copy_(self._from_ref(state_dict[prefix + name].to(v.device, v.dtype)))
            elif strict:
                missing_keys.append(prefix + name)


# =============================================================== conv stacks
class _Conv(nn.Module):
    """one conv: native weight (taps, Cout, Cin) + bias; reference weight (Cout, Cin*Fh, k[, k])."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size, flatten_height=1):
        super().__init__()
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.flatten_height = flatten_height
        if ndim == 2:
            kf, kt = (kernel_size, kernel_size) if isinstance(kernel_size, int) else kernel_size
            self.kernel = (kf, kt)
            self.taps = [(i - (kf - 1) // 2, j - (kt - 1) // 2) for i in range(kf) for j in range(kt)]
            fan = kf * kt
        else:
            k = int(kernel_size)
            self.kernel = (k,)
            self.taps = [(f, j - (k - 1) // 2) for f in range(flatten_height) for j in range(k)]
            fan = k
        cin_ref = in_channels * flatten_height
        bound = math.sqrt(6. / ((cin_ref + out_channels) * fan))          # xavier_uniform_
        self.weight = nn.Parameter(torch.empty(len(self.taps), out_channels, in_channels)
                                   .uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def _to_ref(self, w):
        Co, Ci, Fh = self.out_channels, self.in_channels, self.flatten_height
        if self.ndim == 2:
            kf, kt = self.kernel
            return w.reshape(kf, kt, Co, Ci).permute(2, 3, 0, 1).contiguous()
        k, = self.kernel
        return w.reshape(Fh, k, Co, Ci).permute(2, 3, 0, 1).reshape(Co, Ci * Fh, k).contiguous()

    def _from_ref(self, w):
        Co, Ci, Fh = self.out_channels, self.in_channels, self.flatten_height
        if self.ndim == 2:
            kf, kt = self.kernel
            return w.reshape(Co, Ci, kf, kt).permute(2, 3, 0, 1).reshape(kf * kt, Co, Ci)
        k, = self.kernel
        return w.reshape(Co, Ci, Fh, k).permute(2, 3, 0, 1).reshape(Fh * k, Co, Ci)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        destination[prefix + 'weight'] = self._to_ref(self.weight.detach())
        destination[prefix + 'bias'] = self.bias.detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        for name in ('weight', 'bias'):
            if prefix + name not in state_dict:
                if strict:
                    missing_keys.append(prefix + name)
                continue
            v = state_dict[prefix + name]
            p = getattr(self, name)
            with torch.no_grad():
                p.copy_((self._from_ref(v) if name == 'weight' else v).to(p.device, p.dtype))


class _ConvBlock(nn.Module):
    """one layer of a padertorch CNN: ``conv`` (weight, bias) and its ``norm`` (or None), so that the
    state-dict keys read ``convs.<i>.conv.{weight,bias}`` / ``convs.<i>.norm.{scale,shift,running_*}`` like
    padertorch's Conv modules -- the reference's init-checkpoint code takes the output layer's index from
    ``sorted(keys)[-1].split('.')[1]`` (weak_label_crnn/training.py:331-340)."""

    def __init__(self, conv, norm):
        super().__init__()
        self.conv = conv
        self.norm = norm


class _CNN(nn.Module):
    """CNN1d / CNN2d of padertorch.contrib.je.modules.conv (kwargs: training.py:218-242,250-259).

    pre_activation: layer i = norm -> relu -> 'same' zero pad -> conv -> pool (layer 0 bare when
    input_layer).  post-activation (output_net): conv -> norm -> relu, last layer bare when
    output_layer.  In both cases norm+relu are applied by the NEXT conv's operand load.
    """
    ndim = None

    def __init__(self, in_channels, out_channels, kernel_size, pool_size=1,
                 residual_connections=None, norm='batch', norm_kwargs=None, activation_fn='relu',
                 pre_activation=False, dropout=0., input_layer=True, output_layer=True,
                 flatten_height=1, **unused):
        super().__init__()
        if activation_fn != 'relu' or dropout != 0.:
            raise NotImplementedError('hot path covers relu / dropout 0 (training.py:226-241)')
        if residual_connections is not None and any(r is not None for r in residual_connections):
            raise NotImplementedError("residual 'deep' config: SURVEY 8f row 4")
        n = len(out_channels)
        ks = list(kernel_size) if isinstance(kernel_size, list) else [kernel_size] * n   # list = per layer
        assert len(ks) == n, (kernel_size, n)
        ps = list(pool_size) if isinstance(pool_size, (list, tuple)) and len(pool_size) == n \
            else [pool_size] * n
        self.in_channels = in_channels * flatten_height
        self.out_channels = list(out_channels)
        self.kernel_sizes, self.pool_sizes = ks, ps
        self.pre_activation, self.input_layer, self.output_layer = pre_activation, input_layer, output_layer
        self.flatten_height = flatten_height
        if pre_activation and output_layer:
            raise NotImplementedError('pre-activation stacks end bare on the hot path (output_layer=False)')
        if not pre_activation and not output_layer:
            raise NotImplementedError('post-activation stacks end with a bare conv on the hot path')
        nk = dict(norm_kwargs or {})
        eps = nk.get('eps', 1e-3)
        momentum = nk.get('momentum', 0.95)
        self.convs = nn.ModuleList()
        c = in_channels
        for i, co in enumerate(self.out_channels):
            fh = flatten_height if i == 0 else 1
            if pre_activation:
                has = not (i == 0 and input_layer)
                nc, fl = c * fh, ((c, fh) if fh > 1 else None)
            else:
                has = not (i == n - 1 and output_layer)
                nc, fl = co, None
            if norm not in ('batch', None):
                raise NotImplementedError(norm)
            self.convs.append(_ConvBlock(
                _Conv(self.ndim, c, co, ks[i], flatten_height=fh),
                Normalization(nc, self.ndim, eps=eps, momentum=momentum, flatten=fl)
                if (norm == 'batch' and has) else None))
            c = co

    @property
    def norms(self):
        """per layer: the Normalization stored with conv i (pre-activation: applied to its input;
        post-activation: to its output), or None."""
        return [blk.norm for blk in self.convs]

    def _pool(self, p):
        if p in (1, None, (1, 1)):
            return 1
        if self.ndim == 2:
            p = (p, p) if isinstance(p, int) else tuple(p)
            if p[1] != 1:
                raise NotImplementedError('time pooling is not on the hot path')
            return int(p[0])
        raise NotImplementedError('1-D pooling is not on the hot path')

    def _input_norm(self, i):
        """the Normalization applied to layer i's INPUT (fused into its operand load), or None."""
        if self.pre_activation:
            return self.norms[i]
        return self.norms[i - 1] if i > 0 else None

    @staticmethod
    def _bf16_ok(conv):
        """can this layer read AND write bf16 activation maps?  (the tensor-core / narrow kernels: channel counts
        that are multiples of 16, a tap table the kernels hold)"""
        return conv.in_channels % 16 == 0 and conv.out_channels % 16 == 0 and len(conv.taps) <= ops._lib.MAX_TAPS

    def forward_native(self, x, seq, stats_in=None, next_stats=None, act_dtype=torch.float32, last_act_dtype=torch.float32):
        """x (B,F,T,C) native -> (B,F',T,C').

        act_dtype: storage type of the maps BETWEEN the layers (``ops.act_dtype()``: bf16 in the 'bf16' mode; a
        layer writes bf16 only if it and its consumer have kernels for it); last_act_dtype: of the stack's output.

        stats_in: batch statistics of x already produced by the previous stack's last conv.
        next_stats: None, or 'c' / 'fc' -- also return the statistics of the output (per channel /
        per (f, c)) for a following stack's first norm; then the result is (y, stats)."""
        n = len(self.convs)
        stats = stats_in
        for i, blk in enumerate(self.convs):
            conv = blk.conv
            B, F_in, T, _ = x.shape
            if self.pre_activation:
                norm = self.norms[i]
                relu = not (i == 0 and self.input_layer)
            else:
                norm = self.norms[i - 1] if i > 0 else None
                relu = i > 0
            fh = conv.flatten_height
            assert F_in == fh or fh == 1, (F_in, fh)
            weight, taps = conv.weight, conv.taps
            if fh > 1 and len(taps) > ops._lib.MAX_TAPS:
                # a tall flatten ('b c f t -> b (c f) t' with f * k taps beyond the kernel's tap table, e.g. the
                # un-pooled doctest net of weak_label/crnn.py:16-34: 80 bands x k = 3): materialise it once --
                # (B,F,T,C) -> (B,1,T,F*C), channel index f*C + c, which is exactly the per-(f,c) norm index --
                # and run the same contraction as a plain 1-D conv over F*C channels
                k = len(taps) // fh
                x = x.permute(0, 2, 1, 3).reshape(B, 1, T, fh * x.shape[3])
                weight = weight.view(fh, k, weight.shape[1], weight.shape[2]).permute(1, 2, 0, 3) \
                    .reshape(k, weight.shape[1], fh * weight.shape[2])
                taps = [(0, dt) for (_, dt) in taps[:k]]
                if stats is not None and stats.shape[0] < fh * conv.in_channels:
                    stats = None
                F_in, fh = 1, 1
            # does the consumer of this layer's output normalise with batch statistics?  then the conv
            # epilogue accumulates them (pooled layers: the statistics are of the pooled map -> own pass)
            if i + 1 < n:
                nxt = self._input_norm(i + 1)
                want, want_pf = self.training and nxt is not None and not nxt.frozen_stats, False
            else:
                want, want_pf = self.training and next_stats is not None, next_stats == 'fc'
            # a frozen layer (freeze(..., freeze_norm_stats=True)) normalises with its running statistics even in
            # train mode and does not update them (padertorch Normalization.freeze)
            live = self.training and not (norm is not None and norm.frozen_stats)
            if stats is not None and not live:
                stats = None
            out_dtype = torch.float32
            if act_dtype == torch.bfloat16:
                writes = self._bf16_ok(conv) or (conv.in_channels == 1 and conv.out_channels in (16, 32))
                if i + 1 < n:
                    out_dtype = torch.bfloat16 if (writes and self._bf16_ok(self.convs[i + 1].conv)) else torch.float32
                else:
                    out_dtype = last_act_dtype if writes else torch.float32
            if x.dtype == torch.bfloat16 and not self._bf16_ok(conv):
                x = x.float()                   # a layer without bf16 kernels (odd channel counts): widen its input
            cfg = dict(out_dtype=out_dtype, F_in=F_in, F_out=1 if fh > 1 else F_in, taps=taps, relu=relu,
                       per_f=fh > 1, pool=self._pool(self.pool_sizes[i]), norm=norm is not None,
                       eps=norm.eps if norm is not None else 0.,
                       momentum=norm.momentum if norm is not None else 0.,
                       training=live, want_stats=want, stats_per_f=want_pf)
            if norm is not None:
                x, stats = ops.ConvLayerFn.apply(x, weight, conv.bias, norm.scale, norm.shift,
                                                 norm.running_mean, norm.running_power,
                                                 norm.num_tracked_values, seq, cfg, stats)
            else:
                x, stats = ops.ConvLayerFn.apply(x, weight, conv.bias, None, None, None, None, None,
                                                 seq, cfg, None)
            if stats.numel() == 0:
                stats = None
        if next_stats is not None:
            return x, stats
        return x

    def forward(self, x, seq_len=None):
        xn = to_native(x) if self.ndim == 2 else to_native(x).unsqueeze(1)
        seq = SeqLen.make(seq_len, xn.shape[0], xn.shape[2], xn.device)
        y = self.forward_native(xn, seq)
        return (from_native(y) if self.ndim == 2 else from_native(y.squeeze(1))), seq_len

    def freeze(self, num_layers=None, freeze_norm_stats=True):
        """training.py:343-350: stop the gradients of the first ``num_layers`` layers; with
        ``freeze_norm_stats`` (the reference experiments' default) their norms also stop tracking: they use
        the running statistics in train mode and leave them untouched."""
        n = len(self.convs) if num_layers is None else num_layers
        for i in range(n):
            for p in self.convs[i].conv.parameters():
                p.requires_grad = False
            norm = self.convs[i].norm
            if norm is not None:
                for p in norm.parameters():
                    p.requires_grad = False
                norm.frozen_stats = bool(freeze_norm_stats)


class CNN2d(_CNN):
    ndim = 2


class CNN1d(_CNN):
    ndim = 1


class CNN(nn.Module):
    """padertorch.contrib.je.modules.hybrid.CNN: cnn_2d -> 'b c f t -> b (c f) t' -> cnn_1d
    (call sites weak_label/crnn.py:93, strong_label/crnn.py:86; sizes :326-330 / :180-184).

    ``cnn_2d`` / ``cnn_1d`` are kwargs dicts (as in the reference config) or built modules."""

    def __init__(self, cnn_2d, cnn_1d, input_height=None, positional_encoding=False,
                 conditional_dims=0, **unused):
        super().__init__()
        if positional_encoding:
            raise NotImplementedError('positional_encoding is off on the hot path')
        self.input_height, self.conditional_dims = input_height, conditional_dims
        if isinstance(cnn_2d, dict):
            cnn_2d = CNN2d(**{k: v for k, v in cnn_2d.items() if k != 'factory'})
        self.cnn_2d = cnn_2d
        height = input_height
        for p in cnn_2d.pool_sizes:
            height //= cnn_2d._pool(p)
        if isinstance(cnn_1d, dict):
            kw = {k: v for k, v in cnn_1d.items() if k not in ('factory', 'in_channels', 'input_layer')}
            cnn_1d = CNN1d(in_channels=cnn_2d.out_channels[-1], input_layer=False,
                           flatten_height=height, **kw)
        self.cnn_1d = cnn_1d
        assert cnn_1d.flatten_height == height, (cnn_1d.flatten_height, height)

    def forward_native(self, x, seq, condition=None):
        if self.conditional_dims:
            cond = condition.reshape(condition.shape[0], -1)
            x = ops.ConcatCondFn.apply(x, cond)
        # the first cnn_1d norm is indexed per (f, c) of the cnn_2d output: its statistics come out of
        # the last cnn_2d conv's epilogue
        n0 = self.cnn_1d._input_norm(0)
        want = 'fc' if (self.training and n0 is not None and not n0.frozen_stats) else None
        # 'bf16' mode: the maps between the conv layers (and their gradients) live in HBM as bf16; the stack's
        # output (the GRU input) is fp32
        adt = ops.act_dtype()
        mid = adt if (adt == torch.bfloat16 and _CNN._bf16_ok(self.cnn_1d.convs[0].conv)) else torch.float32
        if want:
            x, stats = self.cnn_2d.forward_native(x, seq, next_stats=want, act_dtype=adt, last_act_dtype=mid)
        else:
            x, stats = self.cnn_2d.forward_native(x, seq, act_dtype=adt, last_act_dtype=mid), None
        x = self.cnn_1d.forward_native(x, seq, stats_in=stats, act_dtype=adt)            # (B,1,T,D) fp32
        return x.squeeze(1)

    def forward(self, x, seq_len=None, condition=None):
        xn = to_native(x)
        seq = SeqLen.make(seq_len, xn.shape[0], xn.shape[2], xn.device)
        return from_native(self.forward_native(xn, seq, condition)), seq_len


# =============================================================== GRU
class GRUCore(nn.Module):
    """parameter holder + driver of the persistent GRU kernels; state-dict compatible with
    ``torch.nn.GRU`` (``weight_ih_l0`` ... ``bias_hh_l1_reverse``; training.py:331-332)."""

    def __init__(self, input_size, hidden_size, num_layers=1, bias=True, dropout=0.,
                 bidirectional=False, batch_first=True, **unused):
        super().__init__()
        if not bias or dropout != 0.:
            raise NotImplementedError('hot path: bias=True, dropout=0 (training.py:245-249)')
        if hidden_size % 32 or hidden_size > 256:
            raise NotImplementedError('GRU kernel: hidden_size multiple of 32, <= 256')
        self.input_size, self.hidden_size = input_size, hidden_size
        self.num_layers, self.bidirectional = num_layers, bidirectional
        nd = 2 if bidirectional else 1
        k = 1. / math.sqrt(hidden_size)
        for l in range(num_layers):
            inp = input_size if l == 0 else hidden_size * nd
            # stacked per direction: the kernel runs both directions of a layer in one launch
            self.register_parameter(f'w_ih_{l}', nn.Parameter(torch.empty(nd, 3 * hidden_size, inp).uniform_(-k, k)))
            self.register_parameter(f'w_hh_{l}', nn.Parameter(torch.empty(nd, 3 * hidden_size, hidden_size).uniform_(-k, k)))
            self.register_parameter(f'b_ih_{l}', nn.Parameter(torch.empty(nd, 3 * hidden_size).uniform_(-k, k)))
            self.register_parameter(f'b_hh_{l}', nn.Parameter(torch.empty(nd, 3 * hidden_size).uniform_(-k, k)))

    _REF = (('weight_ih', 'w_ih'), ('weight_hh', 'w_hh'), ('bias_ih', 'b_ih'), ('bias_hh', 'b_hh'))

    def _names(self):
        for l in range(self.num_layers):
            for d in range(2 if self.bidirectional else 1):
                yield l, d, f'_l{l}' + ('_reverse' if d else '')

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        for l, d, sfx in self._names():
            for ref, nat in self._REF:
                destination[prefix + ref + sfx] = getattr(self, f'{nat}_{l}')[d].detach().clone()

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        with torch.no_grad():
            for l, d, sfx in self._names():
                for ref, nat in self._REF:
                    key = prefix + ref + sfx
                    p = getattr(self, f'{nat}_{l}')
                    if key in state_dict:
                        p[d].copy_(state_dict[key].to(p.device, p.dtype))
                    elif strict:
                        missing_keys.append(key)

    def layer_params(self, l):
        return tuple(getattr(self, f'{n}_{l}') for n in ('w_ih', 'w_hh', 'b_ih', 'b_hh'))

    def forward_native(self, x, seq, reverse=False):
        """x (B,T,In) -> (B,T,H*ndir)."""
        return gru_stack([self], [x], seq, [reverse])[0]


def gru_stack(cores, xs, seq, reverses):
    """run several GRUCore modules of identical shape side by side: every layer is ONE launch that
    advances all of them concurrently (the reference runs rnn_fwd and rnn_bwd back to back)."""
    c0 = cores[0]
    assert all((c.hidden_size, c.num_layers, c.bidirectional) ==
               (c0.hidden_size, c0.num_layers, c0.bidirectional) for c in cores)
    meta = [[False, True] if c.bidirectional else [bool(r)] for c, r in zip(cores, reverses)]
    xs = list(xs)
    for l in range(c0.num_layers):
        flat = []
        for c, x in zip(cores, xs):
            flat += [x, *c.layer_params(l)]
        xs = list(ops.GruMultiFn.apply(seq, meta, *flat))
    return xs


class GRU(nn.Module):
    """padertorch.contrib.je.modules.rnn.GRU (kwargs training.py:243-260; weak_label/crnn.py:340
    ``reverse=True``; strong_label/crnn.py:189-195 ``bidirectional``).  ``rnn`` / ``output_net`` are
    kwargs dicts or built modules; ``rnn=None`` keeps only the output_net (weak_label/crnn.py:333)."""

    def __init__(self, rnn, output_net, reverse=False, **unused):
        super().__init__()
        if isinstance(rnn, dict):
            rnn = GRUCore(**{k: v for k, v in rnn.items() if k != 'factory'})
        if isinstance(output_net, dict):
            kw = {k: v for k, v in output_net.items() if k != 'factory'}
            if kw.get('in_channels') is None and rnn is not None:
                kw['in_channels'] = rnn.hidden_size * (2 if rnn.bidirectional else 1)
            output_net = CNN1d(**kw)
        self.rnn, self.output_net, self.reverse = rnn, output_net, reverse
        if reverse and rnn is not None and rnn.bidirectional:
            raise NotImplementedError('reverse=True with a bidirectional rnn is not used by pb_sed')

    def forward_native(self, x, seq):
        """x (B,T,F) -> logits (B,T,K)."""
        if self.rnn is not None:
            x = self.rnn.forward_native(x, seq, self.reverse)
        elif self.reverse:
            raise NotImplementedError('reverse=True needs an rnn')
        return self.output_net.forward_native(x.unsqueeze(1), seq).squeeze(1)

    def forward(self, x, seq_len=None):
        xn = to_native(x)
        seq = SeqLen.make(seq_len, xn.shape[0], xn.shape[1], xn.device)
        return from_native(self.forward_native(xn, seq)), seq_len


# =============================================================== feature extractor
def _sampler_cfg(cfg, *names, default=None):
    """read ``scale`` / ``truncation`` ... from a padertorch sampling-fn config dict or object."""
    out = []
    for n in names:
        v = cfg.get(n, default) if isinstance(cfg, dict) else getattr(cfg, n, default)
        out.append(v)
    return out


class NormalizedLogMelExtractor(nn.Module):
    """padertorch.contrib.je.modules.features.NormalizedLogMelExtractor on the GPU (kwargs
    training.py:190-217; call weak_label/crnn.py:86-90).

    Accepts the reference's 5-D ``stft`` (B,1,T,F,2) OR the raw waveform (B,1,S) / (B,S), in which
    case the STFT of ``data_preparation/provider.py:315-323`` (``stft_kwargs``) runs fused in the same
    kernel.  Output (B,1,n_mels,T): log-mel, cumulative running mean/var normalisation per band
    (eps 1e-5, no affine), clamp(+-6), padded frames zeroed.

    Train-time augmentation (all random draws are made ON THE DEVICE with torch's
    generator, so the step stays CUDA-graph capturable, and are kept in ``last_augmentation``):
      * ``frequency_warping_fn`` = MelWarping config (training.py:195-208): per-example warped
        filterbanks, built by ``pbsed_make_warped_fbank`` and consumed by the fused STFT/mel kernel;
      * ``n_time_masks`` / ``n_frequency_masks`` (+ max steps / rates, training.py:210-215) and
        ``max_noise_scale`` (:216): fused into the normalise/clamp pass;
      * ``time_warping`` = dict(anchor=(.4,.6), anchor_shift=(-.1,.1)) (provider.py:329-338, the
        TimeWarpedSTFT wrapper of transform.py:36-45): non-uniform frame onsets inside the STFT kernel
        (raw-audio input only); frame-level targets are re-sampled on the same grid.
    """

    def __init__(self, sample_rate, stft_size, number_of_filters, num_channels=1,
                 lowest_frequency=50., highest_frequency=None, add_deltas=False,
                 add_delta_deltas=False, norm_eps=1e-5, clamp=6., stft_kwargs=None,
                 frequency_warping_fn=None, blur_sigma=0., n_time_masks=0, max_masked_time_steps=70,
                 max_masked_time_rate=.2, n_frequency_masks=0, max_masked_frequency_bands=20,
                 max_masked_frequency_rate=.2, max_noise_scale=0., time_warping=None, **unused):
        super().__init__()
        if add_deltas or add_delta_deltas or num_channels != 1:
            raise NotImplementedError('deltas / multi-channel input are off on the hot path')
        if blur_sigma:
            raise NotImplementedError('blur_sigma is commented out in the reference config (training.py:209)')
        self.sample_rate, self.stft_size, self.number_of_filters = sample_rate, stft_size, number_of_filters
        self.add_deltas, self.add_delta_deltas = add_deltas, add_delta_deltas
        self.norm_eps, self.clamp = norm_eps, clamp
        self.lowest_frequency = lowest_frequency
        self.highest_frequency = sample_rate / 2 if highest_frequency is None else highest_frequency
        fb = mel_filterbank(sample_rate, stft_size, number_of_filters, lowest_frequency, highest_frequency)
        self.register_buffer('fbanks', torch.from_numpy(fb.T.copy()).float(), persistent=False)   # (F, n_mels); not a checkpoint key
        lo, hi, w, stride = sparse_filterbank(fb.astype(np.float32))
        self.register_buffer('_fb_lo', torch.from_numpy(lo), persistent=False)
        self.register_buffer('_fb_hi', torch.from_numpy(hi), persistent=False)
        self.register_buffer('_fb_w', torch.from_numpy(w), persistent=False)
        self._fb_stride = stride
        kw = dict(shift=320, window_length=960, size=stft_size, fading='half', pad=True)
        kw.update(stft_kwargs or {})
        kw.setdefault('window_length', kw['size'])
        assert kw['size'] == stft_size
        self.stft_kwargs = kw
        self.register_buffer('_window', torch.from_numpy(blackman_window(kw['window_length'])).float(),
                             persistent=False)
        self.norm = Normalization(number_of_filters, 2, eps=norm_eps, momentum=-1., affine=False)
        # reference buffer shape is (1, 1, n_mels, 1): statistics axis 'bt' of 'bcft'
        self.norm._ref_shape = lambda: (1, 1, number_of_filters, 1)
        # ---- augmentation config
        self.warping = None
        if frequency_warping_fn is not None:
            wf, bf, hf = _sampler_cfg(frequency_warping_fn, 'warp_factor_sampling_fn',
                                      'boundary_frequency_ratio_sampling_fn', 'highest_frequency')
            (w_scale, w_trunc), (b_scale, b_trunc) = _sampler_cfg(wf, 'scale', 'truncation'), _sampler_cfg(bf, 'scale', 'truncation')
            self.warping = dict(scale=float(w_scale), truncation=float(w_trunc), ratio_scale=float(b_scale),
                                ratio_truncation=float(b_trunc),
                                highest_frequency=float(hf if hf is not None else sample_rate / 2))
        self.n_time_masks, self.max_masked_time_steps, self.max_masked_time_rate = \
            n_time_masks, max_masked_time_steps, max_masked_time_rate
        self.n_frequency_masks, self.max_masked_frequency_bands, self.max_masked_frequency_rate = \
            n_frequency_masks, max_masked_frequency_bands, max_masked_frequency_rate
        self.max_noise_scale = max_noise_scale
        self.time_warping = time_warping
        self.last_augmentation = None
        self.next_augmentation = None      # test hook: draws to use instead of sampling (consumed once)

    def _fb(self):
        return dict(lo=self._fb_lo, hi=self._fb_hi, w=self._fb_w, stride=self._fb_stride,
                    n_mels=self.number_of_filters)

    def augments(self):
        return bool(self.warping or self.n_time_masks or self.n_frequency_masks or self.max_noise_scale
                    or self.time_warping)

    # ---- random draws (device-side, graph capturable)
    @staticmethod
    def _masks(n, lens, max_steps, max_rate, dev):
        """(B, n, 2) int32 (onset, width): width ~ U{0..min(max_steps, floor(rate*len))}, onset ~ U{0..len-width}."""
        B = lens.shape[0]
        max_w = torch.minimum(torch.full_like(lens, float(max_steps)), torch.floor(lens * max_rate))
        w = torch.floor(torch.rand((B, n), device=dev) * (max_w[:, None] + 1.)).clamp_(max=max_w[:, None])
        on = torch.floor(torch.rand((B, n), device=dev) * (lens[:, None] - w + 1.)).clamp_(max=(lens[:, None] - w))
        return torch.stack([on, w], -1).to(torch.int32).contiguous()

    def sample_augmentation(self, B, T, seq, dev, from_audio):
        aug = {}
        if self.warping:
            c = self.warping
            a = c['truncation'] / c['scale']
            p_lo = .5 * (1. + math.erf(-a / math.sqrt(2.)))
            u = p_lo + torch.rand(B, device=dev) * (1. - 2. * p_lo)
            z = math.sqrt(2.) * torch.erfinv((2. * u - 1.).clamp(-1. + 1e-7, 1. - 1e-7))
            aug['alpha'] = torch.exp(c['scale'] * z.clamp(-a, a))                        # LogTruncatedNormal
            u = torch.rand(B, device=dev)
            aug['ratio'] = -c['ratio_scale'] * torch.log1p(-u * (1. - math.exp(-c['ratio_truncation'] / c['ratio_scale'])))
        lens = seq.dev.float() if seq.dev is not None else torch.full((B,), float(T), device=dev)
        if self.n_time_masks:
            aug['time_masks'] = self._masks(self.n_time_masks, lens, self.max_masked_time_steps,
                                            self.max_masked_time_rate, dev)
        if self.n_frequency_masks:
            bands = torch.full((B,), float(self.number_of_filters), device=dev)
            aug['freq_masks'] = self._masks(self.n_frequency_masks, bands, self.max_masked_frequency_bands,
                                            self.max_masked_frequency_rate, dev)
        if self.max_noise_scale:
            aug['noise_scale'] = torch.rand(B, device=dev) * self.max_noise_scale
            aug['noise'] = torch.randn((B, self.number_of_filters, T), device=dev)
        if self.time_warping and from_audio:
            (a0, a1), (s0, s1) = self.time_warping['anchor'], self.time_warping['anchor_shift']
            aug['anchor'] = a0 + torch.rand(B, device=dev) * (a1 - a0)
            aug['anchor_shift'] = s0 + torch.rand(B, device=dev) * (s1 - s0)
        return aug

    def _time_warp_grid(self, aug, T, shift):
        """source position (frames, float64) and first sample (int32) of every output frame."""
        a_in = aug['anchor'].double()[:, None] * T
        a_out = ((aug['anchor'] + aug['anchor_shift']).double()[:, None] * T).clamp(1., T - 1.)
        t = torch.arange(T, device=a_in.device, dtype=torch.float64)[None]
        src = torch.where(t <= a_out, t * a_in / a_out, a_in + (t - a_out) * (T - a_in) / (T - a_out))
        return src, torch.floor(src * shift + .5).to(torch.int32).contiguous()

    def forward(self, x, seq_len=None, targets=None):
        with torch.no_grad():
            n = self.norm
            train = self.training
            from_audio = x.dim() != 5
            if from_audio:
                a = x.reshape(x.shape[0], -1)
                B, S = a.shape
                kw = self.stft_kwargs
                T = stft_num_frames(S, kw['shift'], kw['window_length'], kw['fading'], kw['pad'])
            else:
                B, C, T = x.shape[:3]
                assert C == 1
            seq = SeqLen.make(seq_len, B, T, x.device)
            aug = {}
            if train and self.augments():
                aug = self.next_augmentation if self.next_augmentation is not None \
                    else self.sample_augmentation(B, T, seq, x.device, from_audio)
                self.next_augmentation = None
            self.last_augmentation = aug
            fb = self._fb()
            if 'alpha' in aug:
                c = self.warping
                fb = ops.make_warped_fbank(aug['alpha'], aug['ratio'], self.number_of_filters, self.stft_size // 2 + 1,
                                           float(hz2mel(self.lowest_frequency)), float(hz2mel(self.highest_frequency)),
                                           float(hz2mel(c['highest_frequency'])), self.stft_size / self.sample_rate,
                                           self.stft_size // 2 + 1)
            stats = ops._stats_buffer(self.number_of_filters, x.device) if train else None
            if from_audio:
                pad_front = {'half': (kw['window_length'] - kw['shift']) // 2, 'full': kw['window_length'] - kw['shift'],
                             True: kw['window_length'] - kw['shift']}.get(kw['fading'], 0)
                cfg = dict(shift=kw['shift'], window_length=kw['window_length'], size=kw['size'],
                           pad_front=pad_front, T=T, window=self._window)
                frame_start = None
                if 'anchor' in aug:
                    src, frame_start = self._time_warp_grid(aug, T, kw['shift'])
                    if targets is not None:     # frame-level targets follow the warped grid (nearest frame)
                        idx = torch.floor(src + .5).long().clamp_(0, T - 1)
                        targets = tuple(tg if tg.dim() < 3 else
                                        torch.gather(tg, 2, idx[:, None, :].expand(-1, tg.shape[1], -1))
                                        for tg in targets)
                y = ops.logmel_from_audio(a, cfg, fb, seq, stats, frame_start)
            else:
                y = ops.logmel_from_stft(x.reshape(B, T, x.shape[3], 2), fb, seq, stats)
            count = ops._sync_count_(stats, self.number_of_filters, seq.frames()) if train else 1.
            scale, shift = ops.norm_finalize(stats, count, self.number_of_filters,
                                             None, None, n.eps, -1., train, n.running_mean,
                                             n.running_power, n.num_tracked_values, x.device)
            ops.logmel_normalize_(y, scale, shift, self.clamp, seq, aug.get('time_masks'), aug.get('freq_masks'),
                                  aug.get('noise'), aug.get('noise_scale'))
            y = y.unsqueeze(1)                                            # (B,1,n_mels,T)
        if targets is None:
            return y, seq_len
        return y, seq_len, targets
