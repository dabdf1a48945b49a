"""Oracle restatement of the pb_sed-owned model arithmetic + builders + train step.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Plain PyTorch on CPU.

* ``FBCRNN``  restates ``pb_sed/models/weak_label/crnn.py:58-206`` (sigmoid,
  forward order, weak / strong forward-backward BCE) and the heads ``:223-239``.
* ``BiCRNN``  restates ``pb_sed/models/strong_label/crnn.py:60-138,200-210``.
* ``build_fbcrnn`` / ``build_bicrnn`` wire module sizes the way
  ``finalize_dogmatic_config`` does (``weak_label/crnn.py:304-340``,
  ``strong_label/crnn.py:155-198``) with the shallow-net constants of
  ``pb_sed/experiments/weak_label_crnn/training.py:158-260``.
* ``train_step`` restates the padertorch trainer step body (SURVEY App. A
  [R]): forward -> review -> backward -> clip_grad_norm_ -> Adam -> zero_grad.

These restatements are pinned against the REAL pb_sed classes (executed
through ``oracle/ref_loader.py``) by ``tests/golden/make_golden.py``; the
third-party module arithmetic underneath (``oracle/pt_port.py``) is unpinned.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import pt_port as P

SHALLOW_2D = [16, 16, 32, 32, 64, 64, 128, 128, 256]          # training.py:161-165
SHALLOW_POOL_2D = 4 * [1, (2, 1)] + [1]                       # training.py:167
SHALLOW_K1D = [1, 3, 3, 3, 1]                                 # training.py:168


# ------------------------------------------------------------------ models
class FBCRNN(nn.Module):
    def __init__(self, feature_extractor, cnn, rnn_fwd, rnn_bwd, *, minimum_score=1e-5,
                 label_smoothing=0., slat=False, strong_fwd_bwd_loss_weight=1.,
                 class_weights=None):
        super().__init__()
        self.feature_extractor, self.cnn = feature_extractor, cnn
        self.rnn_fwd, self.rnn_bwd = rnn_fwd, rnn_bwd
        self.minimum_score = minimum_score
        self.label_smoothing = label_smoothing
        self.slat = slat
        self.strong_fwd_bwd_loss_weight = strong_fwd_bwd_loss_weight
        self.class_weights = None if class_weights is None else torch.Tensor(class_weights)

    def sigmoid(self, y):                                     # weak_label/crnn.py:58-59
        return self.minimum_score + (1 - 2 * self.minimum_score) * torch.sigmoid(y)

    def logits(self, inputs):
        """pre-sigmoid output_net outputs (the quantity BASELINE.json's
        'frame-logit max|delta|' is taken on) + features."""
        x = inputs['stft']
        seq_len = None if inputs.get('seq_len') is None else np.array(inputs['seq_len'])
        x, seq_len_x = self.feature_extractor(x, seq_len=seq_len)
        h, seq_len_h = self.cnn(x, seq_len_x)
        z_fwd, seq_len_y = self.rnn_fwd(h, seq_len_h)
        z_bwd = None if self.rnn_bwd is None else self.rnn_bwd(h, seq_len_h)[0]
        return z_fwd, z_bwd, seq_len_y, x, h

    def forward(self, inputs):                                # weak_label/crnn.py:69-100
        z_fwd, z_bwd, seq_len_y, x, _ = self.logits(inputs)
        targets = None
        if 'weak_targets' in inputs:
            targets = (inputs['weak_targets'],)
            if 'boundary_targets' in inputs:
                targets = targets + (inputs['boundary_targets'],)
        y_bwd = None if z_bwd is None else self.sigmoid(z_bwd)
        return self.sigmoid(z_fwd), y_bwd, seq_len_y, x, seq_len_y, targets

    def loss(self, y_fwd, y_bwd, seq_len, targets):           # weak_label/crnn.py:117-153
        wt = targets[0]
        m_w = ((wt < .01) | (wt > .99)).to(y_fwd.dtype)
        wt = wt * m_w
        tw = wt
        if self.label_smoothing > 0.:
            tw = tw.clamp(self.label_smoothing, 1 - self.label_smoothing)
        if y_bwd is None:
            y_weak = P.TakeLast(axis=2)(y_fwd, seq_len)
            loss = F.binary_cross_entropy(y_weak, tw, reduction='none')[..., None].expand(y_fwd.shape)
        else:
            loss = F.binary_cross_entropy(torch.maximum(y_fwd, y_bwd),
                                          tw[..., None].expand(y_fwd.shape), reduction='none')
        loss = loss * m_w[..., None]
        if self.strong_fwd_bwd_loss_weight > 0.:
            bt = wt[..., None].expand(y_fwd.shape) if self.slat else targets[1]
            m_b = ((bt > .99) | (bt < .01)).to(y_fwd.dtype)
            m_b = m_b * (m_b.mean(-1, keepdim=True) > .999) * (wt > .99)[..., None]
            if (m_b == 1).any():
                tb = bt
                if self.label_smoothing > 0.:
                    tb = tb.clamp(self.label_smoothing, 1 - self.label_smoothing)
                t_fwd = torch.cummax(tb, dim=-1)[0]
                t_bwd = torch.cummax(tb.flip(-1), dim=-1)[0].flip(-1)
                ls = F.binary_cross_entropy(y_fwd, t_fwd, reduction='none')
                if y_bwd is not None:
                    ls = ls / 2 + F.binary_cross_entropy(y_bwd, t_bwd, reduction='none') / 2
                w = m_b * self.strong_fwd_bwd_loss_weight
                loss = w * ls + (1. - w) * loss
        loss = P.Mean(axis=-1)(loss, seq_len)
        weights = m_w if self.class_weights is None else m_w * self.class_weights
        return (loss * weights).sum() / weights.sum()

    def review(self, inputs, outputs):
        y_fwd, y_bwd, seq_len, x, _, targets = outputs
        return dict(loss=self.loss(y_fwd, y_bwd, seq_len, targets))

    # heads, weak_label/crnn.py:223-239
    def tagging(self, inputs):
        y_fwd, y_bwd, seq_len_y, *_ = self.forward(inputs)
        last = P.TakeLast(axis=-1, keepdims=True)(y_fwd, seq_len_y)
        if y_bwd is None:
            return last, np.ones_like(seq_len_y)
        return (last + y_bwd[..., :1]) / 2, np.ones_like(seq_len_y)

    def boundaries_detection(self, inputs):
        y_fwd, y_bwd, seq_len_y, *_ = self.forward(inputs)
        m = P.compute_mask(y_fwd, seq_len_y, 0, -1)
        return torch.minimum(y_fwd * m, y_bwd * m), seq_len_y


class BiCRNN(nn.Module):
    def __init__(self, feature_extractor, cnn, rnn, *, tag_conditioning=False):
        super().__init__()
        self.feature_extractor, self.cnn, self.rnn = feature_extractor, cnn, rnn
        self.tag_conditioning = tag_conditioning

    def logits(self, inputs):                                 # strong_label/crnn.py:60-92
        x = inputs['stft']
        seq_len = np.array(inputs['seq_len'])
        x, seq_len_x = self.feature_extractor(x, seq_len=seq_len)
        tag = inputs['tag_condition'].unsqueeze(-1) if self.tag_conditioning else None
        h, seq_len_h = self.cnn(x, seq_len_x, tag if self.cnn.conditional_dims else None)
        if self.tag_conditioning:
            b, _, t = h.shape
            h = torch.cat([h, tag.to(h.dtype).expand(b, tag.shape[1], t)], dim=1)
        z, seq_len_y = self.rnn(h, seq_len_h)
        return z, seq_len_y, x

    def forward(self, inputs):
        z, seq_len_y, x = self.logits(inputs)
        targets = None
        if 'strong_targets' in inputs:
            targets = (inputs['weak_targets'], inputs['strong_targets'])
        return torch.sigmoid(z), seq_len_y, x, seq_len_y, targets

    def loss(self, y, seq_len_y, targets):                    # strong_label/crnn.py:107-112
        st = targets[1]
        m = ((st > .99) | (st < .01)).to(y.dtype)
        bce = F.binary_cross_entropy(y, st, reduction='none') * m
        return P.Sum(axis=-1)(bce, seq_len_y).sum() / m.sum()

    def review(self, inputs, outputs):
        y, seq_len_y, x, _, targets = outputs
        return dict(loss=self.loss(y, seq_len_y, targets))

    def sound_event_detection(self, inputs):                  # strong_label/crnn.py:207-210
        y, seq_len_y, *_ = self.forward(inputs)
        return y * P.compute_mask(y, seq_len_y, 0, -1), seq_len_y

    def tagging(self, inputs):                                # strong_label/crnn.py:200-202
        y, seq_len_y, *_ = self.forward(inputs)
        return y.max(-1, keepdim=True)[0], np.ones_like(seq_len_y)


# ---------------------------------------------------------------- builders
def _modules(num_events, n_mels, stft_size, sample_rate, channels_2d, pool_2d, channels_1d,
             k1d, hidden, rnn_layers, out_hidden, in_channels_2d=1, conditional_dims=0,
             bidirectional=False, rnn_extra_in=0):
    fe = P.NormalizedLogMelExtractor(sample_rate, stft_size, n_mels)
    nk = dict(eps=1e-3)
    cnn_2d = P.CNN2d(in_channels_2d, channels_2d, 3, pool_size=pool_2d, norm='batch',
                     norm_kwargs=nk, pre_activation=True, output_layer=False)
    height = n_mels
    for p in cnn_2d.pool_sizes:
        if p not in (1, None):
            height //= (p if isinstance(p, int) else p[0])
    cnn_1d = P.CNN1d(channels_2d[-1] * height, channels_1d, k1d, norm='batch', norm_kwargs=nk,
                     pre_activation=True, input_layer=False, output_layer=False)
    cnn = P.CNN(cnn_2d, cnn_1d, input_height=n_mels, conditional_dims=conditional_dims)

    def rnn(reverse=False):
        gru = nn.GRU(channels_1d[-1] + rnn_extra_in, hidden, num_layers=rnn_layers,
                     batch_first=True, bidirectional=bidirectional)
        out = P.CNN1d(hidden * (2 if bidirectional else 1), [out_hidden, num_events], 1,
                      norm='batch', norm_kwargs=nk, pre_activation=False, output_layer=True)
        return P.GRU(gru, out, reverse=reverse)
    return fe, cnn, rnn


def build_fbcrnn(num_events=10, n_mels=128, stft_size=1024, sample_rate=16000,
                 channels_2d=SHALLOW_2D, pool_2d=SHALLOW_POOL_2D, channels_1d=5 * [256],
                 k1d=SHALLOW_K1D, hidden=256, rnn_layers=2, out_hidden=256, seed=0, **kw):
    """the reference's default ('shallow', DESED) FBCRNN; ~3.49 M parameters."""
    torch.manual_seed(seed)
    fe, cnn, rnn = _modules(num_events, n_mels, stft_size, sample_rate, list(channels_2d),
                            list(pool_2d), list(channels_1d), list(k1d), hidden, rnn_layers,
                            out_hidden)
    return FBCRNN(fe, cnn, rnn(False), rnn(True), **kw)


def build_bicrnn(num_events=10, n_mels=128, stft_size=1024, sample_rate=16000,
                 channels_2d=SHALLOW_2D, pool_2d=SHALLOW_POOL_2D, channels_1d=5 * [256],
                 k1d=SHALLOW_K1D, hidden=256, rnn_layers=2, out_hidden=256, seed=0,
                 tag_conditioning=True):
    """tag-conditioned BiCRNN (strong_label_crnn/training.py:245-262: 2-layer BiGRU)."""
    torch.manual_seed(seed)
    k = num_events if tag_conditioning else 0
    fe, cnn, rnn = _modules(num_events, n_mels, stft_size, sample_rate, list(channels_2d),
                            list(pool_2d), list(channels_1d), list(k1d), hidden, rnn_layers,
                            out_hidden, in_channels_2d=1 + k, conditional_dims=k,
                            bidirectional=True, rnn_extra_in=k)
    return BiCRNN(fe, cnn, rnn(False), tag_conditioning=tag_conditioning)


def tiny_fbcrnn(num_events=10, seed=0, **kw):
    """the doctest-sized net of weak_label/crnn.py:16-30 (+2 pool layers so pooling is covered)."""
    return build_fbcrnn(num_events, n_mels=16, stft_size=64, channels_2d=[8, 8, 16],
                        pool_2d=[1, (2, 1), (2, 1)], channels_1d=[32, 32], k1d=[3, 1],
                        hidden=32, rnn_layers=2, out_hidden=16, seed=seed, **kw)


# ------------------------------------------------------------ synthetic data
def synthetic_audio(batch, num_samples=160000, seed=1234):
    """low-passed noise + gated sinusoid 'events', peak-normalised (float32 numpy)."""
    rng = np.random.RandomState(seed)
    x = rng.randn(batch, num_samples).astype(np.float64)
    out = np.empty_like(x)
    t = np.arange(num_samples) / 16000.
    for b in range(batch):
        a = rng.uniform(0.5, 0.98)
        # one-pole low-pass via FFT-domain response (vectorised; same for every caller)
        spec = np.fft.rfft(x[b])
        w = np.exp(-2j * np.pi * np.arange(spec.shape[0]) / num_samples)
        y = np.fft.irfft(spec * (1 - a) / (1 - a * w), n=num_samples)
        for _ in range(rng.randint(1, 4)):
            f0 = rng.uniform(200., 6000.)
            on = rng.randint(0, max(num_samples - 1600, 1))
            off = min(on + rng.randint(1600, max(num_samples // 2, 1601)), num_samples)
            y[on:off] += rng.uniform(0.2, 2.) * y.std() * np.sin(2 * np.pi * f0 * t[on:off])
        out[b] = y / np.abs(y).max()
    return out.astype(np.float32)[:, None, :]                 # (B, 1, S)


def synthetic_targets(batch, num_events, num_frames, seed=1234, seq_len=None):
    rng = np.random.RandomState(seed + 1)
    weak = (rng.rand(batch, num_events) < 0.2).astype(np.float32)
    for b in range(batch):
        if weak[b].sum() == 0:
            weak[b, rng.randint(num_events)] = 1.
    boundary = np.zeros((batch, num_events, num_frames), np.float32)
    for b in range(batch):
        n = num_frames if seq_len is None else int(seq_len[b])
        for k in np.nonzero(weak[b])[0]:
            on = rng.randint(0, max(n - 1, 1))
            off = rng.randint(on + 1, n + 1)
            boundary[b, k, on:off] = 1.
    return weak, boundary


def synthetic_batch(batch, num_events=10, num_samples=160000, seed=1234, stft_kwargs=None,
                    seq_len=None):
    """example dict as Collate() would hand it to the model (transform.py:65-72,115,124)."""
    kw = dict(shift=320, window_length=960, size=1024, fading='half', pad=True)
    kw.update(stft_kwargs or {})
    audio = synthetic_audio(batch, num_samples, seed)
    spec = P.stft(audio, **kw)
    stft = np.stack([spec.real, spec.imag], -1).astype(np.float32)      # (B,1,T,F,2)
    T = stft.shape[2]
    seq_len = [T] * batch if seq_len is None else list(seq_len)
    weak, boundary = synthetic_targets(batch, num_events, T, seed, seq_len)
    return dict(audio_data=torch.from_numpy(audio), stft=torch.from_numpy(stft), seq_len=seq_len,
                weak_targets=torch.from_numpy(weak), boundary_targets=torch.from_numpy(boundary))


# ---------------------------------------------------------------- train step
def make_adam(model, lr=5e-4):
    """padertorch Adam wrapper == torch.optim.Adam defaults (training.py:264-269)."""
    return torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=lr,
                            betas=(0.9, 0.999), eps=1e-8, weight_decay=0.)


def train_step(model, optimizer, batch, gradient_clipping=1e10):
    """one trainer iteration; returns (loss, grad_norm, outputs)."""
    model.train()
    outputs = model(dict(batch))
    loss = model.review(batch, outputs)['loss']
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    grad_norm = torch.nn.utils.clip_grad_norm_(
        [p for p in model.parameters() if p.requires_grad], gradient_clipping)
    optimizer.step()
    return loss.detach(), grad_norm.detach(), outputs
