"""FBCRNN -- mirror of ``pb_sed.models.weak_label.CRNN`` (pb_sed/models/weak_label/crnn.py).

Same constructor, ``forward`` tuple, ``review`` dict and inference heads; the arithmetic runs in
the sm_100a kernels (feature extraction K1, tap-GEMM conv stack K2, persistent GRU K3, fused
sigmoid / loss K4).  ``forward`` additionally accepts the raw waveform under ``'audio_data'``
(when no ``'stft'`` is given) so that the STFT runs on the GPU too.
"""
import numpy as np
import torch

from .. import ops
from ..modules import CNN, GRU, NormalizedLogMelExtractor, TakeLast, compute_mask, gru_stack
from ..ops import SeqLen, to_native, from_native
from . import base


class CRNN(base.SoundEventModel):
    def __init__(self, feature_extractor, cnn, rnn_fwd, rnn_bwd, *, minimum_score=1e-5,
                 label_smoothing=0., labelwise_metrics=(), label_mapping=None, test_labels=None,
                 slat=False, strong_fwd_bwd_loss_weight=1., class_weights=None):
        super().__init__(labelwise_metrics=labelwise_metrics, label_mapping=label_mapping,
                         test_labels=test_labels)
        self.feature_extractor, self.cnn = feature_extractor, cnn
        self.rnn_fwd, self.rnn_bwd = rnn_fwd, rnn_bwd
        self.minimum_score = minimum_score
        self.label_smoothing = label_smoothing
        self.slat = slat
        self.strong_fwd_bwd_loss_weight = strong_fwd_bwd_loss_weight
        self.class_weights = None if class_weights is None else torch.Tensor(class_weights)
        self.emit_buffers = True     # review(): D2H copies of y_weak / targets_weak (crnn.py:155-162)

    # ---- scores
    def sigmoid(self, y):
        """bounded sigmoid on reference-layout logits (B,K,T) (crnn.py:58-59)."""
        return ops.SigmoidScoresFn.apply(to_native(y), self.minimum_score)

    def _scores_native(self, rnn, h, seq):
        z = rnn.forward_native(h, seq)                        # logits (B,T,K)
        return ops.SigmoidScoresFn.apply(z, self.minimum_score), z

    def fwd_tagging(self, h, seq_len):
        hn = to_native(h)
        seq = SeqLen.make(seq_len, hn.shape[0], hn.shape[1], hn.device)
        return self._scores_native(self.rnn_fwd, hn, seq)[0], seq_len

    def bwd_tagging(self, h, seq_len):
        hn = to_native(h)
        seq = SeqLen.make(seq_len, hn.shape[0], hn.shape[1], hn.device)
        return self._scores_native(self.rnn_bwd, hn, seq)[0], seq_len

    # ---- forward / review
    def _features(self, inputs, pop=False):
        key = 'stft' if 'stft' in inputs else 'audio_data'
        return inputs.pop(key) if pop else inputs[key]

    def encode(self, inputs, pop=False, with_targets=True):
        """features + CNN on native maps -> (h (B,T,D), seq, x (B,1,F,T), seq_len, targets)."""
        x = self._features(inputs, pop)
        seq_len = inputs.get('seq_len')
        seq_len = None if seq_len is None else np.array(seq_len)
        targets = None
        if with_targets and 'weak_targets' in inputs:
            targets = self.read_targets(inputs)
            x, seq_len_x, targets = self.feature_extractor(x, seq_len=seq_len, targets=targets)
        else:
            x, seq_len_x = self.feature_extractor(x, seq_len=seq_len)
        B, _, F, T = x.shape
        seq = SeqLen.make(seq_len_x, B, T, x.device)
        h = self.cnn.forward_native(x.reshape(B, F, T, 1), seq)
        return h, seq, x, seq_len_x, targets

    def forward(self, inputs):
        h, seq, x, seq_len_x, targets = self.encode(inputs, pop=self.training)
        y_bwd = None
        if self._can_pair():
            # both recurrences advance in the same persistent-kernel launches
            hf, hb = gru_stack([self.rnn_fwd.rnn, self.rnn_bwd.rnn], [h, h], seq,
                               [self.rnn_fwd.reverse, self.rnn_bwd.reverse])
            z = self.rnn_fwd.output_net.forward_native(hf.unsqueeze(1), seq).squeeze(1)
            y_fwd, self._z_fwd = ops.SigmoidScoresFn.apply(z, self.minimum_score), z.detach()
            z = self.rnn_bwd.output_net.forward_native(hb.unsqueeze(1), seq).squeeze(1)
            y_bwd, self._z_bwd = ops.SigmoidScoresFn.apply(z, self.minimum_score), z.detach()
        else:
            y_fwd, z = self._scores_native(self.rnn_fwd, h, seq)
            self._z_fwd = z.detach()          # frame logits (B,T,K), kept for the parity metric
            if self.rnn_bwd is not None:
                y_bwd, z = self._scores_native(self.rnn_bwd, h, seq)
                self._z_bwd = z.detach()
        if seq_len_x is None:
            seq_len_x = np.full(x.shape[0], x.shape[-1])
        return y_fwd, y_bwd, seq_len_x, x, seq_len_x, targets

    def _can_pair(self):
        a, b = self.rnn_fwd, self.rnn_bwd
        return (isinstance(a, GRU) and isinstance(b, GRU) and a.rnn is not None and b.rnn is not None
                and not a.rnn.bidirectional and not b.rnn.bidirectional
                and (a.rnn.hidden_size, a.rnn.num_layers, a.rnn.input_size) ==
                    (b.rnn.hidden_size, b.rnn.num_layers, b.rnn.input_size))

    def read_targets(self, inputs, subsample_idx=None):
        if 'boundary_targets' in inputs:
            return inputs['weak_targets'], inputs['boundary_targets']
        return inputs['weak_targets'],

    def loss(self, y_fwd, y_bwd, seq_len, targets):
        """crnn.py:117-153 in one fused kernel (value + gradient)."""
        B, K, T = y_fwd.shape
        seq = SeqLen.make(seq_len, B, T, y_fwd.device)
        weak = targets[0]
        boundary = None
        if self.strong_fwd_bwd_loss_weight > 0.:
            if self.slat:
                wm = ((weak < .01) | (weak > .99)).to(weak.dtype)
                boundary = (weak * wm)[..., None].expand(B, K, T).contiguous()
            else:
                assert len(targets) == 2, len(targets)
                boundary = targets[1]
        cw = None if self.class_weights is None else self.class_weights.to(y_fwd.device)
        return ops.FbcrnnLossFn.apply(y_fwd, y_bwd, weak, boundary, cw, seq,
                                      self.strong_fwd_bwd_loss_weight, self.label_smoothing)

    def review(self, inputs, outputs):
        y_fwd, y_bwd, seq_len, x, _, targets = outputs
        assert targets is not None
        loss = self.loss(y_fwd, y_bwd, seq_len, targets)
        review = dict(loss=loss, scalars=dict(seq_len=np.mean(inputs['seq_len'])),
                      images=dict(features=x[:3]), buffers=dict())
        if self.emit_buffers:
            weak = targets[0]
            wm = (weak < .01) | (weak > .99)
            review['scalars']['weak_label_rate'] = wm.float().mean().item()
            if self.strong_fwd_bwd_loss_weight > 0.:
                wt = weak * wm
                bt = wt[..., None].expand(y_fwd.shape) if self.slat else targets[1]      # crnn.py:130-134
                bm = ((bt > .99) | (bt < .01))
                bm = bm * (bm.float().mean(-1, keepdim=True) > .999) * (wt > .99)[..., None]
                review['scalars']['boundary_label_rate'] = bm.float().mean().item()
            else:
                review['scalars']['boundary_label_rate'] = 0.
            labeled = wm.all(-1).cpu().numpy()
            y_weak = TakeLast(axis=2)(y_fwd.detach(), seq_len=seq_len)
            if y_bwd is not None:
                y_weak = y_weak / 2 + y_bwd.detach()[..., 0] / 2
            review['buffers'] = dict(y_weak=y_weak.cpu().numpy()[labeled],
                                     targets_weak=(weak * wm).cpu().numpy()[labeled])
        return review

    # ---- inference heads (crnn.py:223-302)
    def tagging(self, inputs):
        y_fwd, y_bwd, seq_len_y, *_ = self.forward(inputs)
        last = TakeLast(axis=-1, keepdims=True)(y_fwd, seq_len_y)
        ones = np.ones_like(seq_len_y)
        if y_bwd is None:
            return last, ones
        return (last + y_bwd[..., :1]) / 2, ones

    def boundaries_detection(self, inputs):
        y_fwd, y_bwd, seq_len_y, *_ = self.forward(inputs)
        m = compute_mask(y_fwd, seq_len_y, batch_axis=0, sequence_axis=-1)
        return torch.minimum(y_fwd * m, y_bwd * m), seq_len_y

    def sound_event_detection(self, inputs, window_length, window_shift=1):
        """sliding-window SED: both GRUs over every frame's +-window context (crnn.py:241-302)."""
        window_length = np.array(window_length, dtype=int)
        h, seq, x, seq_len, _ = self.encode(inputs, with_targets=False)
        if seq_len is None:
            seq_len = np.full(h.shape[0], h.shape[1])
        if window_length.ndim == 0:
            return self._single_window_length_sed(h, seq_len, int(window_length), window_shift)
        y = None
        for win_len in np.unique(window_length.flatten()):
            yi, seq_len_y = self._single_window_length_sed(h, seq_len, int(win_len), window_shift)
            b, k, t = yi.shape
            wl = window_length
            if wl.ndim == 1:
                assert wl.shape[0] in [1, k], wl.shape
            elif wl.ndim == 2:
                assert wl.shape[1] in [1, k], wl.shape
                wl = np.broadcast_to(wl, (wl.shape[0], k))
                yi = yi[:, None]
            else:
                raise ValueError('window_length.ndim must not be greater than 2.')
            if y is None:
                y = torch.zeros((b, *wl.shape, t), device=yi.device)
            y += (torch.from_numpy(wl.copy()).to(yi.device) == win_len)[..., None] * yi
        return y, seq_len_y

    def _single_window_length_sed(self, h, seq_len, window_length, window_shift):
        """h native (B,T,D)."""
        B, T, D = h.shape
        front = end = 0
        if window_length > window_shift:
            p = window_length - window_shift
            front, end = p // 2, int(np.ceil(p / 2))
        end += window_shift - 1
        hp = torch.nn.functional.pad(h, (0, 0, front, end))
        starts = np.arange(0, T, window_shift)
        win = torch.cat([hp[:, i:i + window_length] for i in starts], dim=0).contiguous()   # (n*B, W, D)
        n = len(starts)
        seq = SeqLen.make(None, n * B, window_length, h.device)
        y = self._scores_native(self.rnn_fwd, win, seq)[0][..., -1]                       # (n*B, K)
        y = y.reshape(n, B, -1).permute(1, 2, 0)
        if self.rnn_bwd is not None:
            yb = self._scores_native(self.rnn_bwd, win, seq)[0][..., 0]
            y = (y + yb.reshape(n, B, -1).permute(1, 2, 0)) / 2
        return y, 1 + (np.asarray(seq_len) - 1) // window_shift

    # ---- construction from the reference's config layout
    @classmethod
    def from_config_dict(cls, config):
        """build from the ``trainer.model`` dict of pb_sed/experiments/weak_label_crnn/training.py:
        188-262 (factory keys ignored; sizes wired like finalize_dogmatic_config, crnn.py:304-340)."""
        cfg = {k: v for k, v in config.items() if k != 'factory'}
        fe_kw = {k: v for k, v in cfg.pop('feature_extractor').items() if k != 'factory'}
        fe = NormalizedLogMelExtractor(**fe_kw)
        cnn_kw = {k: v for k, v in cfg.pop('cnn').items() if k != 'factory'}
        cnn_kw['cnn_2d'] = dict(cnn_kw['cnn_2d'], in_channels=1)
        cnn_kw.setdefault('input_height', fe.number_of_filters)
        cnn = CNN(**cnn_kw)
        rnn_cfg = {k: v for k, v in cfg.pop('rnn_fwd').items() if k != 'factory'}
        d = cnn.cnn_1d.out_channels[-1]

        def make(reverse):
            rc = dict(rnn_cfg)
            if rc.get('rnn') is not None:
                rc['rnn'] = dict(rc['rnn'], input_size=d)
                rc['rnn'].setdefault('num_layers', 1)
            else:
                rc['output_net'] = dict(rc['output_net'], in_channels=d)
            return GRU(**rc, reverse=reverse)
        rnn_bwd_cfg = cfg.pop('rnn_bwd', {})
        rnn_bwd = None if rnn_bwd_cfg is None else make(True)
        return cls(fe, cnn, make(False), rnn_bwd, **cfg)
