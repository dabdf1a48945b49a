"""(tag-conditioned) BiCRNN -- mirror of ``pb_sed.models.strong_label.CRNN``
(pb_sed/models/strong_label/crnn.py)."""
import numpy as np
import torch

from .. import ops
from ..modules import CNN, GRU, NormalizedLogMelExtractor, Mean, compute_mask
from ..ops import SeqLen
from . import base


class CRNN(base.SoundEventModel):
    def __init__(self, feature_extractor, cnn, rnn, *, tag_conditioning=False,
                 labelwise_metrics=(), label_mapping=None, eval_segment_length=1):
        super().__init__(labelwise_metrics=labelwise_metrics, label_mapping=label_mapping)
        self.feature_extractor, self.cnn, self.rnn = feature_extractor, cnn, rnn
        self.tag_conditioning = tag_conditioning
        self.eval_segment_length = eval_segment_length
        self.emit_buffers = True

    def forward(self, inputs):
        """crnn.py:60-93: features -> CNN(+tag channels) -> cat(h, tags) -> BiGRU -> sigmoid."""
        key = 'stft' if 'stft' in inputs else 'audio_data'
        x = inputs.pop(key) if self.training else inputs[key]
        seq_len = np.array(inputs['seq_len'])
        targets = None
        if 'strong_targets' in inputs:
            x, seq_len_x, targets = self.feature_extractor(
                x, seq_len=seq_len, targets=(inputs['weak_targets'], inputs['strong_targets']))
        else:
            x, seq_len_x = self.feature_extractor(x, seq_len=seq_len)
        B, _, F, T = x.shape
        seq = SeqLen.make(seq_len_x, B, T, x.device)
        tags = inputs['tag_condition'].float() if self.tag_conditioning else None
        h = self.cnn.forward_native(x.reshape(B, F, T, 1), seq,
                                    tags if self.cnn.conditional_dims else None)      # (B,T,D)
        if self.tag_conditioning:
            h = ops.ConcatCondFn.apply(h, tags)
        z = self.rnn.forward_native(h, seq)
        self._z = z.detach()
        return ops.SigmoidScoresFn.apply(z, 0.), seq_len_x, x, seq_len_x, targets

    def loss(self, y, seq_len_y, targets):
        B, K, T = y.shape
        assert targets[1].shape == y.shape, (targets[1].shape, y.shape)
        return ops.BicrnnLossFn.apply(y, targets[1], SeqLen.make(seq_len_y, B, T, y.device))

    def review(self, inputs, outputs):
        y, seq_len_y, x, _, targets = outputs
        assert targets is not None
        st = targets[1]
        review = dict(loss=self.loss(y, seq_len_y, targets),
                      scalars=dict(seq_len=np.mean(inputs['seq_len'])),
                      images=dict(features=x[:3], strong_targets=st[:3]), buffers=dict())
        if self.emit_buffers:
            m = ((st > .99) | (st < .01)).float()
            review['scalars']['strong_label_rate'] = m.mean().item()
            full = (Mean(axis=-1)(m, seq_len_y) > .999).all(-1).cpu().numpy()
            yc, tc = y.detach().cpu().numpy(), st.cpu().numpy()
            L = self.eval_segment_length

            def seg(a, n):
                a = a[:, :n].T
                n_seg = (a.shape[0] - L) // L + 1
                return a[:n_seg * L].reshape(n_seg, L, -1).max(1)
            idx = np.nonzero(full)[0]
            if len(idx):
                review['buffers'] = dict(
                    y_strong=np.concatenate([seg(yc[i], seq_len_y[i]) for i in idx]),
                    targets_strong=np.concatenate([seg(tc[i], seq_len_y[i]) for i in idx]))
        return review

    def tagging(self, inputs):
        y, seq_len_y, *_ = self.forward(inputs)
        return y.max(-1, keepdim=True)[0], np.ones_like(seq_len_y)

    def boundaries_detection(self, inputs):
        return self.sound_event_detection(inputs)

    def sound_event_detection(self, inputs):
        y, seq_len_y, *_ = self.forward(inputs)
        return y * compute_mask(y, seq_len_y, batch_axis=0, sequence_axis=-1), seq_len_y

    @classmethod
    def from_config_dict(cls, config):
        """sizes wired like finalize_dogmatic_config (crnn.py:155-198)."""
        cfg = {k: v for k, v in config.items() if k != 'factory'}
        tagc = cfg.get('tag_conditioning', False)
        fe = NormalizedLogMelExtractor(**{k: v for k, v in cfg.pop('feature_extractor').items() if k != 'factory'})
        rnn_cfg = {k: v for k, v in cfg.pop('rnn').items() if k != 'factory'}
        k_events = rnn_cfg['output_net']['out_channels'][-1]
        cnn_kw = {k: v for k, v in cfg.pop('cnn').items() if k != 'factory'}
        cond = k_events if tagc else 0
        cnn_kw['cnn_2d'] = dict(cnn_kw['cnn_2d'], in_channels=1 + cond)
        cnn_kw['conditional_dims'] = cond
        cnn_kw.setdefault('input_height', fe.number_of_filters)
        cnn = CNN(**cnn_kw)
        rnn_kw = dict(num_layers=1, bias=True, dropout=0., bidirectional=True)
        rnn_kw.update(rnn_cfg.get('rnn') or {})
        rnn_kw['input_size'] = cnn.cnn_1d.out_channels[-1] + cond
        rnn_cfg['rnn'] = rnn_kw
        return cls(fe, cnn, GRU(**rnn_cfg), **cfg)
