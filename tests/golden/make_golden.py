"""Generate tests/golden/*.npz by running the REAL pb_sed model classes (imported from
/root/reference through oracle/ref_loader.py) on top of the oracle's module restatements.

    python tests/golden/make_golden.py          (build container only: needs /root/reference)

What the vectors pin: the pb_sed-owned arithmetic (CRNN.forward order, bounded sigmoid, weak /
strong forward-backward BCE, strong-label BCE, tagging / boundary / sliding-window heads) as
executed by pb_sed's own source.  The module arithmetic underneath is the oracle's restatement
of padertorch/paderbox (parity unpinned there, see oracle/__init__.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models as OM, ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TINY_STFT = dict(shift=16, window_length=48, size=64)


def npz(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def fbcrnn_case(name, seq_len, strong_weight, fractional_targets, seed):
    weak_mod, _ = ref_loader.load()
    m = OM.tiny_fbcrnn(seed=seed, strong_fwd_bwd_loss_weight=strong_weight)
    ref = weak_mod.CRNN(m.feature_extractor, m.cnn, m.rnn_fwd, m.rnn_bwd,
                        strong_fwd_bwd_loss_weight=strong_weight)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    batch = OM.synthetic_batch(4, num_samples=16 * 40 + 5, stft_kwargs=TINY_STFT, seq_len=seq_len, seed=seed)
    if fractional_targets:          # 'unknown' labels encoded as .5 (transform.py:62-63,111-123)
        batch['weak_targets'][1, 3] = .5
        batch['boundary_targets'][2, :, 5:9] = .5
    ref.train()
    out = ref(dict(batch))
    review = ref.review(batch, out)
    review['loss'].backward()
    grads = {'grad.' + k: p.grad.clone() for k, p in ref.named_parameters()}
    ref.eval()
    with torch.no_grad():
        tag, _ = ref.tagging(batch)
        bnd, _ = ref.boundaries_detection(batch)
        sed, sed_len = ref.sound_event_detection(batch, window_length=5, window_shift=2)
    d = dict(audio=batch['audio_data'], stft=batch['stft'], seq_len=np.array(batch['seq_len']),
             weak_targets=batch['weak_targets'], boundary_targets=batch['boundary_targets'],
             y_fwd=out[0], y_bwd=out[1], features=out[3], loss=review['loss'],
             y_weak=review['buffers']['y_weak'], tagging=tag, boundaries=bnd, sed=sed, sed_len=sed_len,
             strong_weight=strong_weight, seed=seed)
    d.update({'state.' + k: v for k, v in state.items()})
    d.update(grads)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **npz(d))
    print(name, 'loss', float(review['loss']))


def bicrnn_case(name, seq_len, seed):
    _, strong_mod = ref_loader.load()
    m = OM.build_bicrnn(n_mels=16, stft_size=64, channels_2d=[8, 8, 16], pool_2d=[1, (2, 1), (2, 1)],
                        channels_1d=[32, 32], k1d=[3, 1], hidden=32, rnn_layers=2, out_hidden=16, seed=seed)
    ref = strong_mod.CRNN(m.feature_extractor, m.cnn, m.rnn, tag_conditioning=True)
    state = {k: v.clone() for k, v in m.state_dict().items()}
    batch = OM.synthetic_batch(4, num_samples=16 * 40 + 5, stft_kwargs=TINY_STFT, seq_len=seq_len, seed=seed)
    batch['strong_targets'] = batch.pop('boundary_targets')
    batch['strong_targets'][3, 2, 1:4] = .5
    batch['tag_condition'] = (batch['weak_targets'] > .5)
    ref.train()
    out = ref(dict(batch))
    review = ref.review(batch, out)
    review['loss'].backward()
    grads = {'grad.' + k: p.grad.clone() for k, p in ref.named_parameters()}
    ref.eval()
    with torch.no_grad():
        sed, _ = ref.sound_event_detection(batch)
    d = dict(audio=batch['audio_data'], stft=batch['stft'], seq_len=np.array(batch['seq_len']),
             weak_targets=batch['weak_targets'], strong_targets=batch['strong_targets'],
             tag_condition=batch['tag_condition'], y=out[0], features=out[2], loss=review['loss'],
             sed=sed, seed=seed)
    d.update({'state.' + k: v for k, v in state.items()})
    d.update(grads)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **npz(d))
    print(name, 'loss', float(review['loss']))


if __name__ == '__main__':
    fbcrnn_case('fbcrnn_tiny_full', [41, 41, 41, 41], 1., False, 0)
    fbcrnn_case('fbcrnn_tiny_ragged', [41, 40, 33, 17], 1., True, 1)
    fbcrnn_case('fbcrnn_tiny_weakonly', [41, 37, 30, 22], 0., True, 2)
    bicrnn_case('bicrnn_tiny_ragged', [41, 40, 33, 17], 3)
