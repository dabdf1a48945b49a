"""Generate tests/golden/inference_tiny.npz by running the REAL pb_sed inference drivers
(pb_sed/models/base/inference.py: tagging / boundaries_detection / sound_event_detection -> inference,
filtering, boundariesfilt, tag masking) on the REAL pb_sed FBCRNN class, both executed unmodified from
/root/reference through oracle/ref_loader.py on top of the oracle's module restatements.

    python tests/golden/make_golden_inference.py       (build container only)
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


def main():
    weak_mod, _ = ref_loader.load()
    _, inf = ref_loader.load_filters()
    models, states = [], []
    for seed in (0, 1):                                    # a 2-model ensemble (inference.py:134-141)
        m = OM.tiny_fbcrnn(seed=seed)
        states.append({k: v.clone() for k, v in m.state_dict().items()})
        models.append(weak_mod.CRNN(m.feature_extractor, m.cnn, m.rnn_fwd, m.rnn_bwd))
    seq_len = [41, 40, 33, 17]
    batch = OM.synthetic_batch(4, num_samples=16 * 40 + 5, stft_kwargs=TINY_STFT, seq_len=seq_len, seed=4)
    batch['example_id'] = ['a', 'b', 'c', 'd']
    K = 10

    def dataset():
        return [dict(batch)]
    out = dict(stft=batch['stft'].numpy(), audio=batch['audio_data'].numpy(), seq_len=np.array(seq_len))
    for i, st in enumerate(states):
        out.update({f'state{i}.' + k: v.numpy() for k, v in st.items()})
    tag = inf.tagging(models, dataset(), 'cpu', medfilt_length=1)
    out.update({'tagging.' + k: v for k, v in tag.items()})
    tags = {k: (v[0] > .5) for k, v in tag.items()}
    out.update({'tags.' + k: v for k, v in tags.items()})
    step = np.array([0, 2, 4, 6, 8, 0, 2, 4, 6, 8])
    out['stepfilt_length'] = step
    bnd = inf.boundaries_detection(models, dataset(), 'cpu', stepfilt_length=step, apply_mask=True, masks=tags)
    out.update({'boundaries.' + k: v for k, v in bnd.items()})
    win = np.array([[3] * K, [5] * K, [3, 5] * (K // 2)])
    med = np.array([[1] * K, [3] * K, [5, 1] * (K // 2)])
    app = np.array([[0] * K, [1] * K, [1, 0] * (K // 2)])
    out.update(window_length=win, medfilt_length=med, apply_mask=app)
    sed = inf.sound_event_detection(models, dataset(), 'cpu', model_kwargs={'window_length': win, 'window_shift': 2},
                                    medfilt_length=med, apply_mask=app, masks=tags)
    out.update({'sed.' + k: v for k, v in sed.items()})
    np.savez_compressed(os.path.join(HERE, 'inference_tiny.npz'), **out)
    for k in ('tagging.a', 'boundaries.a', 'sed.a', 'sed.d'):
        print(k, out[k].shape, out[k].dtype)


if __name__ == '__main__':
    main()
