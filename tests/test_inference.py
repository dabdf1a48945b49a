"""Inference drivers (mirror of pb_sed/models/base/inference.py + pb_sed/utils/segment.py): host logic on
CPU, the full GPU path (-m gpu) against golden vectors produced by the REAL reference drivers running the
REAL pb_sed FBCRNN class (tests/golden/make_golden_inference.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_loader
from util import GOLDEN, padertorch_key


def test_segment_batch_and_merge_round_trip():
    """segment -> identity scores -> merge reproduces the clip (size-independent property), and the
    example-id / seq_len bookkeeping follows pb_sed/utils/segment.py:34-44."""
    from pb_sed_b200.inference import segment_batch, merge_segments
    T, Fb = 50, 7
    stft = torch.arange(3 * T, dtype=torch.float32).reshape(3, 1, T, 1, 1).expand(3, 1, T, Fb, 2).contiguous()
    batch = {'example_id': ['a', 'b', 'c'], 'stft': stft, 'seq_len': [50, 47, 46]}
    segs = segment_batch(batch, 12, 2)
    assert len(segs) == 5 and segs[0]['example_id'][1] == 'b_!segment!_0_5'
    assert segs[0]['stft'].shape == (3, 1, 12, Fb, 2) and segs[0]['seq_len'] == [12, 12, 12]
    assert segs[4]['seq_len'] == [10, 7, 6] and segs[4]['stft'].shape[2] == 10
    assert segment_batch(batch, 50, 2) == [batch]                      # nothing to split
    cache = {}
    for s in segs:
        for i, (eid, sl) in enumerate(zip(s['example_id'], s['seq_len'])):
            cache[eid] = s['stft'][i, 0, :sl, :1, 0].numpy().repeat(4, -1)      # (t, k) "scores"
    merged = merge_segments(cache, segment_overlap=2)
    for i, (eid, sl) in enumerate(zip(batch['example_id'], batch['seq_len'])):
        assert np.array_equal(merged[eid][:, 0], stft[i, 0, :sl, 0, 0].numpy()), eid


@pytest.mark.skipif(not ref_loader.reference_available(), reason='/root/reference not mounted')
def test_merge_segments_matches_live_reference():
    from pb_sed_b200.inference import merge_segments
    ref = ref_loader.load_segment().merge_segments
    rng = np.random.RandomState(0)
    for overlap in (0, 2, 3):
        cache = {f'x_!segment!_{i}_3': rng.rand(2, 10, 4) for i in range(3)}
        cache['y'] = rng.rand(2, 7, 4)
        a, b = merge_segments(dict(cache), overlap), ref(dict(cache), overlap)
        assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)


@pytest.mark.gpu
@pytest.mark.parametrize('from_audio', [False, True])
def test_gpu_inference_drivers_match_reference_golden(built_lib, from_audio):
    from pb_sed_b200 import config, inference as I
    from pb_sed_b200.models import weak_label
    d = dict(np.load(os.path.join(GOLDEN, 'inference_tiny.npz')))
    models = []
    for i in range(2):
        m = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config())
        m.load_state_dict({padertorch_key(k[len(f'state{i}.'):]): torch.from_numpy(v) for k, v in d.items()
                           if k.startswith(f'state{i}.') and not k.endswith('fbanks')})
        models.append(m)
    ids = ['a', 'b', 'c', 'd']
    batch = {'example_id': ids, 'seq_len': [int(s) for s in d['seq_len']]}
    batch['audio_data' if from_audio else 'stft'] = torch.from_numpy(d['audio'] if from_audio else d['stft'])
    tol = 2e-3 if from_audio else 2e-4            # scores in (0,1); the audio path adds the fp32-FFT difference

    def check(got, prefix):
        assert sorted(got) == ids
        for k in ids:
            ref = d[f'{prefix}.{k}']
            assert got[k].shape == ref.shape, (prefix, k, got[k].shape, ref.shape)
            assert np.abs(got[k] - ref).max() < tol, (prefix, k)
    check(I.tagging(models, [dict(batch)], 'cuda:0', medfilt_length=1), 'tagging')
    tags = {k: d[f'tags.{k}'] for k in ids}
    check(I.boundaries_detection(models, [dict(batch)], 'cuda:0', stepfilt_length=d['stepfilt_length'],
                                 apply_mask=True, masks=tags), 'boundaries')
    check(I.sound_event_detection(models, [dict(batch)], 'cuda:0',
                                  model_kwargs={'window_length': d['window_length'], 'window_shift': 2},
                                  medfilt_length=d['medfilt_length'], apply_mask=d['apply_mask'], masks=tags), 'sed')
