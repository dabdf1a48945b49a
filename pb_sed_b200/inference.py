"""Inference drivers -- mirror of ``pb_sed/models/base/inference.py`` (``tagging`` :13-37,
``boundaries_detection`` :40-63, ``sound_event_detection`` :66-89, ``inference`` :92-222) with the score
path kept on the GPU: the models' scores are averaged, sequence-masked, median / boundary filtered and
tag-masked in HBM (``pb_sed_b200.filters``) and cross PCIe ONCE per segment, already post-processed.
The reference copies every model's raw scores to the host (:137) and filters them row by row in numpy.

Same arguments and return value as the reference (a dict ``audio_id -> (T, K)`` / ``(N, T, K)`` numpy
array).  Segmenting long clips (``max_segment_length``) and merging the segments' scores follow
``pb_sed/utils/segment.py``.  Writing ``sed_scores_eval`` data frames (``timestamps`` /
``event_classes`` / ``score_storage_dir``) is host-side bookkeeping outside the hot path and is left
to the caller (``NotImplementedError`` if requested here).
"""
from math import ceil

import re

import numpy as np
import torch

from . import filters as F


# ------------------------------------------------------------------ pb_sed/utils/segment.py
def segment_batch(batch, max_length, overlap, keys=('stft',), axis=2):
    """pb_sed/utils/segment.py:6-47: split a batch whose longest clip exceeds ``max_length`` frames into
    segments of ``max_length`` with hop ``max_length - overlap`` (zero padded at the end), tagging the
    example ids with ``_!segment!_<i>_<m>``.  (The reference delegates the slicing to padertorch's
    ``Segmenter(length, shift, mode='constant', padding=True)``; restated here.)"""
    seq_lens = list(batch['seq_len'])
    if max(seq_lens) <= max_length:
        return [batch]
    shift = max_length - overlap
    total = max(seq_lens)
    m = max(int(ceil((total - max_length) / shift)) + 1, 1)
    segments = []
    for i in range(m):
        start = i * shift
        seg = {k: v for k, v in batch.items() if k not in keys}
        seg['segment_start'] = start
        seg['example_id'] = [f'{eid}_!segment!_{i}_{m}' for eid in batch['example_id']]
        seg['seq_len'] = [min(max_length, sl - start) for sl in seq_lens]
        for k in keys:
            x = batch[k]
            x = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))
            part = x.narrow(axis, start, min(max_length, x.shape[axis] - start))
            need = max(seg['seq_len']) - part.shape[axis]
            if need > 0:
                pad = [0, 0] * (x.dim() - axis - 1) + [0, need]
                part = torch.nn.functional.pad(part, pad)
            seg[k] = part.narrow(axis, 0, max(seg['seq_len']))
        segments.append(seg)
    return segments


_SEGMENT_ID = re.compile(r'^(?P<base>.*)_!segment!_(?P<index>\d+)_(?P<count>\d+)$')


def merge_segments(segmental_output, segment_overlap):
    """stitch per-segment score arrays (frames on axis -2) back into one array per clip; same contract as
    pb_sed/utils/segment.py:50-72: ids ``<clip>_!segment!_<i>_<n>``, neighbouring segments share
    ``segment_overlap`` frames, of which the left segment keeps the first half (rounded down) and the right one the rest;
    ids without the segment suffix pass through."""
    groups, merged = {}, {}
    for key, value in segmental_output.items():
        m = _SEGMENT_ID.match(key)
        if m is None:
            merged[key] = value
        else:
            groups.setdefault(m['base'], {})[int(m['index'])] = (int(m['count']), value)
    drop_right, drop_left = ((segment_overlap + 1) // 2, segment_overlap // 2) if segment_overlap > 0 else (0, 0)
    for base in sorted(groups):
        parts = groups[base]
        count = parts[0][0]
        pieces = []
        for i in range(count):
            arr = parts[i][1]
            lo = drop_left if i > 0 else 0
            hi = arr.shape[-2] - (drop_right if i < count - 1 else 0)
            pieces.append(arr[..., lo:hi, :])
        merged[base] = np.concatenate(pieces, axis=-2)
    return merged


# ------------------------------------------------------------------ drivers
def tagging(models, dataset, device, max_segment_length=None, segment_overlap=None,
            merge_score_segments=False, score_segment_overlap=None, model_kwargs=None, medfilt_length=1,
            method='tagging', timestamps=None, event_classes=None, score_storage_dir=None):
    return inference(models, method, dataset, device, max_segment_length=max_segment_length,
                     segment_overlap=segment_overlap, merge_score_segments=merge_score_segments,
                     score_segment_overlap=score_segment_overlap, model_kwargs=model_kwargs,
                     medfilt_length=medfilt_length, post_processing_fn=lambda x: x.max(-2, keepdims=True),
                     timestamps=timestamps, event_classes=event_classes, score_storage_dir=score_storage_dir)


def boundaries_detection(models, dataset, device, max_segment_length=None, segment_overlap=None,
                         merge_score_segments=False, score_segment_overlap=None, model_kwargs=None,
                         medfilt_length=1, stepfilt_length=0, apply_mask=False, masks=None,
                         method='boundaries_detection', timestamps=None, event_classes=None,
                         score_storage_dir=None):
    return inference(models, method, dataset, device, max_segment_length=max_segment_length,
                     segment_overlap=segment_overlap, merge_score_segments=merge_score_segments,
                     score_segment_overlap=score_segment_overlap, model_kwargs=model_kwargs,
                     medfilt_length=medfilt_length, stepfilt_length=stepfilt_length, apply_mask=apply_mask,
                     masks=masks, timestamps=timestamps, event_classes=event_classes,
                     score_storage_dir=score_storage_dir)


def sound_event_detection(models, dataset, device, max_segment_length=None, segment_overlap=None,
                          merge_score_segments=False, score_segment_overlap=None, model_kwargs=None,
                          medfilt_length=1, method='sound_event_detection', apply_mask=False, masks=None,
                          timestamps=None, event_classes=None, score_storage_dir=None):
    return inference(models, method, dataset, device, max_segment_length=max_segment_length,
                     segment_overlap=segment_overlap, merge_score_segments=merge_score_segments,
                     score_segment_overlap=score_segment_overlap, model_kwargs=model_kwargs,
                     medfilt_length=medfilt_length, apply_mask=apply_mask, masks=masks,
                     timestamps=timestamps, event_classes=event_classes, score_storage_dir=score_storage_dir)


def inference(model, method, dataset, device, max_segment_length=None, segment_overlap=0,
              merge_score_segments=False, score_segment_overlap=None, model_kwargs=None, medfilt_length=1,
              stepfilt_length=None, apply_mask=False, masks=None, post_processing_fn=None,
              timestamps=None, event_classes=None, score_storage_dir=None):
    if timestamps is not None or event_classes is not None or score_storage_dir is not None:
        raise NotImplementedError('sed_scores_eval data frames are written by the caller (host-side bookkeeping)')
    if not isinstance(model, (list, tuple)):
        model = [model]
    if model_kwargs is None:
        model_kwargs = {}
    if not isinstance(model_kwargs, (list, tuple)):
        model_kwargs = len(model) * [model_kwargs]
    else:
        assert len(model_kwargs) == len(model), (len(model), len(model_kwargs))
    medfilt_length = np.array(medfilt_length, dtype=int)
    apply_mask = np.array(apply_mask, dtype=bool)
    for m in model:
        assert hasattr(m, method), (m, method)
        m.to(device)
        m.eval()
    if post_processing_fn is None:
        def post_processing_fn(x):
            return x
    seg_key = 'stft'
    scores = {}
    with torch.no_grad():
        for batch in dataset:
            score_cache = {}
            batch = {k: v for k, v in batch.items()
                     if k not in ('weak_targets', 'boundary_targets', 'strong_targets')}
            if 'stft' not in batch:
                seg_key = 'audio_data'
            if max_segment_length is not None:
                assert seg_key == 'stft', 'segmenting is defined on the frame axis of the stft input'
                input_segments = segment_batch(batch, max_length=max_segment_length, overlap=segment_overlap)
            else:
                input_segments = [batch]
            for segment in input_segments:
                segment = model[0].example_to_device(segment, device)
                seq_len = None
                y = None
                for i in range(len(model)):
                    yi, seq_len_i = getattr(model[i], method)(segment, **model_kwargs[i])
                    y = yi.float() if y is None else y + yi.float()
                    if i == 0:
                        seq_len = np.asarray(seq_len_i)
                    else:
                        assert (np.asarray(seq_len_i) == seq_len).all(), (seq_len, seq_len_i)
                if len(model) > 1:
                    y = y / len(model)
                tags = None
                if apply_mask.any():
                    assert masks is not None
                    for audio_id in segment['example_id']:
                        assert audio_id in masks, audio_id
                    tags = torch.as_tensor(np.stack([np.asarray(masks[a], dtype=np.float32).reshape(-1)
                                                     for a in segment['example_id']]))
                y = F.post_process(y.contiguous(), seq_len, medfilt_length=medfilt_length,
                                   stepfilt_length=stepfilt_length, apply_mask=apply_mask, tags=tags)
                y = y.cpu().numpy()                       # the ONE device-to-host copy of the segment
                score_cache.update({
                    audio_id: post_processing_fn(y[i, ..., :sl].swapaxes(-2, -1))
                    for i, (audio_id, sl) in enumerate(zip(segment['example_id'], seq_len))})
            if merge_score_segments and '_!segment!_' in input_segments[-1]['example_id'][0]:
                score_cache = merge_segments(
                    score_cache, segment_overlap=segment_overlap if score_segment_overlap is None
                    else score_segment_overlap)
            scores.update(score_cache)
    return scores
