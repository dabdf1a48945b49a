"""Batch assembly and the host->device input pipeline.

Mirrors what the reference does between ``Transform`` and the model:
``pb_sed/data_preparation/fetcher.py:36-51`` (examples of a bucket sorted by ``seq_len`` descending,
then padertorch ``Collate()``: arrays zero-padded to the longest example and stacked, everything else
gathered in lists) and ``Model.example_to_device`` (``pb_sed/models/base/inference.py:130``).  The
differences are the ones the B200 path wants: batches land in PINNED host memory, raw audio
(``Transform(pop_audio_data=False)``, transform.py:17,126-127) can replace the 3.2x larger ``stft``
(20.5 MB instead of 65.7 MB per batch of 32), and ``DeviceLoader`` copies batch i+1
on a copy stream while batch i computes.
"""
import numpy as np
import torch

from .modules import stft_num_frames

_PAD_LAST = ('boundary_targets', 'strong_targets')     # (K, T): time is the last axis
_STACK = ('weak_targets', 'tag_condition')


def _pin(t):
    return t.pin_memory() if torch.cuda.is_available() else t


def collate(examples, keep_stft=True, stft_kwargs=None):
    """list of example dicts (transform.py:65-72,115,124) -> batch dict of pinned tensors.

    ``stft`` (1, T, F, 2) and ``audio_data`` (1, S) are zero-padded along time / samples;
    ``boundary_targets`` / ``strong_targets`` (K, T) along T; ``seq_len`` (frames) is taken from the
    example, or derived from the number of samples with the reference STFT geometry when only audio is
    there.  Examples are sorted by ``seq_len`` descending (fetcher.py:49-50: the GRU packs them)."""
    assert len(examples) > 0
    kw = dict(shift=320, window_length=960, fading='half', pad=True)
    kw.update({k: v for k, v in (stft_kwargs or {}).items() if k in kw})
    ex = []
    for e in examples:
        e = dict(e)
        if 'seq_len' not in e:
            e['seq_len'] = (e['stft'].shape[1] if 'stft' in e else
                            stft_num_frames(np.asarray(e['audio_data']).shape[-1], **kw))
        ex.append(e)
    ex.sort(key=lambda e: -int(e['seq_len']))
    batch = {}
    for key in ex[0]:
        vals = [e[key] for e in ex]
        if key == 'stft':
            if not keep_stft and 'audio_data' in ex[0]:
                continue
            T = max(np.asarray(v).shape[1] for v in vals)
            out = torch.zeros((len(vals),) + tuple(np.asarray(vals[0]).shape[:1]) + (T,) + tuple(np.asarray(vals[0]).shape[2:]))
            for i, v in enumerate(vals):
                v = torch.as_tensor(np.asarray(v), dtype=torch.float32)
                out[i, :, :v.shape[1]] = v
            batch[key] = _pin(out)
        elif key == 'audio_data':
            S = max(np.asarray(v).shape[-1] for v in vals)
            out = torch.zeros((len(vals), 1, S))
            for i, v in enumerate(vals):
                v = torch.as_tensor(np.asarray(v), dtype=torch.float32).reshape(1, -1)
                out[i, :, :v.shape[-1]] = v
            batch[key] = _pin(out)
        elif key in _PAD_LAST:
            T = max(np.asarray(v).shape[-1] for v in vals)
            out = torch.zeros((len(vals), np.asarray(vals[0]).shape[0], T))
            for i, v in enumerate(vals):
                v = torch.as_tensor(np.asarray(v), dtype=torch.float32)
                out[i, :, :v.shape[-1]] = v
            batch[key] = _pin(out)
        elif key in _STACK:
            batch[key] = _pin(torch.as_tensor(np.stack([np.asarray(v) for v in vals])).float()
                              if key == 'weak_targets' else torch.as_tensor(np.stack([np.asarray(v) for v in vals])))
        elif key == 'seq_len':
            batch[key] = [int(v) for v in vals]
        else:
            batch[key] = list(vals)
    return batch


class DeviceLoader:
    """iterate device batches: the H2D copy of batch i+1 (pinned memory, copy stream) overlaps the
    compute of batch i -- the eager counterpart of ``train.GraphedTrainStep.prefetch``."""

    def __init__(self, batches, device):
        self.batches, self.device = batches, torch.device(device)
        self.stream = torch.cuda.Stream(self.device)

    def _upload(self, batch):
        out = {}
        with torch.cuda.stream(self.stream):
            for k, v in batch.items():
                out[k] = v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return out, ev

    def __iter__(self):
        it = iter(self.batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                nxt = self._upload(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for v in cur.values():
                if torch.is_tensor(v):
                    v.record_stream(torch.cuda.current_stream(self.device))
            yield cur


class SyntheticClipStream:
    """endless on-device stream of synthetic 10 s clips for the AudioSet-scale configuration (BASELINE.json
    configs[4]: 2 M clips, K = 527, weak labels only; reference data path AudioSetProvider,
    pb_sed/experiments/weak_label_crnn/training.py:113-128).  Every ``fill_`` draws a NEW batch from
    torch's counter-based (Philox) CUDA generator -- nothing is stored, nothing crosses PCIe -- and is
    CUDA-graph capturable (the generator's offset advances per replay), so ``GraphedTrainStep(input_fn=
    stream.fill_)`` trains on a fresh batch every step.  Clips: 3-tap low-passed Gaussian noise plus one gated
    sinusoid 'event', peak-normalised (audio reader ``normalization_type='max'``, provider.py:309-310);
    labels: Bernoulli(labels_per_clip / K) with at least one active class."""

    def __init__(self, batch, num_events, device, num_samples=160000, sample_rate=16000, labels_per_clip=2.7):
        self.B, self.K, self.S = batch, num_events, num_samples
        self.device = torch.device(device)
        self.p = labels_per_clip / num_events
        self.t = torch.arange(num_samples, device=self.device, dtype=torch.float32) / sample_rate
        self.clips = 0

    def example(self):
        b = {'audio_data': torch.empty((self.B, 1, self.S), device=self.device),
             'weak_targets': torch.empty((self.B, self.K), device=self.device),
             'seq_len': [stft_num_frames(self.S, 320, 960)] * self.B}
        self.fill_(b)
        return b

    def fill_(self, static):
        B, S, dev = self.B, self.S, self.device
        x = torch.randn((B, S + 2), device=dev)
        y = x[:, 2:] + 1.6 * x[:, 1:-1] + .8 * x[:, :-2]
        u = torch.rand((B, 4), device=dev)
        f0 = 200. + 5800. * u[:, 0:1]
        on = u[:, 1:2] * 8.
        off = on + .1 + u[:, 2:3] * 5.
        gate = ((self.t[None] >= on) & (self.t[None] < off)).float()
        y = y + (1. + 4. * u[:, 3:4]) * gate * torch.sin(2. * np.pi * f0 * self.t[None])
        y = y / y.abs().amax(-1, keepdim=True)
        static['audio_data'].copy_(y[:, None])
        weak = (torch.rand((B, self.K), device=dev) < self.p).float()
        first = torch.randint(0, self.K, (B, 1), device=dev)
        weak.scatter_(1, first, 1.)
        static['weak_targets'].copy_(weak)
        self.clips += B
