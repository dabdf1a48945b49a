"""shared helpers for the parity tests (the oracle is the checker; see oracle/__init__.py)."""
import copy
import os
import re

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
TINY_STFT = dict(shift=16, window_length=48, size=64)


def padertorch_key(k):
    """the golden files were frozen with a flat ``convs.<i>.{weight,bias}`` / ``norms.<i>.*`` key layout; the
    modules (and padertorch checkpoints) nest them as ``convs.<i>.conv.*`` / ``convs.<i>.norm.*``."""
    k = re.sub(r'convs\.(\d+)\.(weight|bias)$', r'convs.\1.conv.\2', k)
    return re.sub(r'norms\.(\d+)\.', r'convs.\1.norm.', k)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    d = {padertorch_key(k): z[k] for k in z.files if not k.endswith('feature_extractor.fbanks')}
    state = {k[len('state.'):]: torch.from_numpy(v) for k, v in d.items() if k.startswith('state.')}
    grads = {k[len('grad.'):]: torch.from_numpy(v) for k, v in d.items() if k.startswith('grad.')}
    return d, state, grads


def golden_batch(d, keys, device=None, stft=True):
    b = {k: torch.from_numpy(d[k]) for k in keys}
    b['stft' if stft else 'audio_data'] = torch.from_numpy(d['stft'] if stft else d['audio'])
    if device is not None:
        b = {k: v.to(device) for k, v in b.items()}
    b['seq_len'] = [int(s) for s in d['seq_len']]
    return b


def ref_layout_grads(model):
    """gradients of a pb_sed_b200 model keyed / shaped like the reference state dict."""
    g = copy.deepcopy(model)
    with torch.no_grad():
        for (n, p), (_, q) in zip(model.named_parameters(), g.named_parameters()):
            q.data = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().clone()
    return {k: v.cpu() for k, v in g.state_dict().items()}


def maxdiff(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max()) if a.numel() else 0.


def reldiff(a, b):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))
