"""Score post-processing on the GPU -- mirror of ``pb_sed/filters.py`` (``medfilt``) and of the
post-processing part of ``pb_sed/models/base/inference.py`` (``filtering`` :225-266,
``boundariesfilt`` :269-289, sequence mask / tag mask :143-183).

The reference copies every score tensor to the host and filters it row by row with
``np.apply_along_axis``; here the scores stay in HBM and one kernel launch handles all rows with
their individual (per class / per hyper-parameter set) filter lengths.  Same function names,
argument meaning and result dtypes as the reference; tensors are CUDA tensors, ``axis`` must be the
last (time) axis.
"""

import numpy as np
import torch

from ._lib import call
from .ops import _ptr, _stream, _f32c


def _lengths(filter_length, device):
    fl = np.asarray(filter_length).astype(np.int64)
    return fl, torch.from_numpy(fl.reshape(-1).astype(np.int32)).to(device)


def _rows(x):
    T = x.shape[-1]
    return int(x.numel() // T), T


def _seq_ptr(seq_len, B, device):
    if seq_len is None:
        return None, None
    s = torch.as_tensor(np.asarray(seq_len).astype(np.int32)).to(device)
    assert s.numel() == B, (s.shape, B)
    return s, _ptr(s)


def _medfilt_rows(x, lens_dev, filt_mod, seq_len=None):
    x = _f32c(x)
    R, T = _rows(x)
    y = torch.empty_like(x)
    keep, sp = _seq_ptr(seq_len, x.shape[0], x.device)
    call('pbsed_medfilt', _ptr(x), R, T, _ptr(lens_dev), filt_mod, sp, R // x.shape[0], _ptr(y), _stream())
    return y


def _boundariesfilt_rows(x, lens_dev, filt_mod, seq_len=None):
    x = _f32c(x)
    R, T = _rows(x)
    y = torch.empty(x.shape, device=x.device, dtype=torch.float64)
    keep, sp = _seq_ptr(seq_len, x.shape[0], x.device)
    call('pbsed_boundariesfilt', _ptr(x), R, T, _ptr(lens_dev), filt_mod, sp, R // x.shape[0], _ptr(y), _stream())
    return y


def medfilt(x, n, axis=-1, seq_len=None):
    """pb_sed/filters.py:56-83 (scipy.signal.medfilt per row: zero padded, odd n; n == 1: identity)."""
    assert axis in (-1, x.dim() - 1), 'time must be the last axis'
    n = int(n)
    assert n % 2 == 1, n
    if n == 1 and seq_len is None:
        return x
    _, dev = _lengths(n, x.device)
    return _medfilt_rows(x, dev, 1, seq_len)


def boundariesfilt(score_arr, stepfilt_length, axis=-1, seq_len=None):
    """inference.py:269-289; float64 like the reference when a step filter is applied."""
    assert axis in (-1, score_arr.dim() - 1), 'time must be the last axis'
    n = int(stepfilt_length)
    assert n % 2 == 0 and n >= 0, n
    _, dev = _lengths(n, score_arr.device)
    y = _boundariesfilt_rows(score_arr, dev, 1, seq_len)
    return y if n > 0 else y.to(score_arr.dtype)


_ROW_KERNELS = {medfilt: _medfilt_rows, boundariesfilt: _boundariesfilt_rows}


def filtering(score_arr, filter_fn, filter_length, seq_len=None):
    """inference.py:225-266: scalar, per-class (K,) or per-(n, class) (N, 1|K) filter lengths; all rows
    in ONE launch.  ``filter_fn`` is ``medfilt`` or ``boundariesfilt`` of this module."""
    rows_fn = _ROW_KERNELS[filter_fn]
    fl, _ = _lengths(filter_length, score_arr.device)
    if filter_fn is medfilt:
        assert ((fl % 2) == 1).all(), fl
    else:
        assert ((fl % 2) == 0).all() and (fl >= 0).all(), fl
    b, *_, k, t = score_arr.shape
    if fl.ndim == 0:
        return filter_fn(score_arr, int(fl), axis=-1, seq_len=seq_len)
    if fl.ndim == 1:
        assert fl.shape[0] == k, fl.shape
    elif fl.ndim == 2:
        assert fl.shape[1] in (1, k), fl.shape
        n = fl.shape[0]
        if score_arr.dim() == 3:
            score_arr = score_arr[:, None].expand(b, n, k, t)
        elif score_arr.dim() == 4:
            assert n == score_arr.shape[1], (score_arr.shape, n)
        else:
            raise ValueError('scores returned by model must be 3- or 4-dimensional.')
        fl = np.broadcast_to(fl, (n, k))
    else:
        raise ValueError(f'filter_length.ndim must not be greater than 2 but {filter_length} was given.')
    dev = torch.from_numpy(np.ascontiguousarray(fl).reshape(-1).astype(np.int32)).to(score_arr.device)
    y = rows_fn(score_arr.contiguous(), dev, int(fl.size), seq_len)
    return y.to(score_arr.dtype)           # the reference assigns into score_arr: dtype is kept


def tag_mask_(scores, tags, apply_mask):
    """inference.py:170-183 on the batched layout: scores (B, [N,] K, T) *= max(tags (B,K), 1 - apply)."""
    ap = np.asarray(apply_mask, dtype=np.float32)
    if not ap.any():
        return scores
    squeeze = scores.dim() == 3
    s4 = scores[:, None] if squeeze else scores
    assert s4.is_contiguous() and s4.dtype == torch.float32
    B, N, K, T = s4.shape
    ap = np.broadcast_to(ap if ap.ndim == 2 else ap.reshape(1, -1) if ap.ndim == 1 else ap.reshape(1, 1), (N, K)).copy()
    ap_dev = torch.from_numpy(ap).to(scores.device)
    tg = _f32c(tags.to(scores.device))
    assert tg.shape == (B, K), tg.shape
    call('pbsed_tag_mask', _ptr(s4), _ptr(tg), _ptr(ap_dev), B, N, K, T, _stream())
    return scores


def post_process(scores, seq_len, medfilt_length=1, stepfilt_length=None, apply_mask=False, tags=None):
    """the per-segment pipeline of inference.py:134-183 without leaving the GPU: sequence mask (fused
    into the first filter) -> median filter -> optional boundary filter -> optional tag mask.
    scores (B, [N,] K, T) CUDA float32 (the ensemble mean of the models' scores)."""
    y = filtering(scores, medfilt, np.asarray(medfilt_length, dtype=int), seq_len=seq_len)
    if y is scores:                                   # n == 1 and no mask requested by medfilt(): mask here
        y = filtering(scores, medfilt, np.asarray(1), seq_len=seq_len)
    if stepfilt_length is not None:
        y = filtering(y, boundariesfilt, np.asarray(stepfilt_length, dtype=int))
    if np.asarray(apply_mask).any():
        assert tags is not None
        y = tag_mask_(y.float().contiguous() if y.dtype != torch.float32 else y.contiguous(), tags, apply_mask)
    return y
