"""CPU restatement of pb_sed's score post-processing.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows pb_sed/filters.py:56-83 (``medfilt``), :113-135 (``stepfilt``) and
pb_sed/models/base/inference.py:225-266 (``filtering``), :269-289 (``boundariesfilt``), :143-183
(sequence mask -> median filter -> boundary filter -> tag mask).  **Pinned**: this is pb_sed's own
code, executable here (numpy / scipy / torch only; ``oracle/ref_loader.load_filters``), and
``tests/golden/filters.npz`` was produced by the real functions
(``tests/golden/make_golden_filters.py``).
"""
import numpy as np


def medfilt(x, n, axis=-1):
    """filters.py:56-83: zero-padded running median of odd length n along ``axis`` (n == 1: identity)."""
    n = int(n)
    if n == 1:
        return x
    assert n % 2 == 1, n
    x = np.moveaxis(np.asarray(x), axis, -1)
    h = (n - 1) // 2
    pad = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(h, h)])
    win = np.lib.stride_tricks.sliding_window_view(pad, n, axis=-1)
    y = np.sort(win, axis=-1)[..., h].astype(x.dtype)
    return np.moveaxis(y, -1, axis)


def stepfilt(x, n, axis=-1):
    """filters.py:113-135: correlate with [-1]*(n/2) + [+1]*(n/2), scaled by 2/n; pad n/2 | n/2-1.
    float64 result (the filter is float64)."""
    n = int(n)
    assert n % 2 == 0, n
    h = n // 2
    x = np.moveaxis(np.asarray(x), axis, -1)
    filt = np.concatenate((-np.ones(h), np.ones(h))) / h
    pad = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(h, h - 1)])
    win = np.lib.stride_tricks.sliding_window_view(pad, n, axis=-1)
    y = (win.astype(np.float64) * filt).sum(-1)
    return np.moveaxis(y, -1, axis)


def _cummax(x, axis):
    return np.maximum.accumulate(x, axis=axis)


def boundariesfilt(score_arr, stepfilt_length, axis=-1):
    """inference.py:269-289."""
    if stepfilt_length > 0:
        fwd = stepfilt(score_arr, stepfilt_length, axis=axis)
        bwd = stepfilt(np.flip(score_arr, axis=axis), stepfilt_length, axis=axis)
    else:
        fwd, bwd = score_arr, np.flip(score_arr, axis=axis)
    return np.minimum(_cummax(fwd, axis), np.flip(_cummax(bwd, axis), axis=axis))


def filtering(score_arr, filter_fn, filter_length):
    """inference.py:225-266 (scalar / per-class / per-(n, class) filter lengths; the per-class branches
    assign into ``score_arr`` and therefore keep its dtype)."""
    filter_length = np.asarray(filter_length)
    b, *_, k, t = score_arr.shape
    if filter_length.ndim == 0:
        return filter_fn(score_arr, filter_length, axis=-1)
    score_arr = score_arr.copy()
    if filter_length.ndim == 1:
        assert filter_length.shape[0] == k
        for c, n in enumerate(filter_length):
            score_arr[..., c, :] = filter_fn(score_arr[..., c, :], n, axis=-1)
        return score_arr
    assert filter_length.ndim == 2 and filter_length.shape[1] in (1, k)
    n_sets = filter_length.shape[0]
    if score_arr.ndim == 3:
        score_arr = np.broadcast_to(score_arr[:, None], (b, n_sets, k, t)).copy()
    for j in range(n_sets):
        if filter_length.shape[1] == 1:
            score_arr[:, j] = filter_fn(score_arr[:, j], filter_length[j, 0], axis=-1)
        else:
            for c in range(k):
                score_arr[:, j, c] = filter_fn(score_arr[:, j, c], filter_length[j, c], axis=-1)
    return score_arr


def post_process(scores, seq_len, medfilt_length=1, stepfilt_length=None, apply_mask=False, tags=None):
    """the per-segment score pipeline of inference.py:134-183 on a (B, [N,] K, T) array (before the
    per-example split): ensemble mean is the caller's; sequence mask -> median filter -> optional
    boundary filter -> optional tag mask (tags (B, K) in {0,1})."""
    scores = np.asarray(scores)
    t = scores.shape[-1]
    m = (np.arange(t)[None] < np.asarray(seq_len)[:, None]).astype(scores.dtype)
    scores = scores * m.reshape((scores.shape[0],) + (1,) * (scores.ndim - 2) + (t,))
    scores = filtering(scores, medfilt, np.asarray(medfilt_length, dtype=int))
    if stepfilt_length is not None:
        scores = filtering(scores, boundariesfilt, np.asarray(stepfilt_length, dtype=int))
    apply_mask = np.asarray(apply_mask, dtype=bool)
    if apply_mask.any():
        tg = np.asarray(tags, dtype=scores.dtype)                              # (B, K)
        ap = apply_mask.astype(scores.dtype)
        if scores.ndim == 4:
            ap = np.broadcast_to(ap, scores.shape[1:3]) if ap.ndim == 2 else np.broadcast_to(ap, (scores.shape[2],))[None]
            mask = np.maximum(tg[:, None, :], 1 - ap[None])                     # (B, N, K)
        else:
            mask = np.maximum(tg, 1 - np.broadcast_to(ap, tg.shape[1:])[None])  # (B, K)
        scores = scores * mask[..., None]
    return scores
