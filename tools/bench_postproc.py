"""micro-bench of the GPU score post-processing next to the reference's CPU path
(scipy.signal.medfilt per row through np.apply_along_axis, pb_sed/filters.py:76-80; numpy step filter
+ torch.cummax, inference.py:269-289) on the box's host cores.  Prints one JSON line.

    python tools/bench_postproc.py [--batch 512]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cpu_reference(s, seq_len, med, step):
    from scipy import signal
    from oracle import filters as OF
    t = s.shape[-1]
    m = (np.arange(t)[None] < seq_len[:, None]).astype(s.dtype)
    x = s * m[:, None]
    for c, n in enumerate(med):                      # filtering(): per class, apply_along_axis(medfilt)
        if n > 1:
            x[:, c] = np.apply_along_axis(lambda r: signal.medfilt(r, int(n)), -1, x[:, c])
    return OF.filtering(x, OF.boundariesfilt, step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=512)
    ap.add_argument('--cpu-batch', type=int, default=32)
    args = ap.parse_args()
    from pb_sed_b200 import filters as GF, _lib
    _lib.load()
    rng = np.random.RandomState(0)
    B, K, T = args.batch, 10, 500
    s = (1. / (1. + np.exp(-rng.randn(B, K, T).cumsum(-1) / 3.))).astype(np.float32)
    seq_len = np.full(B, T)
    med = np.array([1, 11, 21, 41, 61, 81, 101, 151, 201, 301])
    step = np.array([0, 2, 4, 10, 20, 40, 80, 100, 200, 400])
    x = torch.from_numpy(s).cuda()

    def gpu():
        y = GF.post_process(x, seq_len, medfilt_length=med, stepfilt_length=step)
        return y
    for _ in range(3):
        y = gpu()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        y = gpu()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # end to end: host scores in, filtered scores back on the host
    hs = torch.from_numpy(s).pin_memory()
    t0 = time.perf_counter()
    for _ in range(5):
        out = GF.post_process(hs.cuda(non_blocking=True), seq_len, medfilt_length=med, stepfilt_length=step).cpu()
    ms_e2e = (time.perf_counter() - t0) / 5 * 1e3
    cb = args.cpu_batch
    t0 = time.perf_counter()
    ref = cpu_reference(s[:cb].copy(), seq_len[:cb], med, step)
    cpu_s = time.perf_counter() - t0
    err = float(np.abs(y[:cb].cpu().numpy() - ref).max())
    print(json.dumps({
        'workload': f'post-processing of (B={B}, K={K}, T={T}) scores: per-class median filters {med.tolist()} '
                    f'+ boundary filters {step.tolist()}',
        'gpu_ms': ms, 'gpu_clips_per_s': B / (ms / 1e3), 'gpu_e2e_ms_host_in_host_out': ms_e2e,
        'gpu_e2e_clips_per_s': B / (ms_e2e / 1e3),
        'cpu_reference_clips_per_s': cb / cpu_s, 'cpu_sample': f'{cb} clips, {cpu_s:.2f} s, '
        f'{len(os.sched_getaffinity(0))} cores available (numpy/scipy path is single-threaded like the reference)',
        'max_abs_diff_vs_cpu_reference': err,
        'algorithmic_bytes': B * K * T * (4 + 4 + 4 + 8), 'achieved_GBs': B * K * T * 20 / (ms / 1e3) / 1e9}))


if __name__ == '__main__':
    main()
