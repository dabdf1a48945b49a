"""Generate tests/golden/filters.npz with the REAL pb_sed post-processing functions
(pb_sed/filters.py medfilt / stepfilt, pb_sed/models/base/inference.py filtering / boundariesfilt),
executed unmodified from /root/reference via oracle/ref_loader.load_filters.

    python tests/golden/make_golden_filters.py      (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    filt, inf = ref_loader.load_filters()
    rng = np.random.RandomState(42)
    B, K, T, N = 3, 4, 57, 2
    # sigmoid-like scores with plateaus and exact ties (ties exercise the selection)
    s = 1. / (1. + np.exp(-3. * rng.randn(B, K, T).cumsum(-1) / 4.))
    s[0, 1, 10:20] = s[0, 1, 10]
    s[2, 3, :] = np.round(s[2, 3, :], 1)
    s = s.astype(np.float32)
    seq_len = np.array([57, 44, 21])
    m = (np.arange(T)[None] < seq_len[:, None]).astype(np.float32)
    sm = s * m[:, None]
    out = dict(scores=s, seq_len=seq_len, masked=sm)
    out['med_scalar_7'] = inf.filtering(sm.copy(), filt.medfilt, np.array(7))
    out['med_scalar_1'] = inf.filtering(sm.copy(), filt.medfilt, np.array(1))
    out['med_scalar_101'] = inf.filtering(sm.copy(), filt.medfilt, np.array(101))      # window > T
    mk = np.array([1, 3, 11, 41])
    out['med_len_k'] = mk
    out['med_per_class'] = inf.filtering(sm.copy(), filt.medfilt, mk)
    mnk = np.array([[3, 5, 7, 9], [21, 1, 15, 57]])
    out['med_len_nk'] = mnk
    out['med_per_nk'] = inf.filtering(sm.copy(), filt.medfilt, mnk)
    mn1 = np.array([[5], [13]])
    out['med_len_n1'] = mn1
    out['med_per_n1'] = inf.filtering(sm.copy(), filt.medfilt, mn1)
    out['step_4'] = filt.stepfilt(sm.astype(np.float64), 4, axis=-1)
    out['bnd_scalar_0'] = inf.filtering(sm.copy(), inf.boundariesfilt, np.array(0))
    out['bnd_scalar_6'] = inf.filtering(sm.copy(), inf.boundariesfilt, np.array(6))
    out['bnd_scalar_80'] = inf.filtering(sm.copy(), inf.boundariesfilt, np.array(80))   # 2h > T
    bk = np.array([0, 2, 10, 30])
    out['bnd_len_k'] = bk
    out['bnd_per_class'] = inf.filtering(sm.copy(), inf.boundariesfilt, bk)
    # median -> boundaries chain as inference.py:149-158 applies them
    out['chain_med5_bnd8'] = inf.filtering(inf.filtering(sm.copy(), filt.medfilt, np.array(5)),
                                           inf.boundariesfilt, np.array(8))
    np.savez_compressed(os.path.join(HERE, 'filters.npz'), **out)
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == '__main__':
    main()
