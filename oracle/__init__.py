"""CPU oracle for the pb_sed FBCRNN / BiCRNN hot path.  TEST INFRASTRUCTURE ONLY.

This package is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``pb_sed_b200/``
imports ``oracle``.

PARITY STATUS — read before trusting a number checked against this oracle:

* pb_sed-owned arithmetic (``CRNN.sigmoid/forward/review``, the weak and
  strong forward-backward losses, the tagging / boundary / sliding-window
  heads; reference ``pb_sed/models/weak_label/crnn.py:58-302`` and
  ``pb_sed/models/strong_label/crnn.py:60-210``) is **pinned**: the golden
  vectors under ``tests/golden/`` were produced by importing the *real*
  pb_sed classes from ``/root/reference`` (``tests/golden/make_golden.py``)
  on top of the module restatements in this package.
* pb_sed-owned post-processing and inference drivers (``pb_sed/filters.py`` medfilt / stepfilt,
  ``pb_sed/models/base/inference.py`` filtering / boundariesfilt / masking / the tagging, boundaries and
  SED drivers, ``pb_sed/utils/segment.py`` merge_segments) is **pinned** the same way:
  ``oracle/filters.py`` is checked against ``tests/golden/filters.npz`` (real functions,
  ``tests/golden/make_golden_filters.py``) and the GPU drivers against
  ``tests/golden/inference_tiny.npz`` (real drivers on the real FBCRNN class,
  ``tests/golden/make_golden_inference.py``).
* train-time augmentation (mel warping, time / frequency masks, noise, time-warped STFT grid) is a
  restatement from the call-site kwargs only; where the upstream formula could not be recalled the
  oracle DEFINES it (see ``pt_port.warp_mel`` / ``time_warp_grid``)  ->  **parity unpinned**.
* third-party arithmetic (padertorch@b7ba24a / paderbox@809b272: STFT,
  mel filterbank, Normalization, CNN2d/CNN1d, GRU wrapper, reductions) is a
  restatement from the published algorithm; neither package is vendored in
  ``/root/reference`` nor installed here and the reference holds no test
  that pins their values  ->  **parity unpinned** at that boundary
  (SURVEY.md section 8c, Appendix A).  Everything that bottoms out in
  ``torch.nn.{Conv2d,Conv1d,GRU,BCELoss}`` / ``torch.cummax`` executes the
  same CPU kernels the reference would.
"""
