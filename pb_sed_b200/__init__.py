"""pb_sed_b200 -- B200-native (sm_100a) hot path for fgnt/pb_sed's FBCRNN / BiCRNN.

Drop-in module surface of the padertorch blocks pb_sed composes (``modules``) and of
``pb_sed.models`` (``models``), backed by hand-written CUDA kernels behind a C ABI
(``include/pbsed_b200.h``, ``pb_sed_b200/csrc``).  No CPU / eager fallback.
"""
from . import _lib, ops, modules, models, filters, inference, data  # noqa: F401
from .modules import (CNN, CNN1d, CNN2d, GRU, NormalizedLogMelExtractor, Pad, TakeLast, Mean,  # noqa: F401
                      Sum, Max, compute_mask)

__all__ = ['CNN', 'CNN1d', 'CNN2d', 'GRU', 'NormalizedLogMelExtractor', 'Pad', 'TakeLast', 'Mean',
           'Sum', 'Max', 'compute_mask', 'models', 'modules', 'ops', 'filters', 'inference', 'data']
