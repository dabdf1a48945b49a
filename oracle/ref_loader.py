"""Load the REAL pb_sed model classes from /root/reference on top of the oracle shims.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Only usable in the build
container (``/root/reference`` does not exist on the GPU box); used by
``tests/golden/make_golden.py`` to pin the pb_sed-owned arithmetic
(``CRNN.forward/review/sigmoid`` and the inference heads) and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).

pb_sed imports ``padertorch`` / ``paderbox`` (absent here).  We register stub
modules under those names that expose the restatements from
``oracle/pt_port.py`` and then exec the reference source files *unmodified*:

    pb_sed/models/base/model.py          (SoundEventModel)
    pb_sed/models/weak_label/crnn.py     (FBCRNN)
    pb_sed/models/strong_label/crnn.py   (tag-conditioned BiCRNN)
    pb_sed/evaluation/instance_based.py  (numpy only)

``pb_sed.models.base.__init__`` also imports the inference / tuning drivers
(sed_scores_eval, sacred ...), which are out of scope, so that package is
stubbed with just ``SoundEventModel``.
"""
import importlib.util
import os
import sys
import types

import numpy as np

from . import pt_port

_HERE = os.path.dirname(os.path.abspath(__file__))
# /root/reference in the build container; on the GPU box the UNMODIFIED pip install of the reference under
# baseline/_ref (git-ignored, travels with the snapshot; `pip install --no-deps --target baseline/_ref`)
REFERENCE_ROOT = os.environ.get('PB_SED_REFERENCE') or next(
    (r for r in ('/root/reference', os.path.join(os.path.dirname(_HERE), 'baseline', '_ref'))
     if os.path.isfile(os.path.join(r, 'pb_sed', 'models', 'weak_label', 'crnn.py'))), '/root/reference')


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'pb_sed', 'models', 'weak_label', 'crnn.py'))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _exec(name, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def segment_axis(x, length, shift, axis=-1, end='cut'):
    """paderbox.array.segment_axis for the one call pattern pb_sed uses
    (strong_label/crnn.py:126-131: length == shift, axis=0, end='cut')."""
    x = np.asarray(x)
    axis = axis % x.ndim
    n = (x.shape[axis] - length) // shift + 1
    idx = np.arange(n)[:, None] * shift + np.arange(length)[None, :]
    return np.take(x, idx, axis=axis)


_loaded = {}


def load(modules=None):
    """returns (weak_label_crnn_module, strong_label_crnn_module): the reference's model sources executed
    unmodified.  ``modules``: the namespace the ``padertorch.contrib.je.modules.*`` imports resolve to --
    default the oracle restatements (``pt_port``); pass ``pb_sed_b200.modules`` to run the REAL pb_sed CRNN
    classes over the sm_100a modules (the drop-in check of tests/test_gpu_reference_dropin.py)."""
    key = 'oracle' if modules is None else modules.__name__
    if ('weak', key) in _loaded:
        return _loaded[('weak', key)], _loaded[('strong', key)]
    assert reference_available(), REFERENCE_ROOT
    if not hasattr(np, 'int'):       # pb_sed uses the removed alias (weak_label/crnn.py:252)
        np.int = int
    P = pt_port if modules is None else modules
    if not hasattr(P, 'Model'):
        P = types.SimpleNamespace(**{k: getattr(P, k) for k in dir(P) if not k.startswith('__')}, Model=pt_port.Model)
    _stub('padertorch', Model=P.Model)
    _stub('padertorch.ops')
    _stub('padertorch.ops.sequence')
    _stub('padertorch.ops.sequence.mask', compute_mask=P.compute_mask)
    _stub('padertorch.contrib')
    _stub('padertorch.contrib.je')
    _stub('padertorch.contrib.je.modules')
    _stub('padertorch.contrib.je.modules.conv', Pad=P.Pad, CNN1d=P.CNN1d, CNN2d=P.CNN2d)
    _stub('padertorch.contrib.je.modules.hybrid', CNN=P.CNN)
    _stub('padertorch.contrib.je.modules.features',
          NormalizedLogMelExtractor=P.NormalizedLogMelExtractor)
    _stub('padertorch.contrib.je.modules.reduce',
          TakeLast=P.TakeLast, Mean=P.Mean, Sum=P.Sum, Max=P.Max)
    _stub('padertorch.contrib.je.modules.rnn', GRU=P.GRU, TransformerEncoder=None)
    _stub('paderbox')
    _stub('paderbox.array', segment_axis=segment_axis)
    _stub('pb_sed')
    _stub('pb_sed.evaluation')
    inst = _exec('pb_sed.evaluation.instance_based', 'pb_sed/evaluation/instance_based.py')
    sys.modules['pb_sed.evaluation'].instance_based = inst
    _stub('pb_sed.models')
    model_mod = _exec('pb_sed.models.base.model', 'pb_sed/models/base/model.py')
    base = _stub('pb_sed.models.base', SoundEventModel=model_mod.SoundEventModel)
    sys.modules['pb_sed.models'].base = base
    weak = _exec('pb_sed.models.weak_label.crnn', 'pb_sed/models/weak_label/crnn.py')
    strong = _exec('pb_sed.models.strong_label.crnn', 'pb_sed/models/strong_label/crnn.py')
    _loaded[('weak', key)], _loaded[('strong', key)] = weak, strong
    return weak, strong


def load_filters():
    """the REAL ``pb_sed/filters.py`` and the post-processing functions of
    ``pb_sed/models/base/inference.py`` (``filtering``, ``boundariesfilt``), executed unmodified;
    everything they import that is absent here (sed_scores_eval, padertorch, paderbox) is stubbed --
    the functions used run on numpy / scipy / torch only."""
    if 'filters' in _loaded:
        return _loaded['filters'], _loaded['inference']
    assert reference_available(), REFERENCE_ROOT
    if not hasattr(np, 'int'):
        np.int = int
    if not hasattr(np, 'bool'):
        np.bool = bool
    for name in ('paderbox', 'paderbox.array', 'pb_sed', 'pb_sed.utils', 'sed_scores_eval',
                 'sed_scores_eval.utils', 'padertorch', 'padertorch.ops', 'padertorch.ops.sequence'):
        if name not in sys.modules:
            _stub(name)
    _stub('paderbox.array.segment', segment_axis=segment_axis)
    _stub('pb_sed.utils.segment', segment_batch=None, merge_segments=None)
    _stub('sed_scores_eval.utils.scores', create_score_dataframe=None)
    sys.modules['sed_scores_eval'].io = None
    if 'padertorch.ops.sequence.mask' not in sys.modules:
        _stub('padertorch.ops.sequence.mask', compute_mask=pt_port.compute_mask)
    filt = _exec('pb_sed.filters', 'pb_sed/filters.py')
    sys.modules['pb_sed'].filters = filt
    inf = _exec('pb_sed.models.base.inference', 'pb_sed/models/base/inference.py')
    _loaded.update(filters=filt, inference=inf)
    return filt, inf


def load_segment():
    """the REAL ``pb_sed/utils/segment.py`` (``merge_segments`` is pure numpy; ``segment_batch`` needs
    padertorch's Segmenter, which is stubbed out and therefore NOT usable through this loader)."""
    if 'segment' in _loaded:
        return _loaded['segment']
    assert reference_available(), REFERENCE_ROOT
    for name in ('padertorch', 'padertorch.data'):
        if name not in sys.modules:
            _stub(name)
    _stub('padertorch.data.segment', Segmenter=None)
    seg = _exec('pb_sed_ref_utils_segment', 'pb_sed/utils/segment.py')
    _loaded['segment'] = seg
    return seg
