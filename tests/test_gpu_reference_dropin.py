"""GPU (-m gpu): the drop-in boundary itself.  The UNMODIFIED reference model sources
(pb_sed/models/weak_label/crnn.py, pb_sed/models/strong_label/crnn.py, executed through oracle/ref_loader.py
from /root/reference or from the pip install under baseline/_ref) run with ``pb_sed_b200.modules`` in the place
of ``padertorch.contrib.je.modules.*``:

  * the module kwargs are the ones pb_sed's experiment config passes (weak_label_crnn/training.py:190-260) after
    the reference's own ``finalize_dogmatic_config`` (weak_label/crnn.py:304-340, strong_label/crnn.py:155-198)
    wired the sizes -- run here on a minimal stand-in for padertorch's nested config dict;
  * the reference ``CRNN.forward`` / ``review`` / heads then drive the sm_100a kernels through the modules'
    reference-layout call signatures; scores, loss and gradients must equal those of this package's own model
    classes (same kernels underneath) and of the CPU oracle.

Skipped when no copy of the reference is present."""
import inspect

import numpy as np
import pytest
import torch

from oracle import models as OM, pt_port as P, ref_loader
from util import maxdiff, ref_layout_grads, TINY_STFT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.reference_available(), reason='no copy of the reference sources')]
DEV = 'cuda:0'


class Cfg(dict):
    """the two behaviours of padertorch's config dict that finalize_dogmatic_config relies on: assigning a
    dict with a 'factory' MERGES into the existing sub-config and fills the factory's keyword defaults;
    ``to_dict()`` gives a plain copy."""

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, Cfg):
            cur = self.get(key)
            cur = cur if isinstance(cur, Cfg) else Cfg()
            factory = value.get('factory', cur.get('factory'))
            if factory is not None and inspect.isclass(factory):
                for name, par in inspect.signature(factory.__init__).parameters.items():
                    if par.default is not inspect.Parameter.empty and name not in cur:
                        dict.__setitem__(cur, name, par.default)
            for k, v in value.items():
                cur[k] = v
            value = cur
        dict.__setitem__(self, key, value)

    def update(self, other=(), **kw):
        """values set inside finalize_dogmatic_config are DEFAULTS: what the user configured wins."""
        for k, v in dict(other, **kw).items():
            if k not in self:
                self[k] = v

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Cfg) else v) for k, v in self.items()}


def build(node):
    """instantiate a finalized (sub-)config: {'factory': cls, **kwargs}; nested module configs stay dicts (the
    modules take them as kwargs dicts, like padertorch's from_config does recursively)."""
    kw = {k: v for k, v in node.items() if k != 'factory'}
    return node['factory'](**kw)


def tiny_user_config(num_events=10, strong=False):
    """the kwargs of weak_label_crnn/training.py:190-260 at doctest size (weak_label/crnn.py:16-30)."""
    nk = {'eps': 1e-3}
    cfg = Cfg()
    cfg['feature_extractor'] = {'sample_rate': 16000, 'stft_size': 64, 'number_of_filters': 16,
                                'stft_kwargs': dict(TINY_STFT)}
    cfg['cnn'] = {'cnn_2d': {'out_channels': [8, 8, 16], 'pool_size': [1, (2, 1), (2, 1)], 'kernel_size': 3,
                             'norm': 'batch', 'norm_kwargs': nk, 'activation_fn': 'relu', 'pre_activation': True,
                             'dropout': 0., 'output_layer': False},
                  'cnn_1d': {'out_channels': [32, 32], 'kernel_size': [3, 1], 'norm': 'batch', 'norm_kwargs': nk,
                             'activation_fn': 'relu', 'pre_activation': True, 'dropout': 0., 'output_layer': False}}
    rnn = {'rnn': {'hidden_size': 32, 'num_layers': 2, 'dropout': 0.},
           'output_net': {'out_channels': [16, num_events], 'kernel_size': 1, 'norm': 'batch', 'norm_kwargs': nk,
                          'activation_fn': 'relu', 'dropout': 0.}}
    if strong:
        rnn['rnn']['bidirectional'] = True
        cfg['rnn'] = rnn
    else:
        cfg['rnn_fwd'] = rnn
    return cfg


def test_real_weak_label_crnn_runs_over_the_b200_modules():
    import pb_sed_b200.modules as M
    from pb_sed_b200 import config
    from pb_sed_b200.models import weak_label
    weak, _ = ref_loader.load(modules=M)
    cfg = tiny_user_config()
    for k in ('rnn_bwd',):
        cfg[k] = {}
    cfg['cnn']['cnn_2d']['in_channels'] = None
    weak.CRNN.finalize_dogmatic_config(cfg)                       # the reference's own size wiring
    assert cfg['cnn']['cnn_2d']['in_channels'] == 1 and cfg['cnn']['input_height'] == 16
    assert cfg['rnn_fwd']['rnn']['input_size'] == 32 and cfg['rnn_bwd']['reverse'] is True
    ref_model = weak.CRNN(feature_extractor=build(cfg['feature_extractor']), cnn=build(cfg['cnn']),
                          rnn_fwd=build(cfg['rnn_fwd']), rnn_bwd=build(cfg['rnn_bwd']))
    ora = OM.tiny_fbcrnn(seed=3)
    ref_model.load_state_dict(ora.state_dict())                   # padertorch-layout checkpoint, strict
    own = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config())
    own.load_state_dict(ora.state_dict())
    ref_model.to(DEV).train(); own.to(DEV).train()
    batch = OM.synthetic_batch(4, num_samples=645, stft_kwargs=TINY_STFT, seed=3, seq_len=[41, 40, 33, 17])
    gb = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items() if k != 'audio_data'}
    cb = {k: v for k, v in batch.items() if k != 'audio_data'}
    out = ref_model(dict(gb))                                     # REAL pb_sed forward (pops 'stft' in training)
    review = ref_model.review(gb, out)                            # REAL pb_sed review / losses (torch ops on the GPU)
    review['loss'].backward()
    out_own = own(dict(gb))
    loss_own = own.review(gb, out_own)['loss']
    loss_own.backward()
    out_ora = ora.train()(dict(cb))
    loss_ora = ora.review(cb, out_ora)['loss']
    loss_ora.backward()
    mask = P.compute_mask(out_ora[0], np.array(batch['seq_len']), 0, -1)
    for i in (0, 1):
        assert maxdiff(out[i].cpu() * mask, out_own[i].cpu() * mask) < 1e-6       # same kernels underneath
        assert maxdiff(out[i].cpu() * mask, out_ora[i] * mask) < 1e-4             # CPU oracle
    assert abs(float(review['loss']) - float(loss_own)) < 1e-5
    assert abs(float(review['loss']) - float(loss_ora)) < 1e-4
    assert set(review) >= {'loss', 'scalars', 'images', 'buffers'}
    g_ref, g_own = ref_layout_grads(ref_model), ref_layout_grads(own)
    for k, p in ora.named_parameters():
        tol = 5e-4 * max(1., float(p.grad.abs().max()))
        assert maxdiff(g_ref[k], p.grad) < tol, k
        assert maxdiff(g_ref[k], g_own[k]) < tol, k
    # the reference's inference heads over the same modules
    ref_model.eval(); ora.eval()
    with torch.no_grad():
        tag, _ = ref_model.tagging(dict(gb))
        tag_o, _ = ora.tagging(dict(cb))
        assert maxdiff(tag, tag_o) < 1e-4
        bd, _ = ref_model.boundaries_detection(dict(gb))
        bd_o, _ = ora.boundaries_detection(dict(cb))
        assert maxdiff(bd.cpu(), bd_o) < 1e-4


def test_real_strong_label_crnn_runs_over_the_b200_modules():
    import pb_sed_b200.modules as M
    _, strong = ref_loader.load(modules=M)
    cfg = tiny_user_config(strong=True)
    cfg['tag_conditioning'] = True
    cfg['cnn']['cnn_2d']['in_channels'] = None
    strong.CRNN.finalize_dogmatic_config(cfg)
    assert cfg['cnn']['conditional_dims'] == 10 and cfg['cnn']['cnn_2d']['in_channels'] == 11
    model = strong.CRNN(feature_extractor=build(cfg['feature_extractor']), cnn=build(cfg['cnn']), rnn=build(cfg['rnn']),
                        tag_conditioning=True)
    ora = OM.build_bicrnn(n_mels=16, stft_size=64, channels_2d=[8, 8, 16], pool_2d=[1, (2, 1), (2, 1)],
                          channels_1d=[32, 32], k1d=[3, 1], hidden=32, rnn_layers=2, out_hidden=16, seed=4)
    model.load_state_dict(ora.state_dict())
    model.to(DEV).train(); ora.train()
    batch = OM.synthetic_batch(4, num_samples=645, stft_kwargs=TINY_STFT, seed=4, seq_len=[41, 40, 33, 17])
    cb = dict(stft=batch['stft'], seq_len=batch['seq_len'], weak_targets=batch['weak_targets'],
              strong_targets=batch['boundary_targets'], tag_condition=batch['weak_targets'] > .5)
    gb = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in cb.items()}
    out = model(dict(gb))
    loss = model.review(gb, out)['loss']
    loss.backward()
    out_o = ora(dict(cb))
    loss_o = ora.review(cb, out_o)['loss']
    loss_o.backward()
    mask = P.compute_mask(out_o[0], np.array(batch['seq_len']), 0, -1)
    assert maxdiff(out[0].cpu() * mask, out_o[0] * mask) < 1e-4
    assert abs(float(loss) - float(loss_o)) < 2e-4 * max(1., abs(float(loss_o)))
    g = ref_layout_grads(model)
    for k, p in ora.named_parameters():
        assert maxdiff(g[k], p.grad) < 5e-4 * max(1., float(p.grad.abs().max())), k
