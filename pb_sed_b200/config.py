"""The reference's default model configurations as plain dicts (no sacred / Configurable needed).

``fbcrnn_config``  = ``trainer.model`` of pb_sed/experiments/weak_label_crnn/training.py:158-262
(net_config 'shallow', DESED).  ``bicrnn_config`` = pb_sed/experiments/strong_label_crnn/
training.py:160-262 (same CNN, 2-layer bidirectional GRU, tag conditioning).
Feed them to ``models.weak_label.CRNN.from_config_dict`` / ``models.strong_label.CRNN.from_config_dict``.
"""

STFT = dict(shift=320, window_length=960, size=1024, fading='half', pad=True)   # provider.py:315-323


def _cnn(out_channels_2d, pool_sizes_2d, out_channels_1d, kernel_size_1d):
    common = dict(norm='batch', norm_kwargs={'eps': 1e-3}, activation_fn='relu', pre_activation=True,
                  dropout=0., output_layer=False)
    return dict(
        cnn_2d=dict(out_channels=list(out_channels_2d), pool_size=list(pool_sizes_2d), kernel_size=3,
                    residual_connections=None, **common),
        cnn_1d=dict(out_channels=list(out_channels_1d), kernel_size=list(kernel_size_1d),
                    residual_connections=None, **common))


def fbcrnn_config(num_events=10, sample_rate=16000, stft_size=1024, number_of_filters=128, width=1,
                  out_channels_2d=None, pool_sizes_2d=None, out_channels_1d=None,
                  kernel_size_1d=None, hidden_size=None, num_layers=2, out_hidden=None,
                  strong_fwd_bwd_loss_weight=1., stft_kwargs=None, **model_kwargs):
    c2 = out_channels_2d or [16 * width, 16 * width, 32 * width, 32 * width, 64 * width, 64 * width,
                             128 * width, 128 * width, min(256 * width, 512)]
    p2 = pool_sizes_2d or 4 * [1, (2, 1)] + [1]
    k1 = kernel_size_1d or [1, 3, 3, 3, 1]
    c1 = out_channels_1d or len(k1) * [256 * width]
    hidden = hidden_size or 256 * width
    return dict(
        feature_extractor=dict(sample_rate=sample_rate, stft_size=stft_size,
                               number_of_filters=number_of_filters,
                               stft_kwargs=dict(STFT, size=stft_size, **(stft_kwargs or {}))),
        cnn=_cnn(c2, p2, c1, k1),
        rnn_fwd=dict(rnn=dict(hidden_size=hidden, num_layers=num_layers, dropout=0.),
                     output_net=dict(out_channels=[out_hidden or 256 * width, num_events], kernel_size=1,
                                     norm='batch', norm_kwargs={'eps': 1e-3}, activation_fn='relu',
                                     dropout=0.)),
        rnn_bwd={},
        strong_fwd_bwd_loss_weight=strong_fwd_bwd_loss_weight, **model_kwargs)


def bicrnn_config(num_events=10, tag_conditioning=True, **kw):
    cfg = fbcrnn_config(num_events=num_events, **kw)
    cfg.pop('rnn_bwd')
    cfg.pop('strong_fwd_bwd_loss_weight')
    rnn = cfg.pop('rnn_fwd')
    rnn['rnn'] = dict(rnn['rnn'], bidirectional=True)
    cfg['rnn'] = rnn
    cfg['tag_conditioning'] = tag_conditioning
    return cfg


def tiny_fbcrnn_config(num_events=10, **kw):
    """doctest-sized net (weak_label/crnn.py:16-30) + two pooling layers."""
    return fbcrnn_config(num_events=num_events, stft_size=64, number_of_filters=16,
                         out_channels_2d=[8, 8, 16], pool_sizes_2d=[1, (2, 1), (2, 1)],
                         out_channels_1d=[32, 32], kernel_size_1d=[3, 1], hidden_size=32,
                         out_hidden=16, stft_kwargs=dict(shift=16, window_length=48), **kw)
