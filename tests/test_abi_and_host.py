"""CPU: the C-ABI library builds, loads and exports every symbol include/pbsed_b200.h declares;
host-side logic (state-dict layouts, filterbank tables, configs, LR schedule)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import models as OM, pt_port as P


def test_library_exports_every_declared_symbol(built_lib):
    from pb_sed_b200 import _lib
    protos = _lib.parse_header()
    assert len(protos) >= 20
    for name in protos:
        assert hasattr(built_lib, name), name
    assert built_lib.pbsed_abi_version() == 6
    assert built_lib.pbsed_launch_count() >= 0
    assert isinstance(built_lib.pbsed_last_kernel(), bytes)      # const char* entry point (not an int prototype)


def test_bad_arguments_return_einval_without_touching_the_gpu(built_lib):
    from pb_sed_b200 import _lib
    assert built_lib.pbsed_tapgemm(None, None, None, None, None, None, None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    d = _lib.TapGemmDesc()
    d.B = d.F_in = d.F_out = d.T = d.Cin = d.Cout = 1
    d.ntaps = 99
    assert built_lib.pbsed_tapgemm_wgrad(ctypes.byref(d), None, None, None, None, None, 0, None, None, None) == -1
    assert built_lib.pbsed_gru_fwd(None, None, None, None, 1, 1, 48, 1, None, None, 48, None, None) == -1
    assert built_lib.pbsed_stft_logmel(None, 1, 1, 1, 1, 8, 0, 1, None, None, None, None, 1, 1, 0, None, None, None, None, None) == -1
    assert built_lib.pbsed_medfilt(None, 1, 1, None, 1, None, 1, None, None) == -1
    assert built_lib.pbsed_make_warped_fbank(None, None, 1, 1, 2, 0., 1., 1., 1., None, None, None, 1, None) == -1
    with pytest.raises(_lib.PbsedError):
        _lib.call('pbsed_adam_step', None, None, None, None, 0, None, None, None, 0, None)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from pb_sed_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(RuntimeError, match='no CPU / eager fallback'):
        _lib.load()


def test_ops_refuse_cpu_tensors(built_lib):
    from pb_sed_b200 import config
    from pb_sed_b200.models import weak_label
    m = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config())
    b = OM.synthetic_batch(2, num_samples=16 * 20, stft_kwargs=dict(shift=16, window_length=48, size=64))
    with pytest.raises(AssertionError, match='CUDA tensors only'):
        m(dict(b))


def test_state_dict_interchanges_with_reference_layout():
    from pb_sed_b200 import config
    from pb_sed_b200.models import weak_label, strong_label
    for prod, ora in ((weak_label.CRNN.from_config_dict(config.fbcrnn_config()), OM.build_fbcrnn(seed=3)),
                      (strong_label.CRNN.from_config_dict(config.bicrnn_config()), OM.build_bicrnn(seed=4))):
        assert sum(p.numel() for p in prod.parameters()) == sum(p.numel() for p in ora.parameters())
        prod.load_state_dict(ora.state_dict(), strict=True)
        sd = prod.state_dict()
        for k, v in ora.state_dict().items():
            assert sd[k].shape == v.shape and torch.equal(sd[k], v), k
        ora2 = type(ora).__name__
        fresh = OM.build_fbcrnn(seed=9) if ora2 == 'FBCRNN' else OM.build_bicrnn(seed=9)
        fresh.load_state_dict(sd, strict=True)          # and back into torch modules
    assert sum(p.numel() for p in weak_label.CRNN.from_config_dict(config.fbcrnn_config()).parameters()) == 3493188


def test_state_dict_keys_follow_padertorch_and_the_init_checkpoint_recipe():
    """checkpoint keys nest as ``convs.<i>.conv.{weight,bias}`` / ``convs.<i>.norm.*`` (padertorch Conv modules);
    the reference's init-checkpoint code (weak_label_crnn/training.py:327-342) pops the output layer by the
    second dotted component of the LAST sorted key -- run that recipe verbatim on our state dict."""
    from pb_sed_b200 import config
    from pb_sed_b200.models import weak_label
    src = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config(num_events=10))
    dst = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config(num_events=7))   # other label set
    sd = src.state_dict()
    keys = set(sd)
    assert 'cnn.cnn_2d.convs.0.conv.weight' in keys and 'cnn.cnn_2d.convs.1.norm.running_mean' in keys
    assert 'cnn.cnn_2d.convs.0.norm.scale' not in keys               # bare input layer
    assert 'rnn_fwd.output_net.convs.0.norm.scale' in keys and 'rnn_fwd.output_net.convs.1.conv.bias' in keys
    assert not any('fbanks' in k or k.split('.')[-2] == 'norms' for k in keys)
    assert 'rnn_fwd.rnn.weight_hh_l1' in keys and 'feature_extractor.norm.running_power' in keys

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    dst.cnn.load_state_dict(sub('cnn.'))
    dst.rnn_fwd.rnn.load_state_dict(sub('rnn_fwd.rnn.'))
    out_fwd, out_bwd = sub('rnn_fwd.output_net.'), sub('rnn_bwd.output_net.')
    param_keys = sorted(out_fwd.keys())
    layer_idx = [key.split('.')[1] for key in param_keys]
    last_layer_idx = layer_idx[-1]
    assert last_layer_idx == '1'                                     # the output layer, not a norm of layer 0
    for key, idx in zip(param_keys, layer_idx):
        if idx == last_layer_idx:
            out_fwd.pop(key)
            out_bwd.pop(key)
    dst.rnn_fwd.output_net.load_state_dict(out_fwd, strict=False)
    dst.rnn_bwd.output_net.load_state_dict(out_bwd, strict=False)
    assert torch.equal(dst.rnn_fwd.output_net.convs[0].conv.weight, src.rnn_fwd.output_net.convs[0].conv.weight)
    assert dst.rnn_fwd.output_net.convs[1].conv.weight.shape[1] == 7
    assert torch.equal(dst.cnn.cnn_2d.convs[2].norm.running_power, src.cnn.cnn_2d.convs[2].norm.running_power)


def test_native_conv_weight_layout_is_the_tap_contraction():
    """native (taps, Cout, Cin) weights + tap offsets == torch conv2d / conv1d / flatten-conv1d."""
    from pb_sed_b200.modules import _Conv
    torch.manual_seed(0)
    x = torch.randn(2, 3, 6, 7)                                     # b c f t
    c2 = _Conv(2, 3, 5, 3)
    w_ref = c2._to_ref(c2.weight.detach())
    y_ref = torch.nn.functional.conv2d(torch.nn.functional.pad(x, (1, 1, 1, 1)), w_ref)
    xn = x.permute(0, 2, 3, 1)
    y = torch.zeros(2, 6, 7, 5)
    for tap, (df, dt) in enumerate(c2.taps):
        xs = torch.zeros_like(xn)
        f0, f1 = max(0, -df), min(6, 6 - df)
        t0, t1 = max(0, -dt), min(7, 7 - dt)
        xs[:, f0:f1, t0:t1] = xn[:, f0 + df:f1 + df, t0 + dt:t1 + dt]
        y += xs @ c2.weight[tap].t()
    assert torch.allclose(y.permute(0, 3, 1, 2), y_ref, atol=1e-5)
    assert torch.equal(c2._from_ref(w_ref), c2.weight.detach())
    # flatten conv1d: reference sees (c f) features
    c1 = _Conv(1, 3, 4, 3, flatten_height=6)
    w_ref = c1._to_ref(c1.weight.detach())
    assert w_ref.shape == (4, 18, 3)
    y_ref = torch.nn.functional.conv1d(torch.nn.functional.pad(x.reshape(2, 18, 7), (1, 1)), w_ref)
    y = torch.zeros(2, 7, 4)
    for tap, (f, dt) in enumerate(c1.taps):
        xs = torch.zeros(2, 7, 3)
        t0, t1 = max(0, -dt), min(7, 7 - dt)
        xs[:, t0:t1] = xn[:, f, t0 + dt:t1 + dt]
        y += xs @ c1.weight[tap].t()
    assert torch.allclose(y.transpose(1, 2), y_ref, atol=1e-5)
    assert torch.equal(c1._from_ref(w_ref), c1.weight.detach())


def test_feature_tables_match_oracle():
    from pb_sed_b200 import modules as M
    fb = M.mel_filterbank(16000, 1024, 128)
    assert np.allclose(fb, P.get_fbanks(16000, 1024, 128), atol=1e-12)
    lo, hi, w, stride = M.sparse_filterbank(fb.astype(np.float32))
    dense = np.zeros_like(fb, dtype=np.float32)
    for m in range(128):
        dense[m, lo[m]:hi[m]] = w[m, :hi[m] - lo[m]]
    assert np.array_equal(dense, fb.astype(np.float32))
    assert np.allclose(M.blackman_window(960), P.blackman_periodic(960))
    for n in (160000, 645, 47, 1):
        assert M.stft_num_frames(n, 320, 960) == P.stft_frames(n, 320, 960)
    assert M.stft_num_frames(160000, 320, 960) == 500


def test_lr_schedule_matches_training_script():
    """breakpoints of pb_sed/experiments/weak_label_crnn/training.py:377-387 at batch size 32."""
    from pb_sed_b200.train import lr_schedule
    bp = [(0, 0.), (1000, 1.), (10000, 1.), (10000, .2)]
    assert lr_schedule(0, bp) == 0. and lr_schedule(500, bp) == 0.5
    assert lr_schedule(5000, bp) == 1. and lr_schedule(10001, bp) == .2


def test_reduce_and_mask_helpers_match_oracle():
    from pb_sed_b200 import modules as M
    x = torch.randn(3, 4, 9)
    sl = np.array([9, 5, 1])
    for name in ('Sum', 'Mean', 'TakeLast'):
        a = getattr(M, name)(axis=-1)(x, sl)
        b = getattr(P, name)(axis=-1)(x, sl)
        assert torch.equal(a, b), name
    assert torch.equal(M.Max(axis=-1)(x, sl)[0], P.Max(axis=-1)(x, sl)[0])
    assert torch.equal(M.compute_mask(x, sl, 0, -1), P.compute_mask(x, sl, 0, -1))
    assert torch.equal(M.Pad('both')(x, 5), P.Pad('both')(x, 5))


def test_collate_pads_sorts_and_keeps_lists():
    """fetcher.py:36-51 + Collate(): sorted by seq_len descending, zero padded, lists kept."""
    from pb_sed_b200.data import collate
    rng = np.random.RandomState(0)
    ex = []
    for i, (S, T) in enumerate([(3000, 10), (5000, 16), (4000, 13)]):
        ex.append({'example_id': f'e{i}', 'dataset': 'd', 'audio_data': rng.randn(1, S).astype(np.float32),
                   'stft': rng.randn(1, T, 5, 2).astype(np.float32), 'seq_len': T,
                   'weak_targets': np.eye(4, dtype=np.float32)[i], 'boundary_targets': rng.rand(4, T).astype(np.float32)})
    b = collate(ex)
    assert b['example_id'] == ['e1', 'e2', 'e0'] and b['seq_len'] == [16, 13, 10] and b['dataset'] == ['d'] * 3
    assert b['stft'].shape == (3, 1, 16, 5, 2) and b['audio_data'].shape == (3, 1, 5000)
    assert b['boundary_targets'].shape == (3, 4, 16) and b['weak_targets'].shape == (3, 4)
    assert float(b['stft'][2, :, 10:].abs().max()) == 0. and float(b['audio_data'][2, :, 3000:].abs().max()) == 0.
    assert np.array_equal(b['stft'][2, :, :10].numpy(), ex[0]['stft'])
    assert np.array_equal(b['weak_targets'].numpy(), np.eye(4, dtype=np.float32)[[1, 2, 0]])
    # audio-only examples: seq_len from the reference STFT geometry (provider.py:315-323), stft dropped
    b2 = collate([{k: v for k, v in e.items() if k not in ('stft', 'seq_len', 'boundary_targets')} for e in ex])
    from oracle import pt_port as P
    assert b2['seq_len'] == [P.stft_frames(5000), P.stft_frames(4000), P.stft_frames(3000)] and 'stft' not in b2
    assert 'stft' not in collate(ex, keep_stft=False)


def test_augmentation_samplers_follow_the_configured_distributions():
    """the device-side draws (here on CPU tensors) of NormalizedLogMelExtractor.sample_augmentation:
    LogTruncatedNormal(scale .08, truncation log 1.3), TruncatedExponential(scale .5, truncation 5),
    uniform mask widths / onsets, uniform noise scale (training.py:195-216) -- checked against scipy."""
    from scipy import stats
    from pb_sed_b200.modules import NormalizedLogMelExtractor

    class Seq:
        dev = None
    torch.manual_seed(0)
    fe = NormalizedLogMelExtractor(
        16000, 1024, 128, n_time_masks=2, max_masked_time_steps=70, max_masked_time_rate=.2,
        n_frequency_masks=1, max_masked_frequency_bands=20, max_masked_frequency_rate=.2, max_noise_scale=.2,
        frequency_warping_fn={'warp_factor_sampling_fn': {'scale': .08, 'truncation': np.log(1.3)},
                              'boundary_frequency_ratio_sampling_fn': {'scale': .5, 'truncation': 5.},
                              'highest_frequency': 8000.})
    n = 20000
    aug = fe.sample_augmentation(n, 8, Seq, 'cpu', False)          # T = 8 keeps the noise tensor small
    la = np.log(aug['alpha'].numpy().astype(np.float64))
    a = np.log(1.3) / .08
    assert np.abs(la).max() <= np.log(1.3) + 1e-6
    assert stats.kstest(la, stats.truncnorm(-a, a, scale=.08).cdf).pvalue > 1e-3
    r = aug['ratio'].numpy().astype(np.float64)
    assert r.min() >= 0 and r.max() <= 5.
    assert stats.kstest(r, stats.truncexpon(b=5. / .5, scale=.5).cdf).pvalue > 1e-3
    ns = aug['noise_scale'].numpy()
    assert stats.kstest(ns, stats.uniform(0, .2).cdf).pvalue > 1e-3
    # masks on a 500-frame clip: width ~ U{0..70}, onset ~ U{0..500-width}
    lens = torch.full((n,), 500.)
    m = NormalizedLogMelExtractor._masks(1, lens, 70, .2, 'cpu').numpy()[:, 0]
    w, on = m[:, 1], m[:, 0]
    assert w.min() == 0 and w.max() == 70 and abs(w.mean() - 35.) < 1. and (on + w <= 500).all() and on.min() >= 0
    counts = np.bincount(w, minlength=71)
    assert stats.chisquare(counts).pvalue > 1e-3
    short = NormalizedLogMelExtractor._masks(1, torch.full((n,), 40.), 70, .2, 'cpu').numpy()[:, 0]
    assert short[:, 1].max() == 8 and (short.sum(-1) <= 40).all()      # min(70, floor(.2 * 40)) = 8


def test_bench_counts_only_valid_taps_as_algorithmic_flops():
    """bench.py's roofline numerator (round-1 review: the transposed flatten data gradient F 1 -> 8 was credited all 8
    taps per output row although exactly ONE source row exists): 2 * Cin * Cout per VALID (output row, tap) pair."""
    import ctypes
    import bench
    from pb_sed_b200 import ops
    taps3 = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)]
    d = ops.make_desc(2, 4, 4, 10, 16, 32, taps3, precision=1)
    pairs = (3 + 4 + 3) * (9 + 10 + 9)                       # rows with a source row x frames with a source frame
    assert bench.tapgemm_flops((ctypes.byref(d),)) == 2. * 2 * pairs * 16 * 32
    flat = ops.make_desc(2, 1, 8, 10, 256, 256, [(-f, 0) for f in range(8)], precision=1)
    assert bench.tapgemm_flops((ctypes.byref(flat),)) == 2. * 2 * 8 * 10 * 256 * 256      # one tap per output row


def test_traffic_table_uses_the_kernel_names_the_library_reports():
    """bench.py fills ``roofline.traffic`` by looking the dominant kernel's ``pbsed_last_kernel()`` name up in
    profiles/traffic.json (measured DRAM bytes per launch, tools/ncu_summary.py traffic): the tensor-core kernels that
    can dominate a step must be present under exactly the names the sources pass to ``pbsed_note_kernel``."""
    import glob
    import json
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    noted = set()
    for path in glob.glob(os.path.join(root, 'pb_sed_b200', 'csrc', '*.cu')):
        noted.update(re.findall(r'"((?:tapgemm|wgrad|conv_cin1)[a-z0-9_]*(?:<[0-9,]+>)?)"', open(path).read()))
    table = json.load(open(os.path.join(root, 'profiles', 'traffic.json')))
    for name in ('wgrad_tma_kernel', 'tapgemm_tc_kernel<2,2,256>', 'tapgemm_tc_kernel<2,4,256>', 'tapgemm_tc_kernel<1,4,256>',
                 'tapgemm_fw_kernel'):
        assert name in noted, name
        assert table[name]['bytes_per_launch'] > 1e6 and table[name]['launches_captured'] >= 1, name
