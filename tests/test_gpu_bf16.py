"""GPU (-m gpu): the 'bf16' mode (BASELINE.json configs[2] / [4]): conv-stack activation maps and their gradients are
stored in HBM as bf16, every kernel on the path reads / writes them directly, products are bf16 activations x
TF32-rounded weights with fp32 accumulation.  Kernel-level tests feed bf16-rounded inputs to the bf16-I/O kernels and
to the exact-fp32 kernels and allow one bf16 rounding of the result (2^-8 relative); the model-level test states
the tolerance of the whole mode against the fp32 CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import models as OM, pt_port as P
from util import maxdiff, reldiff

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
BF = torch.bfloat16
TAPS_3x3 = [(i - 1, j - 1) for i in range(3) for j in range(3)]


@pytest.fixture(scope='module', autouse=True)
def _lib_loaded(built_lib):
    assert torch.cuda.is_available()
    yield


@pytest.mark.parametrize('B,F,T,Cin,Cout,taps', [
    (4, 16, 500, 16, 16, TAPS_3x3),           # frequency-walking kernel
    (4, 8, 300, 32, 32, TAPS_3x3),
    (3, 4, 500, 64, 128, TAPS_3x3),           # generic tcgen05 kernel
    (2, 1, 500, 256, 256, [(0, -1), (0, 0), (0, 1)]),
])
@pytest.mark.parametrize('io', [(BF, BF), (torch.float32, BF), (BF, torch.float32)])
def test_bf16_maps_forward_dgrad_wgrad_vs_fp32_kernels(B, F, T, Cin, Cout, taps, io):
    from pb_sed_b200 import ops
    idt, odt = io
    torch.manual_seed(Cin + Cout)
    x = torch.randn(B, F, T, Cin, device=DEV).to(idt)
    W = torch.randn(len(taps), Cout, Cin, device=DEV) / np.sqrt(Cin * len(taps))
    bias = torch.randn(Cout, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    sl = np.array(sorted([T] + [int(v) for v in np.linspace(T - 1, 77, B - 1)], reverse=True))
    seq = ops.SeqLen.make(sl, B, T, DEV)
    dims = (B, F, F, T, Cin, Cout)
    # forward + fused statistics
    ref_desc = ops.make_desc(*dims, taps, relu=True, precision=0)
    ref = ops.tapgemm(x.float(), W, bias, ref_desc, scale, shift, seq)
    desc = ops.make_desc(*dims, taps, relu=True, precision=3, in_dtype=ops._dt(idt), out_dtype=ops._dt(odt))
    stats = torch.zeros(Cout, 2, device=DEV, dtype=torch.float64)
    out = ops.tapgemm(x, W, bias, desc, scale, shift, seq, out_stats=stats)
    assert out.dtype == odt
    # bf16 activations x bf16 weights (kind::f16, where the input map is bf16 and Cin % 32 == 0; TF32-rounded weights
    # otherwise) + one bf16 rounding of the output
    assert reldiff(out.float(), ref) < 2e-2
    st_ref = torch.zeros(Cout, 2, device=DEV, dtype=torch.float64)
    ops.call('pbsed_channel_stats', ops._ptr(ref), B, F, T, Cout, 0, seq.ptr, ops._ptr(st_ref), 0, ops._stream())
    assert reldiff(stats, st_ref) < 5e-3                  # statistics come from the fp32 accumulators
    # data gradient with ReLU-mask epilogue: dz in `odt`, x (mask source) and g in `idt`
    dz = torch.randn(B, F, T, Cout, device=DEV).to(odt)
    rtaps = [(-a, -b) for a, b in taps]
    rdesc = ops.make_desc(B, F, F, T, Cout, Cin, rtaps, transpose_w=True, precision=0)
    g_ref = ops.tapgemm(dz.float(), W, None, rdesc, None, None, seq, ep_src=x.float(), ep_scale=scale, ep_shift=shift)
    ddesc = ops.make_desc(B, F, F, T, Cout, Cin, rtaps, transpose_w=True, precision=3, in_dtype=ops._dt(odt),
                          out_dtype=ops._dt(idt))
    g = ops.tapgemm(dz, W, None, ddesc, None, None, seq, ep_src=x, ep_scale=scale, ep_shift=shift)
    assert g.dtype == idt
    assert reldiff(g.float(), g_ref) < 2e-2
    # weight gradient
    res = []
    for prec, xx, zz, kw in ((0, x.float(), dz.float(), {}), (3, x, dz, dict(in_dtype=ops._dt(idt), out_dtype=ops._dt(odt)))):
        d = ops.make_desc(*dims, taps, relu=True, precision=prec, **kw)
        dW = torch.zeros(len(taps), Cout, Cin, device=DEV)
        db = torch.zeros(Cout, device=DEV)
        ops.tapgemm_wgrad(xx, zz, d, dW, db, scale, shift, seq, mask_out=True)
        res.append((dW, db))
    assert reldiff(res[1][0], res[0][0]) < 3e-3
    assert reldiff(res[1][1], res[0][1]) < 1e-4


def test_bf16_maps_first_layer_pool_and_norm_backward():
    """first conv layer writing a bf16 map (fp32 log-mel in), frequency max-pool (+ its backward) and the batch-norm
    backward apply on bf16 maps, each against the fp32 kernel on the same (rounded) data."""
    from pb_sed_b200 import ops
    torch.manual_seed(3)
    B, F, T, C = 3, 8, 500, 16
    x = torch.randn(B, F, T, 1, device=DEV)
    W = torch.randn(9, C, 1, device=DEV) / 3.
    bias = torch.randn(C, device=DEV)
    ref = ops.tapgemm(x, W, bias, ops.make_desc(B, F, F, T, 1, C, TAPS_3x3, precision=0, no_input_mask=True))
    out = ops.tapgemm(x, W, bias, ops.make_desc(B, F, F, T, 1, C, TAPS_3x3, precision=3, no_input_mask=True, out_dtype=1))
    assert out.dtype == BF and reldiff(out.float(), ref) < 5e-3
    dzb = torch.randn(B, F, T, C, device=DEV).to(BF)
    res = []
    for zz, od in ((dzb.float(), 0), (dzb, 1)):
        dW = torch.zeros(9, C, 1, device=DEV)
        db = torch.zeros(C, device=DEV)
        ops.tapgemm_wgrad(x, zz, ops.make_desc(B, F, F, T, 1, C, TAPS_3x3, precision=3, out_dtype=od), dW, db, mask_out=False)
        res.append((dW, db))
    assert reldiff(res[1][0], res[0][0]) < 1e-5 and reldiff(res[1][1], res[0][1]) < 1e-5
    # pool
    seq = ops.SeqLen.make(np.array([T, T - 9, 200]), B, T, DEV)
    z = out.view(B, F, T, C)
    outs = []
    for zz in (z.float(), z):
        dt = ops._dt(zz)
        y = torch.empty((B, F // 2, T, C), device=DEV, dtype=zz.dtype)
        idx = torch.empty(y.shape, device=DEV, dtype=torch.uint8)
        st = torch.zeros(C, 2, device=DEV, dtype=torch.float64)
        ops.call('pbsed_maxpool_f', ops._ptr(zz), B, F, T, C, 2, ops._ptr(y), ops._ptr(idx), seq.ptr, ops._ptr(st), dt, dt, ops._stream())
        dy = torch.randn(B, F // 2, T, C, device=DEV).to(BF).to(zz.dtype) if not outs else outs[0][3].to(zz.dtype)
        dx = torch.empty((B, F, T, C), device=DEV, dtype=zz.dtype)
        ops.call('pbsed_maxpool_f_bwd', ops._ptr(dy), ops._ptr(idx), B, F, T, C, 2, ops._ptr(dx), dt, dt, ops._stream())
        outs.append((y, idx, st, dy, dx))
    assert torch.equal(outs[0][0], outs[1][0].float()) and torch.equal(outs[0][1], outs[1][1])
    assert reldiff(outs[1][2], outs[0][2]) < 1e-6
    assert torch.equal(outs[0][4], outs[1][4].float())
    # batch-norm backward: reduce + apply
    g = torch.randn(B, F, T, C, device=DEV).to(BF)
    xx = torch.randn(B, F, T, C, device=DEV).to(BF)
    mean, rstd, gamma = torch.randn(C, device=DEV) * .1, torch.rand(C, device=DEV) + .5, torch.rand(C, device=DEV) + .5
    res = []
    for gg, xq in ((g.float(), xx.float()), (g, xx)):
        dt = ops._dt(gg)
        sums = torch.zeros(C, 2, device=DEV, dtype=torch.float64)
        ops.call('pbsed_norm_bwd_reduce', ops._ptr(gg), ops._ptr(xq), B, F, T, C, 0, seq.ptr, ops._ptr(mean), ops._ptr(rstd),
                 ops._ptr(sums), dt, ops._stream())
        dx = torch.empty_like(gg)
        dga, dbe = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
        ops.call('pbsed_norm_bwd_apply', ops._ptr(gg), ops._ptr(xq), B, F, T, C, 0, seq.ptr, ops._ptr(mean), ops._ptr(rstd),
                 ops._ptr(gamma), ops._ptr(sums), float(seq.frames() * F), ops._ptr(dx), ops._ptr(dga), ops._ptr(dbe), dt,
                 ops._stream())
        res.append((sums, dx.float(), dga))
    assert reldiff(res[1][0], res[0][0]) < 1e-9
    assert reldiff(res[1][1], res[0][1]) < 5e-3 and reldiff(res[1][2], res[0][2]) < 1e-5


def test_full_size_fbcrnn_bf16_mode_stated_tolerance():
    """the whole 'bf16' mode on the full-size FBCRNN (B = 4): STATED tolerance against the fp32 CPU oracle --
    frame-logit max|delta| <= 0.5 (|logit| ~ 6-8; measured ~0.1-0.3), loss within 5 %, gradient norm within 10 % --
    and it must really be reduced precision (logit delta > 1e-3) with bf16 maps inside the stack."""
    from pb_sed_b200 import config, train, ops
    from pb_sed_b200.models import weak_label
    ora = OM.build_fbcrnn(seed=0)
    model = weak_label.CRNN.from_config_dict(config.fbcrnn_config())
    model.load_state_dict(ora.state_dict())
    model.to(DEV)
    model.emit_buffers = False
    batch = OM.synthetic_batch(4, seed=13, seq_len=[500, 500, 431, 277])
    gb = {k: (v.to(DEV) if torch.is_tensor(v) else v) for k, v in batch.items() if k != 'stft'}
    opt = train.Adam(model, lr=5e-4)
    model.train()
    seen = []
    real = ops.tapgemm

    def spy(x, W, bias, desc, *a, **kw):
        seen.append((desc.in_dtype, desc.out_dtype))
        return real(x, W, bias, desc, *a, **kw)
    ops.set_default_precision('bf16')
    ops.tapgemm = spy
    try:
        out = model(dict(gb))
        loss = model.review(gb, out)['loss']
        loss.backward()
        z_fwd = model._z_fwd.detach().cpu()
        gnorm = opt.step()
    finally:
        ops.tapgemm = real
        ops.set_default_precision('tf32x3')
    assert sum(1 for i, o in seen if o == 1) >= 12 and sum(1 for i, o in seen if i == 1) >= 12   # bf16 maps were used
    cb = {k: v for k, v in batch.items() if k != 'audio_data'}
    ora.train()
    zr_fwd, *_ = ora.logits(cb)
    ora2 = OM.build_fbcrnn(seed=0)
    ref_loss, ref_gnorm, _ = OM.train_step(ora2, OM.make_adam(ora2), cb)
    mask = P.compute_mask(zr_fwd, np.array(batch['seq_len']), 0, -1)
    d = maxdiff(z_fwd.transpose(1, 2) * mask, zr_fwd.detach() * mask)
    print(f'bf16 mode: logit max|d| {d:.3f}, loss {float(loss):.5f} vs {float(ref_loss):.5f}, '
          f'grad norm {float(gnorm):.5f} vs {float(ref_gnorm):.5f}')
    assert 1e-3 < d < 0.5, d
    assert abs(float(loss) - float(ref_loss)) < 5e-2 * float(ref_loss)
    assert abs(float(gnorm) - float(ref_gnorm)) < 1e-1 * float(ref_gnorm)
