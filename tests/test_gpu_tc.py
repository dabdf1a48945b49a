"""GPU (-m gpu): the tcgen05 (3xTF32) tap-GEMM against the exact-fp32 FFMA kernel on identical inputs.
Split-TF32 keeps ~21 mantissa bits per product, so the two must agree to ~1e-5 relative."""
import numpy as np
import pytest
import torch

from util import maxdiff, reldiff

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

TAPS_3x3 = [(i - 1, j - 1) for i in range(3) for j in range(3)]
TAPS_1x3 = [(0, -1), (0, 0), (0, 1)]
TAPS_1x1 = [(0, 0)]


def run(prec, x, W, bias, dims, taps, scale=None, shift=None, relu=False, per_f=False, seq=None,
        transpose_w=False, ep=None):
    from pb_sed_b200 import ops
    B, F_in, F_out, T, Cin, Cout = dims
    desc = ops.make_desc(B, F_in, F_out, T, Cin, Cout, taps, relu=relu, per_f=per_f,
                         transpose_w=transpose_w, precision=prec)
    kw = {}
    if ep is not None:
        kw = dict(ep_src=ep[0], ep_scale=ep[1], ep_shift=ep[2])
    return ops.tapgemm(x, W, bias, desc, scale, shift, seq, **kw)


@pytest.mark.parametrize('B,F,T,Cin,Cout,taps', [
    (1, 1, 128, 16, 16, TAPS_1x1),
    (2, 3, 37, 16, 16, TAPS_3x3),
    (2, 4, 500, 32, 64, TAPS_3x3),
    (1, 2, 600, 64, 128, TAPS_3x3),
    (1, 2, 260, 128, 256, TAPS_3x3),
    (3, 1, 500, 256, 768, TAPS_1x1),
    (2, 1, 300, 256, 256, TAPS_1x3),
])
def test_tc_forward_matches_ffma(built_lib, B, F, T, Cin, Cout, taps):
    from pb_sed_b200 import ops
    torch.manual_seed(Cin + Cout + T)
    x = torch.randn(B, F, T, Cin, device=DEV)
    W = torch.randn(len(taps), Cout, Cin, device=DEV) / np.sqrt(Cin * len(taps))
    bias = torch.randn(Cout, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    sl = np.minimum(np.array([T, max(T - 3, 1), max(T // 2, 1)][:B]), T)
    seq = ops.SeqLen.make(sl, B, T, DEV)
    dims = (B, F, F, T, Cin, Cout)
    for kw in (dict(), dict(scale=scale, shift=shift, relu=True, seq=seq)):
        ref = run(0, x, W, bias, dims, taps, **kw)
        out = run(1, x, W, bias, dims, taps, **kw)
        assert torch.isfinite(out).all()
        assert reldiff(out, ref) < 2e-5, kw.keys()


def test_tc_flatten_conv_per_f_affine(built_lib):
    """first CNN1d layer: 8 frequency taps over the (B,8,T,256) map, batch norm indexed per (f, c)."""
    torch.manual_seed(0)
    B, Fh, T, Cin, Cout = 2, 8, 200, 64, 128
    taps = [(f, 0) for f in range(Fh)]
    x = torch.randn(B, Fh, T, Cin, device=DEV)
    W = torch.randn(len(taps), Cout, Cin, device=DEV) / np.sqrt(Cin * Fh)
    scale = torch.rand(Fh * Cin, device=DEV) + .5
    shift = torch.randn(Fh * Cin, device=DEV) * .3
    dims = (B, Fh, 1, T, Cin, Cout)
    ref = run(0, x, W, None, dims, taps, scale=scale, shift=shift, relu=True, per_f=True)
    out = run(1, x, W, None, dims, taps, scale=scale, shift=shift, relu=True, per_f=True)
    assert reldiff(out, ref) < 2e-5


def test_tc_dgrad_with_relu_mask_epilogue(built_lib):
    from pb_sed_b200 import ops
    torch.manual_seed(1)
    B, F, T, Cin, Cout = 2, 3, 300, 32, 64            # forward layer Cin -> Cout
    dz = torch.randn(B, F, T, Cout, device=DEV)
    W = torch.randn(9, Cout, Cin, device=DEV) / np.sqrt(Cin * 9)
    xprev = torch.randn(B, F, T, Cin, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    seq = ops.SeqLen.make(np.array([T, T - 40]), B, T, DEV)
    rtaps = [(-a, -b) for a, b in TAPS_3x3]
    dims = (B, F, F, T, Cout, Cin)
    ref = run(0, dz, W, None, dims, rtaps, seq=seq, transpose_w=True, ep=(xprev, scale, shift))
    out = run(1, dz, W, None, dims, rtaps, seq=seq, transpose_w=True, ep=(xprev, scale, shift))
    assert reldiff(out, ref) < 2e-5
    assert float((out == 0).float().mean()) > .3      # the ReLU mask really zeroed entries


@pytest.mark.parametrize('B,F,T,Cin,Cout,taps,relu', [
    (2, 3, 37, 16, 16, TAPS_3x3, True),
    (3, 13, 45, 16, 16, TAPS_3x3, False),      # narrow layers: row-stacked tiles with a ragged last row group
    (2, 7, 100, 16, 32, TAPS_3x3, True),
    (3, 5, 70, 32, 32, TAPS_3x3, True),
    (2, 5, 90, 64, 64, TAPS_3x3, True),        # two dout rows x two input rows per tile, two CTA groups
    (2, 4, 500, 32, 64, TAPS_3x3, True),
    (2, 2, 300, 128, 256, TAPS_3x3, True),
    (3, 1, 500, 256, 768, TAPS_1x1, False),
    (2, 1, 300, 256, 256, TAPS_1x3, True),
    (2, 1, 200, 256, 128, [(0, -1)], False),
])
def test_tc_wgrad_matches_ffma(built_lib, B, F, T, Cin, Cout, taps, relu):
    from pb_sed_b200 import ops
    torch.manual_seed(Cin + Cout + T)
    x = torch.randn(B, F, T, Cin, device=DEV)
    dz = torch.randn(B, F, T, Cout, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    sl = np.minimum(np.array([T, max(T - 3, 1), max(T // 2, 1)][:B]), T)
    seq = ops.SeqLen.make(sl, B, T, DEV)
    res = []
    for prec in (0, 1):
        desc = ops.make_desc(B, F, F, T, Cin, Cout, taps, relu=relu, precision=prec)
        dW = torch.zeros(len(taps), Cout, Cin, device=DEV)
        db = torch.zeros(Cout, device=DEV)
        ops.tapgemm_wgrad(x, dz, desc, dW, db, scale if relu else None, shift if relu else None, seq, mask_out=True)
        res.append((dW, db))
    assert reldiff(res[1][0], res[0][0]) < 5e-5
    assert reldiff(res[1][1], res[0][1]) < 5e-5
    assert float(res[0][0].abs().max()) > 0


@pytest.mark.parametrize('B,F,T,Cin,Cout,relu', [
    (2, 3, 37, 16, 16, True), (3, 13, 45, 16, 16, False), (2, 7, 100, 16, 32, True), (3, 5, 70, 32, 32, True),
    (2, 4, 500, 32, 64, True),
])
def test_narrow_wgrad_fallback_walk_kernel_matches_ffma(built_lib, monkeypatch, B, F, T, Cin, Cout, relu):
    """the frequency-walking mma.sync kernel behind the row-stacked tcgen05 tiles (PBSED_WG_STACK=0, read per call):
    same ragged cases, same bar against the exact-fp32 kernels; the last kernel name proves which one ran."""
    from pb_sed_b200 import ops, _lib
    monkeypatch.setenv('PBSED_WG_STACK', '0')
    torch.manual_seed(Cin + Cout + T)
    x = torch.randn(B, F, T, Cin, device=DEV)
    dz = torch.randn(B, F, T, Cout, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    sl = np.minimum(np.array([T, max(T - 3, 1), max(T // 2, 1)][:B]), T)
    seq = ops.SeqLen.make(sl, B, T, DEV)
    res = []
    for prec in (0, 1):
        desc = ops.make_desc(B, F, F, T, Cin, Cout, TAPS_3x3, relu=relu, precision=prec)
        dW = torch.zeros(9, Cout, Cin, device=DEV)
        db = torch.zeros(Cout, device=DEV)
        ops.tapgemm_wgrad(x, dz, desc, dW, db, scale if relu else None, shift if relu else None, seq, mask_out=True)
        res.append((dW, db))
    assert _lib.load().pbsed_last_kernel().decode() == 'wgrad_walk_kernel'
    assert reldiff(res[1][0], res[0][0]) < 5e-5
    assert reldiff(res[1][1], res[0][1]) < 5e-5


def test_tc_wgrad_flatten_and_strided_input(built_lib):
    """per-(f,c) affine with 8 frequency taps; and the GRU W_hh gradient reading one direction's half of a
    (B,T,2H) map (in_stride = 2H)."""
    from pb_sed_b200 import ops
    import ctypes
    torch.manual_seed(3)
    B, Fh, T, Cin, Cout = 2, 8, 200, 64, 128
    taps = [(f, 0) for f in range(Fh)]
    x = torch.randn(B, Fh, T, Cin, device=DEV)
    dz = torch.randn(B, 1, T, Cout, device=DEV)
    scale = torch.rand(Fh * Cin, device=DEV) + .5
    shift = torch.randn(Fh * Cin, device=DEV) * .3
    res = []
    for prec in (0, 1):
        desc = ops.make_desc(B, Fh, 1, T, Cin, Cout, taps, relu=True, per_f=True, precision=prec)
        dW = torch.zeros(len(taps), Cout, Cin, device=DEV)
        ops.tapgemm_wgrad(x, dz, desc, dW, None, scale, shift, None, mask_out=False)
        res.append(dW)
    assert reldiff(res[1], res[0]) < 5e-5
    H = 64
    h = torch.randn(B, T, 2 * H, device=DEV)
    dgh = torch.randn(B, T, 3 * H, device=DEV)
    seq = ops.SeqLen.make(np.array([T, T - 50]), B, T, DEV)
    res = []
    for prec in (0, 1):
        desc = ops.make_desc(B, 1, 1, T, H, 3 * H, [(0, 1)], in_stride=2 * H, precision=prec)
        dW = torch.zeros(1, 3 * H, H, device=DEV)
        db = torch.zeros(3 * H, device=DEV)
        ops.tapgemm_wgrad(None, dgh, desc, dW, db, None, None, seq, mask_out=True,
                          x_ptr=ctypes.c_void_p(h.data_ptr() + 4 * H))
        res.append((dW, db))
    assert reldiff(res[1][0], res[0][0]) < 5e-5 and reldiff(res[1][1], res[0][1]) < 5e-5


# ------------------------------------------------------------------ precision 3: one TF32 pass
@pytest.mark.parametrize('B,F,T,Cin,Cout,taps', [
    (2, 4, 500, 32, 64, TAPS_3x3),
    (1, 2, 260, 128, 256, TAPS_3x3),
    (3, 1, 500, 256, 768, TAPS_1x1),
])
def test_single_pass_tf32_forward_and_wgrad(built_lib, B, F, T, Cin, Cout, taps):
    """reduced-precision mode (BASELINE 'bf16' configurations): operands rounded to nearest TF32
    (unit round-off 2^-11), fp32 accumulation.  Tolerance: 3e-3 of the output range -- and it must be
    measurably different from the fp32-equivalent split, i.e. the single pass is really what ran."""
    from pb_sed_b200 import ops
    torch.manual_seed(Cin + T)
    x = torch.randn(B, F, T, Cin, device=DEV)
    W = torch.randn(len(taps), Cout, Cin, device=DEV) / np.sqrt(Cin * len(taps))
    bias = torch.randn(Cout, device=DEV)
    scale = torch.rand(Cin, device=DEV) + .5
    shift = torch.randn(Cin, device=DEV) * .3
    dims = (B, F, F, T, Cin, Cout)
    ref = run(0, x, W, bias, dims, taps, scale=scale, shift=shift, relu=True)
    out = run(3, x, W, bias, dims, taps, scale=scale, shift=shift, relu=True)
    err = reldiff(out, ref)
    assert 2e-5 < err < 3e-3, err
    dz = torch.randn(B, F, T, Cout, device=DEV)
    res = []
    for prec in (0, 3):
        desc = ops.make_desc(B, F, F, T, Cin, Cout, taps, relu=True, precision=prec)
        dW = torch.zeros(len(taps), Cout, Cin, device=DEV)
        db = torch.zeros(Cout, device=DEV)
        ops.tapgemm_wgrad(x, dz, desc, dW, db, scale, shift, None, mask_out=False)
        res.append((dW, db))
    # (narrow layers route the weight gradient to the exact FFMA kernel in every mode: no lower bound here)
    assert reldiff(res[1][0], res[0][0]) < 3e-3
    assert reldiff(res[1][1], res[0][1]) < 5e-5          # the bias gradient is summed in fp32 either way
