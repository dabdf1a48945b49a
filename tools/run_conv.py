"""driver for profiling single conv layers (forward, dgrad, wgrad) of the shallow FBCRNN at B = 32 under ncu."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from pb_sed_b200 import ops
TAPS = [(i - 1, j - 1) for i in range(3) for j in range(3)]
B, T = 32, 500
shapes = [(16, 128, 128), (8, 128, 256), (32, 64, 64), (128, 16, 16)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for F, Cin, Cout in shapes:
    x = torch.randn(B, F, T, Cin, device='cuda')
    W = torch.randn(9, Cout, Cin, device='cuda') / np.sqrt(9 * Cin)
    bias = torch.zeros(Cout, device='cuda')
    scale = torch.ones(Cin, device='cuda'); shift = torch.zeros(Cin, device='cuda')
    dz = torch.randn(B, F, T, Cout, device='cuda')
    seq = ops.SeqLen.make(None, B, T, x.device)
    for _ in range(2):
        d = ops.make_desc(B, F, F, T, Cin, Cout, TAPS, relu=True, precision=1)
        y = ops.tapgemm(x, W, bias, d, scale, shift, seq)
        dd = ops.make_desc(B, F, F, T, Cout, Cin, [(-a, -b) for a, b in TAPS], transpose_w=True, precision=1)
        g = ops.tapgemm(dz, W, None, dd, None, None, seq, ep_src=x, ep_scale=scale, ep_shift=shift)
        dW = torch.zeros_like(W); db = torch.zeros(Cout, device='cuda')
        ops.tapgemm_wgrad(x, dz, d, dW, db, scale, shift, seq, mask_out=False)
torch.cuda.synchronize()
