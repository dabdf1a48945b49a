"""time the weight gradient of the four narrow 3x3 layers of the B = 32 step (kernel picked by the library / env)."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from pb_sed_b200 import ops, _lib
TAPS = [(a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)]
B, T = 32, 500
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bf = len(sys.argv) > 2 and sys.argv[2] == 'bf16'
out = []
for F, Cin, Cout in [(128, 16, 16), (64, 16, 32), (64, 32, 32), (32, 32, 64), (32, 64, 64), (16, 128, 128)]:
    torch.manual_seed(0)
    x = torch.randn(B, F, T, Cin, device='cuda')
    dz = torch.randn(B, F, T, Cout, device='cuda')
    dt = 1 if bf else 0
    if bf:
        x, dz = x.bfloat16(), dz.bfloat16()
    scale = torch.rand(Cin, device='cuda') + .5
    shift = torch.randn(Cin, device='cuda') * .3
    seq = ops.SeqLen.make(np.full(B, T), B, T, 'cuda')
    desc = ops.make_desc(B, F, F, T, Cin, Cout, TAPS, relu=True, precision=prec, in_dtype=dt, out_dtype=dt)
    dW = torch.zeros(9, Cout, Cin, device='cuda'); db = torch.zeros(Cout, device='cuda')
    for _ in range(3):
        ops.tapgemm_wgrad(x, dz, desc, dW, db, scale, shift, seq, mask_out=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.tapgemm_wgrad(x, dz, desc, dW, db, scale, shift, seq, mask_out=False)
    e1.record(); torch.cuda.synchronize()
    out.append('%dx%d>%d %s %.3f' % (F, Cin, Cout, _lib.load().pbsed_last_kernel().decode(), e0.elapsed_time(e1) / 10))
print(' | '.join(out))
