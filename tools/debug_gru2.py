"""localise the tf32x3-vs-fp32 discrepancy of the B = 64 GRU test: same model / data in both precisions on the
GPU, compare every gradient; several seeds.  (debug helper, GPU)"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch
from oracle import pt_port as P
from pb_sed_b200 import modules as M, ops

def build(B, In, bidir, seed, T=120, H=256):
    torch.manual_seed(seed)
    out_kw = dict(out_channels=[256, 10], kernel_size=1, norm='batch', norm_kwargs={'eps': 1e-3})
    gru = torch.nn.GRU(In, H, num_layers=2, batch_first=True, bidirectional=bidir)
    ora = P.GRU(gru, P.CNN1d(H * (2 if bidir else 1), **out_kw, pre_activation=False, output_layer=True), reverse=not bidir).train()
    sl = np.array(sorted([T] * 7 + [int(v) for v in np.linspace(T - 1, 3, B - 7)], reverse=True))
    x = torch.randn(B, In, T)
    g = torch.randn(B, 10, T) * P.compute_mask(torch.zeros(B, 10, T), sl, 0, -1)
    return ora, sl, x, g, out_kw

def run_gpu(ora, sl, x, g, out_kw, In, bidir, prec, H=256):
    ops.set_default_precision(prec)
    prod = M.GRU(dict(input_size=In, hidden_size=H, num_layers=2, bidirectional=bidir), out_kw, reverse=not bidir)
    prod.load_state_dict(ora.state_dict())
    prod.to('cuda').train()
    xg = x.to('cuda').requires_grad_(True)
    y, _ = prod(xg, sl)
    y.backward(g.to('cuda'))
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in prod.named_parameters()}
    return y.detach(), xg.grad.detach(), grads

for (B, In, bidir, seed) in [(64, 266, False, 64), (64, 256, False, 1), (64, 256, False, 2), (64, 256, False, 3), (32, 256, False, 4), (64, 266, False, 5)]:
    ora, sl, x, g, out_kw = build(B, In, bidir, seed)
    y0, dx0, g0 = run_gpu(ora, sl, x, g, out_kw, In, bidir, 'fp32')
    y1, dx1, g1 = run_gpu(ora, sl, x, g, out_kw, In, bidir, 'tf32x3')
    d = (dx1 - dx0).abs()
    per_b = d.amax((1, 2))
    print(f'B={B} In={In} seed={seed}: y {float((y1 - y0).abs().max()):.2e}  dx {float(d.max()):.2e}  bad clips {[int(i) for i in torch.nonzero(per_b > 1e-3).flatten()]}')
    for n in g0:
        e = float((g1[n] - g0[n]).abs().max()) / max(1e-6, float(g0[n].abs().max()))
        if e > 2e-4:
            print(f'    {n}: rel {e:.2e}')
