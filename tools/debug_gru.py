"""where does the bidirectional GRU input gradient differ from torch.nn.GRU?  (debug helper, GPU)"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import torch
from oracle import pt_port as P
from pb_sed_b200 import modules as M
from util import ref_layout_grads

def run(B, In, bidir, T=120, H=256, ragged=True):
    torch.manual_seed(B)
    out_kw = dict(out_channels=[256, 10], kernel_size=1, norm='batch', norm_kwargs={'eps': 1e-3})
    gru = torch.nn.GRU(In, H, num_layers=2, batch_first=True, bidirectional=bidir)
    ora = P.GRU(gru, P.CNN1d(H * (2 if bidir else 1), **out_kw, pre_activation=False, output_layer=True), reverse=not bidir).train()
    prod = M.GRU(dict(input_size=In, hidden_size=H, num_layers=2, bidirectional=bidir), out_kw, reverse=not bidir)
    prod.load_state_dict(ora.state_dict())
    prod.to('cuda').train()
    sl = np.array(sorted([T] * 7 + [int(v) for v in np.linspace(T - 1, 3, B - 7)], reverse=True)) if ragged else None
    x = torch.randn(B, In, T)
    xr = x.clone().requires_grad_(True)
    xg = x.to('cuda').requires_grad_(True)
    y_ref, _ = ora(xr, sl)
    y, _ = prod(xg, sl)
    mask = P.compute_mask(y_ref, sl, 0, -1)
    g = torch.randn_like(y_ref) * mask
    y_ref.backward(g)
    y.backward(g.to('cuda'))
    xmask = P.compute_mask(x, sl, 0, -1)
    d = ((xg.grad.cpu() - xr.grad) * xmask).abs()
    per_b = d.amax((1, 2))
    worst_b = int(per_b.argmax())
    per_t = d[worst_b].amax(0)
    print(f'B={B} In={In} bidir={bidir} ragged={ragged}: y diff {float(((y.cpu()-y_ref)*mask).abs().max()):.2e}  dx diff {float(d.max()):.2e} '
          f'(|dx| max {float(xr.grad.abs().max()):.2f}) worst clip {worst_b} len {None if sl is None else sl[worst_b]} worst t {int(per_t.argmax())}; '
          f'clips with diff>1e-3: {[int(i) for i in torch.nonzero(per_b > 1e-3).flatten()][:20]}')
    grads = ref_layout_grads(prod)
    for k, p in ora.named_parameters():
        e = float((grads[k] - p.grad).abs().max()) / max(1., float(p.grad.abs().max()))
        if e > 2e-4:
            print('   param', k, f'{e:.2e}')

import os
cases = {'small': [(64, 266, False)], 'all': [(64, 266, True), (64, 256, True), (64, 266, False), (64, 272, False)]}[os.environ.get('GRU_CASES', 'all')]
for args in cases:
    run(*args)
