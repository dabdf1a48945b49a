"""tiny driver for profiling the persistent GRU kernels in isolation (used with ncu)."""
import sys
sys.path.insert(0, '.')
import torch
from pb_sed_b200 import modules as M, ops
torch.manual_seed(0)
B, T, H = 32, 500, 256
a = M.GRUCore(256, H, 2).cuda()
b = M.GRUCore(256, H, 2).cuda()
x = torch.randn(B, T, 256, device='cuda', requires_grad=True)
seq = ops.SeqLen.make(None, B, T, x.device)
for _ in range(2):
    hf, hb = M.gru_stack([a, b], [x, x], seq, [False, True])
    (hf.sum() + hb.sum()).backward()
torch.cuda.synchronize()
