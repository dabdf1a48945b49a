import sys, ctypes; sys.path.insert(0, '.')
import numpy as np, torch
from pb_sed_b200 import ops
DEV='cuda:0'
torch.manual_seed(0)
for (B,F,T,Cin,Cout,taps) in [(1,1,32,16,16,[(0,0)]), (1,1,32,128,128,[(0,0)]), (1,1,64,16,16,[(0,-1),(0,0),(0,1)])]:
    x = torch.randn(B,F,T,Cin,device=DEV); dz = torch.randn(B,F,T,Cout,device=DEV)
    out=[]
    for prec in (0,1):
        desc = ops.make_desc(B,F,F,T,Cin,Cout,taps,precision=prec)
        dW = torch.zeros(len(taps),Cout,Cin,device=DEV); db=torch.zeros(Cout,device=DEV)
        ops.tapgemm_wgrad(x,dz,desc,dW,db,None,None,None,mask_out=False)
        torch.cuda.synchronize()
        out.append((dW.cpu(),db.cpu()))
    print((B,F,T,Cin,Cout,len(taps)), 'ffma', out[0][0].flatten()[:6], 'tc', out[1][0].flatten()[:6], 'nonzero frac', float((out[1][0]!=0).float().mean()), 'bias', out[0][1][:4], out[1][1][:4])
    ref = torch.einsum('btn,btc->nc', dz[:,0].cpu(), x[:,0].cpu())
    print(' ref', ref.flatten()[:6], ' tc^T?', out[1][0][len(taps)//2].t().flatten()[:6])
