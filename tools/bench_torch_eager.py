"""Secondary comparator of SURVEY 8d ("the honest bar"): the oracle restatement of the reference train
step moved to ``cuda`` -- torch eager -> cuDNN / cuBLAS Blackwell kernels -- on the same B200, same
batch-32 FBCRNN workload (STFT precomputed on the host and excluded, like the reference feeds it).
Not part of bench.py's contract; results are recorded under profiles/.

    python tools/bench_torch_eager.py [--batch 32] [--steps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--steps', type=int, default=10)
    args = ap.parse_args()
    from oracle import models as OM, pt_port as P
    dev = torch.device('cuda:0')
    audio, weak, boundary = bench.synthetic_clips(args.batch, 1234)
    spec = P.stft(audio)
    stft = torch.from_numpy(np.stack([spec.real, spec.imag], -1).astype(np.float32)).to(dev)
    batch = dict(stft=stft, seq_len=[bench.T_FRAMES] * args.batch, weak_targets=torch.from_numpy(weak).to(dev),
                 boundary_targets=torch.from_numpy(boundary).to(dev))
    out = {}
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32          # torch default: True for cuDNN convs
        torch.backends.cuda.matmul.allow_tf32 = False   # torch default
        model = OM.build_fbcrnn(seed=0).to(dev)
        opt = OM.make_adam(model)
        for _ in range(3):
            OM.train_step(model, opt, batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            OM.train_step(model, opt, batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out['cudnn_tf32' if tf32 else 'cudnn_fp32'] = {'ms_per_step': ms, 'clips_per_s': args.batch / (ms / 1e3)}
    print(json.dumps({'workload': f'oracle FBCRNN train step on cuda (torch eager, cuDNN/cuBLAS), batch {args.batch}, '
                                  'STFT excluded', **out}))


if __name__ == '__main__':
    main()
