"""CPU, world_size 2 over gloo: the data-parallel host logic (clip sharding, flat-bucket gradient
all-reduce, identical replicas after the update).  The per-rank gradients come from the oracle model
(CPU); the all-reduce is the product helper ``train.allreduce_flat`` the NCCL path uses."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import models as OM
    from pb_sed_b200 import train
    torch.set_num_threads(2)
    model = OM.tiny_fbcrnn(seed=0)
    batch = OM.synthetic_batch(4, num_samples=645, stft_kwargs=dict(shift=16, window_length=48, size=64), seed=3)
    lo, hi = train.shard_bounds(4, rank, world)
    shard = {k: (v[lo:hi] if torch.is_tensor(v) else v[lo:hi]) for k, v in batch.items()}
    model.train()
    out = model(dict(shard))
    model.review(shard, out)['loss'].backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    local = flat.clone()
    train.allreduce_flat(flat)
    flat *= 1. / world                                   # what hyper[6] does inside the Adam kernel
    torch.save({'local': local, 'reduced': flat, 'bounds': (lo, hi)}, os.path.join(out_dir, f'r{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f'r{i}.pt') for i in range(world)]
    assert [x['bounds'] for x in r] == [(0, 2), (2, 4)]
    assert torch.equal(r[0]['reduced'], r[1]['reduced'])                 # replicas stay identical
    mean = (r[0]['local'] + r[1]['local']) / 2
    assert torch.allclose(r[0]['reduced'], mean, atol=1e-7)
    assert float((r[0]['local'] - r[1]['local']).abs().max()) > 0.       # shards really differed


def test_shard_bounds_cover_everything():
    from pb_sed_b200.train import shard_bounds
    for n in (1, 7, 32, 1024):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
