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


def _sync_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pb_sed_b200 import ops, train
    ops.set_sync_stats('exact')
    assert ops.sync_stats_on()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 3, 7, 4, generator=g, dtype=torch.float64)        # (B, F, T, C) rows x channels
    lens = [7, 7, 5, 4, 3, 2]
    lo, hi = train.shard_bounds(6, rank, world)
    nch = 4
    stats = ops._stats_buffer(nch, 'cpu')
    assert stats.shape == (nch + 1, 2)
    count = 0.
    for b in range(lo, hi):
        v = x[b, :, :lens[b]].reshape(-1, nch)
        stats[:nch, 0] += v.sum(0)
        stats[:nch, 1] += (v * v).sum(0)
        count += v.shape[0]
    assert ops._sync_count_(stats, nch, count) == 0.0                      # kernel reads the count on device
    # loss: replica value N_r / D_r with D_r weights -> global loss and the replica's gradient scale
    N, D = [3.0, 1.0][rank], [4.0, 1.0][rank]
    loss, gscale = ops._sync_loss(torch.tensor([N / D, D], dtype=torch.float64))
    torch.save({'stats': stats, 'loss': loss, 'gscale': gscale}, os.path.join(out_dir, f's{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()


def test_exact_statistics_allreduce_world2(tmp_path):
    """SURVEY 8e 'exact' mode host logic: (sum, sum of squares, count) of every replica in one buffer;
    loss numerator / normaliser all-reduced, gradient scale D_r * world / D_global."""
    world = 2
    mp.spawn(_sync_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f's{i}.pt') for i in range(world)]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 3, 7, 4, generator=g, dtype=torch.float64)
    lens = [7, 7, 5, 4, 3, 2]
    rows = torch.cat([x[b, :, :lens[b]].reshape(-1, 4) for b in range(6)])
    for s in r:
        assert torch.allclose(s['stats'][:4, 0], rows.sum(0)) and torch.allclose(s['stats'][:4, 1], (rows * rows).sum(0))
        assert float(s['stats'][4, 0]) == rows.shape[0]
        assert abs(float(s['loss']) - (3.0 + 1.0) / (4.0 + 1.0)) < 1e-12
    assert abs(float(r[0]['gscale']) - 4.0 * 2 / 5.0) < 1e-12 and abs(float(r[1]['gscale']) - 1.0 * 2 / 5.0) < 1e-12
