"""torchrun worker (one process per GPU, NCCL): data-parallel 'exact' statistics.

Every rank trains on its shard of ONE ragged batch with ``Adam(sync_stats='exact')``; rank 0 then
repeats the step single-process on the whole batch and compares loss, gradient norm, updated
parameters and running statistics.  Also covers the CUDA-graph path (NCCL all-reduces captured
inside the forward/backward graph).  Prints ``DP_EXACT_OK`` on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dp_exact_worker.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_batch(B):
    from oracle import models as OM
    T = 41
    seq_len = sorted([T - (3 * i) % 17 for i in range(B)], reverse=True)
    seq_len[0] = T
    batch = OM.synthetic_batch(B, num_samples=16 * 40 + 5, stft_kwargs=dict(shift=16, window_length=48, size=64),
                               seq_len=seq_len, seed=11)
    # 'unknown' weak labels (0.5) on some clips: the loss normaliser sum(m_w) then differs per replica
    w = batch['weak_targets']
    w[1, 2] = .5
    w[B - 1, 0] = .5
    w[B - 1, 7] = .5
    batch.pop('stft')
    return batch


def build(dev, sync, distributed, lr=5e-4):
    from pb_sed_b200 import config, train
    from pb_sed_b200.models import weak_label
    torch.manual_seed(0)
    model = weak_label.CRNN.from_config_dict(config.tiny_fbcrnn_config()).to(dev)
    model.emit_buffers = False
    opt = train.Adam(model, lr=lr, gradient_clipping=1e10, distributed=distributed, sync_stats=sync)
    return model, opt


def to_dev(batch, dev, lo=None, hi=None):
    out = {}
    for k, v in batch.items():
        v = v[lo:hi] if lo is not None else v
        out[k] = v.to(dev) if torch.is_tensor(v) else v
    return out


def state(model):
    return {k: v.detach().double().cpu() for k, v in model.state_dict().items()}


def main():
    from pb_sed_b200 import train, _lib
    _lib.load()
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    B = 4 * world
    batch = make_batch(B)
    lo, hi = train.shard_bounds(B, rank, world)

    # --- eager DP step, exact statistics
    model, opt = build(dev, 'exact', True)
    shard = to_dev(batch, dev, lo, hi)
    loss, gnorm = train.train_step(model, opt, shard)
    loss, gnorm, state1 = float(loss), float(gnorm), state(model)
    loss2, gnorm2 = train.train_step(model, opt, shard)          # second step: running stats / Adam moments
    torch.cuda.synchronize()
    dp = dict(loss=loss, gnorm=gnorm, loss2=float(loss2), gnorm2=float(gnorm2), state=state(model), state1=state1)

    # --- graphed DP step (all-reduces captured)
    model_g, opt_g = build(dev, 'exact', True)
    step = train.GraphedTrainStep(model_g, opt_g, shard, warmup=1)   # runs 1 eager step + capture (no replay yet)
    # the warm-up already advanced the model by one step; replay once more = 2 steps in total
    lg, gg = step(shard)
    torch.cuda.synchronize()
    graph = dict(loss2=float(lg), gnorm2=float(gg), state=state(model_g))

    # --- per-replica statistics: must differ from the single-process result (the check has teeth)
    model_n, opt_n = build(dev, 'none', True)
    loss_n, _ = train.train_step(model_n, opt_n, shard)
    torch.cuda.synchronize()
    none_loss = float(loss_n)

    ok = True
    if rank == 0:
        model_s, opt_s = build(dev, 'none', False)
        full = to_dev(batch, dev)
        l1, g1 = train.train_step(model_s, opt_s, full)
        l1, g1, ref1 = float(l1), float(g1), state(model_s)
        l2, g2 = train.train_step(model_s, opt_s, full)
        torch.cuda.synchronize()
        ref = state(model_s)

        def cmp(name, a, b, tol):
            nonlocal ok
            d = abs(a - b) / max(abs(b), 1e-12)
            good = d < tol
            ok &= good
            print(f'{name}: dp {a:.8f} single {b:.8f} rel {d:.2e} {"ok" if good else "FAIL"}')

        cmp('loss step1', dp['loss'], float(l1), 2e-5)
        cmp('grad-norm step1', dp['gnorm'], float(g1), 2e-4)
        cmp('loss step2', dp['loss2'], float(l2), 2e-4)
        cmp('grad-norm step2', dp['gnorm2'], float(g2), 1e-3)
        cmp('graph loss step2', graph['loss2'], float(l2), 2e-4)
        cmp('graph grad-norm step2', graph['gnorm2'], float(g2), 1e-3)
        for tag, st, ref in (('eager step 1', dp['state1'], ref1), ('eager step 2', dp['state'], ref),
                             ('graph step 2', graph['state'], ref)):
            # running statistics: plain averages -> tight relative bound.  Parameters: Adam's first steps
            # move every weight by ~lr * sign(g), and the conv biases in front of a batch-norm have an
            # analytically ZERO gradient (rounding noise only), so their sign is arbitrary: the bound is
            # absolute, 2 steps * 2 * lr.
            stat_keys = [k for k in ref if 'running_' in k or 'num_tracked' in k]
            par_keys = [k for k in ref if k not in stat_keys]
            rel = {k: float((st[k] - ref[k]).abs().max() / (ref[k].abs().max() + 1e-6)) for k in stat_keys}
            wk = max(rel, key=rel.get)
            w_stat = rel[wk]
            w_par = max(float((st[k] - ref[k]).abs().max()) for k in par_keys)
            # step 1 starts from identical parameters: the running statistics must agree tightly; after
            # step 2 the running means inherit the arbitrary-sign bias updates described above (a conv
            # bias moves its output mean one-to-one), so they are reported but not bounded
            good = (w_stat < 1e-4 or not tag.endswith('1')) and w_par < 2.1e-3
            ok &= good
            print(f'{tag}: max rel running-stat diff {w_stat:.2e} ({wk}), max abs parameter diff '
                  f'{w_par:.2e} {"ok" if good else "FAIL"}')
        teeth = abs(none_loss - float(l1)) / abs(float(l1))
        print(f"sync_stats='none' loss differs from single-process by {teeth:.2e} (expected: clearly non-zero)")
        ok &= teeth > 1e-4
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    # replicas identical after the exact step
    p = torch.cat([v.reshape(-1) for v in model.state_dict().values()]).float()
    pm = p.clone()
    dist.all_reduce(pm, op=dist.ReduceOp.MAX)
    same = bool((pm == p).all())
    if rank == 0:
        print('replicas bit-identical:', same)
        if flag.item() and same:
            print('DP_EXACT_OK')
    code = 0 if (flag.item() and same) else 1
    import faulthandler
    faulthandler.dump_traceback_later(25, exit=True)        # a stuck teardown must not eat GPU minutes
    torch.cuda.synchronize()
    step.close()                                            # graphs that hold NCCL kernels go first
    del step
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    dist.destroy_process_group()
    sys.exit(code)


if __name__ == '__main__':
    main()
