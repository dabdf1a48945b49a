"""Train-step plumbing: flat parameter arena, fused clip+Adam, data-parallel gradient allreduce,
CUDA-graph capture of the whole step.

Mirrors the slice of ``padertorch.train`` pb_sed drives
(pb_sed/experiments/weak_label_crnn/training.py:264-269 optimizer, :377-396 LR schedule,
:397-400 ``trainer.train``; step body per SURVEY App. A):
    forward -> review -> loss.backward() -> clip_grad_norm_ -> Adam.step -> zero_grad

* ``FlatArena`` re-homes every trainable parameter (and its ``.grad``) as a view into one
  contiguous fp32 buffer, so the weight-gradient kernels accumulate straight into the arena,
  one NCCL all-reduce covers all gradients, and one kernel does norm + clip + Adam.
* ``Adam`` keeps the padertorch optimizer surface (``lr``, ``gradient_clipping``,
  ``clip_grad``/``step``/``zero_grad``) on top of ``pbsed_grad_sumsq`` / ``pbsed_adam_step``.
* ``GraphedTrainStep`` captures forward+loss+backward+optimizer into one CUDA graph over static
  input buffers (shapes are static for fixed-length clips) and replays it per batch.
"""

import numpy as np
import torch
import torch.distributed as dist

from . import ops
from ._lib import call
from .ops import _ptr, _stream


def allreduce_flat(flat, group=None):
    """sum-all-reduce one flat gradient buffer in place (NCCL on GPUs; gloo in the CPU tests).
    The 1/world scaling is folded into the fused Adam kernel (``hyper[6]``)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def shard_bounds(n_items, rank, world):
    """contiguous, balanced shard [lo, hi) of ``n_items`` clips for ``rank``."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatArena:
    def __init__(self, module):
        params = [p for p in module.parameters() if p.requires_grad]
        assert params and all(p.is_cuda and p.dtype == torch.float32 for p in params), \
            'move the model to the GPU (fp32) before building the arena'
        dev = params[0].device
        sizes = [p.numel() for p in params]
        pad = [(-(s) % 4) for s in sizes]                  # keep every view 16-byte aligned
        self.n = int(sum(s + q for s, q in zip(sizes, pad)))
        self.params = torch.zeros(self.n, device=dev)
        self.grads = torch.zeros(self.n, device=dev)
        self.exp_avg = torch.zeros(self.n, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, device=dev)
        off = 0
        for p, s, q in zip(params, sizes, pad):
            view = self.params[off:off + s].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grads[off:off + s].view(p.shape)
            off += s + q
        self.module_params = params

    def rebind_grads(self):
        """re-attach the arena views after something set ``.grad = None``."""
        off = 0
        for p in self.module_params:
            s = p.numel()
            p.grad = self.grads[off:off + s].view(p.shape)
            off += s + (-s % 4)


class Adam:
    """padertorch.train.optimizer.Adam surface over the fused kernel (betas/eps = torch defaults)."""

    def __init__(self, model, lr=5e-4, gradient_clipping=1e10, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, distributed=None, sync_stats='none'):
        """sync_stats (data parallel only): 'none' = per-replica batch statistics; 'exact' = every
        batch-norm / running-norm / loss normaliser reduces over all replicas (``ops.set_sync_stats``),
        so that N replicas reproduce the single-process step on the concatenated batch."""
        self.arena = FlatArena(model)
        dev = self.arena.params.device
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, gradient_clipping, 0., 1., 0.],
                                  device=dev, dtype=torch.float32)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.grad_norm = torch.zeros(1, device=dev)
        self.lr, self.gradient_clipping = lr, gradient_clipping
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.process_group = process_group
        ops.enable_wgrad_stream(True)      # this optimizer joins the weight-gradient stream before it reads grads
        if self.distributed:
            self.set_grad_scale(1. / dist.get_world_size(process_group))
        self.sync_stats = sync_stats
        ops.set_sync_stats(sync_stats if self.distributed else 'none', process_group)

    def set_lr(self, lr):
        """LRAnnealingHook equivalent: the LR is a device scalar, valid under graph replay."""
        self.lr = lr
        self.hyper[0:1].copy_(torch.tensor([lr], dtype=torch.float32), non_blocking=True)

    def set_grad_scale(self, s):
        self.hyper[6:7].copy_(torch.tensor([s], dtype=torch.float32))

    def zero_grad(self):
        self.arena.grads.zero_()

    def allreduce_grads(self):
        ops.join_wgrad_stream()
        if self.distributed:
            allreduce_flat(self.arena.grads, self.process_group)

    def step(self):
        """all-reduce (if DP) -> global norm -> clip -> Adam -> zero the gradient arena."""
        self.allreduce_grads()
        return self.update()

    def update(self, join=True):
        a = self.arena
        if join:
            ops.join_wgrad_stream()
        call('pbsed_grad_sumsq', _ptr(a.grads), a.n, _ptr(self.hyper), _ptr(self.sumsq), _stream())
        call('pbsed_adam_step', _ptr(a.params), _ptr(a.grads), _ptr(a.exp_avg), _ptr(a.exp_avg_sq),
             a.n, _ptr(self.hyper), _ptr(self.sumsq), _ptr(self.grad_norm), 1, _stream())
        return self.grad_norm

    clip_grad = step     # norm + clip happen inside the fused step


def lr_schedule(iteration, breakpoints):
    """piecewise-linear LR factor of LRAnnealingHook (training.py:377-396)."""
    xs = [b[0] for b in breakpoints]
    ys = [b[1] for b in breakpoints]
    if iteration >= xs[-1]:
        return ys[-1]
    return float(np.interp(iteration, xs, ys))


def train_step(model, optimizer, batch):
    """one eager trainer iteration; returns (loss, grad_norm) device tensors."""
    model.train()
    outputs = model(dict(batch))
    loss = model.review(batch, outputs)['loss']
    loss.backward()
    return loss.detach(), optimizer.step()


class GraphedTrainStep:
    """whole train step as ONE CUDA graph.  ``example`` is a batch dict of CUDA tensors whose shapes
    stay fixed; ``__call__(batch)`` copies the new batch into the static buffers and replays."""

    def __init__(self, model, optimizer, example, warmup=2, input_fn=None):
        """input_fn(static): optional, captured at the head of the graph -- refills the static input buffers
        in place on the device (``data.SyntheticClipStream.fill_``), so every replay trains on a new batch."""
        self.model, self.optimizer, self.input_fn = model, optimizer, input_fn
        model.train()
        saved = getattr(model, 'emit_buffers', None)
        if saved is not None:
            model.emit_buffers = False        # no D2H syncs inside the captured region
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # DP: the NCCL all-reduce runs eagerly between two captured graphs
        self.split = self.optimizer.distributed
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
            if not self.split:
                self.grad_norm = self.optimizer.update()
            else:
                ops.join_wgrad_stream()       # a capture must end with every forked stream joined
        if self.split:
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update):
                self.grad_norm = self.optimizer.update(join=False)     # joined at the end of the first graph
        self.optimizer.arena.rebind_grads()
        self._static_out = (self.loss, self.grad_norm)      # the graph's own result buffers

    def _fwd_bwd(self):
        if self.input_fn is not None:
            self.input_fn(self.static)
        outputs = self.model(dict(self.static))
        loss = self.model.review(self.static, outputs)['loss']
        loss.backward()
        # keep only detached results: a live autograd graph would pin its AccumulateGrad nodes to the
        # warm-up stream and poison the capture with a cross-stream dependency
        self.outputs = tuple(o.detach() if torch.is_tensor(o) else o for o in outputs)
        return loss.detach()

    def _body(self):
        loss = self._fwd_bwd()
        return loss, self.optimizer.step()

    def close(self):
        """release the captured graphs (do this before tearing down a process group whose collectives
        were captured: 'exact' statistics put NCCL kernels inside the forward/backward graph)."""
        torch.cuda.synchronize()
        for name in ('graph', 'graph_update'):
            g = getattr(self, name, None)
            if g is not None:
                g.reset()
                setattr(self, name, None)

    def _same_lengths(self, batch):
        """sequence lengths (a host list) are baked into the captured graph: masks, batch-norm counts, loss
        denominators and the augmentation limits all derive from them at capture time."""
        a, b = batch.get('seq_len'), self.static.get('seq_len')
        return a is None or b is None or [int(v) for v in a] == [int(v) for v in b]

    def load(self, batch):
        assert self._same_lengths(batch), \
            'GraphedTrainStep: seq_len differs from the captured batch (use __call__, which falls back to an eager step)'
        for k, v in batch.items():
            if torch.is_tensor(v):
                self.static[k].copy_(v, non_blocking=True)

    # ---- double-buffered input pipeline: H2D of batch i+1 overlaps the compute of batch i
    def prefetch(self, host_batch):
        """start the host->device copy of the NEXT batch (pinned tensors) on a copy stream."""
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream()
            self._stage = [{k: torch.empty_like(v) for k, v in self.static.items() if torch.is_tensor(v)}
                           for _ in range(2)]
            self._stage_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._stage_idx = 0
            for e in self._stage_free:
                e.record()
        i = self._stage_idx
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[i])
            for k, v in host_batch.items():
                if torch.is_tensor(v):
                    self._stage[i][k].copy_(v, non_blocking=True)
            self._stage_ready[i].record()
        self._pending, self._pending_seq_len = i, host_batch.get('seq_len')
        self._stage_idx = i ^ 1

    def step_prefetched(self):
        """run one step on the batch handed to the last prefetch()."""
        i = self._pending
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready[i])
        if not self._same_lengths({'seq_len': self._pending_seq_len}):
            batch = dict({k: v.clone() for k, v in self._stage[i].items()}, seq_len=self._pending_seq_len)
            self._stage_free[i].record(cur)
            return self(batch)                                  # other clip lengths: eager step (see __call__)
        for k, v in self._stage[i].items():
            self.static[k].copy_(v, non_blocking=True)          # device-to-device, microseconds
        self._stage_free[i].record(cur)
        return self()

    def __call__(self, batch=None):
        """replay the captured step on ``batch`` (None: whatever the static buffers hold).  A batch whose
        ``seq_len`` differs from the captured one (ragged batches of ``data.collate``) cannot reuse the graph:
        it runs as an eager ``train_step`` -- same arithmetic, same optimizer state, un-graphed speed."""
        if batch is not None and not self._same_lengths(batch):
            if not getattr(self, '_warned', False):
                import warnings
                warnings.warn('GraphedTrainStep: batch with different seq_len -> eager step (bucket clips by '
                              'length, or capture one graph per length pattern, to stay on the graph path)')
                self._warned = True
            merged = {k: v for k, v in self.static.items() if not torch.is_tensor(v)}
            merged.update(batch)
            self.loss, self.grad_norm = train_step(self.model, self.optimizer, merged)
            self.optimizer.arena.rebind_grads()
            return self.loss, self.grad_norm
        if batch is not None:
            self.load(batch)
        self.graph.replay()
        if self.split:
            self.optimizer.allreduce_grads()
            self.graph_update.replay()
        self.loss, self.grad_norm = self._static_out
        return self.loss, self.grad_norm
