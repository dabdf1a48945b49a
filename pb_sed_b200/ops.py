"""Thin torch-facing wrappers + autograd Functions over the C ABI (``include/pbsed_b200.h``).

Device memory and streams come from torch; all arithmetic happens in the hand-written
sm_100a kernels of ``libpbsed_b200.so``.  There is no eager / CPU fallback: every function
asserts CUDA tensors and calls through ``_lib.call``.

Native activation layout ("rows x channels"): 2-D maps ``(B, F, T, C)``, 1-D maps
``(B, T, C)``; channels contiguous.  The reference's ``(B, C, F, T)`` / ``(B, C, T)`` tensors
are exposed as permuted *views* of these (``to_native`` / ``from_native`` are zero-copy when
the tensor was produced by this package).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import TapGemmDesc, call

import os

PRECISION = {'fp32': 0, 'tf32x3': 1, 'tf32': 3, 'bf16': 3}
_default_precision = PRECISION[os.environ.get('PBSED_PRECISION', 'tf32x3')]
# storage type of the conv-stack activation maps and their gradients in HBM: torch.float32, or torch.bfloat16 in
# the 'bf16' mode (BASELINE configs[2] / [4]).  In that mode the conv stacks run ONE tensor-core pass with fp32
# accumulation: kind::f16 (bf16 x bf16, K = 16) MMAs wherever the input map is bf16 with a multiple of 32 channels
# (forward, data and weight gradients of the wide layers), one kind::tf32 pass on the bf16-exact values elsewhere; master
# weights, batch statistics, the GRU and the optimizer stay fp32.
_act_dtype = torch.bfloat16 if os.environ.get('PBSED_PRECISION') == 'bf16' else torch.float32


def set_default_precision(p):
    """precision of the tap-GEMMs: 'fp32' (exact FFMA), 'tf32x3' (tcgen05, 3-pass split TF32 =
    fp32-equivalent, the default) or 'tf32' (tcgen05, ONE TF32 pass with fp32 accumulation: the
    reduced-precision mode offered for the BASELINE "bf16" configurations -- 10 mantissa bits, i.e.
    at least bf16's 7, at a third of the tensor work)."""
    global _default_precision, _act_dtype
    _default_precision = PRECISION[p] if isinstance(p, str) else int(p)
    _act_dtype = torch.bfloat16 if p == 'bf16' else torch.float32


def act_dtype():
    """storage dtype of conv-stack activation maps (torch.bfloat16 in the 'bf16' mode)."""
    return _act_dtype


def _dt(t):
    """C-ABI dtype code of an activation tensor / torch dtype: 0 = fp32, 1 = bf16."""
    d = t if isinstance(t, torch.dtype) else t.dtype
    return 1 if d == torch.bfloat16 else 0


def _actc(t):
    """contiguous activation map in one of the two storage types."""
    if t.dtype not in (torch.float32, torch.bfloat16):
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------ weight-gradient side stream
# Weight gradients feed nothing but the optimizer, so (when a consumer that joins the stream is in
# charge, i.e. train.Adam) they are launched on a side stream: they fill the SMs the latency-bound GRU
# recurrences leave idle and overlap the data-gradient chain instead of sitting on its critical path.
_wgrad_stream = None
_wgrad_stream_enabled = False


def enable_wgrad_stream(flag=True):
    global _wgrad_stream_enabled
    _wgrad_stream_enabled = bool(flag) and os.environ.get('PBSED_WGRAD_STREAM', '1') != '0'


class _WgradStream:
    """context: fork the side stream off the current one and keep the tensors it reads alive."""

    def __init__(self, *tensors, direct=True):
        """direct: every gradient of this block is accumulated straight into an existing leaf ``.grad``
        (the flat arena, joined by ``train.Adam``).  A gradient buffer that is RETURNED to autograd is
        consumed on the current stream, so such blocks stay on the current stream."""
        self.tensors = [t for t in tensors if t is not None]
        self.ctx = None
        self.direct = direct

    def __enter__(self):
        global _wgrad_stream
        if not _wgrad_stream_enabled or not self.direct:
            return self
        if _wgrad_stream is None:
            _wgrad_stream = torch.cuda.Stream()
        _wgrad_stream.wait_stream(torch.cuda.current_stream())
        for t in self.tensors:
            t.record_stream(_wgrad_stream)
        self.ctx = torch.cuda.stream(_wgrad_stream)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def join_wgrad_stream():
    """make the current stream wait for every weight gradient launched so far."""
    if _wgrad_stream is not None:
        torch.cuda.current_stream().wait_stream(_wgrad_stream)


# ------------------------------------------------------------------ data-parallel exact statistics
# clips shard over ranks, but the 15 batch-norm layers (and the feature extractor's running
# normalisation, and the loss normaliser) reduce over the WHOLE batch in the single-process reference.
# 'exact' all-reduces (sum, sum of squares, count) per norm layer -- one small buffer, the count rides
# in its last row and stays on the device -- in forward, the mirrored (sum g, sum g*xhat, count) in
# backward, and (loss numerator, sum of weights) in the loss, so that N replicas reproduce the
# single-process step.  'none' keeps per-replica statistics (what DistributedDataParallel would do).
_sync = {'on': False, 'group': None}


def set_sync_stats(mode='none', group=None):
    assert mode in ('none', 'exact'), mode
    import torch.distributed as dist
    on = mode == 'exact' and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    _sync['on'], _sync['group'] = on, group


def sync_stats_on():
    return _sync['on']


def allreduce_stats_(t):
    """in-place sum over the replicas of a small statistics buffer (NCCL on GPUs, gloo in CPU tests)."""
    import torch.distributed as dist
    if _sync['on']:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_sync['group'])
    return t


def _stats_buffer(nch, device):
    """(nch [+1], 2) float64 zeros; the extra row carries the replica's count in 'exact' mode."""
    return torch.zeros((nch + (1 if _sync['on'] else 0), 2), device=device, dtype=torch.float64)


def _sync_count_(stats, nch, count):
    """'exact': put the local count in the last row, all-reduce, and tell the kernel (count = 0) to read
    the global count from the buffer.  Otherwise return the host count unchanged."""
    if not _sync['on']:
        return float(count)
    stats[nch:nch + 1, 0:1].fill_(float(count))       # fill kernel: no H2D copy, graph-capturable
    allreduce_stats_(stats)
    return 0.0


def _ptr(t):
    if t is None:
        return None
    assert t.is_cuda, 'pb_sed_b200 ops run on CUDA tensors only (no CPU fallback)'
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------ sequence lengths
class SeqLen:
    """host + device copy of the per-clip number of valid frames (None == all valid)."""
    _cache = {}

    def __init__(self, host, T, device):
        self.host = host                      # np.int64 array or None
        self.T = T
        self.dev = None
        if host is not None:
            self.dev = torch.from_numpy(np.minimum(host, T).astype(np.int32)).to(device)

    @classmethod
    def make(cls, seq_len, B, T, device):
        if isinstance(seq_len, SeqLen):
            assert seq_len.T == T, (seq_len.T, T)
            return seq_len
        if seq_len is None:
            key = (None, B, T, str(device))
        else:
            host = np.asarray(seq_len).astype(np.int64).reshape(-1)
            assert host.shape[0] == B, (host.shape, B)
            if (host >= T).all():
                host = None                   # nothing to mask: skip the predicate everywhere
                key = (None, B, T, str(device))
            else:
                key = (host.tobytes(), B, T, str(device))
        s = cls._cache.get(key)
        if s is None:
            if len(cls._cache) > 256:
                cls._cache.clear()
            s = cls(None if key[0] is None else host, T, device)
            s.B = B
            cls._cache[key] = s
        return s

    @property
    def ptr(self):
        return _ptr(self.dev)

    def frames(self):
        """total number of valid frames over the batch."""
        return float(self.B * self.T) if self.host is None else float(np.minimum(self.host, self.T).sum())


def to_native(x):
    """(B,C,F,T) -> (B,F,T,C) / (B,C,T) -> (B,T,C), contiguous (zero-copy for our own views)."""
    if x.dim() == 4:
        return x.permute(0, 2, 3, 1).contiguous()
    return x.transpose(1, 2).contiguous()


def from_native(x):
    if x.dim() == 4:
        return x.permute(0, 3, 1, 2)
    return x.transpose(1, 2)


# ------------------------------------------------------------------ raw kernels
def make_desc(B, F_in, F_out, T, Cin, Cout, taps, relu=False, per_f=False, transpose_w=False,
              in_stride=0, out_stride=0, precision=None, no_input_mask=False, in_dtype=0, out_dtype=0):
    d = TapGemmDesc()
    d.B, d.F_in, d.F_out, d.T, d.Cin, d.Cout = B, F_in, F_out, T, Cin, Cout
    d.ntaps = len(taps)
    assert 1 <= d.ntaps <= _lib.MAX_TAPS, d.ntaps
    for i, (df, dt) in enumerate(taps):
        d.df[i], d.dt[i] = int(df), int(dt)
    d.relu, d.per_f = int(relu), int(per_f)
    d.w_tap_stride = Cin * Cout
    if transpose_w:          # this call's (Cout, Cin) are the forward layer's (Cin, Cout)
        d.w_sn, d.w_sc = 1, Cout
    else:
        d.w_sn, d.w_sc = Cin, 1
    d.in_stride, d.out_stride = in_stride, out_stride
    d.precision = _default_precision if precision is None else precision
    d.no_input_mask = int(no_input_mask)
    d.in_dtype, d.out_dtype = int(in_dtype), int(out_dtype)
    return d


def tapgemm(x, W, bias, desc, scale=None, shift=None, seq=None, ep_src=None, ep_scale=None,
            ep_shift=None, out=None, x_ptr=None, out_stats=None, ep_mean=None, ep_rstd=None, ep_sums=None):
    rows = desc.B * desc.F_out * desc.T
    if out is None:
        out = torch.empty((rows, desc.Cout), device=W.device,
                          dtype=torch.bfloat16 if desc.out_dtype == 1 else torch.float32)
    ws, ws_bytes = None, 0
    if desc.precision != 0:
        ws_bytes = int(_lib.load().pbsed_tapgemm_workspace_bytes(ctypes.byref(desc)))
        ws = torch.empty(ws_bytes, device=W.device, dtype=torch.uint8)
    call('pbsed_tapgemm', ctypes.byref(desc), x_ptr if x_ptr is not None else _ptr(x), _ptr(scale),
         _ptr(shift), seq.ptr if seq is not None else None, _ptr(W), _ptr(bias), _ptr(out),
         _ptr(ep_src), _ptr(ep_scale), _ptr(ep_shift), _ptr(out_stats), _ptr(ep_mean), _ptr(ep_rstd),
         _ptr(ep_sums), _ptr(ws), ws_bytes, _stream())
    return out


def tapgemm_wgrad(x, dout, desc, dW, dbias, scale=None, shift=None, seq=None, mask_out=True,
                  x_ptr=None, dout_ptr=None, dW_ptr=None, dbias_ptr=None):
    call('pbsed_tapgemm_wgrad', ctypes.byref(desc), x_ptr if x_ptr is not None else _ptr(x),
         _ptr(scale), _ptr(shift), seq.ptr if seq is not None else None,
         dout_ptr if dout_ptr is not None else _ptr(dout), int(mask_out),
         dW_ptr if dW_ptr is not None else _ptr(dW),
         dbias_ptr if dbias_ptr is not None else _ptr(dbias), _stream())


def _grad_target(p):
    """accumulate straight into ``p.grad`` when it exists (flat-arena training), else into a
    fresh zero buffer that is returned to autograd."""
    if p is None or not p.requires_grad:
        return None, None
    if p.is_leaf and p.grad is not None:
        return p.grad, None
    g = torch.zeros_like(p)
    return g, g


# ------------------------------------------------------------------ conv layer
class ConvLayerFn(torch.autograd.Function):
    """[norm -> ReLU ->] zero-pad -> tap-conv(+bias) [-> frequency max-pool] on native maps.

    x: (B, F_in, T, Cin) native.  weight: (ntaps, Cout, Cin).  Returns (B, F_out', T, Cout).
    cfg: dict(F_in, F_out, taps, relu, per_f, pool, eps, momentum, training, norm: bool)
    norm buffers (running_mean, running_power, num_tracked) are updated in place.
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, rmean, rpower, ntracked, seq, cfg, stats_in=None):
        """returns (y, stats_out): stats_out = (Cout[*F_out], 2) float64 sum / sum-of-squares of y over
        valid frames when cfg['want_stats'] (fused into the conv epilogue), else an empty tensor.
        stats_in: the same quantity for x, produced by the previous layer (skips the statistics pass)."""
        x = _actc(x)
        B, F_in, T, Cin = x.shape
        ntaps, Cout, Cin_w = weight.shape
        assert Cin_w == Cin and F_in == cfg['F_in'], (weight.shape, x.shape, cfg)
        idt, odt = _dt(x), _dt(cfg.get('out_dtype', torch.float32))
        F_out, pool = cfg['F_out'], cfg.get('pool', 1)
        per_f = cfg.get('per_f', False)
        scale = shift = smean = srstd = None
        nch = (F_in if per_f else 1) * Cin
        count = seq.frames() * (1 if per_f else F_in)
        if cfg['norm']:
            scale = torch.empty(nch, device=x.device)
            shift = torch.empty(nch, device=x.device)
            if cfg['training']:
                if stats_in is not None and stats_in.shape[0] == nch + (1 if _sync['on'] else 0):
                    stats = stats_in
                else:
                    stats = _stats_buffer(nch, x.device)
                    call('pbsed_channel_stats', _ptr(x), B, F_in, T, Cin, int(per_f), seq.ptr,
                         _ptr(stats), idt, _stream())
                smean = torch.empty(nch, device=x.device)
                srstd = torch.empty(nch, device=x.device)
                call('pbsed_norm_finalize', _ptr(stats), _sync_count_(stats, nch, count), nch, _ptr(gamma), _ptr(beta),
                     cfg['eps'], cfg['momentum'], 1, _ptr(rmean), _ptr(rpower), _ptr(ntracked),
                     _ptr(scale), _ptr(shift), _ptr(smean), _ptr(srstd), _stream())
            else:
                call('pbsed_norm_finalize', None, 1.0, nch, _ptr(gamma), _ptr(beta), cfg['eps'],
                     cfg['momentum'], 0, _ptr(rmean), _ptr(rpower), _ptr(ntracked), _ptr(scale),
                     _ptr(shift), None, None, _stream())
        # the reference masks padded frames inside Normalization only: a bare conv (layer 0, whose
        # tag-condition channels are NOT zero at padded frames) reads its input unmasked
        desc = make_desc(B, F_in, F_out, T, Cin, Cout, cfg['taps'], relu=cfg['relu'], per_f=per_f,
                         no_input_mask=not cfg['norm'], in_dtype=idt, out_dtype=odt)
        stats_out = None
        if cfg.get('want_stats') and pool == 1:
            sp = cfg.get('stats_per_f', False)
            desc_s = make_desc(B, F_in, F_out, T, Cin, Cout, cfg['taps'], relu=cfg['relu'], per_f=per_f,
                               no_input_mask=not cfg['norm'], in_dtype=idt, out_dtype=odt)
            stats_out = _stats_buffer((F_out if sp else 1) * Cout, x.device)
            if sp == per_f:
                z = tapgemm(x, weight, bias, desc_s, scale, shift, seq, out_stats=stats_out)
            else:      # statistics indexed differently from the load affine: separate pass
                z = tapgemm(x, weight, bias, desc, scale, shift, seq)
                call('pbsed_channel_stats', _ptr(z), B, F_out, T, Cout, int(sp), seq.ptr, _ptr(stats_out), odt, _stream())
            z = z.view(B, F_out, T, Cout)
        else:
            z = tapgemm(x, weight, bias, desc, scale, shift, seq).view(B, F_out, T, Cout)
        idx = None
        if pool > 1:
            y = torch.empty((B, F_out // pool, T, Cout), device=x.device, dtype=z.dtype)
            idx = torch.empty(y.shape, device=x.device, dtype=torch.uint8)
            if cfg.get('want_stats') and not cfg.get('stats_per_f', False):
                stats_out = _stats_buffer(Cout, x.device)      # statistics of the pooled map, same pass
            call('pbsed_maxpool_f', _ptr(z), B, F_out, T, Cout, pool, _ptr(y), _ptr(idx), seq.ptr,
                 _ptr(stats_out), odt, odt, _stream())
        else:
            y = z
        ctx.cfg, ctx.seq, ctx.count = cfg, seq, count
        ctx.save_for_backward(x, weight, gamma, scale, shift, smean, srstd, idx)
        ctx.params = (weight, bias, gamma, beta)
        if stats_out is None:
            stats_out = torch.empty(0, device=x.device, dtype=torch.float64)
        ctx.mark_non_differentiable(stats_out)
        return y, stats_out

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        x, weight, gamma, scale, shift, smean, srstd, idx = ctx.saved_tensors
        w_p, b_p, g_p, be_p = ctx.params
        cfg, seq = ctx.cfg, ctx.seq
        B, F_in, T, Cin = x.shape
        ntaps, Cout, _ = weight.shape
        F_out, pool, per_f = cfg['F_out'], cfg.get('pool', 1), cfg.get('per_f', False)
        idt, odt = _dt(x), _dt(cfg.get('out_dtype', torch.float32))
        dy = _actc(dy)
        if _dt(dy) != odt:
            dy = dy.to(torch.bfloat16 if odt else torch.float32)
        if pool > 1:
            dz = torch.empty((B, F_out, T, Cout), device=x.device, dtype=dy.dtype)
            call('pbsed_maxpool_f_bwd', _ptr(dy), _ptr(idx), B, F_out, T, Cout, pool, _ptr(dz), odt, odt, _stream())
        else:
            dz = dy
        desc = make_desc(B, F_in, F_out, T, Cin, Cout, cfg['taps'], relu=cfg['relu'], per_f=per_f,
                         in_dtype=idt, out_dtype=odt)
        dW, dW_ret = _grad_target(w_p)
        db, db_ret = _grad_target(b_p)
        if dW is not None:
            with _WgradStream(x, dz, scale, shift, direct=dW_ret is None and db_ret is None):
                tapgemm_wgrad(x, dz, desc, dW, db, scale, shift, seq if cfg['norm'] else None, mask_out=False)
        dx = dg_ret = dbe_ret = None
        # the data-gradient pass is also what yields this layer's OWN norm gradients (dgamma, dbeta): it must
        # run when only they are wanted (input produced by frozen layers, training.py:343-350)
        norm_grads = cfg['norm'] and cfg['training'] and g_p is not None and g_p.requires_grad
        if ctx.needs_input_grad[0] or norm_grads:
            rtaps = [(-df, -dt) for df, dt in cfg['taps']]
            ddesc = make_desc(B, F_out, F_in, T, Cout, Cin, rtaps, per_f=per_f, transpose_w=True,
                              in_dtype=odt, out_dtype=idt)       # reads dz, writes g in the dtype of x (= ep_src)
            train_norm = cfg['norm'] and cfg['training']
            sums = None
            fuse = train_norm and cfg['relu']          # batch-norm backward pass 1 rides in the dgrad epilogue
            if train_norm:
                nch = (F_in if per_f else 1) * Cin
                sums = _stats_buffer(nch, x.device)
            g = tapgemm(dz, weight, None, ddesc, None, None, seq,
                        ep_src=x if cfg['relu'] else None,
                        ep_scale=scale if cfg['relu'] else None,
                        ep_shift=shift if cfg['relu'] else None,
                        ep_mean=smean if fuse else None, ep_rstd=srstd if fuse else None,
                        ep_sums=sums if fuse else None).view(B, F_in, T, Cin)
            if train_norm:
                if not fuse:
                    call('pbsed_norm_bwd_reduce', _ptr(g), _ptr(x), B, F_in, T, Cin, int(per_f), seq.ptr,
                         _ptr(smean), _ptr(srstd), _ptr(sums), idt, _stream())
                dga, dg_ret = _grad_target(g_p)
                dbe, dbe_ret = _grad_target(be_p)
                if _sync['on']:
                    # dx needs the sums over ALL replicas; dgamma / dbeta take the local ones (the flat
                    # gradient all-reduce adds the other replicas' later)
                    local = sums[:nch].float()
                    cnt = _sync_count_(sums, nch, ctx.count)
                    call('pbsed_norm_bwd_apply', _ptr(g), _ptr(x), B, F_in, T, Cin, int(per_f), seq.ptr,
                         _ptr(smean), _ptr(srstd), _ptr(gamma), _ptr(sums), cnt, _ptr(g), None, None, idt, _stream())
                    if dga is not None:
                        dga.add_(local[:, 1])
                        dbe.add_(local[:, 0])
                else:
                    call('pbsed_norm_bwd_apply', _ptr(g), _ptr(x), B, F_in, T, Cin, int(per_f), seq.ptr,
                         _ptr(smean), _ptr(srstd), _ptr(gamma), _ptr(sums), ctx.count, _ptr(g),
                         _ptr(dga), _ptr(dbe), idt, _stream())
                dx = g
            elif cfg['norm']:
                # eval-mode norm is a fixed affine: dx = g * scale  (rare: frozen-stat finetuning)
                sc = scale.view(F_in, 1, Cin) if per_f else scale
                dx = (g * sc).to(g.dtype)
            else:
                dx = g
            if not ctx.needs_input_grad[0]:
                dx = None
        return dx, dW_ret, db_ret, dg_ret, dbe_ret, None, None, None, None, None, None


# ------------------------------------------------------------------ GRU layer
def _ptr_array(tensors_or_ptrs):
    vals = [(t if isinstance(t, int) else (0 if t is None else t.data_ptr())) for t in tensors_or_ptrs]
    return (ctypes.c_void_p * len(vals))(*vals)


class GruMultiFn(torch.autograd.Function):
    """ONE layer of G independent GRU modules ("groups"), each with ``nd`` directions, in a single
    persistent-kernel launch (G*nd <= 4 recurrences run concurrently on disjoint SM clusters).

    apply(seq, meta, *tensors):  meta = [reverse flags per group], tensors = per group
    (x (B,T,In), w_ih (nd,3H,In), w_hh (nd,3H,H), b_ih (nd,3H), b_hh (nd,3H)).
    Returns one (B,T,nd*H) map per group.  The reference's rnn_fwd / rnn_bwd pair
    (weak_label/crnn.py:338-340) is G = 2, nd = 1; its bidirectional GRU (strong_label/crnn.py:
    189-195) is G = 1, nd = 2.
    """

    @staticmethod
    def forward(ctx, seq, meta, *tensors):
        G = len(meta)
        groups = [tensors[5 * g:5 * g + 5] for g in range(G)]
        x0 = groups[0][0]
        B, T, _ = x0.shape
        nd, H3, H = groups[0][2].shape
        dev = x0.device
        gis, hs, saves, xs = [], [], [], []
        p_gi, p_whh, p_bhh, p_h, p_save, rev = [], [], [], [], [], []
        for g in range(G):
            x, w_ih, w_hh, b_ih, b_hh = groups[g]
            x = _f32c(x)
            xs.append(x)
            assert w_hh.shape == (nd, H3, H) and x.shape[:2] == (B, T)
            In = x.shape[2]
            gi = torch.empty((nd, B, T, H3), device=dev)
            for d in range(nd):
                tapgemm(x, w_ih[d], b_ih[d], make_desc(B, 1, 1, T, In, H3, [(0, 0)]), seq=None, out=gi[d])
            h = torch.empty((B, T, nd * H), device=dev)
            save = torch.empty((nd, B, T, 4 * H), device=dev)
            gis.append(gi); hs.append(h); saves.append(save)
            for d in range(nd):
                p_gi.append(gi[d]); p_whh.append(w_hh[d]); p_bhh.append(b_hh[d])
                p_h.append(h.data_ptr() + 4 * d * H); p_save.append(save[d])
                rev.append(int(meta[g][d]))
        n = len(rev)
        call('pbsed_gru_fwd', _ptr_array(p_gi), _ptr_array(p_whh), _ptr_array(p_bhh), seq.ptr, B, T, H,
             n, (ctypes.c_int * n)(*rev), _ptr_array(p_h), nd * H, _ptr_array(p_save), _stream())
        ctx.seq, ctx.meta, ctx.rev, ctx.dims = seq, meta, rev, (G, nd, B, T, H)
        ctx.save_for_backward(*xs, *hs, *saves, *[groups[g][1] for g in range(G)],
                              *[groups[g][2] for g in range(G)])
        ctx.params = [groups[g][1:] for g in range(G)]
        return tuple(hs)

    @staticmethod
    def backward(ctx, *dhs):
        G, nd, B, T, H = ctx.dims
        H3 = 3 * H
        sv = ctx.saved_tensors
        xs, hs, saves = sv[0:G], sv[G:2 * G], sv[2 * G:3 * G]
        w_ihs, w_hhs = sv[3 * G:4 * G], sv[4 * G:5 * G]
        seq, rev = ctx.seq, ctx.rev
        dev = xs[0].device
        dgis = [torch.empty((nd, B, T, H3), device=dev) for _ in range(G)]
        dghs = [torch.empty((nd, B, T, H3), device=dev) for _ in range(G)]
        dhs = [(_f32c(dh) if dh is not None else torch.zeros((B, T, nd * H), device=dev)) for dh in dhs]
        p_dh, p_h, p_save, p_whh, p_dgi, p_dgh = [], [], [], [], [], []
        for g in range(G):
            for d in range(nd):
                p_dh.append(dhs[g].data_ptr() + 4 * d * H); p_h.append(hs[g].data_ptr() + 4 * d * H)
                p_save.append(saves[g][d]); p_whh.append(w_hhs[g][d])
                p_dgi.append(dgis[g][d]); p_dgh.append(dghs[g][d])
        n = G * nd
        call('pbsed_gru_bwd', _ptr_array(p_dh), _ptr_array(p_h), _ptr_array(p_save), _ptr_array(p_whh),
             seq.ptr, B, T, H, n, (ctypes.c_int * n)(*rev), _ptr_array(p_dgi), _ptr_array(p_dgh),
             nd * H, _stream())
        grads = []
        for g in range(G):
            p_wih, p_whh_, p_bih, p_bhh = ctx.params[g]
            x, h = xs[g], hs[g]
            In = x.shape[2]
            dwih, r0 = _grad_target(p_wih)
            dwhh, r1 = _grad_target(p_whh_)
            dbih, r2 = _grad_target(p_bih)
            dbhh, r3 = _grad_target(p_bhh)
            dx = None
            with _WgradStream(x, h, dgis[g], dghs[g], direct=all(r is None for r in (r0, r1, r2, r3))):
                for d in range(nd):
                    if dwih is not None:
                        tapgemm_wgrad(x, dgis[g][d], make_desc(B, 1, 1, T, In, H3, [(0, 0)]), dwih[d],
                                      dbih[d] if dbih is not None else None, seq=seq, mask_out=True)
                    if dwhh is not None:
                        desc = make_desc(B, 1, 1, T, H, H3, [(0, 1 if rev[g * nd + d] else -1)], in_stride=nd * H)
                        tapgemm_wgrad(None, dghs[g][d], desc, dwhh[d], dbhh[d] if dbhh is not None else None,
                                      seq=seq, mask_out=True, x_ptr=ctypes.c_void_p(h.data_ptr() + 4 * d * H))
            for d in range(nd):
                if ctx.needs_input_grad[2 + 5 * g]:
                    ddesc = make_desc(B, 1, 1, T, H3, In, [(0, 0)], transpose_w=True)
                    gx = tapgemm(dgis[g][d], w_ihs[g][d], None, ddesc, seq=None).view(B, T, In)
                    dx = gx if dx is None else dx.add_(gx)
            grads += [dx, r0, r1, r2, r3]
        return (None, None, *grads)


# ------------------------------------------------------------------ misc element-wise ops
class ConcatCondFn(torch.autograd.Function):
    """native rows (B,F,T,C0) ++ cond (B,K) broadcast -> (B,F,T,C0+K)  (strong_label/crnn.py:86-91)."""

    @staticmethod
    def forward(ctx, x, cond):
        x = _f32c(x)
        cond = _f32c(cond)
        shp = x.shape
        B, C0 = shp[0], shp[-1]
        FT = int(np.prod(shp[1:-1]))
        K = cond.shape[1]
        out = torch.empty(shp[:-1] + (C0 + K,), device=x.device)
        call('pbsed_concat_cond', _ptr(x), _ptr(cond), B, 1, FT, C0, K, _ptr(out), _stream())
        ctx.dims = (B, FT, C0, K, shp)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, FT, C0, K, shp = ctx.dims
        dx = torch.empty(shp, device=dout.device)
        call('pbsed_split_cond_bwd', _ptr(_f32c(dout)), B, 1, FT, C0, K, _ptr(dx), _stream())
        return dx, None


class SigmoidScoresFn(torch.autograd.Function):
    """logits (B,T,K) native -> scores (B,K,T) = min + (1 - 2 min) * sigmoid (weak_label/crnn.py:58-59)."""

    @staticmethod
    def forward(ctx, z, min_score):
        z = _f32c(z)
        B, T, K = z.shape
        y = torch.empty((B, K, T), device=z.device)
        call('pbsed_sigmoid_btk_to_bkt', _ptr(z), B, T, K, float(min_score), _ptr(y), _stream())
        ctx.save_for_backward(z)
        ctx.min_score = float(min_score)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, = ctx.saved_tensors
        B, T, K = z.shape
        dz = torch.empty_like(z)
        call('pbsed_sigmoid_bwd', _ptr(_f32c(dy)), _ptr(z), B, T, K, ctx.min_score, _ptr(dz), _stream())
        return dz, None


def _sync_loss(out):
    """out = (local loss = N_r / D_r, D_r = local sum of weights).  'exact' data parallelism: the
    single-process loss is sum_r N_r / sum_r D_r, so the replica's gradient is rescaled by
    D_r * world / D_global (the flat all-reduce + 1/world then yields the exact gradient)."""
    if not _sync['on']:
        return out[0], torch.ones((), device=out.device)
    import torch.distributed as dist
    world = dist.get_world_size(_sync['group'])
    packed = torch.stack([out[0] * out[1], out[1]])
    allreduce_stats_(packed)
    return packed[0] / packed[1], out[1] * world / packed[1]


class FbcrnnLossFn(torch.autograd.Function):
    """pb_sed weak_label CRNN.review loss (weak_label/crnn.py:117-153); value and gradient
    come out of the same kernel pass."""

    @staticmethod
    def forward(ctx, y_fwd, y_bwd, weak, boundary, class_weights, seq, strong_weight, smoothing):
        y_fwd = _f32c(y_fwd)
        B, K, T = y_fwd.shape
        y_bwd = None if y_bwd is None else _f32c(y_bwd)
        weak = _f32c(weak)
        boundary = None if boundary is None else _f32c(boundary)
        cw = None if class_weights is None else _f32c(class_weights)
        out = torch.empty(2, device=y_fwd.device)
        ws = torch.empty(2 * B * K + 8, device=y_fwd.device)
        dyf = torch.empty_like(y_fwd)
        dyb = None if y_bwd is None else torch.empty_like(y_bwd)
        call('pbsed_fbcrnn_loss', _ptr(y_fwd), _ptr(y_bwd), _ptr(weak), _ptr(boundary), _ptr(cw),
             seq.ptr, B, K, T, float(strong_weight), float(smoothing), _ptr(out), _ptr(dyf),
             _ptr(dyb), _ptr(ws), _stream())
        loss, gscale = _sync_loss(out)
        ctx.save_for_backward(dyf, dyb, gscale)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dyf, dyb, gscale = ctx.saved_tensors
        dl = dl * gscale
        return (dyf * dl, None if dyb is None else dyb * dl, None, None, None, None, None, None)


class BicrnnLossFn(torch.autograd.Function):
    """pb_sed strong_label CRNN.review loss (strong_label/crnn.py:107-112)."""

    @staticmethod
    def forward(ctx, y, strong, seq):
        y = _f32c(y)
        B, K, T = y.shape
        strong = _f32c(strong)
        out = torch.empty(2, device=y.device)
        ws = torch.empty(8, device=y.device)
        dy = torch.empty_like(y)
        call('pbsed_bicrnn_loss', _ptr(y), _ptr(strong), seq.ptr, B, K, T, _ptr(out), _ptr(dy),
             _ptr(ws), _stream())
        loss, gscale = _sync_loss(out)
        ctx.save_for_backward(dy, gscale)
        return loss

    @staticmethod
    def backward(ctx, dl):
        dy, gscale = ctx.saved_tensors
        return dy * (dl * gscale), None, None


# ------------------------------------------------------------------ features
def logmel_from_audio(audio, stft_cfg, fb, seq, stats, frame_start=None):
    """audio (B,S) -> log-mel (B, n_mels, T) (+ accumulates per-band stats).  ``fb['per_clip']``: one
    filterbank per clip (MelWarping); ``frame_start`` int32 (B,T): TimeWarpedSTFT frame onsets."""
    audio = _f32c(audio)
    B, S = audio.shape
    T = stft_cfg['T']
    out = torch.empty((B, fb['n_mels'], T), device=audio.device)
    call('pbsed_stft_logmel', _ptr(audio), B, S, stft_cfg['shift'], stft_cfg['window_length'],
         stft_cfg['size'], stft_cfg['pad_front'], T, _ptr(stft_cfg['window']), _ptr(fb['lo']),
         _ptr(fb['hi']), _ptr(fb['w']), fb['stride'], fb['n_mels'], int(fb.get('per_clip', False)),
         _ptr(frame_start), seq.ptr, _ptr(out), _ptr(stats), _stream())
    return out


def logmel_from_stft(stft, fb, seq, stats):
    """stft (B,T,F,2) -> log-mel (B, n_mels, T)."""
    stft = _f32c(stft)
    B, T, n_bins, two = stft.shape
    assert two == 2
    out = torch.empty((B, fb['n_mels'], T), device=stft.device)
    call('pbsed_spec_logmel', _ptr(stft), B, T, n_bins, _ptr(fb['lo']), _ptr(fb['hi']), _ptr(fb['w']),
         fb['stride'], fb['n_mels'], int(fb.get('per_clip', False)), seq.ptr, _ptr(out), _ptr(stats), _stream())
    return out


def make_warped_fbank(alpha, ratio, n_mels, n_bins, mel_lo, mel_hi, mel_warp_hi, bins_per_hz, stride):
    """per-example warped filterbank tables (B = len(alpha)) for ``logmel_from_*`` (``per_clip``)."""
    B = alpha.shape[0]
    dev = alpha.device
    lo = torch.empty((B, n_mels), device=dev, dtype=torch.int32)
    hi = torch.empty((B, n_mels), device=dev, dtype=torch.int32)
    w = torch.empty((B, n_mels, stride), device=dev)
    call('pbsed_make_warped_fbank', _ptr(_f32c(alpha)), _ptr(_f32c(ratio)), B, n_mels, n_bins, float(mel_lo),
         float(mel_hi), float(mel_warp_hi), float(bins_per_hz), _ptr(lo), _ptr(hi), _ptr(w), stride, _stream())
    return dict(lo=lo, hi=hi, w=w, stride=stride, n_mels=n_mels, per_clip=True)


def norm_finalize(stats, count, nch, gamma, beta, eps, momentum, training, rmean, rpower, ntracked,
                  device):
    scale = torch.empty(nch, device=device)
    shift = torch.empty(nch, device=device)
    call('pbsed_norm_finalize', _ptr(stats), float(count), nch, _ptr(gamma), _ptr(beta), float(eps),
         float(momentum), int(training), _ptr(rmean), _ptr(rpower), _ptr(ntracked), _ptr(scale),
         _ptr(shift), None, None, _stream())
    return scale, shift


def logmel_normalize_(x, scale, shift, clamp, seq, time_masks=None, freq_masks=None, noise=None,
                      noise_scale=None):
    """in place: normalise, clamp, zero padded frames; train-time: (B,n,2) int32 time / frequency masks
    and ``noise_scale[b] * noise``."""
    B, F, T = x.shape
    nt = 0 if time_masks is None else time_masks.shape[1]
    nf = 0 if freq_masks is None else freq_masks.shape[1]
    call('pbsed_logmel_normalize', _ptr(x), B, F, T, _ptr(scale), _ptr(shift), float(clamp or 0.),
         seq.ptr, _ptr(time_masks), nt, _ptr(freq_masks), nf, _ptr(noise), _ptr(noise_scale), _stream())
    return x
