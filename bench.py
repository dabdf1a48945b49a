#!/usr/bin/env python
"""bench.py -- 10 s-clips/sec of one FBCRNN train step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision fp32]

One "step" = forward + pb_sed loss + backward + (DP all-reduce) + clip + Adam on one batch of
synthetic 10 s / 16 kHz clips (STFT on the GPU, raw audio in).  N = 1 workload = BASELINE.json
configs[1] (FBCRNN 128-mel, batch 32, fp32).  N > 1: one process per GPU (torchrun), 32 clips per
GPU (weak scaling), flat-bucket NCCL all-reduce.  Prints ONE JSON line (rank 0).

--impl reference: the reference's CPU implementation of the same path (the oracle restatement:
plain PyTorch fp32 on all host cores, numpy-rfft STFT included) on the SAME batch size, for the
requested --steps / --warmup.

Other workloads (not the headline): --workload bicrnn_infer (configs[3]), --workload audioset_stream
(configs[4]: K = 527, weak labels only, clip 0.1, lr 1e-4, batch 128/GPU, clips generated on the device by a
counter-based stream), --global-batch G (strong scaling: G clips split over the N ranks).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_SAMPLES, NUM_EVENTS, T_FRAMES = 160000, 10, 500
FLOP_PER_CLIP_TRAIN = 35.26e9          # fwd 11.75 GFLOP, fwd+bwd 35.26 GFLOP per clip


def synthetic_clips(batch, seed, num_events=None):
    """low-passed noise + gated sinusoid events, peak normalised; float32 (B,1,S)."""
    NUM_EVENTS = num_events or globals()['NUM_EVENTS']
    rng = np.random.RandomState(seed)
    x = rng.randn(batch, NUM_SAMPLES)
    spec = np.fft.rfft(x, axis=-1)
    a = rng.uniform(.5, .98, size=(batch, 1))
    w = np.exp(-2j * np.pi * np.arange(spec.shape[-1]) / NUM_SAMPLES)[None]
    y = np.fft.irfft(spec * (1 - a) / (1 - a * w), n=NUM_SAMPLES, axis=-1)
    t = np.arange(NUM_SAMPLES) / 16000.
    for b in range(batch):
        for _ in range(rng.randint(1, 4)):
            on = rng.randint(0, NUM_SAMPLES - 1600)
            off = min(on + rng.randint(1600, NUM_SAMPLES // 2), NUM_SAMPLES)
            y[b, on:off] += rng.uniform(.2, 2.) * y[b].std() * np.sin(2 * np.pi * rng.uniform(200., 6000.) * t[on:off])
    y /= np.abs(y).max(-1, keepdims=True)
    weak = (rng.rand(batch, NUM_EVENTS) < .2).astype(np.float32)
    boundary = np.zeros((batch, NUM_EVENTS, T_FRAMES), np.float32)
    for b in range(batch):
        if weak[b].sum() == 0:
            weak[b, rng.randint(NUM_EVENTS)] = 1.
        for k in np.nonzero(weak[b])[0]:
            on = rng.randint(0, T_FRAMES - 1)
            boundary[b, k, on:rng.randint(on + 1, T_FRAMES + 1)] = 1.
    return y.astype(np.float32)[:, None], weak, boundary


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'sw_thermal_slowdown': 0x20,
                 'hw_thermal_slowdown': 0x40, 'hw_power_brake_slowdown': 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.reasons |= {n for n, bit in names.items() if r & bit}
            except Exception:
                pass
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        d = json.load(open(path))
        return d.get('hbm_gbs', 6650.), d.get('bf16_tflops', 1590.), d.get('bf16_tflops_sustained', 1400.), 'measured'
    return 6650., 1590., 1400., 'fallback'


# ------------------------------------------------------------------------------ reference / CPU arm
def cpu_reference_run(steps, warmup, batch, seed=1234, parity_out=None):
    """oracle restatement of the reference path on the host cores (all of them); returns clips/s and details.
    parity_out (dict): also keep the frame logits / loss of the FIRST step's forward (fresh seed-0 weights) so
    that the GPU arm can compare its own first step on the same batch against them."""
    import torch
    from oracle import models as OM, pt_port as P
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    model = OM.build_fbcrnn(seed=0)
    opt = OM.make_adam(model)
    audio, weak, boundary = synthetic_clips(batch, seed)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        spec = P.stft(audio)                                   # the reference's CPU STFT (transform.py:53)
        stft = torch.from_numpy(np.stack([spec.real, spec.imag], -1).astype(np.float32))
        b = dict(stft=stft, seq_len=[T_FRAMES] * batch, weak_targets=torch.from_numpy(weak),
                 boundary_targets=torch.from_numpy(boundary))
        if it == 0 and parity_out is not None:
            model.train()
            with torch.no_grad():
                ref = OM.build_fbcrnn(seed=0).train()
                z_fwd, z_bwd, seq_len_y, _, _ = ref.logits(b)
                parity_out['z_fwd'], parity_out['z_bwd'] = z_fwd, z_bwd
                parity_out['loss'] = float(ref.loss(ref.sigmoid(z_fwd), ref.sigmoid(z_bwd), seq_len_y,
                                                    (b['weak_targets'], b['boundary_targets'])))
        OM.train_step(model, opt, b)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = float(np.mean(times)) * 1e3
    return batch / (ms / 1e3), ms, cores, batch


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    batch = args.batch
    value, ms, cores, batch = cpu_reference_run(steps, warmup, batch)
    sample = (f'{steps} timed train steps of batch {batch} after {warmup} warm-up steps (the same batch size and step '
              f'count as the GPU arm), numpy-rfft STFT included, torch.set_num_threads({cores})')
    print(json.dumps({
        'impl': 'reference', 'metric': 'fbcrnn_train_clips_per_sec', 'value': value, 'unit': '10s-clips/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name(args, batch), 'global_batch': batch, 'parallelism': 'cpu',
                   'precision': 'fp32 (torch CPU kernels)'},
        'cpu_baseline': {'value': value, 'unit': '10s-clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': '10s-clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_name(args, B):
    if args.num_events == 10 and B == 32:
        return ('BASELINE configs[1]: FBCRNN shallow (3.49M params) 128-mel, batch 32/GPU of 10 s / 16 kHz clips, '
                'K=10, fp32, full train step (STFT+logmel, CNN, fwd+bwd GRU, pb_sed loss, backward, clip+Adam)')
    return (f'FBCRNN shallow 128-mel, batch {B}/GPU of 10 s / 16 kHz clips, K={args.num_events}, '
            f'strong_fwd_bwd_loss_weight={args.strong_weight}, gradient_clipping={args.grad_clip}, '
            f'lr={args.lr}, full train step')


# ------------------------------------------------------------------------------ GPU arm
def tapgemm_flops(args):
    """algorithmic FLOPs of one tap-GEMM / weight-gradient call: 2 x Cin x Cout per VALID (output row, tap)
    pair -- a tap whose source row f = fo + df or source frame t + dt falls outside the map is zero padding
    and counts nothing (the transposed flatten data gradient F 1 -> 8 has ONE valid tap per output row)."""
    d = args[0]._obj if hasattr(args[0], '_obj') else None
    if d is None:
        return 0.
    pairs = 0
    for i in range(d.ntaps):
        df, dt = d.df[i], d.dt[i]
        n_f = sum(1 for fo in range(d.F_out) if 0 <= fo + df < d.F_in)
        pairs += n_f * max(d.T - abs(dt), 0)
    return 2. * d.B * pairs * d.Cin * d.Cout


def measure_tensor_peaks(dev):
    """dense GEMM throughput of this GPU, measured in this run with cuBLAS through torch: the TF32 figure
    is the roofline denominator of the kind::tf32 kernels ('measure it'), bf16 for reference."""
    import torch
    out = {}
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    for name, dt, tf32 in (('tf32', torch.float32, True), ('bf16', torch.bfloat16, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a = torch.randn(n, n, device=dev, dtype=dt)
        b = torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            a @ b
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 20
        for _ in range(reps):
            a @ b
        e1.record()
        torch.cuda.synchronize()
        out[name] = 2. * n ** 3 * reps / (e0.elapsed_time(e1) / 1e3) / 1e12
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = old
    return out


def kernel_breakdown(model, opt, batch, train, _lib):
    """one eager, per-call-timed train step: device time + algorithmic FLOPs per C-ABI entry point AND per
    main kernel (``pbsed_last_kernel``)."""
    import torch
    from pb_sed_b200 import ops
    # the un-graphed drop-in step (what a stock trainer loop runs): 1 warm-up (allocator), then the MEDIAN of 5 timed
    # steps -- the eager step is bound by host issue time, and one descheduled Python thread tripled a 3-step mean
    train.train_step(model, opt, batch)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        train.train_step(model, opt, batch)
        t1.record()
        torch.cuda.synchronize()
        times.append(t0.elapsed_time(t1))
    eager_ms = sorted(times)[len(times) // 2]
    ops.enable_wgrad_stream(False)          # isolate the per-call timings (no concurrent side-stream kernels)
    sink = []
    _lib.profile_sink = sink
    train.train_step(model, opt, batch)
    _lib.profile_sink = None
    ops.enable_wgrad_stream(True)
    torch.cuda.synchronize()
    agg, kern, detail = {}, {}, []
    for name, e0, e1, a, kname in sink:
        ms = e0.elapsed_time(e1)
        gemm = name in ('pbsed_tapgemm', 'pbsed_tapgemm_wgrad')
        fl = tapgemm_flops(a) if gemm else 0.
        if gemm:
            d = a[0]._obj
            detail.append((name, kname, f'B{d.B} F{d.F_in}>{d.F_out} T{d.T} C{d.Cin}>{d.Cout} taps{d.ntaps} '
                                        f'relu{d.relu} ws{d.w_sn}', round(ms, 3), round(fl / (ms / 1e3) / 1e12, 1)))
        else:
            detail.append((name, '', '', round(ms, 3), None))
        for table, key in ((agg, name), (kern, kname if gemm else name)):
            r = table.setdefault(key, {'ms': 0., 'calls': 0, 'flop': 0.})
            r['ms'] += ms
            r['calls'] += 1
            r['flop'] += fl
    if os.environ.get('PBSED_BENCH_DETAIL'):
        with open(os.environ['PBSED_BENCH_DETAIL'], 'w') as f:
            for row in detail:
                f.write(' '.join(str(x) for x in row) + '\n')
    return agg, kern, eager_ms


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pb_sed_b200 import _lib, config, train, ops, data
    from pb_sed_b200.models import weak_label
    _lib.load()
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus or world == 1, (world, args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    stream_mode = args.workload == 'audioset_stream'
    if stream_mode:              # BASELINE configs[4] (training.py:128,139,150-151): AudioSet settings
        args.num_events, args.strong_weight, args.grad_clip, args.lr = 527, 0., .1, 1e-4
        if args.batch == 32:
            args.batch = 128
    ops.set_default_precision(args.precision)
    B = args.batch
    scaling = 'weak'
    if args.global_batch:        # strong scaling: the global batch is fixed, every rank takes its shard
        lo, hi = train.shard_bounds(args.global_batch, rank, world)
        B, scaling = hi - lo, 'strong'
    torch.manual_seed(0)
    torch.cuda.manual_seed(1234 + rank)
    model = weak_label.CRNN.from_config_dict(config.fbcrnn_config(
        num_events=args.num_events, strong_fwd_bwd_loss_weight=args.strong_weight)).to(dev)
    model.emit_buffers = False
    opt = train.Adam(model, lr=args.lr, gradient_clipping=args.grad_clip, sync_stats=args.sync_stats)

    n_sets = 3
    host, resident, h2d, stream = [], [], 0, None
    if stream_mode:
        stream = data.SyntheticClipStream(B, args.num_events, dev)
        resident = [stream.example()]
    else:
        for i in range(n_sets):
            audio, weak, boundary = synthetic_clips(B, 1234 + 1000 * rank + i, args.num_events)
            host.append({'audio_data': torch.from_numpy(audio).pin_memory(),
                         'weak_targets': torch.from_numpy(weak).pin_memory(),
                         'boundary_targets': torch.from_numpy(boundary).pin_memory()})
        resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
        for r in resident:
            r['seq_len'] = [T_FRAMES] * B
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())

    n0 = _lib.launch_count()
    step = train.GraphedTrainStep(model, opt, resident[0], warmup=1,
                                  input_fn=stream.fill_ if stream_mode else None)
    launches_per_step = (_lib.launch_count() - n0) // 2          # 1 eager warm-up + 1 capture
    loss_host = torch.zeros(1).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    def resident_step(i):
        step(None if stream_mode else resident[i % n_sets])

    def e2e_step(i):
        # public pipeline API: the H2D copy of batch i+1 (pinned host memory, copy stream) overlaps the
        # compute of batch i; every step's inputs cross PCIe inside the timed region, its loss comes back
        if stream_mode:
            step()               # inputs are generated on the device by the captured stream: nothing to upload
        else:
            step.step_prefetched()
            step.prefetch(host[(i + 1) % n_sets])
        loss_host.copy_(step.loss.reshape(1), non_blocking=True)  # D2H of the step's loss

    for i in range(args.warmup):
        resident_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(resident_step, args.steps)
    clocks = sampler.result()
    if not stream_mode:
        step.prefetch(host[0])
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    final_loss = float(step.loss)
    assert np.isfinite(final_loss), final_loss

    ms_step = ms_total / args.steps
    clips = torch.tensor([float(B)], device=dev)
    if world > 1:
        dist.all_reduce(clips)
    gB = int(clips.item())
    value = gB * args.steps / (ms_total / 1e3)
    e2e = gB * args.steps / (ms_e2e / 1e3)
    out = None
    if rank == 0:
        hbm, tf_burst, tf_sus, which = measured_peaks()
        out = {
            'metric': 'fbcrnn_train_clips_per_sec', 'value': value, 'unit': '10s-clips/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': scaling, 'vs_baseline': None, 'dtype': {'fp32': 'f32', 'tf32x3': 'f32 via 3xTF32 split (tcgen05, fp32 accumulate)',
                      'tf32': 'tf32 single pass (tcgen05, fp32 accumulate; reduced precision >= bf16 mantissa)',
                      'bf16': 'bf16 (conv-stack activations + gradients stored as bf16 in HBM; tcgen05 kind::f16 bf16 x bf16 MMAs on the wide '
                              'layers, one kind::tf32 pass on the narrow ones, fp32 accumulate; fp32 master weights, statistics, GRU, '
                              'optimizer)'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': ('BASELINE configs[4] shape: ' if stream_mode else '') + workload_name(args, B),
                       'global_batch': gB, 'parallelism': f'dp{world}', 'precision': args.precision,
                       'sync_stats': args.sync_stats if world > 1 else 'n/a (1 GPU)',
                       'inputs': ('generated on the device inside the captured step by data.SyntheticClipStream '
                                  '(counter-based Philox stream, a new batch every step)') if stream_mode else
                                 'pinned host batches (3 rotate)',
                       'l2': f'{1 if stream_mode else n_sets} distinct input batches; per-step activation working set '
                             '(several GB) >> 126 MB L2', 'cuda_graph': True, 'final_loss': final_loss},
            'e2e': {'value': e2e, 'unit': '10s-clips/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches_per_step * args.steps,
            'clocks': clocks,
            'step_tflops': value / world * FLOP_PER_CLIP_TRAIN / 1e12,
        }
        if stream_mode:
            out['clips_streamed'] = gB * (args.steps * 2 + args.warmup + 4)       # every replay drew a new batch
    if rank == 0 and world == 1 and not stream_mode:
        peaks = measure_tensor_peaks(dev)
        agg, kern, eager_ms = kernel_breakdown(model, opt, resident[1], train, _lib)
        tot = sum(r['ms'] for r in agg.values())
        passes = 3. if args.precision == 'tf32x3' else 1.
        # the dominant KERNEL (not entry point) by device time; its algorithmic FLOPs against the TF32 dense
        # peak measured in this run
        name, r = max(((k, v) for k, v in kern.items() if v['flop']), key=lambda kv: kv[1]['ms'])
        achieved = r['flop'] / (r['ms'] / 1e3) / 1e12
        gemm = [v for v in kern.values() if v['flop']]
        gemm_ms, gemm_fl = sum(v['ms'] for v in gemm), sum(v['flop'] for v in gemm)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(tpath):
            tj = json.load(open(tpath)).get(name)
            if tj:
                traffic, traffic_src = tj['bytes_per_launch'], tj['source']
        peak = peaks['tf32'] if args.precision != 'fp32' else None
        out['roofline'] = {
            'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
            'frac': achieved / peak if peak else None, 'traffic': traffic, 'traffic_source': traffic_src,
            'peak_source': 'dense TF32 GEMM (cuBLAS via torch.matmul, 8192^3) measured in this run; '
                           f'MEASURED_PEAKS.json ({which}) bf16 burst/sustained = {tf_burst:.0f}/{tf_sus:.0f}, '
                           f'bf16 measured in this run = {peaks["bf16"]:.0f} TFLOP/s',
            'algorithmic_flop_per_launch': r['flop'] / r['calls'],
            'avg_launch_ms': r['ms'] / r['calls'], 'launches_per_step': r['calls'],
            'share_of_step': r['ms'] / tot,
            'tensor_passes': passes, 'executed_frac': passes * achieved / peak if peak else None,
            'note': 'achieved = ALGORITHMIC FLOPs (2*Cin*Cout per valid output-row x tap pair) / CUDA-event time of the '
                    'launches of this kernel in one eager step; the fp32 configuration executes 3 TF32 passes per '
                    'product (hi*hi + lo*hi + hi*lo), so the tensor pipe does tensor_passes x that work',
            'all_tapgemm': {'tflops': gemm_fl / (gemm_ms / 1e3) / 1e12, 'ms': gemm_ms,
                            'frac': gemm_fl / (gemm_ms / 1e3) / 1e12 / peak if peak else None},
        }
        out['measured_peaks_this_run'] = peaks
        out['kernel_breakdown_ms'] = {k: {'ms': round(v['ms'], 3), 'calls': v['calls'],
                                          'tflops': round(v['flop'] / (v['ms'] / 1e3) / 1e12, 2) if v['flop'] else None}
                                      for k, v in sorted(kern.items(), key=lambda kv: -kv[1]['ms'])}
        out['entry_point_ms'] = {k: round(v['ms'], 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])}
        out['eager_step_ms'] = eager_ms
        # per-stage rooflines the north star asks for (BASELINE.md algorithmic bytes per clip)
        st = {}
        if 'pbsed_stft_logmel' in agg:
            gbs = 896000. * B / (agg['pbsed_stft_logmel']['ms'] / 1e3) / 1e9
            st['stft_logmel'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm}
        if 'pbsed_gru_fwd' in agg:
            gbs = 8.19e6 * B / (agg['pbsed_gru_fwd']['ms'] / 1e3) / 1e9
            st['gru_fwd'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                             'note': '500 dependent time steps per layer: latency-bound'}
        out['stage_rooflines'] = st
        if not args.no_cpu_baseline:
            # CPU baseline leg: the oracle restatement on the SAME first batch; its first forward also yields the
            # frame logits / loss this GPU path is checked against (BASELINE metric: clips/s AND logit max|delta|)
            par = {}
            v, ms, cores, cb = cpu_reference_run(4, 1, B, seed=1234, parity_out=par)
            out['cpu_baseline'] = {'value': v, 'unit': '10s-clips/s', 'cores': cores, 'kind': 'port',
                                   'sample': f'4 timed oracle train steps of batch {cb} (the benchmark batch) after 1 warm-up '
                                             f'step, STFT included, {ms:.0f} ms each'}
            if B <= 64 and args.num_events == 10:
                out['parity'] = gpu_parity(args, dev, host[0], par)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        step.close()
        dist.barrier()
        dist.destroy_process_group()


def gpu_parity(args, dev, host_batch, par):
    """first train-mode forward + loss of a fresh seed-0 model on the first timed batch, GPU vs the CPU oracle
    (``par`` from cpu_reference_run): frame-logit max|delta| (BASELINE.json: <= 1e-3) and loss |delta|."""
    import torch
    from oracle import models as OM
    from pb_sed_b200 import config
    from pb_sed_b200.models import weak_label
    ora = OM.build_fbcrnn(seed=0)
    model = weak_label.CRNN.from_config_dict(config.fbcrnn_config(
        num_events=args.num_events, strong_fwd_bwd_loss_weight=args.strong_weight))
    model.load_state_dict(ora.state_dict())
    model.to(dev).train()
    model.emit_buffers = False
    b = {k: v.to(dev) for k, v in host_batch.items()}
    b['seq_len'] = [T_FRAMES] * b['audio_data'].shape[0]
    with torch.no_grad():
        outp = model(dict(b))
        loss = float(model.review(b, outp)['loss'])
    z_fwd = model._z_fwd.transpose(1, 2).cpu()
    z_bwd = model._z_bwd.transpose(1, 2).cpu()
    d = max(float((z_fwd - par['z_fwd']).abs().max()), float((z_bwd - par['z_bwd']).abs().max()))
    return {'logit_max_abs': d, 'loss_abs': abs(loss - par['loss']), 'logit_abs_max_value': float(par['z_fwd'].abs().max()),
            'tolerance': 1e-3, 'batch': int(z_fwd.shape[0]), 'precision': args.precision,
            'what': 'frame logits (pre-sigmoid output_net outputs, both directions) and review loss of the first '
                    'train-mode step on the first timed batch, fresh seed-0 weights, vs the CPU oracle'}


# ------------------------------------------------------------------------------ BASELINE configs[3]
def run_bicrnn_infer(args):
    """strong_label_crnn tag-conditioned BiCRNN inference (BASELINE.json configs[3]): frame-score
    throughput, eval mode, raw audio in, scores + post-processing out.  N > 1: independent replicas
    (clips shard, no collective)."""
    import torch
    from pb_sed_b200 import _lib, config, ops, filters as GF
    from pb_sed_b200.models import strong_label
    _lib.load()
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    ops.set_default_precision(args.precision)
    B = args.batch if args.batch != 32 else 512
    torch.manual_seed(0)
    model = strong_label.CRNN.from_config_dict(config.bicrnn_config(num_events=NUM_EVENTS)).to(dev).eval()
    chunk = 64
    host = []
    for i in range(2):
        audio, weak, _ = synthetic_clips(chunk, 77 + i + 10 * rank)
        audio = np.tile(audio, (B // chunk, 1, 1))
        host.append({'audio_data': torch.from_numpy(audio).pin_memory(),
                     'tag_condition': torch.from_numpy(np.tile(weak, (B // chunk, 1)) > .5).pin_memory()})
    seq_len = [T_FRAMES] * B
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    med = np.array(NUM_EVENTS * [21])

    def infer(batch):
        with torch.no_grad():
            y, sl = model.sound_event_detection(dict(batch, seq_len=seq_len))
            return GF.post_process(y.contiguous(), sl, medfilt_length=med)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n0 = _lib.launch_count()
    for i in range(max(args.warmup, 3)):
        infer(resident[i % 2])
    sync()
    launches = (_lib.launch_count() - n0) // max(args.warmup, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    e0.record()
    for i in range(args.steps):
        infer(resident[i % 2])
    e1.record()
    sync()
    clocks = sampler.result()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    out_host = torch.empty((B, NUM_EVENTS, T_FRAMES)).pin_memory()
    e0.record()
    for i in range(args.steps):
        out_host.copy_(infer({k: v.to(dev, non_blocking=True) for k, v in host[i % 2].items()}), non_blocking=True)
    e1.record()
    sync()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    if rank == 0:
        v = world * B * args.steps / (float(ms) / 1e3)
        print(json.dumps({
            'metric': 'bicrnn_inference_clips_per_sec', 'value': v, 'unit': '10s-clips/s', 'frames_per_sec': v * T_FRAMES,
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': float(ms) / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': {'fp32': 'f32', 'tf32x3': 'f32 via 3xTF32 split (tcgen05, fp32 accumulate)',
                      'tf32': 'tf32 single pass (tcgen05, fp32 accumulate)',
                      'bf16': 'bf16 activation maps in HBM, fp32 accumulate'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': f'BASELINE configs[3]: strong_label tag-conditioned BiCRNN inference, batch {B}/GPU of '
                                   '10 s / 16 kHz clips, eval mode, raw audio -> frame scores -> sequence mask + '
                                   'median filter (21) on the GPU', 'parallelism': f'replicas x{world}',
                       'l2': 'per-step activations (GBs) >> 126 MB L2; 2 input batches rotate'},
            'e2e': {'value': world * B * args.steps / (float(ms2) / 1e3), 'unit': '10s-clips/s',
                    'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': out_host.numel() * 4},
            'gpu_launches': launches * args.steps, 'clocks': clocks}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32', 'bf16'],
                    help="tf32x3 = fp32-equivalent split (headline); tf32 = one TF32 pass; bf16 = bf16 activation maps in HBM, "
                         "bf16 x bf16 products, fp32 accumulation (BASELINE configs[2] / [4])")
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--num-events', type=int, default=NUM_EVENTS, help='K (527 = AudioSet, training.py:128)')
    ap.add_argument('--strong-weight', type=float, default=1., help='strong_fwd_bwd_loss_weight (0 for AudioSet, :151)')
    ap.add_argument('--grad-clip', type=float, default=1e10, help='gradient_clipping (0.1 for AudioSet, :150)')
    ap.add_argument('--lr', type=float, default=5e-4)
    ap.add_argument('--workload', default='fbcrnn_train', choices=['fbcrnn_train', 'bicrnn_infer', 'audioset_stream'],
                    help='fbcrnn_train = BASELINE configs[1] (the headline); bicrnn_infer = configs[3]; '
                         'audioset_stream = configs[4] shape with an on-device clip stream')
    ap.add_argument('--global-batch', type=int, default=0,
                    help='strong scaling: fixed global batch split over the ranks (0 = weak scaling, --batch per GPU)')
    ap.add_argument('--sync-stats', default='none', choices=['none', 'exact'],
                    help="N > 1: per-replica batch statistics ('none') or all-reduced ('exact', SURVEY 8e)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload == 'bicrnn_infer':
        run_bicrnn_infer(args)
    else:
        run_gpu(args)


if __name__ == '__main__':
    main()
