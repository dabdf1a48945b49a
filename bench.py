#!/usr/bin/env python
"""bench.py -- 10 s-clips/sec of one FBCRNN train step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--precision fp32]

One "step" = forward + pb_sed loss + backward + (DP all-reduce) + clip + Adam on one batch of
synthetic 10 s / 16 kHz clips (STFT on the GPU, raw audio in).  N = 1 workload = BASELINE.json
configs[1] (FBCRNN 128-mel, batch 32, fp32).  N > 1: one process per GPU (torchrun), 32 clips per
GPU (weak scaling), flat-bucket NCCL all-reduce.  Prints ONE JSON line (rank 0).

--impl reference: the reference's CPU implementation of the same path (the oracle restatement:
plain PyTorch fp32 on the host cores, numpy-rfft STFT included) on a bounded sample of the workload.
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
CPU_SAMPLE_BATCH = 8


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
def cpu_reference_run(steps, warmup, batch=CPU_SAMPLE_BATCH):
    """oracle restatement of the reference path on the host cores; returns clips/s and details."""
    import torch
    from oracle import models as OM, pt_port as P
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    model = OM.build_fbcrnn(seed=0)
    opt = OM.make_adam(model)
    audio, weak, boundary = synthetic_clips(batch, 1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        spec = P.stft(audio)                                   # the reference's CPU STFT (transform.py:53)
        stft = torch.from_numpy(np.stack([spec.real, spec.imag], -1).astype(np.float32))
        b = dict(stft=stft, seq_len=[T_FRAMES] * batch, weak_targets=torch.from_numpy(weak),
                 boundary_targets=torch.from_numpy(boundary))
        OM.train_step(model, opt, b)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = float(np.mean(times)) * 1e3
    return batch / (ms / 1e3), ms, cores, batch


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 1)
    steps, warmup = min(steps, 10), min(warmup, 2)             # bounded: ~1.1 s per B=8 step on 16 cores -> ~13 s
    value, ms, cores, batch = cpu_reference_run(steps, warmup)
    sample = f'{steps} timed train steps of batch {batch} (bounded sample of the batch-32 workload), STFT included'
    print(json.dumps({
        'impl': 'reference', 'metric': 'fbcrnn_train_clips_per_sec', 'value': value, 'unit': '10s-clips/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'FBCRNN shallow (3.49M params), 128 mel, 10 s / 16 kHz clips, K=10, train step; '
                               f'CPU sample batch {batch}'},
        'cpu_baseline': {'value': value, 'unit': '10s-clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': '10s-clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ------------------------------------------------------------------------------ GPU arm
def tapgemm_flops(args):
    d = args[0]._obj if hasattr(args[0], '_obj') else None
    if d is None:
        return 0.
    return 2. * d.B * d.F_out * d.T * d.ntaps * d.Cin * d.Cout


def kernel_breakdown(model, opt, batch, train, _lib):
    """one eager, per-call-timed train step: device time + algorithmic FLOPs per C-ABI entry point."""
    import torch
    from pb_sed_b200 import ops
    ops.enable_wgrad_stream(False)          # isolate the per-call timings (no concurrent side-stream kernels)
    sink = []
    _lib.profile_sink = sink
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    train.train_step(model, opt, batch)
    t1.record()
    _lib.profile_sink = None
    ops.enable_wgrad_stream(True)
    torch.cuda.synchronize()
    agg = {}
    detail = []
    for name, e0, e1, a in sink:
        ms = e0.elapsed_time(e1)
        key = name
        if name in ('pbsed_tapgemm', 'pbsed_tapgemm_wgrad'):
            d = a[0]._obj
            detail.append((name, f'B{d.B} F{d.F_in}>{d.F_out} T{d.T} C{d.Cin}>{d.Cout} taps{d.ntaps} '
                                 f'relu{d.relu} ws{d.w_sn}', round(ms, 3),
                           round(tapgemm_flops(a) / (ms / 1e3) / 1e12, 1)))
        else:
            detail.append((name, '', round(ms, 3), None))
    if os.environ.get('PBSED_BENCH_DETAIL'):
        with open(os.environ['PBSED_BENCH_DETAIL'], 'w') as f:
            for row in detail:
                f.write(' '.join(str(x) for x in row) + '\n')
    for name, e0, e1, a in sink:
        ms = e0.elapsed_time(e1)
        key = name
        fl = tapgemm_flops(a) if name in ('pbsed_tapgemm', 'pbsed_tapgemm_wgrad') else 0.
        r = agg.setdefault(key, {'ms': 0., 'calls': 0, 'flop': 0.})
        r['ms'] += ms
        r['calls'] += 1
        r['flop'] += fl
    return agg, t0.elapsed_time(t1)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from pb_sed_b200 import _lib, config, train, ops
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
    ops.set_default_precision(args.precision)
    B = args.batch
    torch.manual_seed(0)
    model = weak_label.CRNN.from_config_dict(config.fbcrnn_config(
        num_events=args.num_events, strong_fwd_bwd_loss_weight=args.strong_weight)).to(dev)
    model.emit_buffers = False
    opt = train.Adam(model, lr=args.lr, gradient_clipping=args.grad_clip, sync_stats=args.sync_stats)

    n_sets = 3
    host = []
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
    step = train.GraphedTrainStep(model, opt, resident[0], warmup=1)
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
        step(resident[i % n_sets])

    def e2e_step(i):
        # public pipeline API: the H2D copy of batch i+1 (pinned host memory, copy stream) overlaps the
        # compute of batch i; every step's inputs cross PCIe inside the timed region, its loss comes back
        step.step_prefetched()
        step.prefetch(host[(i + 1) % n_sets])
        loss_host.copy_(step.loss.reshape(1), non_blocking=True)  # D2H of the step's loss

    for i in range(args.warmup):
        resident_step(i)
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(resident_step, args.steps)
    clocks = sampler.result()
    step.prefetch(host[0])
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    final_loss = float(step.loss)
    assert np.isfinite(final_loss), final_loss

    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    out = None
    if rank == 0:
        hbm, tf_burst, tf_sus, which = measured_peaks()
        out = {
            'metric': 'fbcrnn_train_clips_per_sec', 'value': value, 'unit': '10s-clips/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': {'fp32': 'f32', 'tf32x3': 'f32 via 3xTF32 split (tcgen05, fp32 accumulate)',
                      'tf32': 'tf32 single pass (tcgen05, fp32 accumulate; reduced precision >= bf16 mantissa)'}[args.precision],
            'data': 'synthetic',
            'config': {'workload': (f'BASELINE configs[1]: FBCRNN shallow (3.49M params) 128-mel, batch {B}/GPU of 10 s / '
                                    '16 kHz clips, K=10, fp32, full train step (GPU STFT+logmel, CNN, fwd+bwd GRU, '
                                    'pb_sed loss, backward, clip+Adam)') if (args.num_events == 10 and B == 32) else
                                   (f'FBCRNN shallow 128-mel, batch {B}/GPU of 10 s / 16 kHz clips, K={args.num_events}, '
                                    f'strong_fwd_bwd_loss_weight={args.strong_weight}, gradient_clipping={args.grad_clip}, '
                                    f'lr={args.lr}, full train step'),
                       'global_batch': B * world, 'parallelism': f'dp{world}', 'precision': args.precision,
                       'sync_stats': args.sync_stats if world > 1 else 'n/a (1 GPU)',
                       'l2': f'{n_sets} distinct input batches rotate; per-step activation working set (several GB) '
                             '>> 126 MB L2', 'cuda_graph': True, 'final_loss': final_loss},
            'e2e': {'value': e2e, 'unit': '10s-clips/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches_per_step * args.steps,
            'clocks': clocks,
            'step_tflops': value / world * FLOP_PER_CLIP_TRAIN / 1e12,
        }
    if rank == 0 and world == 1:
        agg, eager_ms = kernel_breakdown(model, opt, resident[1], train, _lib)
        tot = sum(r['ms'] for r in agg.values())
        top = max(agg.items(), key=lambda kv: kv[1]['ms'])
        gemm_ms = sum(agg[k]['ms'] for k in ('pbsed_tapgemm', 'pbsed_tapgemm_wgrad') if k in agg)
        gemm_fl = sum(agg[k]['flop'] for k in ('pbsed_tapgemm', 'pbsed_tapgemm_wgrad') if k in agg)
        name, r = top
        achieved = r['flop'] / (r['ms'] / 1e3) / 1e12 if r['flop'] else None
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.isfile(tpath):
            tj = json.load(open(tpath)).get(name)
            if tj:
                traffic, traffic_src = tj['bytes_per_launch'], tj['source']
        out['roofline'] = {
            'kernel': name, 'bound': 'tensor', 'achieved': achieved, 'peak': tf_sus, 'unit': 'TFLOP/s',
            'frac': (achieved / tf_sus) if achieved else None, 'traffic': traffic, 'traffic_source': traffic_src,
            'executed_tensor_tflops': (3. * achieved if (achieved and args.precision == 'tf32x3') else achieved),
            'executed_frac_of_tf32_peak': ((3. if args.precision == 'tf32x3' else 1.) * achieved / (tf_sus / 2.)) if achieved else None,
            'peak_source': f'{which} bf16 dense sustained (MEASURED_PEAKS.json); the fp32 config runs '
                           f'{args.precision} arithmetic, TF32 nominal peak is half of bf16',
            'avg_launch_ms': r['ms'] / r['calls'], 'launches_per_step': r['calls'],
            'share_of_step': r['ms'] / tot,
            'all_tapgemm_tflops': gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms else None,
        }
        out['kernel_breakdown_ms'] = {k: {'ms': round(v['ms'], 3), 'calls': v['calls'],
                                          'tflops': round(v['flop'] / (v['ms'] / 1e3) / 1e12, 2) if v['flop'] else None}
                                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1]['ms'])}
        out['eager_step_ms'] = eager_ms
        # per-stage rooflines the north star asks for (BASELINE.md algorithmic bytes per clip)
        st = {}
        if 'pbsed_stft_logmel' in agg:
            gbs = 896000. * B / (agg['pbsed_stft_logmel']['ms'] / 1e3) / 1e9
            st['stft_logmel'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                                 'note': 'fp32 FFT: issue-bound (profiles/r01_s4_logmel1024_ncu_full.txt), see DESIGN.md'}
        if 'pbsed_gru_fwd' in agg:
            gbs = 8.19e6 * B / (agg['pbsed_gru_fwd']['ms'] / 1e3) / 1e9
            st['gru_fwd'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm, 'unit': 'GB/s', 'frac': gbs / hbm,
                             'note': '500 dependent time steps per layer: latency-bound'}
        out['stage_rooflines'] = st
        if not args.no_cpu_baseline:
            v, ms, cores, cb = cpu_reference_run(8, 2)
            out['cpu_baseline'] = {'value': v, 'unit': '10s-clips/s', 'cores': cores, 'kind': 'port',
                                   'sample': f'8 timed oracle train steps of batch {cb} after 2 warm-up steps (STFT '
                                             f'included), {ms:.0f} ms each: a bounded sample of the batch-32 workload'}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        step.close()
        dist.barrier()
        dist.destroy_process_group()


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
                      'tf32': 'tf32 single pass (tcgen05, fp32 accumulate)'}[args.precision],
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
    ap.add_argument('--precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--num-events', type=int, default=NUM_EVENTS, help='K (527 = AudioSet, training.py:128)')
    ap.add_argument('--strong-weight', type=float, default=1., help='strong_fwd_bwd_loss_weight (0 for AudioSet, :151)')
    ap.add_argument('--grad-clip', type=float, default=1e10, help='gradient_clipping (0.1 for AudioSet, :150)')
    ap.add_argument('--lr', type=float, default=5e-4)
    ap.add_argument('--workload', default='fbcrnn_train', choices=['fbcrnn_train', 'bicrnn_infer'],
                    help='fbcrnn_train = BASELINE configs[1] (the headline); bicrnn_infer = configs[3]')
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
