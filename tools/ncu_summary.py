#!/usr/bin/env python
"""Summarise ncu outputs into the text files committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv        > profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep        > profiles/rNN_<kernel>.txt
    python tools/ncu_summary.py pipe     gpurun_out/pipe.csv            > profiles/rNN_tensor_pipe_all_launches.txt
      (pipe.csv: ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,
       dram__bytes_read.sum,dram__bytes_write.sum --csv of one eager step: time-weighted tensor-pipe utilisation
       and DRAM traffic per kernel and over ALL tap-GEMM + weight-gradient launches)
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
           'lts__t_sector_hit_rate.pct', 'smsp__cycles_active.avg']


def launches(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        if v != v:
            continue
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(row['Metric Unit'], v)
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    print(f'# {path}: {sum(n for n, _ in agg.values())} launches, {tot / 1e3:.2f} ms total device time '
          f'(ncu: cold-cache, serialised -- compare SHARES)')
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{v:11.1f} us {100 * v / tot:5.1f}%  n={n:4d}  avg {v / n:9.1f} us  {k[:120]}')


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('## ' + r[hdr.index('Kernel Name')][:150])
        for m in METRICS:
            if m in hdr:
                print(f'   {m:70s} {r[hdr.index(m)]:>16s} {units[hdr.index(m)]}')


def pipe(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    per = collections.OrderedDict()                       # launch id -> {metric: value}
    for row in csv.DictReader(lines):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except (ValueError, KeyError):
            continue
        unit = row.get('Metric Unit', '')
        if row['Metric Name'] == 'gpu__time_duration.sum':
            v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3, 'nsecond': v / 1e3, 'usecond': v, 'msecond': v * 1e3}.get(unit, v)
        if row['Metric Name'].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        d = per.setdefault(row['ID'], {'name': re.sub(r'\(.*', '', row['Kernel Name'])})
        d[row['Metric Name']] = v
    T, P = 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'
    agg = collections.defaultdict(lambda: [0, 0., 0., 0.])
    for d in per.values():
        if T not in d:
            continue
        a = agg[d['name']]
        a[0] += 1
        a[1] += d[T]
        a[2] += d[T] * d.get(P, 0.)
        a[3] += d.get('dram__bytes_read.sum', 0.) + d.get('dram__bytes_write.sum', 0.)
    tot = sum(a[1] for a in agg.values())
    print(f'# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms device time under ncu '
          '(serialised, cold caches: compare shares).  tensor% = sm__pipe_tensor_cycles_active, time-weighted')
    print(f'# {"us":>10s} {"share":>6s} {"n":>5s} {"tensor%":>8s} {"DRAM MB/launch":>15s} {"DRAM GB/s":>10s}  kernel')
    gemm_t = gemm_p = 0.
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{a[1]:12.1f} {100 * a[1] / tot:5.1f}% {a[0]:5d} {a[2] / a[1]:8.1f} {a[3] / a[0] / 1e6:15.1f} '
              f'{a[3] / (a[1] * 1e-6) / 1e9:10.0f}  {k[:100]}')
        if re.search(r'tapgemm_tc|tapgemm_fw|wgrad_tc|wgrad_tma|wgrad_mma', k):
            gemm_t += a[1]
            gemm_p += a[2]
    if gemm_t:
        print(f'# all tensor-core tap-GEMM + weight-gradient launches: {gemm_t / 1e3:.2f} ms, time-weighted tensor pipe '
              f'{gemm_p / gemm_t:.1f} %')


def traffic(path):
    """profiles/traffic.json from the per-launch list of one eager step: measured DRAM bytes per launch of every kernel,
    keyed by the name ``pbsed_last_kernel()`` reports (bench.py copies the dominant kernel's entry into roofline.traffic)."""
    import json
    lines = [l for l in open(path) if not l.startswith('==')]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if not row.get('Metric Name', '').startswith('dram__bytes'):
            continue
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(row.get('Metric Unit', ''), 1)
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '').strip()
        m = re.match(r'(tapgemm_tc_kernel)<(\d+), (\d+), (\d+)', name)
        if m:
            name = '%s<%s,%s,%s>' % m.groups()
        elif name.startswith(('wgrad_tma_kernel', 'wgrad_bf16_kernel', 'tapgemm_fw_kernel', 'wgrad_walk_kernel')):
            name = name.split('<')[0]
        d = per.setdefault(row['ID'], [name, 0.])
        d[1] += v
    agg = collections.OrderedDict()
    for name, b in per.values():
        a = agg.setdefault(name, [0, 0.])
        a[0] += 1
        a[1] += b
    src = ('%s (ncu dram__bytes_read.sum + dram__bytes_write.sum of every launch of one eager B=32 train step, averaged '
           'per kernel)' % path)
    print(json.dumps({k: {'bytes_per_launch': a[1] / a[0], 'launches_captured': a[0], 'source': src}
                      for k, a in agg.items() if a[1] / a[0] > 1e6}, indent=1))


def stalls(path, pattern='.', skip=0):
    """source-level view of ONE launch of a --set full --import-source on capture: the most-sampled SASS
    instructions and the stall-reason totals (python tools/ncu_summary.py stalls X.ncu-rep <kernel regex> <skip>)."""
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pattern,
                          '--launch-skip', str(skip), '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        print('no source page')
        return
    print('## ' + rows[0][1][:120])
    hdr, body = rows[1], rows[2:]
    S, A = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(r[A]) for r in body if len(r) > A and r[A].isdigit())
    reasons = {h: 0 for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    for r in body:
        for h in reasons:
            try:
                reasons[h] += int(r[hdr.index(h)])
            except (ValueError, IndexError):
                pass
    rt = max(sum(reasons.values()), 1)
    print(f'   warp-state samples {tot}; by reason: ' +
          ', '.join(f'{k[6:]} {100 * v / rt:.0f}%' for k, v in sorted(reasons.items(), key=lambda kv: -kv[1]) if v > .02 * rt))
    top = sorted(((int(r[A]), i, r[S].strip()) for i, r in enumerate(body) if len(r) > A and r[A].isdigit()), reverse=True)[:14]
    for n, i, src in top:
        print(f'   {100 * n / max(tot, 1):5.1f}%  #{i:5d}  {src[:110]}')
    # synchronisation sites: samples of each mbarrier wait / barrier / tensor-core / bulk-copy / reduction instruction
    # together with the 5 instructions after it (the polling loop's branch), so "who waits for whom" can be read off
    print('   -- synchronisation / async sites (instruction + following 5):')
    smp = [int(r[A]) if len(r) > A and r[A].isdigit() else 0 for r in body]
    for i, r in enumerate(body):
        src = r[S].strip() if len(r) > S else ''
        if re.search(r'SYNCS\.PHASECHK|BAR\.SYNC|UTCHMMA|UBLKCP|UTMALDG|REDG|LDTM|UTCBAR', src):
            w = sum(smp[i:i + 6])
            if w > .003 * tot:
                print(f'   {100 * w / max(tot, 1):5.1f}%  #{i:5d}  {src[:100]}')


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel, 'pipe': pipe, 'stalls': stalls, 'traffic': traffic}[sys.argv[1]](*sys.argv[2:])
