#!/usr/bin/env python
"""Summarise ncu outputs into the text files committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv        > profiles/rNN_launches.txt
    python tools/ncu_summary.py kernel   gpurun_out/prof.ncu-rep        > profiles/rNN_<kernel>.txt
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


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
