#!/bin/bash
# final evidence refresh for the last kernel changes: full-step launch list (tensor pipe + DRAM bytes), traffic.json,
# --set full capture of the dominant kernel, headline bench line.   usage: bash tools/gpu_final.sh TAG
TAG=$1
O=gpurun_out
mkdir -p $O /tmp/rep
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 330 --csv --log-file $O/${TAG}_pipe_full_step.csv python tools/run_step.py 3 > /tmp/rep/ncu1.log 2>&1
python tools/ncu_summary.py pipe $O/${TAG}_pipe_full_step.csv > $O/${TAG}_tensor_pipe_all_launches.txt
python tools/ncu_summary.py traffic $O/${TAG}_pipe_full_step.csv > $O/${TAG}_traffic.json && cp $O/${TAG}_traffic.json profiles/traffic.json
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:wgrad_tma_kernel" -s 23 -c 23 -o /tmp/rep/wg python tools/run_step.py 2 > /tmp/rep/wg.log 2>&1
python tools/ncu_summary.py kernel /tmp/rep/wg.ncu-rep > $O/${TAG}_wgrad_tma_ncu_full.txt 2>/dev/null
for sk in 0 16 19 20; do python tools/ncu_summary.py stalls /tmp/rep/wg.ncu-rep wgrad_tma_kernel $sk >> $O/${TAG}_wgrad_tma_ncu_full.txt 2>/dev/null; done
PBSED_BENCH_DETAIL=$O/${TAG}_detail.txt timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 300 python bench.py --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_b256.json 2>> $O/${TAG}_bench.err
python - <<PY
import json
for f in ('bench', 'bench_b256'):
    try:
        d = json.load(open('$O/${TAG}_%s.json' % f)); print(f, round(d['value'], 1), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value'], 1), 'eager', d.get('eager_step_ms'), d.get('roofline', {}).get('kernel'), d.get('roofline', {}).get('frac'))
    except Exception as e: print(f, 'FAILED', e)
PY
head -8 $O/${TAG}_tensor_pipe_all_launches.txt | cut -c1-140; tail -1 $O/${TAG}_tensor_pipe_all_launches.txt
