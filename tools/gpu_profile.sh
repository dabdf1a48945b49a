#!/bin/bash
# evidence pass (1 GPU): launch list with tensor-pipe + DRAM bytes of a full eager step, --set full captures of the
# dominant kernels (summarised on the box: the .ncu-rep files are too big to bring back), bench lines of the other
# single-GPU workloads.   usage: bash tools/gpu_profile.sh TAG
TAG=$1
O=gpurun_out
mkdir -p $O /tmp/rep
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 330 --csv --log-file $O/${TAG}_pipe_full_step.csv python tools/run_step.py 3 > /tmp/rep/ncu1.log 2>&1
python tools/ncu_summary.py pipe $O/${TAG}_pipe_full_step.csv > $O/${TAG}_tensor_pipe_all_launches.txt
python tools/ncu_summary.py traffic $O/${TAG}_pipe_full_step.csv > $O/${TAG}_traffic.json && cp $O/${TAG}_traffic.json profiles/traffic.json   # bench.py below reads it
cap() {  # name regex skip count [launch indices for the source-level stall view...]
  local name=$1 re=$2 skip=$3 cnt=$4
  shift 4
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$re" -s $skip -c $cnt -o /tmp/rep/$name python tools/run_step.py 2 > /tmp/rep/$name.log 2>&1
  python tools/ncu_summary.py kernel /tmp/rep/$name.ncu-rep > $O/${TAG}_${name}_ncu_full.txt 2>/dev/null
  for sk in 0 "$@"; do python tools/ncu_summary.py stalls /tmp/rep/$name.ncu-rep "$re" $sk >> $O/${TAG}_${name}_ncu_full.txt 2>/dev/null; done
}
cap wgrad_tma wgrad_tma_kernel 23 23 16 20
cap tapgemm_fw tapgemm_fw_kernel 6 6 3
cap tapgemm_tc tapgemm_tc_kernel 37 37 14
cap gru "gru_(fwd|bwd)_kernel" 4 4 2
timeout 600 python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --workload bicrnn_infer --steps 10 --warmup 3 > $O/${TAG}_bench_bicrnn.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --precision tf32 --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_tf32.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --workload audioset_stream --precision tf32 --steps 10 --warmup 3 > $O/${TAG}_bench_stream.json 2>> $O/${TAG}_bench.err
timeout 600 python bench.py --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_b256.json 2>> $O/${TAG}_bench.err
timeout 600 python -m pytest tests/test_gpu_reference_dropin.py -m gpu -q -s > $O/${TAG}_dropin.log 2>&1
for f in bench bench_reference bench_bicrnn bench_tf32 bench_stream bench_b256; do python - <<PY
import json
try:
    d=json.load(open('$O/${TAG}_$f.json')); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$f FAILED', e)
PY
done
tail -3 $O/${TAG}_dropin.log; du -sh $O
