#!/bin/bash
# evidence pass (1 GPU): launch list with tensor-pipe + DRAM bytes of a full eager step, --set full captures of the
# dominant kernels, bench lines of the other single-GPU workloads.   usage: bash tools/gpu_profile.sh TAG
TAG=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 500 -c 330 --csv --log-file gpurun_out/${TAG}_pipe_full_step.csv python tools/run_step.py 3 > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_tma_kernel -s 19 -c 19 -o gpurun_out/${TAG}_wgrad_tma python tools/run_step.py 2 > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_fw_kernel -s 6 -c 6 -o gpurun_out/${TAG}_fw python tools/run_step.py 2 > gpurun_out/${TAG}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tapgemm_tc_kernel" -s 37 -c 37 -o gpurun_out/${TAG}_tc python tools/run_step.py 2 > gpurun_out/${TAG}_ncu4.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload bicrnn_infer --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_bicrnn.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --precision tf32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_tf32.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload audioset_stream --precision tf32 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_stream.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --batch 256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_b256.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python -m pytest tests/test_gpu_reference_dropin.py -m gpu -q -s > gpurun_out/${TAG}_dropin.log 2>&1
for f in bench bench_reference bench_bicrnn bench_tf32 bench_stream bench_b256; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_$f.json')); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$f FAILED', e)
PY
done
tail -3 gpurun_out/${TAG}_dropin.log; tail -c 600 gpurun_out/${TAG}_bench.err
