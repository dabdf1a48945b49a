#!/bin/bash
# usage: bash tools/gpu_call.sh TAG [pytest-args...]   -- GPU tests (all failures listed), bench, per-launch ncu list
TAG=$1; shift
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -s "$@" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/${TAG}_pytest.log | tail -25
PBSED_BENCH_DETAIL=gpurun_out/${TAG}_detail.txt timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench.json'))
    print('bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'eager', d.get('eager_step_ms'), 'parity', d.get('parity',{}).get('logit_max_abs'))
    print({k:v['ms'] for k,v in d['kernel_breakdown_ms'].items() if v['ms']>0.1})
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench.err').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 540 -c 270 --csv --log-file gpurun_out/${TAG}_pipe.csv python tools/run_step.py 3 > gpurun_out/${TAG}_ncu.log 2>&1
echo done
