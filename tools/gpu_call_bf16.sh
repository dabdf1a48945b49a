#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q -s -x > gpurun_out/${TAG}_bf16_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^E  |bf16 mode" gpurun_out/${TAG}_bf16_pytest.log | cut -c1-300 | tail -20
timeout 600 python bench.py --precision bf16 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_bf16.json'))
    print('bf16 bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'final_loss', d['config']['final_loss'])
    print({k:v['ms'] for k,v in d['kernel_breakdown_ms'].items() if v['ms']>0.1})
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench_bf16.err').read()[-1500:])
PY
timeout 600 python bench.py --workload audioset_stream --precision bf16 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_stream_bf16.json 2> gpurun_out/${TAG}_bench_stream_bf16.err; tail -c 400 gpurun_out/${TAG}_bench_stream_bf16.json; tail -c 600 gpurun_out/${TAG}_bench_stream_bf16.err
timeout 600 python bench.py --global-batch 256 --precision bf16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_gb256_bf16.json 2>> gpurun_out/${TAG}_bench_stream_bf16.err; tail -c 300 gpurun_out/${TAG}_bench_gb256_bf16.json
