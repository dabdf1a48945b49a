#!/bin/bash
# usage: bash tools/gpu_quick.sh TAG  -- A/B check: conv-layer + bf16 tests, fp32 and bf16 bench lines without the CPU legs
TAG=$1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_bench_config.py tests/test_gpu_tc.py -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; grep -E "passed|failed|^FAILED|^E  " gpurun_out/${TAG}_pytest.log | cut -c1-300 | tail -20
for P in tf32x3 bf16; do
timeout 600 python bench.py --precision $P --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$P.json 2> gpurun_out/${TAG}_bench_$P.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${TAG}_bench_$P.json'))
    print('$P bench', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
    print({k:v['ms'] for k,v in d['kernel_breakdown_ms'].items() if v['ms']>0.1})
except Exception as e:
    print('bench failed', e); print(open('gpurun_out/${TAG}_bench_$P.err').read()[-1500:])
PY
done
