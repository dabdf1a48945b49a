#!/bin/bash
# multi-GPU evidence (run with gpurun --gpus N): weak / strong scaling, exact statistics, configs[2] / [4] shapes
N=${1:-8}; TAG=${2:-r2m}
O=gpurun_out; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
run() {  # name, extra args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $O/${TAG}_${name}_n$N.json 2> $O/${TAG}_${name}_n$N.err
  python - <<PY
import json
try:
    d=json.load(open('$O/${TAG}_${name}_n$N.json')); print('$name n=$N', round(d['value'],1), d['unit'], round(d['ms_per_step'],2),'ms', d['scaling'], 'gb', d['config']['global_batch'], 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('$name FAILED', e); print(open('$O/${TAG}_${name}_n$N.err').read()[-800:])
PY
}
run weak_fp32 --steps 20 --warmup 3
run strong_gb256_fp32 --global-batch 256 --steps 20 --warmup 3
run weak_fp32_exact --steps 20 --warmup 3 --sync-stats exact
run cfg2_bf16_gb256 --precision bf16 --global-batch 256 --steps 20 --warmup 3
run cfg4_stream_bf16 --workload audioset_stream --precision bf16 --steps 20 --warmup 3
run cfg4_stream_bf16_exact --workload audioset_stream --precision bf16 --steps 20 --warmup 3 --sync-stats exact
