#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
for K in 32 16; do
echo "fp32 kr=$K";  PBSED_WG_KR=$K python tools/bench_wgrad_narrow.py 1
echo "bf16 kr=$K";  PBSED_WG_KR=$K python tools/bench_wgrad_narrow.py 3 bf16
done
