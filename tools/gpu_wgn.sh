#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
echo "fp32";  python tools/bench_wgrad_narrow.py 1
echo "fp32 nopair";  PBSED_WG_PAIR=0 python tools/bench_wgrad_narrow.py 1
echo "bf16";  python tools/bench_wgrad_narrow.py 3 bf16
