#!/bin/bash
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 600 python tools/debug_gru2.py 2>&1 | grep -v Warn > gpurun_out/r2c5_debug_gru2.txt
echo "== PBSED_TC_2CTA=0" >> gpurun_out/r2c5_debug_gru2.txt
PBSED_TC_2CTA=0 timeout 600 python tools/debug_gru2.py 2>&1 | grep -v Warn | head -12 >> gpurun_out/r2c5_debug_gru2.txt
cut -c1-200 gpurun_out/r2c5_debug_gru2.txt
