#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
echo "== default" > gpurun_out/r2c4_debug_gru.txt; GRU_CASES=all timeout 300 python tools/debug_gru.py 2>&1 | grep -v Warning >> gpurun_out/r2c4_debug_gru.txt
echo "== PBSED_WG_TMA=0" >> gpurun_out/r2c4_debug_gru.txt; PBSED_WG_TMA=0 GRU_CASES=small timeout 300 python tools/debug_gru.py 2>&1 | grep -v Warning >> gpurun_out/r2c4_debug_gru.txt
echo "== PBSED_PRECISION=fp32" >> gpurun_out/r2c4_debug_gru.txt; PBSED_PRECISION=fp32 GRU_CASES=small timeout 300 python tools/debug_gru.py 2>&1 | grep -v Warning >> gpurun_out/r2c4_debug_gru.txt
echo "== memcheck" >> gpurun_out/r2c4_debug_gru.txt; GRU_CASES=small timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/debug_gru.py 2>&1 | grep -v Warning | tail -60 >> gpurun_out/r2c4_debug_gru.txt
echo "== initcheck" >> gpurun_out/r2c4_debug_gru.txt; GRU_CASES=small timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python tools/debug_gru.py 2>&1 | grep -v Warning | tail -60 >> gpurun_out/r2c4_debug_gru.txt
cut -c1-300 gpurun_out/r2c4_debug_gru.txt | tail -70
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tapgemm_fw_kernel -s 3 -c 3 -o gpurun_out/r2c4_fw python tools/run_step.py 2 > gpurun_out/r2c4_ncu_fw.log 2>&1
bash tools/gpu_call.sh r2c4
