#!/bin/bash
timeout 300 python tools/debug_gru.py > gpurun_out/r2c3_debug_gru.txt 2>&1; tail -12 gpurun_out/r2c3_debug_gru.txt
bash tools/gpu_call.sh r2c3
