#!/bin/bash
# usage: bash tools/gpu_cap.sh TAG NAME REGEX SKIP COUNT [STALL_SKIP...] -- one `ncu --set full` capture, summarised on the box
TAG=$1; NAME=$2; RE=$3; SKIP=$4; CNT=$5; shift 5
O=gpurun_out
mkdir -p $O /tmp/rep
python -c "import __graft_entry__ as g; g.build()" >/dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -o /tmp/rep/$NAME python tools/run_step.py 2 > /tmp/rep/$NAME.log 2>&1
python tools/ncu_summary.py kernel /tmp/rep/$NAME.ncu-rep > $O/${TAG}_${NAME}_ncu_full.txt 2>/dev/null
for sk in 0 "$@"; do python tools/ncu_summary.py stalls /tmp/rep/$NAME.ncu-rep "$RE" $sk >> $O/${TAG}_${NAME}_ncu_full.txt 2>/dev/null; done
tail -5 /tmp/rep/$NAME.log; wc -l $O/${TAG}_${NAME}_ncu_full.txt
