#!/bin/bash
# round-2 GPU call 1: full GPU test suite, bench (both arms), eager host profile, per-launch tensor-pipe list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
tail -5 gpurun_out/r2c1_pytest.log
PBSED_BENCH_DETAIL=gpurun_out/r2c1_detail.txt timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -c 1500 gpurun_out/r2c1_bench.json
timeout 300 python tools/profile_eager.py 5 > gpurun_out/r2c1_eager_profile.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 520 -c 270 --csv --log-file gpurun_out/r2c1_pipe.csv python tools/run_step.py 3 > gpurun_out/r2c1_ncu.log 2>&1
echo done
