#!/bin/bash
# session-2 call 1: re-verify restored tree, per-shape trace, ncu launch list of the bench command, ncu full of LPG + 2 conv shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/step_trace.py > gpurun_out/step_trace.log 2>&1; head -40 gpurun_out/step_trace.log
# launch list: skip the warm-up launches, capture > 1 step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 13000 -c 5000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/launches_bench.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lpg_ -c 4 -o gpurun_out/lpg_r8_v2 -f python tools/lpg_micro.py 8 1024 128 1 > gpurun_out/ncu_lpg.log 2>&1; echo "ncu lpg rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 --launch-skip 2 -o gpurun_out/conv_db1_3x3 -f python tools/conv_one.py 192 88 176 48 3 1 0 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 --launch-skip 2 -o gpurun_out/conv_conv1 -f python tools/conv_one.py 36 352 704 32 3 1 0 > gpurun_out/ncu_conv1.log 2>&1; echo "ncu conv1 rc=$?"
