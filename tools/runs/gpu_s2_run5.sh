#!/bin/bash
# session-2 call 5: dense-K packing + epilogue BatchNorm statistics: parity, per-layer, bench, then ncu evidence of the final kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/pytest_gpu5.log
for mode in fwd dgrad; do
  timeout 200 python tools/conv_layers.py $mode tc > gpurun_out/conv_layers_${mode}_s5.log 2>&1; echo "== $mode"; tail -21 gpurun_out/conv_layers_${mode}_s5.log
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run5.json 2> gpurun_out/bench_s2_run5.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_s2_run5.json; tail -3 gpurun_out/bench_s2_run5.err
BTS_B200_EPI_STATS=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu 2>&1 | cut -c1-200
timeout 300 python tools/step_trace.py > gpurun_out/step_trace5.log 2>&1; head -30 gpurun_out/step_trace5.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile5.log 2>&1; head -30 gpurun_out/step_profile5.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 --launch-skip 2 -o gpurun_out/conv5_v4 -f python tools/conv_one.py 896 22 44 512 3 1 0 > gpurun_out/ncu_conv5.log 2>&1; echo "ncu conv5 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 1 --launch-skip 2 -o gpurun_out/conv_db1_3x3_v4 -f python tools/conv_one.py 192 88 176 48 3 1 0 > gpurun_out/ncu_db1.log 2>&1; echo "ncu db1 rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad2_tc_kernel -c 1 --launch-skip 2 -o gpurun_out/wgrad2_conv1_v3 -f python tools/wgrad_one.py 36 352 704 32 3 1 0 > gpurun_out/ncu_wgrad2.log 2>&1; echo "ncu wgrad2 rc=$?"
