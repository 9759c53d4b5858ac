#!/bin/bash
# session-2 call 4: MMA-count reduction (N-stacked hi/lo, 256-wide tiles, tap-packed wgrad2), pointwise wgrad kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu4.log
for mode in fwd dgrad wgrad; do
  timeout 200 python tools/conv_layers.py $mode tc > gpurun_out/conv_layers_${mode}_s4.log 2>&1; echo "== $mode"; tail -21 gpurun_out/conv_layers_${mode}_s4.log
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run4.json 2> gpurun_out/bench_s2_run4.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_s2_run4.json; tail -3 gpurun_out/bench_s2_run4.err
timeout 300 python tools/step_trace.py > gpurun_out/step_trace4.log 2>&1; head -60 gpurun_out/step_trace4.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile4.log 2>&1; head -45 gpurun_out/step_profile4.log
