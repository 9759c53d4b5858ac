#!/bin/bash
# session-2 call 2: lean conv_tc producers + fused decoder glue: parity tests, per-layer timing, bench, step trace
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu2.log
timeout 200 python tools/conv_layers.py fwd > gpurun_out/conv_layers_fwd_s2.log 2>&1; tail -21 gpurun_out/conv_layers_fwd_s2.log
timeout 200 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad_s2.log 2>&1; tail -21 gpurun_out/conv_layers_wgrad_s2.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run2.json 2> gpurun_out/bench_s2_run2.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_s2_run2.json; tail -5 gpurun_out/bench_s2_run2.err
timeout 300 python tools/step_trace.py > gpurun_out/step_trace2.log 2>&1; head -45 gpurun_out/step_trace2.log
