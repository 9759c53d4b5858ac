#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/step_trace.py > gpurun_out/step_trace2.log 2>&1; head -40 gpurun_out/step_trace2.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_tc2.json
