#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad7.log 2>&1; tail -20 gpurun_out/conv_layers_wgrad7.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_v2.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_v2.json
