#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fused_gpu.py -x -q > gpurun_out/pytest_fused.log 2>&1; echo "fused rc=$?"; tail -15 gpurun_out/pytest_fused.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_fused.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_fused.json; tail -3 gpurun_out/bench.err
timeout 600 python tools/step_profile.py > gpurun_out/step_profile_fused.log 2>&1; head -28 gpurun_out/step_profile_fused.log
