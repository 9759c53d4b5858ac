#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
timeout 120 python tools/lpg_sweep.py > gpurun_out/lpg_sweep.json 2>&1; cat gpurun_out/lpg_sweep.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
