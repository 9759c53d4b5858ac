#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
BTS_B200_WGRAD=aten timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_tc_aten.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_tc_aten.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_tc_tc.json 2> gpurun_out/bench_tc.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_tc_tc.json
