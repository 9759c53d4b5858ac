#!/bin/bash
# session-2 call 8: wgrad_tc with 256-wide tiles (A/B against 128-wide), full parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu8.log
timeout 200 python tools/conv_layers.py wgrad tc 2>&1 | tail -21
echo "== wgrad, 128-wide tiles (BTS_B200_WGRAD_WIDE=0)"
BTS_B200_WGRAD_WIDE=0 timeout 200 python tools/conv_layers.py wgrad tc 2>&1 | tail -21
timeout 400 python bench.py --steps 10 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run8.json 2> gpurun_out/bench_s2_run8.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_s2_run8.json; tail -3 gpurun_out/bench_s2_run8.err
BTS_B200_WGRAD_WIDE=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-lpg --no-cpu 2>&1 | cut -c1-200
