#!/bin/bash
# session-2 call 9: software-pipelined epilogue; parity + per-layer fwd/dgrad + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu9.log
for mode in fwd dgrad; do echo "== $mode"; timeout 200 python tools/conv_layers.py $mode tc 2>&1 | tail -21; done
timeout 400 python bench.py --steps 10 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run9.json 2> gpurun_out/bench_s2_run9.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_s2_run9.json; tail -3 gpurun_out/bench_s2_run9.err
