#!/bin/bash
# session-2 call 3: async-producer conv kernel + 16-warp wgrad producers: parity, A/B per-layer timing, plan sweep, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu3.log
for mode in fwd dgrad wgrad; do
  timeout 200 python tools/conv_layers.py $mode tc > gpurun_out/conv_layers_${mode}_s3.log 2>&1; echo "== $mode (async default)"; tail -21 gpurun_out/conv_layers_${mode}_s3.log
done
echo "== fwd, register-prefetch kernel (BTS_B200_CONV_ASYNC=0)"
BTS_B200_CONV_ASYNC=0 timeout 200 python tools/conv_layers.py fwd tc 2>&1 | tail -21
for plan in 2,6,4 3,3,6 2,4,8 3,4,4; do
  echo "== fwd plan sa,sb,ring=$plan"
  BTS_B200_CONV_PLAN=$plan timeout 200 python tools/conv_layers.py fwd tc 2>&1 | tail -21
done
timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu > gpurun_out/bench_s2_run3.json 2> gpurun_out/bench_s2_run3.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/bench_s2_run3.json; tail -3 gpurun_out/bench_s2_run3.err
BTS_B200_CONV_ASYNC=0 timeout 400 python bench.py --steps 5 --warmup 3 --no-lpg --no-cpu 2>&1 | cut -c1-200
timeout 300 python tools/step_trace.py > gpurun_out/step_trace3.log 2>&1; head -50 gpurun_out/step_trace3.log
