#!/bin/bash
# session-2 call 6: final state -- parity, smoke, per-layer table vs cuDNN, full bench line (LPG roofline + CPU baseline), reference arm, traces
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu6.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for mode in fwd dgrad wgrad; do
  timeout 300 python tools/conv_layers.py $mode > gpurun_out/conv_layers_${mode}_s6.log 2>&1; echo "== $mode"; tail -21 gpurun_out/conv_layers_${mode}_s6.log
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_s2_final.json 2> gpurun_out/bench_s2_final.err; echo "bench rc=$?"; cat gpurun_out/bench_s2_final.json; tail -3 gpurun_out/bench_s2_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_s2_reference.json 2> gpurun_out/bench_s2_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_s2_reference.json
timeout 300 python tools/step_trace.py > gpurun_out/step_trace6.log 2>&1; head -12 gpurun_out/step_trace6.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile6.log 2>&1; head -24 gpurun_out/step_profile6.log
