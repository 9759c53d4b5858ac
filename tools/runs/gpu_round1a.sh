#!/bin/bash
# first GPU pass: parity tests, smoke, bench, launch list, ncu full of the LPG kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
for r in 8 4 2; do timeout 120 python tools/lpg_micro.py $r 1024 128 5; timeout 120 python tools/lpg_micro.py $r 256 16 20; done > gpurun_out/lpg_micro.log 2>&1; cat gpurun_out/lpg_micro.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lpg_ -s 4 -c 2 -o gpurun_out/lpg_r8_full python tools/lpg_micro.py 8 1024 128 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --no-lpg --no-cpu > gpurun_out/ncu_step.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out
