#!/bin/bash
mkdir -p gpurun_out
python -c "import os; print('cores', os.cpu_count(), len(os.sched_getaffinity(0))); print(open('/sys/fs/cgroup/cpu.max').read())" > gpurun_out/host.txt 2>&1; cat gpurun_out/host.txt
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q > gpurun_out/pytest_conv.log 2>&1; echo "conv rc=$?"; tail -25 gpurun_out/pytest_conv.log
timeout 120 python -m pytest tests/test_lpg_gpu.py tests/test_heads_gpu.py -x -q > gpurun_out/pytest_lpg.log 2>&1; echo "lpg rc=$?"; tail -4 gpurun_out/pytest_lpg.log
timeout 120 python tools/lpg_sweep.py > gpurun_out/lpg_sweep.json 2>&1; cat gpurun_out/lpg_sweep.json
