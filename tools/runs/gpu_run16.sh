#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad3.log 2>&1; cat gpurun_out/conv_layers_wgrad3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv1_thin python tools/conv_one.py 36 352 704 32 3 1 0 > gpurun_out/ncu_c1thin.log 2>&1; tail -2 gpurun_out/ncu_c1thin.log
