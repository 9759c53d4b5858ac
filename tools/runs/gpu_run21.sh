#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_fused_gpu.py -x -q 2>&1 | tail -5
timeout 300 python tools/conv_layers.py fwd > gpurun_out/conv_layers_fwd6.log 2>&1; tail -20 gpurun_out/conv_layers_fwd6.log
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad6.log 2>&1; tail -20 gpurun_out/conv_layers_wgrad6.log
