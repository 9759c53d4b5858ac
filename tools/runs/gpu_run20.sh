#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py tests/test_fused_gpu.py -x -q 2>&1 | tail -8
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad5.log 2>&1; tail -21 gpurun_out/conv_layers_wgrad5.log
