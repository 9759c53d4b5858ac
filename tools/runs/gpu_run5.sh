#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/conv_layers.py fwd > gpurun_out/conv_layers_fwd2.log 2>&1; cat gpurun_out/conv_layers_fwd2.log
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad.log 2>&1; cat gpurun_out/conv_layers_wgrad.log
