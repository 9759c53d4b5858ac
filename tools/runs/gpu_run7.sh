#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/conv_layers.py fwd > gpurun_out/conv_layers_fwd3.log 2>&1; cat gpurun_out/conv_layers_fwd3.log
timeout 300 python tools/conv_layers.py wgrad > gpurun_out/conv_layers_wgrad2.log 2>&1; cat gpurun_out/conv_layers_wgrad2.log
timeout 300 python tools/conv_layers.py dgrad > gpurun_out/conv_layers_dgrad.log 2>&1; cat gpurun_out/conv_layers_dgrad.log
