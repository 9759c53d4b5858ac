#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/debug_wgrad.py > gpurun_out/debug_wgrad.log 2>&1; cat gpurun_out/debug_wgrad.log | tail -12
timeout 300 python tools/conv_layers.py fwd > gpurun_out/conv_layers_fwd.log 2>&1; cat gpurun_out/conv_layers_fwd.log
