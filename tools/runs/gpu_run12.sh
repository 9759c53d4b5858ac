#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py > gpurun_out/step_trace.log 2>&1; head -64 gpurun_out/step_trace.log
