#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 1 -c 1 -o gpurun_out/wgrad_thin python tools/wgrad_one.py 64 352 704 32 3 1 0 > gpurun_out/ncu_wgthin.log 2>&1; tail -2 gpurun_out/ncu_wgthin.log
