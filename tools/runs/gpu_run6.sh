#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/conv5_v2 python tools/conv_one.py 896 22 44 512 3 1 0 > gpurun_out/ncu_conv5.log 2>&1; tail -2 gpurun_out/ncu_conv5.log
