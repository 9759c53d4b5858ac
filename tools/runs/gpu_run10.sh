#!/bin/bash
mkdir -p gpurun_out
BTS_B200_WGRAD=aten timeout 600 python tools/step_profile.py > gpurun_out/step_profile_aten.log 2>&1; tail -45 gpurun_out/step_profile_aten.log
