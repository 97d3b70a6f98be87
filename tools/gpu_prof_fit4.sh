#!/bin/bash
mkdir -p gpurun_out
python tools/profile_cfg4_fit.py 4096 2 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -c 1 -o gpurun_out/prof_fit4 -f python tools/profile_cfg4_fit.py 1184 > gpurun_out/ncu_fit4.log 2>&1
tail -2 gpurun_out/ncu_fit4.log
