#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_cluster -c 1 -o gpurun_out/prof_fitc -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu_fitc.log 2>&1
tail -2 gpurun_out/ncu_fitc.log
