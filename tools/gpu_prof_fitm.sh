#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_mma -c 1 -o gpurun_out/prof_fitm -f python tools/profile_fit_mma.py 3 4 > gpurun_out/ncu_fitm.log 2>&1
tail -2 gpurun_out/ncu_fitm.log
