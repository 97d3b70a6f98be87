#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_step -s 12 -c 1 -o gpurun_out/prof_step_r12 -f python tools/profile_target.py cfg3 > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
