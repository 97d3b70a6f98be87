#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
tail -2 gpurun_out/launch_run.log
