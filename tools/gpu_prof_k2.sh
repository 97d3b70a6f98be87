#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval -s 3 -c 1 -o gpurun_out/prof_k2 -f python tools/profile_target.py cfg3 > gpurun_out/ncu_k2.log 2>&1
tail -2 gpurun_out/ncu_k2.log
