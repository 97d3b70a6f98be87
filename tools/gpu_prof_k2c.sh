#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval_kernel -s 30 -c 1 -o gpurun_out/prof_k2_r02c -f python tools/profile_target.py cfg3 > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
