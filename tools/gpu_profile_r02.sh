#!/bin/bash
# round-2 profile set: launch list of one cfg-3 BO iteration + full captures of the stepper (round 25),
# the fused kernel (8,192 starts, the regime the dispatcher uses it in) and the fused tail (resume)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_cfg3.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
tail -1 gpurun_out/launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_warp -s 25 -c 1 -o gpurun_out/prof_k3_r02 -f python tools/profile_target.py cfg3 > gpurun_out/ncu_k3.log 2>&1; tail -1 gpurun_out/ncu_k3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_fused -c 1 -o gpurun_out/prof_k3f_tail_r02 -f python tools/profile_target.py cfg3 > gpurun_out/ncu_k3f_tail.log 2>&1; tail -1 gpurun_out/ncu_k3f_tail.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_fused -c 1 -o gpurun_out/prof_k3f_8192_r02 -f python tools/profile_target.py cfg3 8192 > gpurun_out/ncu_k3f.log 2>&1; tail -1 gpurun_out/ncu_k3f.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval_kernel -s 30 -c 1 -o gpurun_out/prof_k2_r02 -f python tools/profile_target.py cfg3 > gpurun_out/ncu_k2.log 2>&1; tail -1 gpurun_out/ncu_k2.log
