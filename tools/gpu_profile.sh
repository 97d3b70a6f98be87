#!/bin/bash
# one gpurun call: parity tests, smoke, launch list, ncu captures of the stepper and K2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_step -s 10 -c 1 -o gpurun_out/prof_step_r10 -f python tools/profile_target.py cfg3 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_step -s 60 -c 1 -o gpurun_out/prof_step_r60 -f python tools/profile_target.py cfg3 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval -s 3 -c 1 -o gpurun_out/prof_k2 -f python tools/profile_target.py cfg3 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -c 1 -o gpurun_out/prof_fit -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out
