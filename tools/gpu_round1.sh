#!/bin/bash
# one gpurun call: parity tests, bench, launch list, ncu captures of the three kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_step -s 6 -c 1 -o gpurun_out/prof_step_r6 -f python tools/profile_target.py cfg3 > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbfgsb_step -s 25 -c 1 -o gpurun_out/prof_step_r25 -f python tools/profile_target.py cfg3 > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_eval -s 3 -c 1 -o gpurun_out/prof_k2 -f python tools/profile_target.py cfg3 > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out
