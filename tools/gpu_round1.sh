#!/bin/bash
# one gpurun call: parity tests, smoke, bench, launch list, ncu captures of the three kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
cp MEASURED_PEAKS.json gpurun_out/ 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_warp -s 6 -c 1 -o gpurun_out/prof_step_r6 -f python tools/profile_target.py cfg3 > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_warp -s 25 -c 1 -o gpurun_out/prof_step_r25 -f python tools/profile_target.py cfg3 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval -s 3 -c 1 -o gpurun_out/prof_k2 -f python tools/profile_target.py cfg3 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -c 1 -o gpurun_out/prof_fit -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out
