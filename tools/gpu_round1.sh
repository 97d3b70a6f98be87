#!/bin/bash
# one gpurun call: all GPU parity tests, smoke, bench (with cpu_baseline), reference arm, launch list, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -10 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_warp -s 25 -c 1 -o gpurun_out/prof_step_r25 -f python tools/profile_target.py cfg3 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_eval -s 3 -c 1 -o gpurun_out/prof_k2 -f python tools/profile_target.py cfg3 > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_cluster -c 1 -o gpurun_out/prof_fitc -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu4.log 2>&1
ls gpurun_out | head -30
