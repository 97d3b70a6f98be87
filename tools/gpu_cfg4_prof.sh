#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 > gpurun_out/launch_cfg4.log 2>&1
tail -2 gpurun_out/launch_cfg4.log | cut -c1-300
