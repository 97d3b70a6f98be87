#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_full.log; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu_full.log | tail -20
nproc; timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
