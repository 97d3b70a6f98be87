#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_plugin.py tests/test_gpu_mlp_eval.py tests/test_gpu_surface.py -m gpu -q -x > gpurun_out/pytest_b.log 2>&1; tail -30 gpurun_out/pytest_b.log
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 1 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 1200 gpurun_out/bench_cfg4.json; tail -5 gpurun_out/bench_cfg4.err
