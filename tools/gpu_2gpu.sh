#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lbfgsb.py -m gpu -q > gpurun_out/pytest_lb.log 2>&1; grep -E "passed|failed|FAILED|agree" gpurun_out/pytest_lb.log | tail -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1200 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
