#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --workload cfg4 > gpurun_out/bench_cfg4_2gpu.json 2> gpurun_out/bench_cfg4_2gpu.err; tail -c 900 gpurun_out/bench_cfg4_2gpu.json; tail -3 gpurun_out/bench_cfg4_2gpu.err
timeout 600 python bench.py --workload cfg4 --steps 2 --warmup 1 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; tail -c 900 gpurun_out/bench_cfg4.json; tail -3 gpurun_out/bench_cfg4.err
