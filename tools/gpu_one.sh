#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -s --durations=8 > gpurun_out/pytest_full.log 2>&1; tail -30 gpurun_out/pytest_full.log
