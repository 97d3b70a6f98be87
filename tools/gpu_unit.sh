#!/bin/bash
# K1u (unit-split cluster fit): parity tests (compile-time shapes and the run-time-shape build), timing against the
# sample-split cluster kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x > gpurun_out/pytest_unit.log 2>&1; tail -3 gpurun_out/pytest_unit.log
BORE_FIT_UNIT_GENERIC=1 timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "4 or persists" > gpurun_out/pytest_unit_g.log 2>&1; tail -3 gpurun_out/pytest_unit_g.log
timeout 300 python tools/fit_time.py unit 2>&1 | tail -12
BORE_FIT_UNIT_GENERIC=1 timeout 300 python tools/fit_time.py unit 2>&1 | grep "mode 4"
