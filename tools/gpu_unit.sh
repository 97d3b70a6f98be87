#!/bin/bash
# K1u (unit-split cluster fit): parity tests, timing against the sample-split cluster kernel, ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "4 or persists" > gpurun_out/pytest_unit.log 2>&1; tail -3 gpurun_out/pytest_unit.log
timeout 300 python tools/fit_time.py unit 2>&1 | tail -12
BORE_FIT_UNIT_ASYNC=0 timeout 300 python tools/fit_time.py unit 2>&1 | grep "mode 4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_unit -c 1 -o gpurun_out/prof_fitu -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu_fitu.log 2>&1
tail -1 gpurun_out/ncu_fitu.log
