#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "4 or persists" 2>&1 | tail -5
BORE_FIT_UNIT_GENERIC=1 timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "4 or persists" 2>&1 | tail -5
timeout 300 python tools/fit_time.py trace 2>&1 | tail -12
