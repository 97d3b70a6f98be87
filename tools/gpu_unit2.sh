#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "4 or persists" 2>&1 | tail -2
timeout 300 python tools/fit_time.py unit 2>&1 | grep "mode 4"
