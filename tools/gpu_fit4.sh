#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_batched.py tests/test_gpu_plugin.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/fit_time.py many 2>&1 | tail -8
