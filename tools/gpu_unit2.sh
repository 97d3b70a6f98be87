#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fit.py tests/test_gpu_variants.py tests/test_gpu_plugin.py tests/test_gpu_surface.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/fit_time.py unit 2>&1 | grep "mode 4\|small batch"
BORE_FIT_UNIT_GENERIC=1 timeout 300 python tools/fit_time.py unit 2>&1 | grep "small batch"
