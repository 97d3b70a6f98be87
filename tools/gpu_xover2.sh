#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_lbfgsb.py tests/test_gpu_surface.py -m gpu -q -x 2>&1 | tail -2
for S in 12288 16384; do timeout 300 python tools/fused_time.py cfg3 $S 2 4 2>&1 | tail -1 | cut -c1-120; done
