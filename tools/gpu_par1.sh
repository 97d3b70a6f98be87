#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity_report.py -m gpu -q -x -s -k "cfg1 or cfg2" 2>&1 | grep -E "^cfg|passed|failed|assert" | cut -c1-900
BORE_FIT_UNIT=0 timeout 600 python -m pytest tests/test_gpu_parity_report.py -m gpu -q -x -s -k "cfg1 or cfg2" 2>&1 | grep -E "^cfg|passed|failed|assert" | cut -c1-900
