#!/bin/bash
timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'tools'); sys.argv=['x','none']
import importlib.util, numpy as np
spec=importlib.util.spec_from_file_location('ft','tools/fit_time.py'); ft=importlib.util.module_from_spec(spec); spec.loader.exec_module(ft)
cfg3 = ([50, 64, 64, 64, 1], ["relu", "relu", "relu", "sigmoid"])
cfg5 = ([8, 32, 32, 32, 1], ["elu", "elu", "elu", "linear"])
for mode in (4,):
    print("cfg3 mode", mode, ft.time_fit(*cfg3, 1, 2000, 31, mode), flush=True)
    print("cfg5 mode", mode, ft.time_fit(*cfg5, 1, 500, 125, mode), flush=True)
PY
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -x -k "cfg3 and 4" 2>&1 | tail -2
