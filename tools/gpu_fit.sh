#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_surface.py tests/test_gpu_plugin.py tests/test_gpu_batched.py -m gpu -q -x > gpurun_out/pytest_fit.log 2>&1; tail -15 gpurun_out/pytest_fit.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print([ (k['name'][:20],round(k['ms_per_step'],2),round(k['frac'],4)) for k in d['kernels']]); print(d['e2e'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
