#!/bin/bash
# one gpurun call: parity tests (full log), smoke, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print(d['roofline']); print([ (k['name'][:20],round(k['ms_per_step'],2),round(k['frac'],4)) for k in d['kernels']]); print(d['e2e']); print(d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
