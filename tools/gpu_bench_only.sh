#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print(d['e2e']); print(d.get('strong') and (d['strong']['ms_per_step'], d['strong']['phases'])); c=d.get('cfg4'); print(c and (c['ms_per_step'], c['value'], c['phases_ms_rank0']))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
