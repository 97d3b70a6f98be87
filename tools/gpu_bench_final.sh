#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_cfg3.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print([ (k['name'][:20],round(k['ms_per_step'],2),round(k['frac'],4)) for k in d['kernels']]); print(d['e2e']); print(d.get('cpu_baseline')); print(d.get('lstm')); print(d['strong']['ms_per_step'], d['strong']['phases']); c=d.get('cfg4'); print(c['ms_per_step'], c['value'], c['phases_ms_rank0']); print(d['clocks'], d['gpu_launches'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_cfg3_reference_arm.json 2>> gpurun_out/bench.err; tail -c 600 gpurun_out/r02c_bench_cfg3_reference_arm.json
