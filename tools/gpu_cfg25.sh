#!/bin/bash
mkdir -p gpurun_out
for w in cfg2 cfg5; do
timeout 600 python bench.py --workload $w --no-cfg4 --steps 5 --warmup 3 > gpurun_out/r02c_bench_$w.json 2> gpurun_out/bench_$w.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_$w.json').read().strip().splitlines()[-1])
    print('$w', d['config']['workload'], round(d['ms_per_step'],2),'ms', round(d['value']/1e6,2),'M evals/s', d['phases'], 'e2e', round(d['e2e']['ms_per_step'],2), d.get('cpu_baseline',{}).get('value'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_$w.err').read()[-1500:])
PY
done
